import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qm_door_b200 as q
from qm_door_b200 import workload
W=workload.WbcWorkload(65536)
ctx=q.WbcContext(W.model,W.wbc,W.B)
ctx.update(W.x_des,W.u_last,W.rbd,W.mode,W.period,W.time)
ctx.kernel_time(reset=True)
for i in range(3): ctx.update(W.x_des,W.u_des if i%2==0 else W.u_last,W.rbd,W.mode,W.period,W.time)
ms,n=ctx.kernel_time()
print("k_wbc ms per batch", ms/n, "solves/s", 65536/(ms/n*1e-3))
