import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import qm_door_b200 as q
from qm_door_b200 import workload
B=65536
W=workload.WbcWorkload(B)
ctx=q.WbcContext(W.model,W.wbc,B)
for mode in (15, 9, 0):
    pick = int(np.nonzero(W.mode == mode)[0][0])
    rep = lambda a: np.repeat(a[pick:pick + 1], B, axis=0)
    ctx.reset()
    ctx.update(rep(W.x_des), rep(W.u_last), rep(W.rbd), rep(W.mode), W.period, W.time)
    ctx.kernel_time(reset=True)
    for i in range(2): ctx.update(rep(W.x_des), rep(W.u_des), rep(W.rbd), rep(W.mode), W.period, W.time)
    ms,n=ctx.kernel_time()
    print("mode", mode, "identical solves: k_wbc ms per batch", ms/n)
