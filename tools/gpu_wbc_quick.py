"""Quick device timing of the WBC batch (development helper)."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import qm_door_b200 as q
from qm_door_b200 import workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
W = workload.WbcWorkload(B)
ctx = q.WbcContext(W.model, W.wbc, B)
ctx.update(W.x_des, W.u_last, W.rbd, W.mode, W.period, W.time)
ctx.kernel_time(reset=True)
for it in range(3):
    t = time.time()
    cmd, st = ctx.update(W.x_des, W.u_des, W.rbd, W.mode, W.period, W.time)
    wall = time.time() - t
    ms, n = ctx.kernel_time(reset=True)
    print("iter %d wall %.2f ms kernel %.2f ms -> %.0f solves/s (kernel), status counts %s" % (it, wall * 1e3, ms, B / (ms * 1e-3), np.bincount(st)))
ctx.close()
