"""Development aid: wall-clock latency of a whole-body-control batch through the host interface (host buffers in, command out) for
small batches, kernel sequence against the single kernel (QMB200_WBC_SPLIT)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qm_door_b200 as q
from qm_door_b200 import workload
for B in (1, 16, 256, 2048):
    W = workload.WbcWorkload(B, seed=3)
    row = []
    for split in ("1", "0"):
        os.environ["QMB200_WBC_SPLIT"] = split
        ctx = q.WbcContext(W.model, W.wbc, B)
        for _ in range(5): ctx.update(W.x_des, W.u_last, W.rbd, W.mode, W.period, W.time)
        ts = []
        for i in range(50):
            t0 = time.perf_counter()
            ctx.update(W.x_des, W.u_des if i % 2 == 0 else W.u_last, W.rbd, W.mode, W.period, W.time)
            ts.append(time.perf_counter() - t0)
        ctx.close()
        row.append(1e3 * float(np.median(ts)))
    print("B = %5d: kernel sequence %.3f ms, single kernel %.3f ms per call (median of 50, host buffers)" % (B, row[0], row[1]))
