"""Development aid: per-phase cycle counts of CTA 0 of k_wbc (library built with -DQM_PHASE_TIMING, Makefile target dbg).
Every solve of the batch is given the inputs of solve `pick`, so CTA 0's time line is that of a known contact pattern."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QMB200_WBC_SPLIT", "0")      # cycle counts of the single kernel k_wbc (one CTA = one whole solve)
os.environ["QMB200_LIB_PATH"] = os.path.join(ROOT, "qm_door_b200", "libqmb200_dbg.so")
import qm_door_b200 as q
from qm_door_b200 import workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
names = {33: "dynamics", 34: "tasks", 35: "L0 build", 36: "L0 householder", 37: "L0 backsub+residual+decide", 38: "kernel basis 0",
         39: "L1 prep", 40: "GI householder + J", 41: "GI iterate", 42: "kernel basis 1 + Z1", 43: "L2 prep", 45: "tail"}
sub = {46: "gi violation + select", 47: "gi d", 48: "gi zd + backsub", 49: "gi step lengths", 50: "gi z, u", 51: "gi add", 52: "gi drop",
       53: "hh inner products", 54: "hh scalars", 55: "hh update", 56: "hh column rewrite"}
W = workload.WbcWorkload(B)
ctx = q.WbcContext(W.model, W.wbc, B)
L = q.lib()
t = (C.c_ulonglong * 64)()
for mode in (15, 9, 6, 0, 7):
    pick = int(np.nonzero(W.mode == mode)[0][0])
    rep = lambda a: np.repeat(a[pick:pick + 1], B, axis=0)
    ctx.reset()
    ctx.update(rep(W.x_des), rep(W.u_last), rep(W.rbd), rep(W.mode), W.period, W.time)
    L.qmb200_debug_ticks(t, 1)
    ctx.update(rep(W.x_des), rep(W.u_des), rep(W.rbd), rep(W.mode), W.period, W.time)
    L.qmb200_debug_ticks(t, 1)
    tot = sum(int(t[i]) for i in names)
    print("mode", mode, "total cycles", tot, {names[i]: int(t[i]) for i in names if t[i]})
    print("   inside the iteration:", {sub[i]: int(t[i]) for i in sub if t[i]})
ctx.close()
