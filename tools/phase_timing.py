"""Development aid: per-phase cycle counts of CTA 0 (library built with -DQM_PHASE_TIMING, see Makefile target dbg)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QMB200_LIB_PATH"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qm_door_b200", "libqmb200_dbg.so")
import qm_door_b200 as q
from qm_door_b200 import workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = workload.Workload(B, horizon=1.0, dt=0.01)
ctx = q.MpcContext(W.model, W.problem, W.solver, B)
out = ctx.alloc_outputs()
L = q.lib()
t = (C.c_ulonglong * 32)()
for c in range(4):
    ctx.cycle(np.full(B, 0.01 * c), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x, out=out)
    L.qmb200_debug_ticks(t, 1)
    print("cycle", c, {i: int(t[i]) for i in range(32) if t[i]})
ctx.close()
