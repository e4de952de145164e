"""Development aid: one GPU cycle against the CPU port, per-node differences."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qm_door_b200 as q
from qm_door_b200 import workload
from oracle import abi_fill
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.12
W = workload.Workload(B, horizon=hor, dt=0.01)
ctx = q.MpcContext(W.model, W.problem, W.solver, B)
cp = abi_fill.CPort(W.model, W.problem, W.solver, B, threads=4)
for c in range(2):
    t0 = np.full(B, 0.01 * c)
    out = ctx.cycle(t0, W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
    ref = cp.cycle(t0, W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
    for b in range(B):
        n = out["n"][b]
        dx = np.abs(out["x"][b, :n] - ref["x"][b, :n]).max(axis=1)
        du = np.abs(out["u"][b, :n] - ref["u"][b, :n]).max(axis=1)
        print("cycle", c, "b", b, "status", out["status"][b], ref["status"][b], "alpha", out["info"][b, 0], ref["info"][b, 0], "modes", out["mode"][b, :n])
        print("   dx", np.array2string(dx, precision=1), "\n   du", np.array2string(du, precision=1))
ctx.close(); cp.close()
