"""Quick device timing of the MPC cycle (development helper; bench.py is the contract)."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import qm_door_b200 as q
from qm_door_b200 import workload

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hor = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 6
W = workload.Workload(B, horizon=hor, dt=0.01)
ctx = q.MpcContext(W.model, W.problem, W.solver, B)
print("device bytes %.2f GB, NMAX %d" % (ctx.device_bytes / 1e9, W.solver.max_nodes))
ctx.set_profiling(True)
out = ctx.alloc_outputs()
for c in range(cycles):
    t = time.time()
    ctx.cycle(np.full(B, 0.01 * c), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x, out=out)
    dt = time.time() - t
    kt = ctx.kernel_times(reset=True)
    st = out["status"]
    print("cycle %d wall %.2f ms  status nonzero %d  alpha mean %.3f iters max %d  n %d..%d" % (
        c, dt * 1e3, int((st & ~32 != 0).sum()), out["info"][:, 0].mean(), int(out["info"][:, 11].max()), out["n"].min(), out["n"].max()))
    print("   ", {k: "%.3f ms x%d" % v for k, v in kt.items() if v[1]})
ctx.close()
