"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck / synccheck) under gpurun:
   compute-sanitizer --tool racecheck python tools/sanitize_case.py
Two warm-started config-2-shaped cycles (N = 100, trot, B = 6, two SQP iterations so k_step runs; one problem is given a
perturbation large enough to backtrack), feedback gains, policy evaluation, state conversion, targets, a WBC batch covering the
16 contact patterns and an actuator tick."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qm_door_b200 as q  # noqa: E402
from qm_door_b200 import workload  # noqa: E402


def main():
    B = 6
    W = workload.Workload(B, horizon=1.0, dt=0.01, seed=123)
    W.solver.sqp_iterations = 2
    W.x0[0, 0:6] += 0.4
    ctx = q.MpcContext(W.model, W.problem, W.solver, B)
    for c in range(2):
        out = ctx.cycle(np.full(B, 0.01 * c), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
    print("mpc status", out["status"], "alpha", out["info"][:, 0], "sqp iterations", out["info"][:, 13])
    ctx.feedback_gains()
    ctx.evaluate_policy(np.full(B, 0.015))
    ctx.evaluate_feedback_policy(np.full(B, 0.015), W.x0)
    WW = workload.WbcWorkload(32, seed=5)
    WW.mode[:16] = np.arange(16)
    x_obs = ctx.rbd_to_state(WW.rbd[:B], np.zeros(B))
    td = q.load_targets()
    ctx.targets(td, 0, np.tile([0.3, 0.0, 0.0, 0.1, 0, 0, 0], (B, 1)), np.zeros(B), x_obs, np.tile([0.6, 0, 0.8, 0, 0, 0, 1.0], (B, 1)),
                np.tile([0.6, 0, 0.8, 0, 0, 0, 1.0], (B, 1)))
    wctx = q.WbcContext(WW.model, WW.wbc, WW.B)
    wctx.update(WW.x_des, WW.u_last, WW.rbd, WW.mode, WW.period, WW.time)
    cmd, st = wctx.update(WW.x_des, WW.u_des, WW.rbd, WW.mode, WW.period, WW.time)
    print("wbc status", st)
    ad = q.actuator_defaults()
    wctx.actuator(ad, np.full(WW.B, 2_000_000, dtype=np.int64), 1_000_000, np.full(WW.B, 11.0), WW.x_des, WW.u_des, cmd,
                  WW.rbd[:, 6:24], WW.rbd[:, 30:48])
    ctx.close()
    wctx.close()
    print("sanitize case done")


if __name__ == "__main__":
    main()
