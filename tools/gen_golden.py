"""Generate golden input/output vectors of the MPC cycle from the NumPy oracle (run in the build container).
Writes tests/golden/mpc_cycle_*.npz; the GPU tests compare the CUDA path against these without running the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import config, gait as G, scenarios, sqp  # noqa: E402

CASES = [  # name, gait, horizon, dt, B, cycles, seed
    ("trot_n20", "trot", 0.2, 0.01, 3, 3, 20261017),                 # BASELINE config 1 shape (N = 20)
    ("flying_trot_n12", "flying_trot", 0.12, 0.01, 2, 2, 11),
    ("static_walk_dt15", "static_walk", 0.15, 0.015, 2, 2, 12),      # reference's own dt (task.info:79)
    ("stance_n10", "stance", 0.1, 0.01, 1, 2, 13),
]


def main():
    m, P = config.load_default()
    tt, ts = scenarios.standing_target(m, P)
    for name, gait, hor, dt, B, cycles, seed in CASES:
        x0s, phase = scenarios.perturbed_states(m, P, B, seed)
        period = P.gaits[gait]["times"][-1]
        EM = 40
        events = np.full((B, EM), 1e30); modes = np.full((B, EM + 1), 15, dtype=np.int32); nev = np.zeros(B, dtype=np.int32)
        outs = []
        for b in range(B):
            ev, md = G.tile_schedule(P.gaits[gait], -np.ceil(1.0 / period) * period - phase[b], 1.2)
            assert len(ev) <= EM, len(ev)
            events[b, :len(ev)] = ev; modes[b, :len(md)] = md; nev[b] = len(ev)
            prob = sqp.MpcProblem(m, P, ev, md, tt, ts, horizon=hor, dt=dt)
            for c in range(cycles):
                tout, xs, us, info = sqp.mpc_cycle(prob, 0.01 * c, x0s[b])
                outs.append((b, c, tout, xs, us, info))
        nmax = max(len(o[2]) for o in outs)
        T = np.zeros((cycles, B, nmax)); X = np.zeros((cycles, B, nmax, 30)); U = np.zeros((cycles, B, nmax, 30))
        NN = np.zeros((cycles, B), dtype=np.int32); MD = np.zeros((cycles, B, nmax), dtype=np.int32)
        AL = np.zeros((cycles, B)); PERF = np.zeros((cycles, B, 7))
        for b, c, tout, xs, us, info in outs:
            n = len(tout)
            T[c, b, :n], X[c, b, :n], U[c, b, :n], NN[c, b], MD[c, b, :n] = tout, xs, us, n, info["modes"]
            AL[c, b] = info["alpha"]
            PERF[c, b] = [info["armijo"], info["base"]["merit"], info["base"]["dyn"], info["base"]["eq"],
                          info["new"]["merit"], info["new"]["dyn"], info["new"]["eq"]]
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mpc_cycle_%s.npz" % name), gait=gait, horizon=hor, dt=dt,
                            x0=x0s, events=events, modes=modes, nevents=nev, target_t=np.tile(tt, (B, 1)),
                            target_x=np.tile(ts, (B, 1, 1)), t=T, x=X, u=U, n=NN, mode=MD, alpha=AL, perf=PERF)
        print(name, "nodes", NN.max(), "alpha", AL.ravel())


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def backtracking_golden():
    """A case where the filter line search rejects the full step (alpha = 0.5): problems 34, 66, 94 of the seed-123 batch
    (horizon 0.3 s, trot), cycles 0..3 with warm start. Found with the CPU port, generated with the NumPy oracle."""
    from qm_door_b200 import workload
    m, P = config.load_default()
    W = workload.Workload(256, horizon=0.3, dt=0.01, seed=123)
    pick = [34, 66, 94]
    cycles = 4
    outs = []
    for b in pick:
        ne = W.nevents[b]
        prob = sqp.MpcProblem(m, P, W.events[b, :ne], W.modes[b, :ne + 1], W.target_t[b], W.target_x[b], horizon=0.3, dt=0.01)
        for c in range(cycles):
            outs.append((pick.index(b), c) + sqp.mpc_cycle(prob, 0.01 * c, W.x0[b]))
    B = len(pick)
    nmax = max(len(o[2]) for o in outs)
    T = np.zeros((cycles, B, nmax)); X = np.zeros((cycles, B, nmax, 30)); U = np.zeros((cycles, B, nmax, 30))
    NN = np.zeros((cycles, B), dtype=np.int32); MD = np.zeros((cycles, B, nmax), dtype=np.int32)
    AL = np.zeros((cycles, B)); PERF = np.zeros((cycles, B, 7))
    for b, c, tout, xs, us, info in outs:
        n = len(tout)
        T[c, b, :n], X[c, b, :n], U[c, b, :n], NN[c, b], MD[c, b, :n] = tout, xs, us, n, info["modes"]
        AL[c, b] = info["alpha"]
        PERF[c, b] = [info["armijo"], info["base"]["merit"], info["base"]["dyn"], info["base"]["eq"],
                      info["new"]["merit"], info["new"]["dyn"], info["new"]["eq"]]
    EM = 40
    events = np.full((B, EM), 1e30); modes = np.full((B, EM + 1), 15, dtype=np.int32)
    events[:, :W.events.shape[1]] = W.events[pick]; modes[:, :W.modes.shape[1]] = W.modes[pick]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mpc_cycle_trot_backtracking.npz"), gait="trot", horizon=0.3, dt=0.01,
                        x0=W.x0[pick], events=events, modes=modes, nevents=W.nevents[pick], target_t=W.target_t[pick],
                        target_x=W.target_x[pick], t=T, x=X, u=U, n=NN, mode=MD, alpha=AL, perf=PERF)
    print("backtracking golden: alpha", AL)
    assert (AL < 1.0).any()


def wbc_golden():
    """WBC golden vectors: 48 solves of config-5 style inputs (all 16 contact patterns, both task stacks, t < 10 s stack)."""
    from oracle import wbc
    from qm_door_b200 import workload
    m, P = config.load_default()
    W = workload.WbcWorkload(48, seed=20261020)
    W.mode[:16] = np.arange(16)
    W.time[32:40] = 5.0                      # arm-joint tracking stack (HierarchicalWbc.cpp:32-36)
    variant = np.zeros(48, dtype=np.int32)
    variant[40:] = 1                         # HierarchicalMpcWbc stack
    cmd = np.zeros((48, 54))
    iters = np.zeros((48, 3), dtype=np.int32)
    for b in range(48):
        O = wbc.Wbc(m, P, mpc_variant=bool(variant[b]))
        O.input_last = W.u_last[b].copy()
        cmd[b], dbg = O.update(W.x_des[b], W.u_des[b], W.rbd[b], int(W.mode[b]), W.period[b], W.time[b], return_debug=True)
        iters[b] = [l.iterations for l in dbg["levels"]]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wbc_config5.npz"), x_des=W.x_des, u_des=W.u_des, rbd=W.rbd,
                        mode=W.mode, period=W.period, time=W.time, u_last=W.u_last, variant=variant, cmd=cmd, iters=iters)
    print("wbc golden: active-set iterations per level (max)", iters.max(0), "solves with active constraints", int((iters.sum(1) > 0).sum()))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "wbc":
    wbc_golden()
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "backtracking":
    backtracking_golden()
