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
    """WBC golden vectors: 72 solves of config-5 style inputs (all 16 contact patterns, both task stacks of the reference, the
    t < 10 s stack, and 24 solves of the synthetic six-level stack over all 16 contact patterns)."""
    from oracle import wbc
    from qm_door_b200 import workload
    m, P = config.load_default()
    NW = 72
    W = workload.WbcWorkload(NW, seed=20261020)
    W.mode[:16] = np.arange(16)
    W.mode[48:64] = np.arange(16)
    W.time[32:40] = 5.0                      # arm-joint tracking stack (HierarchicalWbc.cpp:32-36)
    variant = np.zeros(NW, dtype=np.int32)
    variant[40:48] = 1                       # HierarchicalMpcWbc stack
    variant[48:] = 2                         # six-level synthetic stack
    cmd = np.zeros((NW, 54))
    iters = np.zeros((NW, 3), dtype=np.int32)
    relaxed = np.zeros(NW, dtype=np.int32)    # a level needed the oracle's wider degeneracy margins: the solve is not well posed
    for b in range(NW):
        O = wbc.Wbc(m, P, mpc_variant=int(variant[b]))
        O.input_last = W.u_last[b].copy()
        try:
            cmd[b], dbg = O.update(W.x_des[b], W.u_des[b], W.rbd[b], int(W.mode[b]), W.period[b], W.time[b], return_debug=True)
            iters[b] = [l.iterations for l in dbg["levels"]]
            relaxed[b] = int(any(getattr(l, "relaxed", False) for l in dbg["all_levels"]))
        except RuntimeError:        # degenerate inherited rows beyond the oracle solver's tolerance: recorded as not available (NaN)
            cmd[b] = np.nan
            iters[b] = -1
    print("not available:", np.nonzero(np.isnan(cmd[:, 0]))[0], "variants", variant[np.isnan(cmd[:, 0])])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wbc_config5.npz"), x_des=W.x_des, u_des=W.u_des, rbd=W.rbd,
                        mode=W.mode, period=W.period, time=W.time, u_last=W.u_last, variant=variant, cmd=cmd, iters=iters, relaxed=relaxed)
    print("relaxed per variant:", [int(relaxed[variant == v].sum()) for v in (0, 1, 2)], "of", [int((variant == v).sum()) for v in (0, 1, 2)])
    print("wbc golden: active-set iterations per level (max)", iters.max(0), "solves with active constraints", int((iters.sum(1) > 0).sum()))


def _solve_problem(args):
    """Worker: all warm-started cycles of one problem with the NumPy oracle."""
    ev, md, tt, tx, hor, dt, x0, cycles, iterations = args
    m, P = config.load_default()
    prob = sqp.MpcProblem(m, P, ev, md, tt, tx, horizon=hor, dt=dt)
    return [sqp.mpc_cycle(prob, 0.01 * c, x0, iterations=iterations) for c in range(cycles)]


def write_case(name, gait, hor, dt, x0, events, modes, nev, tt, tx, cycles, extra=None, iterations=1):
    """Golden file of explicit per-problem inputs (padded C-ABI arrays in, oracle outputs of `cycles` warm-started cycles)."""
    import multiprocessing as mp
    B = x0.shape[0]
    jobs = [(events[b, :nev[b]].copy(), modes[b, :nev[b] + 1].copy(), tt[b], tx[b], hor, dt, x0[b], cycles, iterations) for b in range(B)]
    with mp.Pool(min(B, os.cpu_count() or 1)) as pool:
        res = pool.map(_solve_problem, jobs)
    nmax = max(len(r[0]) for per in res for r in per)
    T = np.zeros((cycles, B, nmax)); X = np.zeros((cycles, B, nmax, 30)); U = np.zeros((cycles, B, nmax, 30))
    NN = np.zeros((cycles, B), dtype=np.int32); MD = np.zeros((cycles, B, nmax), dtype=np.int32)
    AL = np.zeros((cycles, B)); PERF = np.zeros((cycles, B, 7)); SQP = np.zeros((cycles, B, 2), dtype=np.int32)
    reasons = {"ITERATIONS": 1, "STEPSIZE": 2, "METRICS": 3, "PRIMAL": 4}
    for b in range(B):
        for c, (tout, xs, us, info) in enumerate(res[b]):
            SQP[c, b] = [len(info["history"]), reasons[info["convergence"]]]
            n = len(tout)
            T[c, b, :n], X[c, b, :n], U[c, b, :n], NN[c, b], MD[c, b, :n] = tout, xs, us, n, info["modes"]
            AL[c, b] = info["alpha"]
            PERF[c, b] = [info["armijo"], info["base"]["merit"], info["base"]["dyn"], info["base"]["eq"],
                          info["new"]["merit"], info["new"]["dyn"], info["new"]["eq"]]
    EM = max(40, events.shape[1])
    evp = np.full((B, EM), 1e30); mdp = np.full((B, EM + 1), 15, dtype=np.int32)
    evp[:, :events.shape[1]] = events; mdp[:, :modes.shape[1]] = modes
    for b in range(B):
        evp[b, nev[b]:] = 1e30
    kw = dict(extra or {})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mpc_cycle_%s.npz" % name), gait=gait, horizon=hor, dt=dt,
                        x0=x0, events=evp, modes=mdp, nevents=nev, target_t=tt, target_x=tx, t=T, x=X, u=U, n=NN, mode=MD,
                        alpha=AL, perf=PERF, max_nodes=int(nmax + 4), max_events=EM, iterations=iterations, sqp=SQP, **kw)
    print(name, "nodes", NN.min(), NN.max(), "alpha", np.unique(AL), "sqp iterations", np.unique(SQP[..., 0]), "reasons", np.unique(SQP[..., 1]))


def contract_goldens():
    """Contract-size cases, NumPy oracle (independent of the CUDA path's source):
      config2_n100      BASELINE config 2: problems 0, 1, 517, 1023 of the bench workload (B = 1024, horizon 1.0 / dt 0.01, trot)
      ref_dt15_h1       the reference's own setting: horizon 1.0 / dt 0.015 (task.info:79,141), trot
      flying_trot_n60   a gait with flight phases at N = 60
      moving_ee_n50     end-effector target moving over the horizon (position lerp + quaternion slerp), trot, N = 50
      gait_library_n100 BASELINE config 4 shape: the 11 moving gaits, one disturbance seed each, N = 100"""
    from qm_door_b200 import workload
    import qm_door_b200 as q
    from oracle import centroidal as ce
    W = workload.Workload(1024, horizon=1.0, dt=0.01)
    pick = [0, 1, 517, 1023]
    write_case("config2_n100", "trot", 1.0, 0.01, W.x0[pick], W.events[pick], W.modes[pick], W.nevents[pick], W.target_t[pick],
               W.target_x[pick], 3, extra=dict(picked=np.array(pick)))
    W = workload.Workload(3, horizon=1.0, dt=0.015, seed=31)
    write_case("ref_dt15_h1", "trot", 1.0, 0.015, W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x, 3)
    W = workload.Workload(2, horizon=0.6, dt=0.01, gait="flying_trot", seed=32, max_events=40)
    write_case("flying_trot_n60", "flying_trot", 0.6, 0.01, W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x, 2)
    W = workload.Workload(2, horizon=0.5, dt=0.01, seed=33)
    tt = np.tile(np.array([0.0, 0.6]), (2, 1))
    tx = W.target_x.copy()
    tx[:, 1, 30:33] += [0.06, -0.04, 0.03]
    ang = 0.35
    for b in range(2):
        qe = ce.quat_slerp(tx[b, 0, 33:37], np.array([np.sin(ang / 2), 0, 0, np.cos(ang / 2)]), 0.6 + 0.2 * b)
        tx[b, 1, 33:37] = qe / np.linalg.norm(qe)
    tx[:, 1, 6] += 0.08                                               # the base reference moves as well
    write_case("moving_ee_n50", "trot", 0.5, 0.01, W.x0, W.events, W.modes, W.nevents, tt, tx, 2)
    gaits = ["trot", "standing_trot", "flying_trot", "pace", "standing_pace", "dynamic_walk", "static_walk", "amble",
             "lindyhop", "skipping", "pawup"]
    B = len(gaits)
    W = workload.Workload(B, horizon=1.0, dt=0.01, seed=20261019, max_events=64)
    rng = np.random.default_rng(20261019)
    W.x0[:, 0:6] += rng.uniform(-0.3, 0.3, (B, 6))                    # external base momentum kick
    for gi, name in enumerate(gaits):
        sw, md = q.load_gait(name)
        ev, ms, ne = q.tile_schedule(sw, md, -np.ceil(1.0 / sw[-1]) * sw[-1] - W.phase[gi], 0.1 + 2.0, 64)
        W.events[gi], W.modes[gi], W.nevents[gi] = ev, ms, ne
    write_case("gait_library_n100", "library", 1.0, 0.01, W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x, 2,
               extra=dict(gaits=np.array(gaits)))


def multi_iteration_golden():
    """sqpIteration = 10 (the reference runs 1, task.info:80): two stance problems and a trot problem close to the nominal pose, so
    that the SQP loop stops early for different reasons / after different numbers of iterations per problem; plus a perturbed trot
    problem that uses the whole budget of 4."""
    from oracle import abi_fill
    m, P = config.load_default()
    tt, ts = scenarios.standing_target(m, P)
    B = 3
    x0s, phase = scenarios.perturbed_states(m, P, B, seed=78)
    x0s = P.x_init + 0.02 * (x0s - P.x_init)
    scheds = [G.tile_schedule(P.gaits["stance"] if b < 2 else P.gaits["trot"], -1.4 - phase[b], 1.2) for b in range(B)]
    ev, md, ne = abi_fill.pack_schedules(scheds, 40)
    write_case("sqp10_early_exit", "stance+trot", 0.1, 0.01, x0s, ev, md, ne, np.tile(tt, (B, 1)), np.tile(ts, (B, 1, 1)), 2, iterations=10)
    x0s, phase = scenarios.perturbed_states(m, P, B, seed=77)
    scheds = [G.tile_schedule(P.gaits["trot"], -1.4 - phase[b], 1.2) for b in range(B)]
    ev, md, ne = abi_fill.pack_schedules(scheds, 40)
    write_case("sqp4_trot_n20", "trot", 0.2, 0.01, x0s, ev, md, ne, np.tile(tt, (B, 1)), np.tile(ts, (B, 1, 1)), 2, iterations=4)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sqp":
    multi_iteration_golden()
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "contract":
    contract_goldens()
def _wbc_solve(args):
    from oracle import wbc
    x_des, u_des, rbd, mode, period, time, u_last = args
    m, P = config.load_default()
    O = wbc.Wbc(m, P)
    O.input_last = u_last.copy()
    try:
        cmd, dbg = O.update(x_des, u_des, rbd, int(mode), period, time, return_debug=True)
    except RuntimeError:        # degenerate inherited rows beyond the oracle solver's tolerance: recorded as not available (NaN)
        return np.full(54, np.nan), [-1, -1, -1]
    return cmd, [l.iterations for l in dbg["levels"]]


def wbc_full_size_golden(n=512):
    """The first `n` solves of the full-size config-5 workload (B = 65 536, the bench's own batch) with the NumPy oracle."""
    import multiprocessing as mp
    from qm_door_b200 import workload
    W = workload.WbcWorkload(65536)
    jobs = [(W.x_des[b], W.u_des[b], W.rbd[b], W.mode[b], W.period[b], W.time[b], W.u_last[b]) for b in range(n)]
    with mp.Pool(os.cpu_count() or 1) as pool:
        res = pool.map(_wbc_solve, jobs, chunksize=8)
    cmd = np.stack([r[0] for r in res]); iters = np.array([r[1] for r in res], dtype=np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wbc_config5_first512.npz"), n=n, cmd=cmd, iters=iters, mode=W.mode[:n])
    print("wbc full-size golden:", n, "solves,", int(np.isnan(cmd[:, 0]).sum()), "not available; active-set iterations per level (max)", iters.max(0))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "wbc512":
    wbc_full_size_golden()
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "wbc":
    wbc_golden()
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "backtracking":
    backtracking_golden()
