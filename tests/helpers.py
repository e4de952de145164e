import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "mpc_cycle_*.npz")))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def solver_for(solver, horizon, dt, max_events=40, max_nodes=None, iterations=1):
    import ctypes
    s = type(solver)()
    ctypes.memmove(ctypes.byref(s), ctypes.byref(solver), ctypes.sizeof(s))
    s.horizon, s.dt = float(horizon), float(dt)
    s.max_nodes = int(round(horizon / dt)) + 1 + 12 if max_nodes is None else int(max_nodes)
    s.max_events = max_events
    s.max_targets = 2
    s.sqp_iterations = iterations
    return s


# perturbed flight-phase problems at N = 100 are the worst conditioned of the suite (summation order of the 100-stage Riccati
# recursion shows at 1e-8 .. 1e-7); everything else holds 1e-8. The contract is 1e-4.
LOOSE = {"gait_library_n100": 1e-6}


def golden_tol(path, default=1e-8):
    name = os.path.basename(path)[len("mpc_cycle_"):-4]
    return LOOSE.get(name, default)


def check_against_golden(make_backend, path, tol_x=1e-8, tol_u=1e-8):
    """make_backend(solver_desc, B) -> object with .cycle(t0, x0, events, modes, nevents, tt, tx) -> dict, .close()."""
    g = np.load(path)
    B = g["x0"].shape[0]
    if "max_nodes" in g.files:      # contract-size cases carry their own capacities (many gait events per horizon)
        be = make_backend(float(g["horizon"]), float(g["dt"]), B, max_events=int(g["max_events"]), max_nodes=int(g["max_nodes"]),
                          iterations=int(g["iterations"]) if "iterations" in g.files else 1)
    else:
        be = make_backend(float(g["horizon"]), float(g["dt"]), B)
    cycles = g["t"].shape[0]
    worst = 0.0
    tol_x, tol_u = max(tol_x, golden_tol(path, tol_x)), max(tol_u, golden_tol(path, tol_u))
    for c in range(cycles):
        out = be.cycle(np.full(B, 0.01 * c), g["x0"], g["events"], g["modes"], g["nevents"], g["target_t"], g["target_x"])
        assert (out["status"] == 0).all(), out["status"]
        assert np.array_equal(out["n"], g["n"][c])                      # node counts: bit exact
        for b in range(B):
            n = out["n"][b]
            assert np.array_equal(out["mode"][b, :n], g["mode"][c, b, :n])   # contact-schedule / mode indices: bit exact
            assert np.array_equal(out["t"][b, :n], g["t"][c, b, :n])         # node times: bit exact
            ex, eu = rel_l2(out["x"][b, :n], g["x"][c, b, :n]), rel_l2(out["u"][b, :n], g["u"][c, b, :n])
            worst = max(worst, ex, eu)
            assert ex < tol_x and eu < tol_u, (path, c, b, ex, eu)
            assert out["info"][b, 0] == g["alpha"][c, b]
            if "sqp" in g.files:      # SQP iterations carried out and why the loop stopped
                assert out["info"][b, 13] == g["sqp"][c, b, 0] and out["info"][b, 14] == g["sqp"][c, b, 1]
            assert np.allclose(out["info"][b, [2, 5, 6, 7, 8, 9, 10]], g["perf"][c, b], rtol=1e-6, atol=1e-12)
    be.close()
    return worst
