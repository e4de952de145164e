"""CPU port (host compilation of the kernels' per-node routines) against the NumPy oracle and its golden vectors.
This is what pins the arithmetic of csrc/qm_core.h / qm_mpc.h without a GPU."""
import ctypes as C

import numpy as np
import pytest

from helpers import GOLDEN, check_against_golden, rel_l2, solver_for
from oracle import abi_fill, centroidal as ce, gait as G, rbd, scenarios, sqp


@pytest.fixture(scope="module")
def cport_factory(descs):
    model, problem, solver, _ = descs

    def make(horizon, dt, B, **caps):
        return abi_fill.CPort(model, problem, solver_for(solver, horizon, dt, **caps), B)
    return make


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.split("mpc_cycle_")[-1][:-4])
def test_cycle_matches_golden(cport_factory, path):
    from helpers import golden_tol
    assert check_against_golden(cport_factory, path) < golden_tol(path)


def test_kinematics_and_analytic_derivatives(descs, oracle_inputs):
    """kin_eval: FK, CMM, frame Jacobians and the analytic d(Av)/dq, d(Jv)/dq against complex-step differentiation."""
    KW = abi_fill.kw_offsets()
    model, _, _, _ = descs
    m, P = oracle_inputs
    lib = abi_fill.load_cport()
    rng = np.random.default_rng(1)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for trial in range(3):
        x = P.x_init + 0.3 * rng.standard_normal(30)
        u = 3 * rng.standard_normal(30)
        w = np.zeros(lib.cport_kin_ws_size())
        lib.cport_kin_eval(C.byref(model), dp(x), dp(u), 1, dp(w))
        nk = ce.node_kinematics(m, x, u)
        get = lambda k, n, shp: w[KW[k]:KW[k] + n].reshape(shp)
        assert rel_l2(get("ACM", 144, (6, 24)), nk["A"]) < 1e-13
        assert rel_l2(get("FPOS", 12, (4, 3)), nk["foot_pos"]) < 1e-13
        assert rel_l2(get("FVEL", 12, (4, 3)), nk["foot_vel"]) < 1e-12
        assert rel_l2(get("VEL", 24, (24,)), nk["v"]) < 1e-12
        assert rel_l2(get("EER", 9, (3, 3)), nk["ee_rot"]) < 1e-13
        v, q = nk["v"], x[6:]

        def hfun(qc):
            A, _ = rbd.centroidal_momentum_matrix(m, rbd.kinematics(m, qc))
            return (A @ v[..., None])[..., 0]

        def fvfun(qc):
            k = rbd.kinematics(m, qc)
            return np.concatenate([(rbd.point_jacobian(m, k, m.foot_joint[i], rbd.frame_position(m, k, m.foot_joint[i], m.foot_off[i]))
                                    @ v[..., None])[..., 0] for i in range(4)], -1)
        assert rel_l2(get("DH", 144, (6, 24)), ce.cstep_jacobian(hfun, q)) < 1e-12
        assert rel_l2(get("DFV", 288, (12, 24)), ce.cstep_jacobian(fvfun, q)) < 1e-12
        # the end-effector Jacobian is a position-level product: its storage is reused by the velocity level (KW_HB)
        w = np.zeros(lib.cport_kin_ws_size())
        lib.cport_kin_eval(C.byref(model), dp(x), None, 0, dp(w))
        assert rel_l2(get("EEJ", 144, (6, 24)), rbd.frame_jacobian6(m, nk["kin"], m.ee_joint, m.ee_off)) < 1e-13


def test_fresh_cycle_against_oracle_with_ee_target_motion(descs, oracle_inputs):
    """A case the golden files do not hold: moving end-effector target (position lerp + quaternion slerp), pace gait."""
    model, problem, solver, _ = descs
    m, P = oracle_inputs
    B, hor = 2, 0.08
    x0s, phase = scenarios.perturbed_states(m, P, B, seed=99)
    tt, ts = scenarios.standing_target(m, P)
    ts = ts.copy()
    tt = np.array([0.0, 0.5])
    ts[1, 30:33] += [0.05, -0.03, 0.02]
    ang = 0.3
    ts[1, 33:37] = ce.quat_slerp(ts[0, 33:37], np.array([np.sin(ang / 2), 0, 0, np.cos(ang / 2)]), 0.7)
    ts[1, 33:37] /= np.linalg.norm(ts[1, 33:37])
    ts[1, 6] += 0.1
    sd = solver_for(solver, hor, 0.01)
    scheds = [G.tile_schedule(P.gaits["pace"], -1.2 - phase[b], 1.2) for b in range(B)]
    ev, md, ne = abi_fill.pack_schedules(scheds, sd.max_events)
    cp = abi_fill.CPort(model, problem, sd, B)
    probs = [sqp.MpcProblem(m, P, ev[b, :ne[b]], md[b, :ne[b] + 1], tt, ts, horizon=hor, dt=0.01) for b in range(B)]
    for c in range(2):
        out = cp.cycle(np.full(B, 0.01 * c), x0s, ev, md, ne, np.tile(tt, (B, 1)), np.tile(ts, (B, 1, 1)))
        for b in range(B):
            _, xs, us, info = sqp.mpc_cycle(probs[b], 0.01 * c, x0s[b])
            n = info["n"] + 1
            assert out["n"][b] == n and np.array_equal(out["mode"][b, :n], info["modes"])
            assert rel_l2(out["x"][b, :n], xs) < 1e-8 and rel_l2(out["u"][b, :n], us) < 1e-8
    cp.close()


def oracle_feedback_gains(info):
    """K = Pu K~ + Px per node in the original input coordinates; pre-event and final nodes repeat the previous node."""
    n = info["n"]
    K = np.zeros((n + 1, 30, 30))
    for k in range(n):
        st = info["stages"][k]
        if info["flags"][k] != G.EV_PRE:
            K[k] = st["Pu"] @ info["Ks"][k] + st["Px"]
        elif k > 0:
            K[k] = K[k - 1]
    K[n] = K[n - 1]
    return K


def test_feedback_gains_against_oracle(descs, oracle_inputs):
    """useFeedbackPolicy (task.info:90): the reduced coordinates depend on the pivot choice of the projection, the gain in the
    original coordinates does not. Also a property of the gain itself: u* + K dx stays on the linearised constraint manifold
    (rows of swing-foot forces are zero)."""
    model, problem, solver, _ = descs
    m, P = oracle_inputs
    B, hor = 2, 0.12
    x0s, phase = scenarios.perturbed_states(m, P, B, seed=5)
    tt, ts = scenarios.standing_target(m, P)
    sd = solver_for(solver, hor, 0.01)
    scheds = [G.tile_schedule(P.gaits["trot"], -1.2 - phase[b], 1.2) for b in range(B)]
    ev, md, ne = abi_fill.pack_schedules(scheds, sd.max_events)
    cp = abi_fill.CPort(model, problem, sd, B)
    probs = [sqp.MpcProblem(m, P, ev[b, :ne[b]], md[b, :ne[b] + 1], tt, ts, horizon=hor, dt=0.01) for b in range(B)]
    for c in range(2):
        out = cp.cycle(np.full(B, 0.01 * c), x0s, ev, md, ne, np.tile(tt, (B, 1)), np.tile(ts, (B, 1, 1)))
        K = cp.feedback_gains()
        for b in range(B):
            _, xs, us, info = sqp.mpc_cycle(probs[b], 0.01 * c, x0s[b], return_debug=True)
            Kref = oracle_feedback_gains(info)
            n = info["n"]
            assert out["n"][b] == n + 1
            assert rel_l2(K[b, :n + 1], Kref) < 1e-7
            for k in range(n):
                swing = [f for f in range(4) if not (info["modes"][k] >> (3 - f)) & 1]
                if info["flags"][k] != G.EV_PRE:
                    for f in swing:
                        assert np.abs(K[b, k, 3 * f:3 * f + 3]).max() < 1e-9
                    # the feedback term keeps the linearised equality constraints C dx + D du = 0 satisfied: D K + C = 0
                    nd = info["nodes"][k]
                    res = nd["D"] @ K[b, k] + nd["C"]
                    assert np.abs(res).max() < 1e-8 * max(1.0, np.abs(nd["C"]).max())
    cp.close()


def _random_rbd(m, P, n, seed):
    """Measured rbd states: nominal pose + perturbations, random generalized velocities, yaw spread over several turns."""
    from oracle import wbc as owbc
    rng = np.random.default_rng(seed)
    out, xs, vs = [], [], []
    for _ in range(n):
        x = P.x_init + np.concatenate([np.zeros(6), 0.05 * rng.standard_normal(3), 0.2 * rng.standard_normal(3), 0.2 * rng.standard_normal(18)])
        x[9] = rng.uniform(-3.1, 3.1)
        v = rng.uniform(-0.5, 0.5, 24)
        out.append(owbc.rbd_from_state(m, x, v)); xs.append(x); vs.append(v)
    return np.array(out), np.array(xs), np.array(vs)


def test_rbd_to_centroidal_state(descs, oracle_inputs):
    """measured rbdState(55) -> MPC state(30) (QMController.cpp:239-244): CPU port against the oracle,
    the round trip state -> rbd -> state, and the yaw unwrapping across +-pi."""
    model, _, _, _ = descs
    m, P = oracle_inputs
    lib = abi_fill.load_cport()
    dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
    rbds, xs, vs = _random_rbd(m, P, 12, seed=3)
    out = np.zeros((12, 30))
    lib.cport_rbd_to_state(C.byref(model), 12, dp(rbds), None, dp(out))
    for i in range(12):
        ref = ce.state_from_rbd(m, rbds[i])
        assert rel_l2(out[i], ref) < 1e-13
        assert rel_l2(out[i, 6:], xs[i, 6:]) < 1e-14                        # generalized coordinates come back unchanged
        A, _ = rbd.centroidal_momentum_matrix(m, rbd.kinematics(m, xs[i, 6:]))
        assert rel_l2(out[i, :6], A @ vs[i] / m.total_mass) < 1e-12         # normalized centroidal momentum of (q, v)
    yaw_last = xs[:, 9] + np.array([0.0, 6.0, -6.0, 12.5, -12.5, 3.0, -3.0, 0.1, -0.1, 6.2, -6.2, 100.0])
    lib.cport_rbd_to_state(C.byref(model), 12, dp(rbds), dp(yaw_last), dp(out))
    for i in range(12):
        ref = ce.state_from_rbd(m, rbds[i], yaw_last[i])
        assert abs(out[i, 9] - ref[9]) < 1e-12 and abs(out[i, 9] - yaw_last[i]) <= np.pi + 1e-12
        assert abs(np.angle(np.exp(1j * (out[i, 9] - xs[i, 9])))) < 1e-9    # same angle modulo 2 pi


def _target_cases(P, n, seed):
    rng = np.random.default_rng(seed)
    obs_state = np.tile(P.x_init, (n, 1)) + 0.1 * rng.standard_normal((n, 30))
    obs_state[:, 9] = rng.uniform(-3.0, 3.0, n)
    obs_time = rng.uniform(0.0, 50.0, n)
    quat = rng.standard_normal((n, 4)); quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    ee_state = np.concatenate([obs_state[:, 6:9] + rng.uniform(-0.5, 0.8, (n, 3)), quat], axis=1)
    last = ee_state + np.concatenate([rng.choice([0.01, 0.3], (n, 1)) * rng.standard_normal((n, 3)), 0.05 * rng.standard_normal((n, 4))], axis=1)
    cmd = np.zeros((n, 7))
    return rng, obs_time, obs_state, ee_state, last, cmd


def test_command_to_target_trajectories(descs, oracle_inputs):
    """velocity commands / end-effector goals -> two-knot reference
    (QmTargetTrajectoriesPublisher_node.cpp:60-257): loader against the oracle's parse, CPU port against the oracle for the
    three converters (lastEeTarget_ state included), and the invariants of the reference layout."""
    import qm_door_b200 as q
    from oracle import targets as ot
    m, P = oracle_inputs
    D = q.load_targets()
    tp = ot.TargetParams(P)
    assert D.com_height == tp.com_height and D.time_to_target == tp.time_to_target and D.arm_dist == ot.ARM_DIST
    assert D.target_displacement_velocity == tp.target_displacement_velocity and D.target_rotation_velocity == tp.target_rotation_velocity
    assert np.array_equal(np.ctypeslib.as_array(D.default_joint_state), tp.default_joint_state)
    lib = abi_fill.load_cport()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    n = 16
    for kind in range(3):
        rng, obs_time, obs_state, ee_state, last, cmd = _target_cases(P, n, seed=40 + kind)
        if kind < 2:
            cmd[:, :4] = rng.uniform(-0.5, 0.5, (n, 4))
        else:
            g = rng.standard_normal((n, 4)); g /= np.linalg.norm(g, axis=1, keepdims=True)
            cmd[:] = np.concatenate([ee_state[:, :3] + rng.uniform(-0.3, 0.3, (n, 3)), g], axis=1)
        last_c = last.copy()
        tt, tx = np.zeros((n, 2)), np.zeros((n, 2, 37))
        lib.cport_targets(C.byref(D), kind, n, dp(cmd), dp(obs_time), dp(obs_state), dp(ee_state), dp(last_c), dp(tt), dp(tx))
        for i in range(n):
            lo = last[i].copy()
            rt, rx = ot.CONVERTERS[kind](tp, cmd[i], lo, obs_time[i], obs_state[i], ee_state[i])
            assert np.abs(tt[i] - rt).max() < 1e-12 and np.abs(tx[i] - rx).max() < 1e-12
            assert np.abs(last_c[i] - lo).max() == 0.0
            assert tt[i, 0] == obs_time[i] and tt[i, 1] > tt[i, 0]
            assert np.array_equal(tx[i, :, 12:30], np.tile(tp.default_joint_state, (2, 1)))      # joints: default posture
            assert (tx[i, :, 8] == tp.com_height).all() and (tx[i, :, 10:12] == 0).all()          # height, pitch / roll


def test_error_statuses(descs):
    """Schedules that do not cover the horizon / node-capacity overflow are flagged per problem, not crashed on."""
    model, problem, solver, x_init = descs
    sd = solver_for(solver, 0.1, 0.01)
    B = 2
    cp = abi_fill.CPort(model, problem, sd, B)
    ev = np.full((B, sd.max_events), 1e30); md = np.full((B, sd.max_events + 1), 15, dtype=np.int32)
    ev[:, 0] = -1.0
    ne = np.array([1, 1], dtype=np.int32)                  # single event in the past: schedule ends before the horizon
    knot = np.concatenate([x_init, [0.6, 0, 0.8, 0, 0, 0, 1]])
    out = cp.cycle(np.zeros(B), np.tile(x_init, (B, 1)), ev, md, ne, np.tile([0.0, 1.0], (B, 1)), np.tile(knot, (B, 2, 1)))
    assert (out["status"] & 2).all()
    cp.close()
    sd2 = solver_for(solver, 0.3, 0.01)
    sd2.max_nodes = 12                                      # too small for 30 intervals
    cp = abi_fill.CPort(model, problem, sd2, B)
    ev[:, 0] = -1.0; ev[:, 1] = 5.0; ne[:] = 2
    out = cp.cycle(np.zeros(B), np.tile(x_init, (B, 1)), ev, md, ne, np.tile([0.0, 1.0], (B, 1)), np.tile(knot, (B, 2, 1)))
    assert (out["status"] & 1).all() and (out["n"] <= 12).all()
    cp.close()


def test_warm_start_across_gait_events(descs, oracle_inputs):
    """Receding horizon over several cycles with dt 0.015 s (task.info:79) while gait events enter, cross and leave the horizon:
    the warm start (segment search once per node, pre-event nodes, nodes beyond the previous solution, initializer inputs)
    must reproduce the oracle's sequential interpolation cycle after cycle."""
    model, problem, solver, _ = descs
    m, P = oracle_inputs
    B, hor, dt = 1, 0.12, 0.015
    x0s, _ = scenarios.perturbed_states(m, P, B, seed=123)
    tt, ts = scenarios.standing_target(m, P)
    sd = solver_for(solver, hor, dt)
    sched = G.tile_schedule(P.gaits["trot"], -0.31, 1.0)         # an event falls inside the first horizons and is passed later
    ev, md, ne = abi_fill.pack_schedules([sched], sd.max_events)
    cp = abi_fill.CPort(model, problem, sd, B)
    prob = sqp.MpcProblem(m, P, ev[0, :ne[0]], md[0, :ne[0] + 1], tt, ts, horizon=hor, dt=dt)
    seen_pre_first = False
    for c in range(7):
        t0 = 0.01 * c
        out = cp.cycle(np.full(B, t0), x0s, ev, md, ne, np.tile(tt, (B, 1)), np.tile(ts, (B, 1, 1)))
        _, xs, us, info = sqp.mpc_cycle(prob, t0, x0s[0])
        n = info["n"] + 1
        assert out["n"][0] == n and np.array_equal(out["mode"][0, :n], info["modes"])
        assert np.array_equal(out["t"][0, :n], np.array([G.interval_start(info["times"][i], info["flags"][i]) for i in range(n)]))
        assert rel_l2(out["x"][0, :n], xs) < 1e-8 and rel_l2(out["u"][0, :n], us) < 1e-8
        seen_pre_first = seen_pre_first or (G.EV_PRE in list(info["flags"][:3]))
    assert seen_pre_first                                          # an event right behind the initial time was part of the run
    cp.close()


def test_grid_sweep_matches_oracle_with_overwritten_event_nodes(descs, oracle_inputs):
    """build_grid (the routine k_schedule runs) against oracle.gait.time_grid over 300 start times with dt = 0.015 and 0.3 s
    phases: node times and counts bit-exact, every pre-event node followed by its post-event node (equal times)."""
    model, problem, solver, x_init = descs
    B = 300
    sd = solver_for(solver, 1.0, 0.015)
    events = [0.3 * k for k in range(1, 9)]
    modes = [15, 9, 6, 9, 6, 9, 6, 9, 15]
    ev, md, ne = abi_fill.pack_schedules([(np.array(events), np.array(modes, dtype=np.int32))] * B, sd.max_events)
    t0 = 0.001 * np.arange(B)
    knot = np.concatenate([x_init, [0.6253031727266175, 0.0, 0.8300452360692332, 0, 0, 0, 1]])
    cp = abi_fill.CPort(model, problem, sd, B, threads=8)
    out = cp.cycle(t0, np.tile(x_init, (B, 1)), ev, md, ne, np.tile([0.0, 1e3], (B, 1)), np.tile(knot, (B, 2, 1)))
    cp.close()
    for b in range(B):
        t, f = G.time_grid(t0[b], t0[b] + 1.0, 0.015, events)
        ts = np.array([G.interval_start(a, c) for a, c in zip(t, f)])
        assert out["n"][b] == len(t)
        assert np.array_equal(out["t"][b, :len(t)], ts)
        pairs = np.nonzero(f == G.EV_PRE)[0]
        assert all(f[k + 1] == G.EV_POST for k in pairs)


def gait_change_scenario(P, B=2, seed=41):
    """Mode schedule that changes mid-run (GaitReceiver, QMController.cpp:297-303): trot for the first two cycles; then a new
    template (pace / flying trot) is inserted at the first trot event after t = 0.1 s, the schedule before it is kept."""
    m = None
    scheds_a, scheds_b = [], []
    rng = np.random.default_rng(seed)
    for b in range(B):
        ph = rng.uniform(0.0, 0.7)
        ev, md = G.tile_schedule(P.gaits["trot"], -1.4 - ph, 1.6)
        scheds_a.append((ev, md))
        k = int(np.searchsorted(ev, 0.1)) + 1                      # events kept: ev[:k]; the new template starts at ev[k - 1]
        new = P.gaits["pace"] if b % 2 == 0 else P.gaits["flying_trot"]
        ev2, md2 = G.tile_schedule(new, ev[k - 1], 1.6)
        scheds_b.append((np.concatenate([ev[:k - 1], ev2]), np.concatenate([md[:k], md2[1:]]).astype(np.int32)))
    return scheds_a, scheds_b


def test_gait_change_between_cycles_against_oracle(descs, oracle_inputs):
    """The mode schedule is an input of every cycle: a gait change between two cycles moves / adds gait-event nodes in the
    horizon while the warm start still comes from the solution under the old schedule."""
    model, problem, solver, _ = descs
    m, P = oracle_inputs
    B, hor = 2, 0.4
    x0s, _ = scenarios.perturbed_states(m, P, B, seed=41)
    tt, ts = scenarios.standing_target(m, P)
    sd = solver_for(solver, hor, 0.01)
    sa, sb = gait_change_scenario(P, B)
    cp = abi_fill.CPort(model, problem, sd, B)
    probs = [sqp.MpcProblem(m, P, sa[b][0], sa[b][1], tt, ts, horizon=hor, dt=0.01) for b in range(B)]
    seen = set()
    for c in range(4):
        scheds = sa if c < 2 else sb
        if c == 2:
            for b in range(B):
                probs[b].set_mode_schedule(*sb[b])
        ev, md, ne = abi_fill.pack_schedules(scheds, sd.max_events)
        out = cp.cycle(np.full(B, 0.01 * c), x0s, ev, md, ne, np.tile(tt, (B, 1)), np.tile(ts, (B, 1, 1)))
        assert (out["status"] == 0).all()
        for b in range(B):
            _, xs, us, info = sqp.mpc_cycle(probs[b], 0.01 * c, x0s[b])
            n = info["n"] + 1
            assert out["n"][b] == n and np.array_equal(out["mode"][b, :n], info["modes"])
            assert rel_l2(out["x"][b, :n], xs) < 1e-8 and rel_l2(out["u"][b, :n], us) < 1e-8
            seen.update(int(v) for v in info["modes"])
    assert {9, 6}.issubset(seen) and (10 in seen or 5 in seen or 0 in seen)       # trot modes, then pace (LF_LH / RF_RH) or flight
