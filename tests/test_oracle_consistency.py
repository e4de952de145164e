"""The reference ships no golden vectors for this path -> the oracle is pinned by the
self-consistency suite of SURVEY.md §8(c) and by the reference's constants (App. A.7)."""
import numpy as np
import pytest

from oracle import centroidal as ce
from oracle import config, gait as G, rbd, scenarios, sqp


def test_constant_pins(oracle_inputs):
    m, P = oracle_inputs
    assert abs(m.total_mass - 27.86796983) < 1e-8                       # sum of the 27 URDF masses
    assert m.nj == 24 and list(m.names[6:9]) == ["LF_HAA", "LF_HFE", "LF_KFE"] and m.names[9] == "LH_HAA"
    assert np.allclose(m.pp[18], [0.2535, 0, 0.056 + 0.0585])           # arm mount + z1_joint_1 offset
    assert np.allclose(m.ee_off, [0.186, 0, 0])                         # gripperStator 0.051 + EE frame 0.135
    assert np.allclose(np.diag(P.Q)[:12], [50, 50, 300, 10, 30, 30, 1000, 1000, 3000, 1000, 2000, 2000])
    assert np.allclose(np.diag(P.R_task)[[0, 12, 24]], [5e-3, 5.0, 1.0])
    assert len(P.gait_list) == 12 and P.gaits["trot"] == dict(modes=[9, 6], times=[0.0, 0.35, 0.70])
    assert P.sqp["dt"] == 0.015 and P.time_horizon == 1.0 and P.fric_mu == 0.7
    assert np.allclose(P.x_init[12:18], [0, 0.8, -1.5, 0, 0.8, -1.5])
    assert np.allclose(m.effort[6:9], [35.278, 35.278, 44.4])


def test_mode_bits_and_gaits(oracle_inputs):
    m, P = oracle_inputs
    assert G.stance_legs(15) == [1, 1, 1, 1] and G.stance_legs(9) == [1, 0, 0, 1] and G.stance_legs(6) == [0, 1, 1, 0]
    ev, md = G.tile_schedule(P.gaits["trot"], 0.0, 1.0)
    assert np.allclose(ev, [0, 0.35, 0.7, 1.05, 1.4]) and list(md) == [15, 9, 6, 9, 6, 15]
    for name, g in P.gaits.items():
        assert len(g["times"]) == len(g["modes"]) + 1 and all(0 <= x <= 15 for x in g["modes"])


def test_time_grid_events():
    t, f = G.time_grid(0.0, 0.1, 0.01, [0.035, 0.5])
    assert list(f) == [0, 0, 0, 0, 1, 2] + [0] * 7
    assert np.allclose(t[:7], [0, 0.01, 0.02, 0.03, 0.035, 0.035, 0.045]) and t[-1] == 0.1
    t, f = G.time_grid(0.0, 0.05, 0.015, [])
    assert np.allclose(t, [0, 0.015, 0.03, 0.045, 0.05]) and not f.any()


def test_every_pre_event_node_is_followed_by_its_post_event_node():
    """[upstream] timeDiscretizationWithEvents appends the PostEvent node after the push-or-overwrite branch: with dt = 0.015
    and 0.3 s phases the accumulated node time lands within dt_min of an event for many start times (1.1999999... vs 1.2),
    the last node is then moved onto the event and must still be followed by a post-event node at the same time."""
    events = [0.3 * k for k in range(1, 9)]
    overwritten = 0
    for i in range(300):
        t0 = 0.001 * i
        t, f = G.time_grid(t0, t0 + 1.0, 0.015, events)
        assert (np.diff(t) >= 0).all() and t[0] == t0 and t[-1] == t0 + 1.0
        pre = np.nonzero(f == G.EV_PRE)[0]
        assert len(pre) == sum(1 for e in events if t0 < e < t0 + 1.0)
        for k in pre:
            assert f[k + 1] == G.EV_POST and t[k + 1] == t[k] and t[k] in events
        for k in np.nonzero(f == G.EV_POST)[0]:
            assert f[k - 1] == G.EV_PRE
        overwritten += int(any(abs((t[k] - t[k - 1]) - 0.015) < 1e-9 and abs((t[k + 2] - t[k]) - 0.015) < 1e-9 for k in pre if k > 0))
    assert overwritten > 20          # the overwrite branch is exercised


def test_swing_spline_boundary_values(oracle_inputs):
    m, P = oracle_inputs
    ev, md = G.tile_schedule(P.gaits["trot"], 0.0, 2.0)
    sp = G.SwingPlanner(ev, md, P.swing)
    # RF (leg 1) swings during LF_RH = [0, 0.35]
    assert abs(sp.z_velocity(1, 1e-12) - P.swing["liftOffVelocity"]) < 1e-9
    assert abs(sp.z_velocity(1, 0.35) - P.swing["touchDownVelocity"]) < 1e-9
    assert abs(sp.z_position(1, 0.175) - P.swing["swingHeight"]) < 1e-12
    assert sp.z_velocity(0, 0.1) == 0.0                                  # LF in stance


def test_rbd_identities(oracle_inputs):
    m, P = oracle_inputs
    rng = np.random.default_rng(3)
    q = P.x_init[6:] + 0.3 * rng.standard_normal(24)
    v = rng.standard_normal(24)
    kin = rbd.kinematics(m, q)
    eps = 1e-6
    # frame Jacobian == central difference of the forward kinematics
    J = rbd.point_jacobian(m, kin, m.ee_joint, rbd.frame_position(m, kin, m.ee_joint, m.ee_off))
    Jfd = np.zeros((3, 24))
    for k in range(24):
        d = np.zeros(24); d[k] = eps
        Jfd[:, k] = (rbd.frame_position(m, rbd.kinematics(m, q + d), m.ee_joint, m.ee_off)
                     - rbd.frame_position(m, rbd.kinematics(m, q - d), m.ee_joint, m.ee_off)) / (2 * eps)
    assert np.abs(J - Jfd).max() < 1e-8
    # momentum: A v == sum of body momenta about the CoM obtained by differencing positions
    A, c = rbd.centroidal_momentum_matrix(m, kin)
    cb = lambda qq: rbd.body_coms(m, rbd.kinematics(m, qq))
    cd = (cb(q + eps * v) - cb(q - eps * v)) / (2 * eps)
    assert np.abs((m.mass[:, None] * cd).sum(0) - (A @ v)[:3]).max() < 1e-7
    # kinetic energy: v' M v == sum_i m |cdot_i|^2 + w' I w ;  M symmetric positive definite
    M = rbd.mass_matrix(m, kin)
    assert np.abs(M - M.T).max() < 1e-12 and np.linalg.eigvalsh(M).min() > 0
    # linear CMM rows == mass * CoM Jacobian ; first 3 columns = m I
    assert np.allclose(A[:3, :3], m.total_mass * np.eye(3)) and np.abs(A[3:, :3]).max() < 1e-12
    # nle at rest equals the gravity torque: -sum_i m_i Jv_i' g
    h0 = rbd.nonlinear_effects(m, q, np.zeros(24))
    g_t = np.zeros(24)
    cbq = rbd.body_coms(m, kin)
    for i in range(24):
        g_t -= m.mass[i] * rbd.point_jacobian(m, kin, i, cbq[i]).T @ rbd.GRAVITY
    assert np.abs(h0 - g_t).max() < 1e-9
    assert abs(h0[2] - m.total_mass * 9.81) < 1e-9


def test_flow_map_static_equilibrium_and_jacobians(oracle_inputs):
    m, P = oracle_inputs
    u = sqp.weight_compensating_input(m, 15)
    f = ce.flow_map(m, P.x_init, u)
    assert np.abs(f[:3]).max() < 1e-12 and np.abs(f[6:]).max() < 1e-12   # zero linear momentum rate, zero velocity
    rng = np.random.default_rng(5)
    x = P.x_init + 0.1 * rng.standard_normal(30)
    u = u + rng.standard_normal(30)
    _, fx, fu = ce.flow_map_linearization(m, x, u)
    eps = 1e-6
    for k in range(30):
        d = np.zeros(30); d[k] = eps
        assert np.abs((ce.flow_map(m, x + d, u) - ce.flow_map(m, x - d, u)) / (2 * eps) - fx[:, k]).max() < 1e-7
        assert np.abs((ce.flow_map(m, x, u + d) - ce.flow_map(m, x, u - d)) / (2 * eps) - fu[:, k]).max() < 1e-7
    # structure the CUDA path relies on: only rows 3..11 of df/dx are non-zero
    assert np.abs(fx[:3]).max() == 0 and np.abs(fx[12:]).max() == 0


def test_quaternion_conventions():
    rng = np.random.default_rng(7)
    for _ in range(20):
        a = rng.standard_normal(3); a /= np.linalg.norm(a)
        th = rng.uniform(-3.1, 3.1)
        R = rbd._axis_rot(a, np.array(th))
        qv = ce.quat_from_matrix(R)
        assert abs(np.linalg.norm(qv) - 1) < 1e-12
        ref = np.concatenate([np.sin(th / 2) * a, [np.cos(th / 2)]])
        assert min(np.abs(qv - ref).max(), np.abs(qv + ref).max()) < 1e-12
    q0 = np.array([0, 0, 0, 1.0]); q1 = np.array([0, 0, np.sin(0.5), np.cos(0.5)])
    assert np.allclose(ce.quat_slerp(q0, q1, 0.5), [0, 0, np.sin(0.25), np.cos(0.25)])
    assert np.allclose(ce.quat_distance(q0, q1), q1[:3])


def test_penalties_match_finite_differences():
    for h in (0.5, 4.0, 6.0, -1.0):
        v, d1, d2 = sqp.relaxed_barrier(h, 0.1, 5.0)
        e = 1e-5
        vp, vm = sqp.relaxed_barrier(h + e, 0.1, 5.0)[0], sqp.relaxed_barrier(h - e, 0.1, 5.0)[0]
        assert abs((vp - vm) / (2 * e) - d1) < 1e-8 and abs((vp - 2 * v + vm) / e ** 2 - d2) < 1e-4


@pytest.mark.parametrize("gait,horizon", [("trot", 0.06), ("static_walk", 0.05)])
def test_riccati_equals_dense_kkt(oracle_inputs, gait, horizon):
    """Projection + Riccati recursion == solution of the assembled dense equality-constrained QP."""
    m, P = oracle_inputs
    ev, md = G.tile_schedule(P.gaits[gait], -1.03, 2.0)
    tt, ts = scenarios.standing_target(m, P)
    prob = sqp.MpcProblem(m, P, ev, md, tt, ts, horizon=horizon, dt=0.01)
    x0 = scenarios.perturbed_states(m, P, 2)[0][0]
    _, _, _, info = sqp.mpc_cycle(prob, 0.0, x0, return_debug=True)
    n, nodes, flags = info["n"], info["nodes"], info["flags"]
    nxv, nz = 30 * (n + 1), 30 * (n + 1) + 30 * n
    H, g = np.zeros((nz, nz)), np.zeros(nz)
    xi = lambda k: slice(30 * k, 30 * k + 30)
    ui = lambda k: slice(nxv + 30 * k, nxv + 30 * k + 30)
    rows, rhs = [], []
    E = np.zeros((30, nz)); E[:, xi(0)] = np.eye(30); rows.append(E); rhs.append(info["dx"][0])
    for k in range(n):
        E = np.zeros((30, nz))
        if flags[k] == G.EV_PRE:
            E[:, xi(k)] = np.eye(30); E[:, xi(k + 1)] = -np.eye(30); rows.append(E); rhs.append(-info["stages"][k]["b"])
            E2 = np.zeros((30, nz)); E2[:, ui(k)] = np.eye(30); rows.append(E2); rhs.append(np.zeros(30))
            continue
        nd = nodes[k]; c = nd["cost"]
        H[xi(k), xi(k)] += c["Q"]; H[ui(k), ui(k)] += c["R"]; g[xi(k)] += c["q"]; g[ui(k)] += c["r"]
        E[:, xi(k)] = nd["A"]; E[:, ui(k)] = nd["B"]; E[:, xi(k + 1)] = -np.eye(30); rows.append(E); rhs.append(-nd["b"])
        E2 = np.zeros((nd["C"].shape[0], nz)); E2[:, xi(k)] = nd["C"]; E2[:, ui(k)] = nd["D"]; rows.append(E2); rhs.append(-nd["e"])
        # the projection satisfies the linearised constraints exactly (item 4)
        st = info["stages"][k]
        assert np.abs(nd["D"] @ st["Pu"]).max() < 1e-10 and np.abs(nd["D"] @ st["Px"] + nd["C"]).max() < 1e-10
        assert np.abs(nd["D"] @ st["Pe"] + nd["e"]).max() < 1e-10
    H[xi(n), xi(n)] += info["terminal"]["Q"]; g[xi(n)] += info["terminal"]["q"]
    Aeq, beq = np.vstack(rows), np.concatenate(rhs)
    K = np.block([[H, Aeq.T], [Aeq, np.zeros((Aeq.shape[0],) * 2)]])
    sol = np.linalg.lstsq(K, np.concatenate([-g, beq]), rcond=None)[0]
    dxd, dud = sol[:nxv].reshape(n + 1, 30), sol[nxv:nz].reshape(n, 30)
    assert np.abs(dxd - info["dx"]).max() / np.abs(dxd).max() < 1e-9
    assert np.abs(dud - info["du"]).max() / np.abs(dud).max() < 1e-9


def test_standing_is_a_fixed_point(oracle_inputs):
    """Stance gait, x0 = nominal, standing reference: the MPC keeps (nearly) standing. Not an exact fixed point: the arm
    shifts the CoM 5.6 cm ahead of the base, so equal weight-compensating forces leave a small pitch moment."""
    m, P = oracle_inputs
    ev, md = G.tile_schedule(P.gaits["stance"], -1.0, 2.0)
    tt, ts = scenarios.standing_target(m, P)
    prob = sqp.MpcProblem(m, P, ev, md, tt, ts, horizon=0.1, dt=0.01)
    for c in range(2):
        _, xs, us, info = sqp.mpc_cycle(prob, 0.01 * c, P.x_init)
    assert np.abs(xs[:, 6:12] - P.x_init[6:12]).max() < 5e-2
    assert abs(us[0, [2, 5, 8, 11]].sum() - m.total_mass * 9.81) < 0.05 * m.total_mass * 9.81


def test_multi_iteration_sqp_specification(oracle_inputs):
    """sqpIteration > 1 (the reference runs 1, task.info:80): the oracle's loop is the specification for the multi-iteration
    path. One iteration of the loop is the single-iteration cycle; an iteration starts from the performance the previous one
    ended with; the constraint violation of a perturbed start does not grow over the iterations."""
    m, P = oracle_inputs
    x0s, phase = scenarios.perturbed_states(m, P, 1, seed=8)
    tt, ts = scenarios.standing_target(m, P)
    ev, md = G.tile_schedule(P.gaits["trot"], -1.2 - phase[0], 1.2)
    mk = lambda: sqp.MpcProblem(m, P, ev, md, tt, ts, horizon=0.06, dt=0.01)
    _, x1, u1, i1 = sqp.mpc_cycle(mk(), 0.0, x0s[0])
    assert i1["convergence"] == "ITERATIONS" and len(i1["history"]) == 1
    _, x3, u3, i3 = sqp.mpc_cycle(mk(), 0.0, x0s[0], iterations=3)
    h = i3["history"]
    assert 1 <= len(h) <= 3 and i3["convergence"] in ("ITERATIONS", "STEPSIZE", "METRICS", "PRIMAL")
    assert h[0]["alpha"] == i1["alpha"] and np.isclose(h[0]["new"]["merit"], i1["new"]["merit"], rtol=1e-12)
    if len(h) == 1:
        assert np.array_equal(x3, x1) and np.array_equal(u3[:-1], u1[:-1])
    viol = lambda p: np.sqrt(p["dyn"] + p["eq"])
    for a, b in zip(h[:-1], h[1:]):
        for key in ("merit", "dyn", "eq"):
            assert np.isclose(b["base"][key], a["new"][key], rtol=1e-9, atol=1e-14)
    assert viol(h[-1]["new"]) <= viol(h[0]["base"]) * (1 + 1e-9)
