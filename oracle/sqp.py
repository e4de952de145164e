"""One MPC cycle of the multiple-shooting SQP solver (TEST INFRASTRUCTURE; see oracle/__init__.py).

Restates [upstream] ocs2_sqp::SqpSolver::runImpl (constructed at qm_controllers/src/QMController.cpp:288-289,
run by advanceMpc() at :323) on the optimal control problem assembled at
qm_interface/src/QMInterface.cpp:79-142, with the settings of qm_controllers/config/task.info:76-93:
  time discretisation with events -> initial guess (warm start / QMInitializer) -> per-node LQ
  transcription (RK2 = Heun sensitivities, cost*dt, equality constraints, LU-type projection) ->
  unconstrained LQ solve (HPIPM == Riccati recursion, strictly convex) -> filter line search.
Derivatives of the dynamics / end-effector kinematics come from complex-step differentiation.
"""
import numpy as np

from . import centroidal as ce
from . import gait as G
from . import rbd


# ----------------------------------------------------------------------------- interpolation
def time_segment(t, times):
    """[upstream] LinearInterpolation::timeSegment -> (index, alpha); value = alpha*d[i] + (1-alpha)*d[i+1]."""
    n = len(times)
    if n <= 1:
        return 0, 1.0
    part = int(np.searchsorted(times, t, side="left"))   # findIndexInTimeArray
    idx = 0 if (part == 0 and t == times[0]) else part - 1
    last = n - 1
    if idx >= 0:
        if idx < last:
            return idx, (times[idx + 1] - t) / (times[idx + 1] - times[idx])
        return max(last - 1, 0), 0.0
    return 0, 1.0


def interp(t, times, data):
    i, a = time_segment(t, times)
    if len(times) <= 1:
        return np.array(data[0])
    return a * data[i] + (1.0 - a) * data[i + 1]


# ----------------------------------------------------------------------------- penalties
def relaxed_barrier(h, mu, delta):
    """[upstream] RelaxedBarrierPenalty value / first / second derivative (real h)."""
    h = np.asarray(h, dtype=float)
    big = h > delta
    hs = np.where(big, h, 1.0)
    val = np.where(big, -mu * np.log(hs), mu * (-np.log(delta) + 0.5 * ((h - 2 * delta) / delta) ** 2 - 0.5))
    d1 = np.where(big, -mu / hs, mu * (h - 2 * delta) / (delta * delta))
    d2 = np.where(big, mu / (hs * hs), mu / (delta * delta))
    return val, d1, d2


def weight_compensating_input(model, mode):
    """[upstream] weightCompensatingInput: m g / n_stance on the z force of every stance foot."""
    u = np.zeros(30)
    flags = G.stance_legs(mode)
    n = sum(flags)
    if n > 0:
        for i in range(4):
            if flags[i]:
                u[3 * i + 2] = model.total_mass * 9.81 / n
    return u


def input_cost_weight(model, P):
    """QMInterface::initializeInputCostWeight (QMInterface.cpp:274-299): leg block J' R J at the initial state."""
    nk = ce.node_kinematics(model, P.x_init)
    J = np.zeros((12, 12))
    for i in range(4):
        Ji = rbd.point_jacobian(model, nk["kin"], model.foot_joint[i], nk["foot_pos"][i])
        J[3 * i:3 * i + 3, :] = Ji[:, 6:18]
    R = P.R_task.copy()
    R[12:24, 12:24] = J.T @ P.R_task[12:24, 12:24] @ J
    return R


# ----------------------------------------------------------------------------- problem instance
class MpcProblem:
    """One MPC problem: model + constants + mode schedule + target trajectories + previous solution."""

    def __init__(self, model, P, events, modes, target_times, target_states, horizon=None, dt=None):
        self.model, self.P = model, P
        self.events, self.modes = np.asarray(events, dtype=float), np.asarray(modes, dtype=np.int32)
        self.swing = G.SwingPlanner(self.events, self.modes, P.swing)
        self.tt, self.ts = np.asarray(target_times, dtype=float), np.asarray(target_states, dtype=float)
        self.R = input_cost_weight(model, P)
        self.horizon = P.time_horizon if horizon is None else horizon
        self.dt = P.sqp["dt"] if dt is None else dt
        self.prev = None  # (times, x, u) of the previous primal solution
        z = np.zeros(6)
        self.box_offset = (relaxed_barrier(z - P.arm_pos_lo, P.pos_bar_mu, P.pos_bar_delta)[0].sum()
                           + relaxed_barrier(P.arm_pos_hi - z, P.pos_bar_mu, P.pos_bar_delta)[0].sum()
                           + relaxed_barrier(z - P.arm_vel_lo, P.vel_bar_mu, P.vel_bar_delta)[0].sum()
                           + relaxed_barrier(P.arm_vel_hi - z, P.vel_bar_mu, P.vel_bar_delta)[0].sum())

    def set_mode_schedule(self, events, modes):
        """A new mode schedule arrives between two cycles ([upstream] GaitReceiver::preSolverRun replacing the gait schedule from
        an insertion time on, qm_controllers/src/QMController.cpp:297-303; published by GaitTopicPublisher.cpp:31-44). The previous
        primal solution is kept as it is and interpolated for the warm start; the re-timing of the warm start across moved
        events ([upstream] trajectorySpread, not vendored) is not restated."""
        self.events, self.modes = np.asarray(events, dtype=float), np.asarray(modes, dtype=np.int32)
        self.swing = G.SwingPlanner(self.events, self.modes, self.P.swing)

    def mode_at(self, t):
        return int(self.modes[G.mode_index(self.events, t)])

    def ee_ref(self, t):
        """EndEffectorConstraint::interpolateEndEffectorPose (EndEffectorConstraint.cpp:82-113)."""
        if len(self.tt) > 1:
            i, a = time_segment(t, self.tt)
            lhs, rhs = self.ts[i][30:37], self.ts[i + 1][30:37]
            return a * lhs[:3] + (1 - a) * rhs[:3], ce.quat_slerp(lhs[3:7], rhs[3:7], 1.0 - a)
        return self.ts[0][30:33], self.ts[0][33:37]

    # ---- value-only node evaluation (line search / metrics)
    def constraints_value(self, t, x, u, nk):
        flags = G.stance_legs(self.mode_at(t))
        rows = []
        for i in range(4):   # order of registration, QMInterface.cpp:116-131
            if not flags[i]:
                rows.append(u[3 * i:3 * i + 3])                                # zeroForce
            if flags[i]:
                rows.append(nk["foot_vel"][i])                                  # zeroVelocity (Av = I, b = 0)
            if not flags[i]:
                rows.append(nk["foot_vel"][i][2:3] - self.swing.z_velocity(i, t))  # normalVelocity
        return np.concatenate(rows) if rows else np.zeros(0)

    def cost_value(self, t, x, u, nk):
        P = self.P
        mode = self.mode_at(t)
        flags = G.stance_legs(mode)
        dx = x - interp(t, self.tt, self.ts)[:30]
        du = u - weight_compensating_input(self.model, mode)
        L = 0.5 * dx @ P.Q @ dx + 0.5 * du @ self.R @ du
        L += self.ee_cost(t, nk, P.mu_ee_pos, P.mu_ee_ori)
        # arm joint limits (QMInterface.cpp:177-259)
        L += relaxed_barrier(x[24:30] - P.arm_pos_lo, P.pos_bar_mu, P.pos_bar_delta)[0].sum()
        L += relaxed_barrier(P.arm_pos_hi - x[24:30], P.pos_bar_mu, P.pos_bar_delta)[0].sum()
        L += relaxed_barrier(u[24:30] - P.arm_vel_lo, P.vel_bar_mu, P.vel_bar_delta)[0].sum()
        L += relaxed_barrier(P.arm_vel_hi - u[24:30], P.vel_bar_mu, P.vel_bar_delta)[0].sum()
        L -= self.box_offset
        for i in range(4):
            if flags[i]:
                L += relaxed_barrier(self.cone(u[3 * i:3 * i + 3])[0], P.fric_bar_mu, P.fric_bar_delta)[0]
        return float(L)

    def ee_cost(self, t, nk, mu_p, mu_o):
        pr, qr = self.ee_ref(t)
        e = np.concatenate([nk["ee_pos"] - pr, ce.quat_distance(ce.quat_from_matrix(nk["ee_rot"]), qr)])
        return 0.5 * mu_p * e[:3] @ e[:3] + 0.5 * mu_o * e[3:] @ e[3:]

    def cone(self, F):
        """[upstream] FrictionConeConstraint value, gradient, Hessian wrt the local (= world) force."""
        P = self.P
        t2 = F[0] * F[0] + F[1] * F[1] + P.fric_reg
        tn = np.sqrt(t2)
        h = P.fric_mu * (F[2] + P.fric_grip) - tn
        g = np.array([-F[0] / tn, -F[1] / tn, P.fric_mu])
        p32 = tn * t2
        H = np.zeros((3, 3))
        H[0, 0] = -(F[1] * F[1] + P.fric_reg) / p32
        H[0, 1] = H[1, 0] = F[0] * F[1] / p32
        H[1, 1] = -(F[0] * F[0] + P.fric_reg) / p32
        return h, g, H

    def rk2(self, x, u, dt):
        """[upstream] SensitivityIntegrator RK2 (Heun): k1=f(x,u), k2=f(x+dt k1,u), x+ = x + dt/2 (k1+k2)."""
        k1 = ce.flow_map(self.model, x, u)
        k2 = ce.flow_map(self.model, x + dt * k1, u)
        return x + 0.5 * dt * (k1 + k2)


# ----------------------------------------------------------------------------- transcription
def initial_guess(prob, t0, x0, times, flags):
    """[upstream] multiple_shooting::initializeStateInputTrajectories (warm start, else QMInitializer.cpp:33-41)."""
    n = len(times) - 1
    prev = prob.prev
    has_prev = prev is not None and len(prev[0]) >= 2
    till_x = prev[0][-1] if has_prev else times[0]
    till_u = prev[0][-2] if has_prev else times[0]
    xs, us = [], []
    t_init = G.interval_start(times[0], flags[0])
    xs.append(interp(t_init, prev[0], prev[1]) if t_init < till_x else np.array(x0, dtype=float))
    for i in range(n):
        if flags[i] == G.EV_PRE:
            us.append(np.zeros(30))
            xs.append(xs[-1].copy())
        else:
            t = G.interval_start(times[i], flags[i])
            tn = G.interval_end(times[i + 1], flags[i + 1])
            if t > till_u or tn > till_x:
                us.append(weight_compensating_input(prob.model, prob.mode_at(t)))
                xs.append(xs[-1].copy())
            else:
                us.append(interp(t, prev[0], prev[2]))
                xs.append(interp(tn, prev[0], prev[1]))
    return np.array(xs), np.array(us)


def transcribe_node(prob, t, dt, x, u, xn):
    """[upstream] multiple_shooting::setupIntermediateNode + projectTranscription for one node."""
    model, P = prob.model, prob.P
    mode = prob.mode_at(t)
    flags = G.stance_legs(mode)
    # -- dynamics (Heun sensitivities)
    f1, A1, B1 = ce.flow_map_linearization(model, x, u)
    x2 = x + dt * f1
    f2, A2, B2 = ce.flow_map_linearization(model, x2, u)
    A2x = A2 + dt * A2 @ A1
    B2u = B2 + dt * A2 @ B1
    A = np.eye(30) + 0.5 * dt * (A1 + A2x)
    B = 0.5 * dt * (B1 + B2u)
    b = x + 0.5 * dt * (f1 + f2) - xn
    # -- kinematics at (x,u) with derivatives by complex step
    xu = np.concatenate([x, u])

    def kin_fun(z):
        nk = ce.node_kinematics(model, z[..., :30], z[..., 30:])
        return np.concatenate([nk["foot_vel"].reshape(z.shape[:-1] + (12,)), nk["ee_pos"],
                               ce.quat_distance(ce.quat_from_matrix(nk["ee_rot"]), qr)], axis=-1)

    pr, qr = prob.ee_ref(t)
    val = kin_fun(xu)
    Jk = ce.cstep_jacobian(kin_fun, xu)
    fv, dfv = val[:12].reshape(4, 3), Jk[:12].reshape(4, 3, 60)
    e = np.concatenate([val[12:15] - pr, val[15:18]])
    Je = Jk[12:18, :30]
    # -- cost (forward Euler * dt)
    W = np.diag([P.mu_ee_pos] * 3 + [P.mu_ee_ori] * 3)
    dx = x - interp(t, prob.tt, prob.ts)[:30]
    du = u - weight_compensating_input(model, mode)
    c0 = 0.5 * dx @ P.Q @ dx + 0.5 * du @ prob.R @ du + 0.5 * e @ W @ e - prob.box_offset
    q = P.Q @ dx + Je.T @ W @ e
    Qm = P.Q + Je.T @ W @ Je
    r = prob.R @ du
    Rm = prob.R.copy()
    for (lo, hi, mu, dl, vec, grad, hess, off) in (
            (P.arm_pos_lo, P.arm_pos_hi, P.pos_bar_mu, P.pos_bar_delta, x, q, Qm, 24),
            (P.arm_vel_lo, P.arm_vel_hi, P.vel_bar_mu, P.vel_bar_delta, u, r, Rm, 24)):
        vl, d1l, d2l = relaxed_barrier(vec[off:off + 6] - lo, mu, dl)
        vh, d1h, d2h = relaxed_barrier(hi - vec[off:off + 6], mu, dl)
        c0 += vl.sum() + vh.sum()
        grad[off:off + 6] += d1l - d1h
        hess[off:off + 6, off:off + 6] += np.diag(d2l + d2h)
    for i in range(4):
        if flags[i]:
            h, g, H = prob.cone(u[3 * i:3 * i + 3])
            pv, p1, p2 = relaxed_barrier(h, P.fric_bar_mu, P.fric_bar_delta)
            c0 += pv
            r[3 * i:3 * i + 3] += p1 * g
            Rm[3 * i:3 * i + 3, 3 * i:3 * i + 3] += p2 * np.outer(g, g) + p1 * H
            Rm[np.diag_indices(30)] += p1 * (-P.fric_hess_shift)      # [upstream] ddhdudu.diagonal() -= shift
            Qm[np.diag_indices(30)] += p1 * (-P.fric_hess_shift)      # [upstream] ddhdxdx.diagonal() -= shift
    cost = dict(c=dt * c0, q=dt * q, Q=dt * Qm, r=dt * r, R=dt * Rm, P=np.zeros((30, 30)))
    # -- state-input equality constraints  C dx + D du + e = 0
    Cs, Ds, es = [], [], []
    for i in range(4):
        if not flags[i]:
            D = np.zeros((3, 30))
            D[:, 3 * i:3 * i + 3] = np.eye(3)
            Cs.append(np.zeros((3, 30))); Ds.append(D); es.append(u[3 * i:3 * i + 3])
        if flags[i]:
            Cs.append(dfv[i][:, :30]); Ds.append(dfv[i][:, 30:]); es.append(fv[i])
        if not flags[i]:
            Cs.append(dfv[i][2:3, :30]); Ds.append(dfv[i][2:3, 30:]); es.append(fv[i][2:3] - prob.swing.z_velocity(i, t))
    C, D, ev = np.vstack(Cs), np.vstack(Ds), np.concatenate(es)
    return dict(A=A, B=B, b=b, cost=cost, C=C, D=D, e=ev, mode=mode, dt=dt)


def full_pivot_columns(D):
    """Pivot columns chosen by Gaussian elimination with full pivoting (Eigen::FullPivLU order)."""
    T = np.array(D, dtype=float)
    nc = T.shape[0]
    piv = []
    for step in range(nc):
        sub = np.abs(T[step:, :])
        r, c = np.unravel_index(int(np.argmax(sub)), sub.shape)
        r += step
        assert sub.max() > 1e-12, "constraint Jacobian lost rank"
        T[[step, r]] = T[[r, step]]
        piv.append(int(c))
        T[step + 1:] -= np.outer(T[step + 1:, c] / T[step, c], T[step])
    return piv


def project(node):
    """[upstream] projectTranscription with luConstraintProjection: du = Pu dut + Px dx + Pe,
    D Pu = 0, D Px = -C, D Pe = -e.  Eigen::FullPivLU semantics: kernel() = [-Dp^-1 Df; I] on the
    non-pivot columns, solve() = particular solution with the non-pivot variables at zero."""
    C, D, e = node["C"], node["D"], node["e"]
    nu = D.shape[1]
    piv = full_pivot_columns(D)
    free = [i for i in range(nu) if i not in piv]
    Dp = D[:, piv]
    Pu = np.zeros((nu, len(free)))
    Px = np.zeros((nu, C.shape[1]))
    Pe = np.zeros(nu)
    Pu[free, np.arange(len(free))] = 1.0
    Pu[piv, :] = -np.linalg.solve(Dp, D[:, free])
    Px[piv, :] = -np.linalg.solve(Dp, C)
    Pe[piv] = -np.linalg.solve(Dp, e)
    out_extra = dict(piv=piv, free=free)
    A, B, b, c = node["A"], node["B"], node["b"], node["cost"]
    out = dict(Pu=Pu, Px=Px, Pe=Pe, nut=Pu.shape[1], **out_extra)
    out["A"] = A + B @ Px
    out["B"] = B @ Pu
    out["b"] = b + B @ Pe
    r1 = c["r"] + c["R"] @ Pe
    out["c"] = c["c"] + c["r"] @ Pe + 0.5 * Pe @ c["R"] @ Pe
    q1 = c["q"] + c["P"].T @ Pe
    out["q"] = q1 + Px.T @ r1
    Pm = c["P"] + c["R"] @ Px
    out["Q"] = c["Q"] + Px.T @ c["P"] + c["P"].T @ Px + Px.T @ c["R"] @ Px
    out["r"] = Pu.T @ r1
    out["P"] = Pu.T @ Pm
    out["R"] = Pu.T @ c["R"] @ Pu
    return out


def riccati(stages, terminal, dx0):
    """Discrete Riccati recursion == the unconstrained OCP-QP HPIPM solves (strictly convex, unique)."""
    S, s = terminal["Q"].copy(), terminal["q"].copy()
    n = len(stages)
    Ks, ks = [None] * n, [None] * n
    for k in range(n - 1, -1, -1):
        st = stages[k]
        A, B, b = st["A"], st["B"], st["b"]
        sb = s + S @ b
        if B.shape[1] > 0:
            Gm = st["R"] + B.T @ S @ B
            H = st["P"] + B.T @ S @ A
            g = st["r"] + B.T @ sb
            L = np.linalg.cholesky(Gm)
            K = -np.linalg.solve(L.T, np.linalg.solve(L, H))
            kf = -np.linalg.solve(L.T, np.linalg.solve(L, g))
            Ks[k], ks[k] = K, kf
            s = st["q"] + A.T @ sb + H.T @ kf
            S = st["Q"] + A.T @ S @ A + H.T @ K
        else:
            Ks[k], ks[k] = np.zeros((0, 30)), np.zeros(0)
            s = st["q"] + A.T @ sb
            S = st["Q"] + A.T @ S @ A
        S = 0.5 * (S + S.T)
    dx = [np.array(dx0, dtype=float)]
    dut = []
    for k in range(n):
        st = stages[k]
        ut = Ks[k] @ dx[-1] + ks[k]
        dut.append(ut)
        dx.append(st["A"] @ dx[-1] + st["B"] @ ut + st["b"])
    return dx, dut, Ks, ks


def performance(prob, x0, times, flags, xs, us):
    """[upstream] SqpSolver::computePerformance -> (merit=cost, dynamicsViolationSSE, equalityConstraintsSSE)."""
    n = len(times) - 1
    cost, dyn, eq = 0.0, float((x0 - xs[0]) @ (x0 - xs[0])), 0.0
    for i in range(n):
        if flags[i] == G.EV_PRE:
            d = xs[i] - xs[i + 1]
            dyn += float(d @ d)
            continue
        t = G.interval_start(times[i], flags[i])
        dt = G.interval_end(times[i + 1], flags[i + 1]) - t
        d = prob.rk2(xs[i], us[i], dt) - xs[i + 1]
        dyn += dt * float(d @ d)
        nk = ce.node_kinematics(prob.model, xs[i], us[i])
        cost += dt * prob.cost_value(t, xs[i], us[i], nk)
        g = prob.constraints_value(t, xs[i], us[i], nk)
        eq += dt * float(g @ g)
    nk = ce.node_kinematics(prob.model, xs[n])
    cost += prob.ee_cost(times[n], nk, prob.P.mu_fee_pos, prob.P.mu_fee_ori)
    return dict(merit=cost, dyn=dyn, eq=eq)


def accept_step(S, base, new, armijo):
    """[upstream] FilterLinesearch::acceptStep."""
    vb = np.sqrt(base["dyn"] + base["eq"])
    vn = np.sqrt(new["dyn"] + new["eq"])
    if vn > S["g_max"]:
        return vn < (1.0 - S["gamma_c"]) * vb
    if vn < S["g_min"] and vb < S["g_min"] and armijo < 0.0:
        return new["merit"] < base["merit"] + S["armijoFactor"] * armijo
    return new["merit"] < base["merit"] - S["gamma_c"] * vb or vn < (1.0 - S["gamma_c"]) * vb


def check_convergence(S, it, iterations, base, new, alpha, dxn, dun):
    """[upstream] SqpSolver::checkConvergence, restated from the published ocs2_sqp sources (not in the reference tree, so this
    ordering of the tests is unverified here): iteration budget, step size below alpha_min, merit change below costTol on a
    feasible iterate, primal step below deltaTol. Returns the reason or None. The reference runs one iteration (task.info:80),
    for which the answer is always "ITERATIONS"."""
    if it + 1 >= iterations:
        return "ITERATIONS"
    if alpha < S["alpha_min"]:
        return "STEPSIZE"
    if abs(new["merit"] - base["merit"]) < S.get("costTol", 1e-4) and np.sqrt(new["dyn"] + new["eq"]) < S["g_min"]:
        return "METRICS"
    if alpha * dxn < S["deltaTol"] and alpha * dun < S["deltaTol"]:
        return "PRIMAL"
    return None


def mpc_cycle(prob, t0, x0, return_debug=False, iterations=1):
    """SqpSolver::runImpl; iterations = sqpIteration (1 in the reference, task.info:80; more than one is the specification for
    the multi-iteration path, which the CUDA library does not have yet). Updates prob.prev. Returns (times, x, u, info)."""
    P, S = prob.P, prob.P.sqp
    times, flags = G.time_grid(t0, t0 + prob.horizon, prob.dt, prob.events)
    n = len(times) - 1
    xs, us = initial_guess(prob, t0, x0, times, flags)
    history = []
    for it in range(iterations):
        stages, nodes = [], []
        for i in range(n):
            if flags[i] == G.EV_PRE:
                z = np.zeros((30, 30))
                st = dict(A=np.eye(30), B=np.zeros((30, 0)), b=xs[i] - xs[i + 1], c=0.0, q=np.zeros(30), Q=z,
                          r=np.zeros(0), P=np.zeros((0, 30)), R=np.zeros((0, 0)), Pu=np.zeros((30, 0)), Px=z,
                          Pe=np.zeros(30), nut=0)
                nodes.append(None)
            else:
                t = G.interval_start(times[i], flags[i])
                dt = G.interval_end(times[i + 1], flags[i + 1]) - t
                nd = transcribe_node(prob, t, dt, xs[i], us[i], xs[i + 1])
                st = project(nd)
                nodes.append(nd)
            stages.append(st)
        # terminal node: finalSoftConstraint "finalEndEffector" (QMInterface.cpp:104)
        tN = times[n]
        pr, qr = prob.ee_ref(tN)

        def ee_fun(z):
            nk = ce.node_kinematics(prob.model, z)
            return np.concatenate([nk["ee_pos"], ce.quat_distance(ce.quat_from_matrix(nk["ee_rot"]), qr)], axis=-1)

        ev = ee_fun(xs[n]) - np.concatenate([pr, np.zeros(3)])
        Je = ce.cstep_jacobian(ee_fun, xs[n])
        W = np.diag([P.mu_fee_pos] * 3 + [P.mu_fee_ori] * 3)
        terminal = dict(Q=Je.T @ W @ Je, q=Je.T @ W @ ev, c=0.5 * ev @ W @ ev)
        # baseline performance
        base = performance(prob, x0, times, flags, xs, us)
        dx, dut, Ks, ks = riccati(stages, terminal, x0 - xs[0])
        armijo = sum(float(st["q"] @ dx[k] + st["r"] @ dut[k]) for k, st in enumerate(stages)) + float(terminal["q"] @ dx[n])
        du = [st["Pu"] @ dut[k] + st["Px"] @ dx[k] + st["Pe"] for k, st in enumerate(stages)]
        dxn = np.sqrt(sum(float(d @ d) for d in dx))
        dun = np.sqrt(sum(float(d @ d) for d in du))
        alpha, accepted = 1.0, False
        while True:
            xn = xs + alpha * np.array(dx)
            un = us + alpha * np.array(du)
            new = performance(prob, x0, times, flags, xn, un)
            if accept_step(S, base, new, alpha * armijo):
                accepted = True
                break
            alpha *= S["alpha_decay"]
            if alpha * dxn < S["deltaTol"] and alpha * dun < S["deltaTol"]:
                break
            if alpha < S["alpha_min"]:
                break
        if accepted:
            xs, us = xn, un
        else:
            alpha, new = 0.0, base
        history.append(dict(alpha=alpha, base=base, new=new))
        convergence = check_convergence(S, it, iterations, base, new, alpha, dxn, dun)
        if convergence is not None:
            break
    # [upstream] toPrimalSolution: inputs at pre-event nodes repeat the previous input, last input repeated
    uo = np.array(us)
    for i in range(n):
        if flags[i] == G.EV_PRE and i > 0:
            uo[i] = uo[i - 1]
    uo = np.vstack([uo, uo[-1:]])
    tout = np.array([G.interval_start(times[i], flags[i]) for i in range(n + 1)])
    prob.prev = (tout, np.array(xs), uo)
    info = dict(alpha=alpha, base=base, new=new, armijo=armijo, flags=flags, times=times, n=n, history=history,
                convergence=convergence,
                modes=np.array([prob.mode_at(G.interval_start(times[i], flags[i])) for i in range(n + 1)], dtype=np.int32))
    if return_debug:
        info.update(stages=stages, nodes=nodes, terminal=terminal, dx=np.array(dx), du=np.array(du), Ks=Ks, ks=ks)
    return tout, np.array(xs), uo, info


def evaluate_policy(prob, t):
    """[upstream] MPC_MRT_Interface::evaluatePolicy with a feed-forward controller (QMController.cpp:140-143)."""
    tt, xs, us = prob.prev
    return interp(t, tt, xs), interp(t, tt, us), prob.mode_at(t)
