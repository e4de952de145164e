"""Whole-body controller oracle (TEST INFRASTRUCTURE; see oracle/__init__.py).

Restates, file by file, the reference's qm_wbc package:
  WbcBase::update / updateMeasured / updateDesired      qm_wbc/src/WbcBase.cpp:123-238
  the 13 task builders formulate*                       qm_wbc/src/WbcBase.cpp:240-578
  WbcBase::updateCmd                                    qm_wbc/src/WbcBase.cpp:580-595
  Task (+, * scalar)                                    qm_wbc/include/qm_wbc/Task.h:17-66
  HoQp (formulation of every level's QP)                qm_wbc/src/HoQp.cpp:12-158
  HierarchicalWbc / HierarchicalMpcWbc::update          qm_wbc/src/HierarchicalWbc.cpp:18-44, HierarchicalMpcWbc.cpp:18-34
qpOASES (pinned 268b2f26, not vendored) is replaced by an exact dual active-set solver (Goldfarb & Idnani 1983) for the
strictly convex QP each level poses; every level's QP has a unique solution, so any exact solver must agree.
Rigid-body quantities come from the definitions in oracle/rbd.py (time variations by complex step).
"""
import numpy as np

from . import centroidal as ce
from . import gait as G
from . import rbd

DEFAULT_GAINS = dict(  # qm_wbc/cfg/wbcWigeht.cfg:7-47
    kp_swing=350.0, kd_swing=37.0, baseHeightKp=400.0, baseHeightKd=140.0, kp_base_linear=400.0, kd_base_linear=100.0,
    kp_base_angular=400.0, kd_base_angular=140.0,
    kp_arm_joint=[4000.0, 4200.0, 4000.0, 4000.0, 4200.0, 6000.0], kd_arm_joint=[75.0] * 6,
    kp_ee_linear=[3000.0] * 3, kd_ee_linear=[75.0] * 3, kp_ee_angular=[2000.0] * 3, kd_ee_angular=[75.0] * 3)


# ----------------------------------------------------------------------------- [upstream] ocs2_robotic_tools rotation helpers
def euler_zyx_map(e):
    """T(euler): world angular velocity = T @ d/dt[yaw, pitch, roll]."""
    z, y = e[0], e[1]
    return np.array([[0.0, -np.sin(z), np.cos(y) * np.cos(z)],
                     [0.0, np.cos(z), np.cos(y) * np.sin(z)],
                     [1.0, 0.0, -np.sin(y)]])


def euler_zyx_map_dot(e, de):
    z, y, dz, dy = e[0], e[1], de[0], de[1]
    return np.array([[0.0, -np.cos(z) * dz, -np.sin(y) * np.cos(z) * dy - np.cos(y) * np.sin(z) * dz],
                     [0.0, -np.sin(z) * dz, -np.sin(y) * np.sin(z) * dy + np.cos(y) * np.cos(z) * dz],
                     [0.0, 0.0, -np.cos(y) * dy]])


def rot_zyx(e):
    return rbd._axis_rot([0, 0, 1], np.array(e[0])) @ rbd._axis_rot([0, 1, 0], np.array(e[1])) @ rbd._axis_rot([1, 0, 0], np.array(e[2]))


def rotation_error_in_world(R_ref, R_cur):
    """[upstream] rotationErrorInWorld: rotation vector of R_ref R_cur^T."""
    R = R_ref @ R_cur.T
    skew = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = 0.5 * (np.trace(R) - 1.0)
    c = min(1.0, max(-1.0, c))
    th = np.arccos(c)
    if th < 1e-8:
        return 0.5 * skew
    return th / (2.0 * np.sin(th)) * skew


# ----------------------------------------------------------------------------- Task (Task.h:17-66)
class Task:
    def __init__(self, a=None, b=None, d=None, f=None, nvar=36):
        self.a = np.zeros((0, nvar)) if a is None else np.asarray(a, dtype=float)
        self.b = np.zeros(0) if b is None else np.asarray(b, dtype=float)
        self.d = np.zeros((0, nvar)) if d is None else np.asarray(d, dtype=float)
        self.f = np.zeros(0) if f is None else np.asarray(f, dtype=float)

    def __add__(self, o):
        return Task(np.vstack([self.a, o.a]), np.concatenate([self.b, o.b]), np.vstack([self.d, o.d]), np.concatenate([self.f, o.f]))

    def __mul__(self, s):
        return Task(self.a * s, self.b * s, self.d * s, self.f * s)


# ----------------------------------------------------------------------------- exact strictly convex QP: min 1/2 x'Hx + c'x  s.t. C x <= d
def solve_qp_gi(Hfac, c, C, d, max_iter=2000, tol=1e-10, tol_degenerate=1e-7):
    """Goldfarb-Idnani dual active set. Hfac = Jm with Jm' H Jm = I (so the nearly singular Hessians of HoQp, which carry a
    1e-12 regularisation, are handled through their square-root factor: step directions are formed in the scaled
    coordinates, the iterate and the constraint residuals are kept in the original ones). Returns (x, active set, iterations)."""
    C = np.asarray(C, dtype=float)
    d = np.asarray(d, dtype=float)
    Ct = -(C @ Hfac)             # normals of  n'y >= b  in the scaled coordinates x = Jm y
    x = -(Hfac @ (Hfac.T @ c))   # unconstrained minimiser
    act, u = [], np.zeros(0)
    it = 0
    scale = 1.0 + np.abs(d)
    ignored = []     # rows violated by less than tol_degenerate whose normal depends on the active set: the hierarchy hands
    #                  the previous level's active rows down with zero slack, so the feasible set is degenerate up to rounding
    while True:
        s = d - C @ x            # >= 0 when satisfied
        s[act] = 0.0
        s[ignored] = 0.0
        p = int(np.argmin(s / scale))
        if s[p] / scale[p] >= -tol:
            return x, act, it
        nplus = Ct[p]
        uplus = np.concatenate([u, [0.0]])
        while True:
            it += 1
            if it > max_iter:
                raise RuntimeError("GI: iteration limit")
            q = len(act)
            if q > 0:
                Q, R = np.linalg.qr(Ct[act].T, mode="complete")
                dv = Q.T @ nplus
                r = np.linalg.solve(R[:q, :q], dv[:q])
                z = Q[:, q:] @ dv[q:]
                dn2 = float(dv[q:] @ dv[q:])
            else:
                r = np.zeros(0)
                z = nplus.copy()
                dn2 = float(z @ z)
            t1, l = np.inf, -1
            for j in range(q):
                if r[j] > 1e-14 * (1 + abs(uplus[j])) and uplus[j] / r[j] < t1:
                    t1, l = uplus[j] / r[j], j
            sp = float(d[p] - C[p] @ x)
            t2 = -sp / dn2 if dn2 > 1e-26 * float(nplus @ nplus) else np.inf
            t = min(t1, t2)
            if not np.isfinite(t):
                if abs(sp) < tol_degenerate * scale[p]:
                    ignored.append(p)
                    uplus = None
                    break
                raise RuntimeError("GI: infeasible QP")
            if np.isfinite(t2):
                x = x + t * (Hfac @ z)
            uplus = uplus + t * np.concatenate([-r, [1.0]])
            if t == t2:
                act.append(p)
                u = uplus
                break
            act.pop(l)
            uplus = np.delete(uplus, l)
        if uplus is None:
            continue


# ----------------------------------------------------------------------------- [upstream] Eigen::FullPivLU<MatrixXd>::kernel()
def full_piv_lu_kernel(A):
    """Kernel basis exactly as Eigen's FullPivLU produces it for a column-major dynamic matrix (HoQp.cpp:129:
    `(task_.a_ * stackedZPrev_).fullPivLu().kernel()`), restated from the published Eigen 3.3/3.4 sources (Eigen is not vendored):
      * elimination with full pivoting; the pivot of step k is the entry of largest magnitude of the remaining corner, the FIRST
        one in column-major order on ties (maxCoeff visitor: outer loop over columns, strict `>`);
      * rank = number of pivots with |U_ii| > eps * diagonalSize * max|pivot| (FullPivLU::threshold() default);
      * kernel = Q [ -U11^-1 U12 ; I ] with Q the accumulated column permutation (kernel_retval::evalTo); a trivial kernel is
        returned as ONE zero column.
    The basis is not orthonormal; HoQp's 1e-12 |z|^2 regularisation acts on the coordinates in THIS basis, which is what selects
    the solution of a rank-deficient level. Returns (N [cols x max(dimker, 1)], rank)."""
    lu = np.array(A, dtype=float)
    rows, cols = lu.shape
    size = min(rows, cols)
    row_t, col_t = list(range(size)), list(range(size))
    nonzero, maxpivot = size, 0.0
    for k in range(size):
        corner = np.abs(lu[k:, k:])
        best, bi, bj = corner[0, 0], 0, 0
        for j in range(corner.shape[1]):              # column-major traversal, strict greater
            for i in range(corner.shape[0]):
                if corner[i, j] > best:
                    best, bi, bj = corner[i, j], i, j
        if best == 0.0:
            nonzero = k
            break
        bi += k; bj += k
        maxpivot = max(maxpivot, best)
        row_t[k], col_t[k] = bi, bj
        if bi != k:
            lu[[k, bi], :] = lu[[bi, k], :]
        if bj != k:
            lu[:, [k, bj]] = lu[:, [bj, k]]
        if k < rows - 1:
            lu[k + 1:, k] /= lu[k, k]
        if k < size - 1:
            lu[k + 1:, k + 1:] -= np.outer(lu[k + 1:, k], lu[k, k + 1:])
    q = list(range(cols))
    for k in range(size):
        q[k], q[col_t[k]] = q[col_t[k]], q[k]
    thr = maxpivot * np.finfo(float).eps * size
    pivots = [i for i in range(nonzero) if abs(lu[i, i]) > thr]
    rank = len(pivots)
    dimker = cols - rank
    if dimker == 0:
        return np.zeros((cols, 1)), rank
    m = np.zeros((rank, cols))
    for i in range(rank):
        m[i, i:] = lu[pivots[i], i:]
    m[:, :rank] = np.triu(m[:, :rank])
    for i in range(rank):
        m[:, [i, pivots[i]]] = m[:, [pivots[i], i]]
    if rank > 0:
        m[:, rank:] = np.linalg.solve(np.triu(m[:, :rank]), m[:, rank:]) if rank > 0 else m[:, rank:]
    for i in range(rank - 1, -1, -1):
        m[:, [i, pivots[i]]] = m[:, [pivots[i], i]]
    N = np.zeros((cols, dimker))
    for i in range(rank):
        N[q[i], :] = -m[i, cols - dimker:]
    for k in range(dimker):
        N[q[rank + k], k] = 1.0
    return N, rank


# ----------------------------------------------------------------------------- HoQp (HoQp.cpp:12-158)
class HoQp:
    def __init__(self, task, higher=None):
        self.task = task
        nv_s = task.d.shape[0]
        if higher is not None:
            Zp, tp, vp, xp = higher.Z, higher.stacked, higher.stacked_slack, higher.x
        else:
            nx = task.a.shape[1]
            Zp, tp, vp, xp = np.eye(nx), Task(nvar=nx), np.zeros(0), np.zeros(nx)
        nz = Zp.shape[1]
        self.stacked = task + tp
        if nz == 0:     # no freedom left (e.g. all four feet swinging: level 1 already has >= 18 rows): x = x_prev
            self.z, self.v, self.x, self.Z = np.zeros(0), np.zeros(nv_s), xp, Zp
            self.stacked_slack = np.concatenate([vp, self.v])
            self.active, self.iterations = [], 0
            return
        AZ = task.a @ Zp
        # H = blkdiag(Z'A'AZ + 1e-12 I, I) -> factor through the SVD of A Z (HoQp.cpp:60-76)
        if AZ.shape[0] > 0:
            _, sv, Vt = np.linalg.svd(AZ, full_matrices=True)
            lam = np.zeros(nz)
            lam[:len(sv)] = sv ** 2
            Jz = Vt.T / np.sqrt(lam + 1e-12)
            cz = AZ.T @ (task.a @ xp - task.b)                                      # HoQp.cpp:78-90
        else:
            Jz = np.zeros((nz, nz))     # zero Hessian block: only reachable with no equality rows (not used by the stacks)
            cz = np.zeros(nz)
            raise NotImplementedError("level without equality rows")
        n = nz + nv_s
        Jm = np.zeros((n, n))
        Jm[:nz, :nz] = Jz
        Jm[nz:, nz:] = np.eye(nv_s)
        c = np.concatenate([cz, np.zeros(nv_s)])
        # constraints (HoQp.cpp:92-124): [0 -I; Dprev Z 0; D Z -I] [z; v] <= [0; fprev - Dprev xprev + vprev; f - D xprev]
        Cm = np.vstack([np.hstack([np.zeros((nv_s, nz)), -np.eye(nv_s)]),
                        np.hstack([tp.d @ Zp, np.zeros((tp.d.shape[0], nv_s))]),
                        np.hstack([task.d @ Zp, -np.eye(nv_s)])])
        dm = np.concatenate([np.zeros(nv_s), tp.f - tp.d @ xp + vp, task.f - task.d @ xp])
        try:
            sol, self.active, self.iterations = solve_qp_gi(Jm, c, Cm, dm)
        except RuntimeError:
            # rows handed down tight from the levels above are violated by accumulated rounding (1e-6 .. 1e-5 in the deeper levels
            # of the six-level stack, whose null-space bases are not orthonormal): accept them as degenerate with a wider margin
            try:
                sol, self.active, self.iterations = solve_qp_gi(Jm, c, Cm, dm, tol_degenerate=1e-5)
            except RuntimeError:
                # last resort, the device code's rule: a violated row whose normal depends on the active set with nothing to drop
                # is skipped whatever its residual (the solve is then flagged degenerate there, WST_DEGENERATE)
                sol, self.active, self.iterations = solve_qp_gi(Jm, c, Cm, dm, tol_degenerate=np.inf)
            self.relaxed = True
        self.z, self.v = sol[:nz], sol[nz:]
        self.x = xp + Zp @ self.z                                                      # HoQp.h:31-34
        self.stacked_slack = np.concatenate([vp, self.v])
        # next null space (HoQp.cpp:126-133): stackedZPrev * (A Zprev).fullPivLu().kernel(), Eigen's own basis (not orthonormal).
        # A trivial kernel comes back from Eigen as one zero column (one dummy variable that moves nothing): no freedom left.
        N, rank = full_piv_lu_kernel(AZ)
        self.Z = Zp @ N if rank < nz else np.zeros((Zp.shape[0], 0))
        self.Zprev, self.xprev, self.AZ = Zp, xp, AZ


# ----------------------------------------------------------------------------- WbcBase
class Wbc:
    def __init__(self, model, P, gains=None, mpc_variant=False):
        self.m, self.P = model, P
        self.g = dict(DEFAULT_GAINS if gains is None else gains)
        self.mu = P.friction_wbc                                                      # task.info:347-350
        # WbcBase.cpp:599-604 + :409-410: effortLimit.segment<3>(6) replicated over the four legs, effortLimit.tail(6) for the arm
        self.tau_max = np.concatenate([np.tile(model.effort[6:9], 4), model.effort[-6:]])
        self.input_last = np.zeros(30)
        self.mpc_variant = mpc_variant

    # -- updateMeasured (WbcBase.cpp:146-203)
    def update_measured(self, rbdm):
        m = self.m
        q, v = np.zeros(24), np.zeros(24)
        q[0:3], q[3:6], q[6:] = rbdm[3:6], rbdm[0:3], rbdm[6:24]
        v[0:3] = rbdm[27:30]
        v[3:6] = np.linalg.solve(euler_zyx_map(q[3:6]), rbdm[24:27])
        v[6:] = rbdm[30:48]
        kin = rbd.kinematics(m, q)
        s = dict(q=q, v=v, kin=kin)
        s["M"] = rbd.mass_matrix(m, kin)
        s["nle"] = rbd.nonlinear_effects(m, q, v)
        s["J"] = np.vstack([rbd.frame_jacobian6(m, kin, m.foot_joint[i], m.foot_off[i])[:3] for i in range(4)])
        s["dJ"] = np.vstack([rbd.frame_jacobian6_dot(m, q, v, m.foot_joint[i], m.foot_off[i])[:3] for i in range(4)])
        s["Jb"] = rbd.frame_jacobian6(m, kin, 5, np.zeros(3))
        s["dJb"] = rbd.frame_jacobian6_dot(m, q, v, 5, np.zeros(3))
        s["Jee"] = rbd.frame_jacobian6(m, kin, m.ee_joint, m.ee_off)
        s["dJee"] = rbd.frame_jacobian6_dot(m, q, v, m.ee_joint, m.ee_off)
        s["foot_pos"] = np.array([rbd.frame_position(m, kin, m.foot_joint[i], m.foot_off[i]) for i in range(4)])
        s["foot_vel"] = (s["J"] @ v).reshape(4, 3)
        s["ee_pos"] = rbd.frame_position(m, kin, m.ee_joint, m.ee_off)
        s["ee_vel"] = s["Jee"] @ v
        s["ee_rot"] = kin["R"][m.ee_joint] @ m.ee_Roff
        return s

    # -- updateDesired (WbcBase.cpp:205-238); stateful input_last
    def update_desired(self, x_des, u_des, period):
        m = self.m
        nk = ce.node_kinematics(m, x_des, u_des)
        q, v = x_des[6:30].copy(), nk["v"]
        s = dict(q=q, v=v)
        joint_acc = (u_des - self.input_last)[12:30] / period
        self.input_last = u_des.copy()
        A = nk["A"]
        Adot = rbd.cmm_dot(m, q, v)
        F = u_des[:12].reshape(4, 3)
        hdot = m.total_mass * np.concatenate([F.sum(0) / m.total_mass + rbd.GRAVITY,
                                              rbd._cross(nk["foot_pos"] - nk["com"], F).sum(0) / m.total_mass])
        rhs = hdot - Adot @ v - A[:, 6:] @ joint_acc
        s["base_acc"] = np.linalg.solve(A[:, :6], rhs)
        kin = nk["kin"]
        s["foot_pos"] = nk["foot_pos"]
        s["foot_vel"] = nk["foot_vel"]
        s["ee_pos"] = nk["ee_pos"]
        s["ee_vel"] = rbd.frame_jacobian6(m, kin, m.ee_joint, m.ee_off) @ v
        s["ee_rot"] = nk["ee_rot"]
        return s

    # -- tasks (WbcBase.cpp:240-578)
    def tasks(self, M, D, mode, u_des):
        g, mu = self.g, self.mu
        flags = G.stance_legs(mode)
        nc = sum(flags)
        Mm, h, J, dJ, v = M["M"], M["nle"], M["J"], M["dJ"], M["v"]
        T = {}
        T["eom"] = Task(np.hstack([Mm[:6], -J.T[:6]]), -h[:6])
        Dj = np.hstack([Mm[6:], -J.T[6:]])
        lim = np.concatenate([self.tau_max])
        T["torque"] = Task(d=np.vstack([Dj, -Dj]), f=np.concatenate([lim - h[6:], lim + h[6:]]))
        a = np.zeros((3 * nc, 36)); b = np.zeros(3 * nc); j = 0
        for i in range(4):
            if flags[i]:
                a[3 * j:3 * j + 3, :24] = J[3 * i:3 * i + 3]
                b[3 * j:3 * j + 3] = -dJ[3 * i:3 * i + 3] @ v
                j += 1
        T["no_contact_motion"] = Task(a, b)
        a = np.zeros((3 * (4 - nc), 36)); j = 0
        for i in range(4):
            if not flags[i]:
                a[3 * j:3 * j + 3, 24 + 3 * i:27 + 3 * i] = np.eye(3); j += 1
        pyr = np.array([[0, 0, -1], [1, 0, -mu], [-1, 0, -mu], [0, 1, -mu], [0, -1, -mu]], dtype=float)
        d = np.zeros((5 * nc + 3 * (4 - nc), 36)); j = 0
        for i in range(4):
            if flags[i]:
                d[5 * j:5 * j + 5, 24 + 3 * i:27 + 3 * i] = pyr; j += 1
        T["friction"] = Task(a, np.zeros(a.shape[0]), d, np.zeros(d.shape[0]))
        qm, qd, vd = M["q"], D["q"], D["v"]
        a = np.zeros((1, 36)); a[0, 2] = 1
        T["base_height"] = Task(a, [D["base_acc"][2] + g["baseHeightKp"] * (qd[2] - qm[2]) + g["baseHeightKd"] * (vd[2] - v[2])])
        a = np.zeros((3, 36)); a[:, :24] = M["Jb"][3:6]
        e = qm[3:6]
        Tm = euler_zyx_map(e)
        w_meas, w_des = Tm @ v[3:6], Tm @ vd[3:6]
        err = rotation_error_in_world(rot_zyx(qd[3:6]), rot_zyx(e))
        acc_des = Tm @ D["base_acc"][3:6] + euler_zyx_map_dot(e, vd[3:6]) @ vd[3:6]
        T["base_angular"] = Task(a, acc_des + g["kp_base_angular"] * err + g["kd_base_angular"] * (w_des - w_meas) - M["dJb"][3:6] @ v)
        a = np.zeros((2, 36)); a[:, :2] = np.eye(2)
        T["base_linear"] = Task(a, D["base_acc"][:2] + g["kp_base_linear"] * (qd[:2] - qm[:2]) + g["kd_base_linear"] * (vd[:2] - v[:2]))
        a = np.zeros((3 * (4 - nc), 36)); b = np.zeros(3 * (4 - nc)); j = 0
        for i in range(4):
            if not flags[i]:
                acc = g["kp_swing"] * (D["foot_pos"][i] - M["foot_pos"][i]) + g["kd_swing"] * (D["foot_vel"][i] - M["foot_vel"][i])
                a[3 * j:3 * j + 3, :24] = J[3 * i:3 * i + 3]
                b[3 * j:3 * j + 3] = acc - dJ[3 * i:3 * i + 3] @ v
                j += 1
        T["swing"] = Task(a, b)
        a = np.zeros((6, 36)); a[:, 18:24] = np.eye(6)
        T["arm_joint"] = Task(a, np.array(g["kp_arm_joint"]) * (qd[18:] - qm[18:]) + np.array(g["kd_arm_joint"]) * (vd[18:] - v[18:]))
        a = np.zeros((3, 36)); a[:, :24] = M["Jee"][:3]
        lin = np.array(g["kp_ee_linear"]) * (D["ee_pos"] - M["ee_pos"]) + np.array(g["kd_ee_linear"]) * (D["ee_vel"][:3] - M["ee_vel"][:3])
        T["ee_linear"] = Task(a, lin - M["dJee"][:3] @ v)
        a = np.zeros((3, 36)); a[:, :24] = M["Jee"][3:6]; a[:, 3:6] = 0
        dj = M["dJee"][3:6].copy(); dj[:, 3:6] = 0
        err = rotation_error_in_world(D["ee_rot"], M["ee_rot"])
        T["ee_angular"] = Task(a, np.array(g["kp_ee_angular"]) * err + np.array(g["kd_ee_angular"]) * (-M["ee_vel"][3:6]) - dj @ v)
        a = np.zeros((12, 36)); a[:, 24:] = np.eye(12)
        T["contact_force"] = Task(a, u_des[:12])
        return T

    def update(self, x_des, u_des, rbd_meas, mode, period, time, return_debug=False):
        """HierarchicalWbc::update (or HierarchicalMpcWbc::update) -> cmd[54] = [x*(36); tau(18)]."""
        M = self.update_measured(rbd_meas)
        D = self.update_desired(x_des, u_des, period)
        T = self.tasks(M, D, mode, u_des)
        task0 = T["eom"] + T["torque"] + T["no_contact_motion"] + T["friction"]
        if int(self.mpc_variant) == 2:
            # SYNTHETIC six-level split of the same tasks (BASELINE config 5 names "6 task levels"; the reference's stacks have
            # three, HierarchicalWbc.cpp:23-43): the construction of every level is HoQp's (HoQp.cpp:12-158). A level without
            # rows (the swing level in full stance) changes nothing and is skipped.
            lower = [T["base_height"] + T["base_angular"], T["ee_linear"] + T["ee_angular"], T["swing"] * 100, T["base_linear"],
                     T["contact_force"]]
        elif self.mpc_variant:
            lower = [T["base_height"] + T["base_angular"] + T["base_linear"] + T["swing"] * 100, T["contact_force"]]
        else:
            lower = [T["arm_joint"] if time < 10 else (T["base_height"] + T["base_angular"] + T["ee_linear"] + T["ee_angular"] + T["swing"] * 100),
                     T["contact_force"] + T["base_linear"]]
        levels = [HoQp(task0)]
        for task in lower:
            if task.a.shape[0] == 0 and task.d.shape[0] == 0:
                continue
            levels.append(HoQp(task, levels[-1]))
        l0, l1, l2 = levels[0], levels[1], levels[-1]
        task1, task2 = lower[0], lower[-1]
        x = levels[-1].x
        tau = np.hstack([M["M"][6:], -M["J"].T[6:]]) @ x + M["nle"][6:]             # updateCmd (WbcBase.cpp:580-595)
        cmd = np.concatenate([x, tau])
        if return_debug:
            return cmd, dict(levels=(l0, l1, l2), all_levels=levels, tasks=(task0, task1, task2), M=M, D=D, T=T)
        return cmd


def rbd_from_state(model, x, v_pin):
    """Measured rbdState(55) layout of qm_estimation/src/StateEstimateBase.cpp:29-102 from (q, generalized velocity)."""
    q = x[6:30]
    r = np.zeros(55)
    r[0:3], r[3:6], r[6:24] = q[3:6], q[0:3], q[6:24]
    r[24:27] = euler_zyx_map(q[3:6]) @ v_pin[3:6]
    r[27:30] = v_pin[0:3]
    r[30:48] = v_pin[6:24]
    return r
