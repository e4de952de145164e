"""Rigid-body algorithms of the oracle (TEST INFRASTRUCTURE; see oracle/__init__.py).

Restates, from their textbook definitions (Featherstone), what the reference obtains from
Pinocchio at qm_interface/src/QMPreComputation.cpp:77-87 and qm_wbc/src/WbcBase.cpp:162-202,214-231:
forward kinematics, frame Jacobians (LOCAL_WORLD_ALIGNED) and their time variation, CRBA mass
matrix, non-linear effects, centroidal momentum matrix (CCRBA) and its time derivative (dCCRBA).

Everything is written from the *definitions* (sums over bodies of m*Jv'Jv + Jw'IJw etc.), not the
recursive O(n) algorithms, vectorised over leading batch dims and complex-safe so derivatives and
time variations can be taken by complex-step differentiation.
"""
import numpy as np

GRAVITY = np.array([0.0, 0.0, -9.81])  # [upstream] pinocchio::Model::gravity981


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _skew(a):
    z = np.zeros_like(a[..., 0])
    return np.stack([np.stack([z, -a[..., 2], a[..., 1]], -1),
                     np.stack([a[..., 2], z, -a[..., 0]], -1),
                     np.stack([-a[..., 1], a[..., 0], z], -1)], -2)


def _axis_rot(axis, th):
    """Rodrigues rotation about a constant unit axis; th (...,) possibly complex."""
    K = _skew(np.asarray(axis, dtype=float))
    s = np.sin(th)[..., None, None]
    c = np.cos(th)[..., None, None]
    return np.eye(3) + s * K + (1.0 - c) * (K @ K)


def kinematics(model, q):
    """World placement of every joint frame.  q (...,nj) -> R (...,nj,3,3), p (...,nj,3), a (...,nj,3)."""
    q = np.asarray(q)
    lead = q.shape[:-1]
    dt = np.result_type(q.dtype, np.float64)
    R = np.zeros(lead + (model.nj, 3, 3), dtype=dt)
    p = np.zeros(lead + (model.nj, 3), dtype=dt)
    a = np.zeros(lead + (model.nj, 3), dtype=dt)
    for j in range(model.nj):
        par = model.parent[j]
        if par < 0:
            Rpar = np.broadcast_to(np.eye(3, dtype=dt), lead + (3, 3))
            ppar = np.zeros(lead + (3,), dtype=dt)
        else:
            Rpar, ppar = R[..., par, :, :], p[..., par, :]
        R0 = Rpar @ model.Rp[j]
        p0 = ppar + (Rpar @ model.pp[j])
        aw = R0 @ model.axis[j]
        if model.jtype[j] == 1:
            R[..., j, :, :] = R0 @ _axis_rot(model.axis[j], q[..., j])
            p[..., j, :] = p0
        else:
            R[..., j, :, :] = R0
            p[..., j, :] = p0 + aw * q[..., j, None]
        a[..., j, :] = aw
    return dict(R=R, p=p, a=a)


def point_jacobian(model, kin, body, r):
    """Linear Jacobian (world-aligned) of world point r (...,3) rigidly attached to joint `body`."""
    a, o = kin["a"], kin["p"]
    rev = _cross(a, r[..., None, :] - o)                       # (...,nj,3)
    col = np.where((model.jtype == 1)[:, None], rev, a)
    col = col * model.path[body][:, None]
    return np.swapaxes(col, -1, -2)                             # (...,3,nj)


def angular_jacobian(model, kin, body):
    col = kin["a"] * ((model.jtype == 1) & model.path[body])[:, None]
    return np.swapaxes(col, -1, -2)


def frame_position(model, kin, body, off):
    return kin["p"][..., body, :] + kin["R"][..., body, :, :] @ off


def body_coms(model, kin):
    return kin["p"] + np.einsum("...jab,jb->...ja", kin["R"], model.com)


def world_inertias(model, kin):
    return np.einsum("...jab,jbc,...jdc->...jad", kin["R"], model.inertia, kin["R"])


def com_position(model, kin):
    return np.einsum("j,...ja->...a", model.mass, body_coms(model, kin)) / model.total_mass


def centroidal_momentum_matrix(model, kin):
    """A(q) (...,6,nj): h = A v, h = [linear momentum; angular momentum about the CoM], world-aligned.
    Definition: h = sum_i [m_i cdot_i ; I_i w_i + m_i (c_i - c) x cdot_i]   (Orin & Goswami; pinocchio::ccrba)"""
    cb = body_coms(model, kin)
    c = com_position(model, kin)
    Iw = world_inertias(model, kin)
    lead = cb.shape[:-2]
    A = np.zeros(lead + (6, model.nj), dtype=cb.dtype)
    for i in range(model.nj):
        if model.mass[i] == 0.0:
            continue
        Jv = point_jacobian(model, kin, i, cb[..., i, :])
        Jw = angular_jacobian(model, kin, i)
        A[..., 0:3, :] += model.mass[i] * Jv
        A[..., 3:6, :] += Iw[..., i, :, :] @ Jw + model.mass[i] * (_skew(cb[..., i, :] - c) @ Jv)
    return A, c


def mass_matrix(model, kin):
    """M(q) = sum_i m_i Jv' Jv + Jw' I Jw  (what pinocchio::crba returns after symmetrisation, WbcBase.cpp:165-167)."""
    cb = body_coms(model, kin)
    Iw = world_inertias(model, kin)
    lead = cb.shape[:-2]
    M = np.zeros(lead + (model.nj, model.nj), dtype=cb.dtype)
    for i in range(model.nj):
        if model.mass[i] == 0.0:
            continue
        Jv = point_jacobian(model, kin, i, cb[..., i, :])
        Jw = angular_jacobian(model, kin, i)
        M += model.mass[i] * np.swapaxes(Jv, -1, -2) @ Jv + np.swapaxes(Jw, -1, -2) @ Iw[..., i, :, :] @ Jw
    return M


def _cstep_dir(fun, q, v, h=1e-30):
    """d/deps fun(q + eps v) at eps=0 by complex step (q, v real)."""
    return np.imag(fun(q + 1j * h * v)) / h


def nonlinear_effects(model, q, v):
    """nle(q,v) = C(q,v) v + g(q)  (pinocchio::nonLinearEffects, WbcBase.cpp:170), from the definition
    tau = sum_i Jv_i' m_i (a_i - g) + Jw_i' (I_i alpha_i + w_i x I_i w_i) with qdd = 0, where the bias
    accelerations a_i = d/dt(Jv_i) v and alpha_i = d/dt(Jw_i) v are complex-step directional derivatives."""
    q = np.asarray(q, dtype=float)
    v = np.asarray(v, dtype=float)
    kin = kinematics(model, q)
    cb = body_coms(model, kin)
    Iw = world_inertias(model, kin)
    tau = np.zeros(q.shape)
    for i in range(model.nj):
        if model.mass[i] == 0.0:
            continue

        def lin_vel(qc, i=i):
            k = kinematics(model, qc)
            c = body_coms(model, k)[..., i, :]
            return (point_jacobian(model, k, i, c) @ v[..., None])[..., 0]

        def ang_vel(qc, i=i):
            k = kinematics(model, qc)
            return (angular_jacobian(model, k, i) @ v[..., None])[..., 0]

        a_lin = _cstep_dir(lin_vel, q, v)
        a_ang = _cstep_dir(ang_vel, q, v)
        Jv = point_jacobian(model, kin, i, cb[..., i, :])
        Jw = angular_jacobian(model, kin, i)
        w = (Jw @ v[..., None])[..., 0]
        Iwi = Iw[..., i, :, :]
        wrench_f = model.mass[i] * (a_lin - GRAVITY)
        wrench_n = (Iwi @ a_ang[..., None])[..., 0] + _cross(w, (Iwi @ w[..., None])[..., 0])
        tau += (np.swapaxes(Jv, -1, -2) @ wrench_f[..., None])[..., 0] + (np.swapaxes(Jw, -1, -2) @ wrench_n[..., None])[..., 0]
    return tau


def frame_jacobian6(model, kin, body, off):
    """6 x nj LOCAL_WORLD_ALIGNED frame Jacobian [linear; angular] (pinocchio::getFrameJacobian)."""
    r = frame_position(model, kin, body, off)
    return np.concatenate([point_jacobian(model, kin, body, r), angular_jacobian(model, kin, body)], axis=-2)


def frame_jacobian6_dot(model, q, v, body, off):
    """d/dt of the LOCAL_WORLD_ALIGNED frame Jacobian (pinocchio::getFrameJacobianTimeVariation)."""
    q = np.asarray(q, dtype=float)
    v = np.asarray(v, dtype=float)
    return _cstep_dir(lambda qc: frame_jacobian6(model, kinematics(model, qc), body, off), q, v)


def cmm_dot(model, q, v):
    """Adot(q,v) (pinocchio::dccrba, WbcBase.cpp:230)."""
    q = np.asarray(q, dtype=float)
    v = np.asarray(v, dtype=float)
    return _cstep_dir(lambda qc: centroidal_momentum_matrix(model, kinematics(model, qc))[0], q, v)
