"""Centroidal model of the oracle (TEST INFRASTRUCTURE; see oracle/__init__.py).

Restates [upstream] ocs2_centroidal_model as used by the reference:
* flow map  xdot = f(x,u): QMDynamicsAD::computeFlowMap (qm_interface/src/dynamics/QMDynamicsAD.cpp:22-25)
  -> PinocchioCentroidalDynamicsAD: [ normalized momentum rate ; pinocchio joint velocity ]
* CentroidalModelPinocchioMapping::getPinocchioJointVelocity (full centroidal model, task.info:1)
* PinocchioEndEffectorKinematicsCppAd position / velocity / orientation error
  (built at qm_interface/src/QMInterface.cpp:363-378)
All functions are complex-safe and vectorised over leading dims; Jacobians by complex step.
"""
import numpy as np

from . import rbd

# input layout: u[0:12] contact forces LF,RF,LH,RH ; u[12:30] joint velocities
# state layout: x[0:6] normalized momentum, x[6:30] generalized coordinates


def pin_velocity(model, kin, x, u, A=None):
    """v = [A_b^{-1}(m h_n - A_j v_j); v_j]."""
    if A is None:
        A, _ = rbd.centroidal_momentum_matrix(model, kin)
    vj = u[..., 12:30]
    rhs = model.total_mass * x[..., 0:6] - (A[..., :, 6:] @ vj[..., None])[..., 0]
    vb = np.linalg.solve(A[..., :, 0:6], rhs[..., None])[..., 0]
    return np.concatenate([vb, vj + 0 * vb[..., :1]], axis=-1)


def node_kinematics(model, x, u=None):
    """Everything a node needs, from one pass. x (...,30) [, u (...,30)]."""
    q = x[..., 6:30]
    kin = rbd.kinematics(model, q)
    A, c = rbd.centroidal_momentum_matrix(model, kin)
    out = dict(kin=kin, A=A, com=c)
    out["foot_pos"] = np.stack([rbd.frame_position(model, kin, model.foot_joint[i], model.foot_off[i])
                                for i in range(4)], axis=-2)
    out["ee_pos"] = rbd.frame_position(model, kin, model.ee_joint, model.ee_off)
    out["ee_rot"] = kin["R"][..., model.ee_joint, :, :] @ model.ee_Roff
    if u is not None:
        v = pin_velocity(model, kin, x, u, A)
        out["v"] = v
        out["foot_vel"] = np.stack(
            [(rbd.point_jacobian(model, kin, model.foot_joint[i], out["foot_pos"][..., i, :]) @ v[..., None])[..., 0]
             for i in range(4)], axis=-2)
    return out


def flow_map(model, x, u, nk=None):
    """f(x,u) (...,30).  [upstream] getNormalizedCentroidalMomentumRate + getPinocchioJointVelocity."""
    if nk is None:
        nk = node_kinematics(model, x, u)
    m = model.total_mass
    F = u[..., 0:12].reshape(u.shape[:-1] + (4, 3))
    lin = F.sum(axis=-2) / m + rbd.GRAVITY
    arm = nk["foot_pos"] - nk["com"][..., None, :]
    ang = rbd._cross(arm, F).sum(axis=-2) / m
    return np.concatenate([lin, ang, nk["v"]], axis=-1)


def cstep_jacobian(fun, z, h=1e-30):
    """Jacobian of fun: (...,n) -> (...,m) by complex step; returns (...,m,n)."""
    z = np.asarray(z, dtype=float)
    n = z.shape[-1]
    zc = z[..., None, :] + 1j * h * np.eye(n)           # (...,n,n): perturbation k in row k
    out = fun(zc)                                        # (...,n,m)
    return np.swapaxes(np.imag(out) / h, -1, -2)


def flow_map_linearization(model, x, u):
    """f, df/dx, df/du — what the CppAD model returns (QMDynamicsAD.cpp:30-33)."""
    xu = np.concatenate([x, u], axis=-1)
    J = cstep_jacobian(lambda z: flow_map(model, z[..., :30], z[..., 30:]), xu)
    return flow_map(model, x, u), J[..., :, :30], J[..., :, 30:]


# ----------------------------------------------------------------------------- quaternions (x,y,z,w)
def quat_from_matrix(R):
    """Eigen::Quaternion(Matrix3) (Shepperd's branches, selected on real parts). Returns (...,4) x,y,z,w."""
    R = np.asarray(R)
    m = lambda a, b: R[..., a, b]
    t = m(0, 0) + m(1, 1) + m(2, 2)
    with np.errstate(all="ignore"):
        cands = []
        s = np.sqrt(t + 1.0)
        si = 0.5 / s
        cands.append(np.stack([(m(2, 1) - m(1, 2)) * si, (m(0, 2) - m(2, 0)) * si, (m(1, 0) - m(0, 1)) * si, 0.5 * s], -1))
        for i in range(3):
            j = (i + 1) % 3
            k = (j + 1) % 3
            s = np.sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0)
            si = 0.5 / s
            qv = [None] * 4
            qv[i] = 0.5 * s
            qv[3] = (m(k, j) - m(j, k)) * si
            qv[j] = (m(j, i) + m(i, j)) * si
            qv[k] = (m(k, i) + m(i, k)) * si
            cands.append(np.stack(qv, -1))
    d = np.stack([m(0, 0).real, m(1, 1).real, m(2, 2).real], -1)
    imax = np.where(d[..., 1] > d[..., 0], 1, 0)
    dsel = np.take_along_axis(d, imax[..., None], -1)[..., 0]
    imax = np.where(d[..., 2] > dsel, 2, imax)
    sel = np.where(t.real > 0, 0, imax + 1)
    out = cands[0]
    for b in range(1, 4):
        out = np.where((sel == b)[..., None], cands[b], out)
    return out


def quat_distance(q, qref):
    """[upstream] ocs2::quaternionDistance: q.w*qref.vec - qref.w*q.vec + q.vec x qref.vec."""
    return q[..., 3:4] * qref[..., :3] - qref[..., 3:4] * q[..., :3] + rbd._cross(q[..., :3], qref[..., :3])


def quat_slerp(q0, q1, t):
    """Eigen::QuaternionBase::slerp(t, other), real inputs, (x,y,z,w)."""
    d = float(np.dot(q0, q1))
    ad = abs(d)
    if ad >= 1.0 - np.finfo(float).eps:
        s0, s1 = 1.0 - t, t
    else:
        th = np.arccos(ad)
        st = np.sin(th)
        s0 = np.sin((1.0 - t) * th) / st
        s1 = np.sin(t * th) / st
    if d < 0:
        s1 = -s1
    return s0 * q0 + s1 * q1


def ee_error(model, x, pos_ref, quat_ref):
    """EndEffectorConstraint::getValue (qm_interface/src/constraint/EndEffectorConstraint.cpp:36-49)."""
    nk = node_kinematics(model, x)
    e_pos = nk["ee_pos"] - pos_ref
    e_ori = quat_distance(quat_from_matrix(nk["ee_rot"]), quat_ref)
    return np.concatenate([e_pos, e_ori], axis=-1)


def state_from_rbd(model, rbd_state, yaw_last=None):
    """[upstream] CentroidalModelRbdConversions::computeCentroidalStateFromRbdModel as called from
    QMController::updateStateEstimation (qm_controllers/src/QMController.cpp:239-243) and the yaw unwrapping of :244.
    rbd_state[55]: estimator layout (qm_estimation/src/StateEstimateBase.cpp:29-102)."""
    r = np.asarray(rbd_state, dtype=float)
    q = np.concatenate([r[3:6], r[0:3], r[6:24]])
    yaw, pitch = q[3], q[4]
    # world angular velocity = T(euler) d/dt[yaw, pitch, roll]  ([upstream] getEulerAnglesZyxDerivativesFromGlobalAngularVelocity)
    T = np.array([[0.0, -np.sin(yaw), np.cos(yaw) * np.cos(pitch)],
                  [0.0, np.cos(yaw), np.sin(yaw) * np.cos(pitch)],
                  [1.0, 0.0, -np.sin(pitch)]])
    v = np.concatenate([r[27:30], np.linalg.solve(T, r[24:27]), r[30:48]])
    kin = rbd.kinematics(model, q)
    A, _ = rbd.centroidal_momentum_matrix(model, kin)
    x = np.concatenate([A @ v / model.total_mass, q])
    if yaw_last is not None:
        d = np.fmod(np.fmod(x[9] - yaw_last, 2 * np.pi) + 2 * np.pi, 2 * np.pi)     # angles::shortest_angular_distance
        if d > np.pi:
            d -= 2 * np.pi
        x[9] = yaw_last + d
    return x

