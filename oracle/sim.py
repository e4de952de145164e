"""Rigid-body forward-dynamics step behind the simulated actuator (TEST INFRASTRUCTURE).

In the reference the loop controller -> actuator -> robot is closed by Gazebo (qm_gazebo/src/QMHWSim.cpp:98-114 writes the joint
efforts, the physics engine integrates). Gazebo is not part of the hot path and is not rebuilt; for batched disturbance studies
the library offers one explicit step of the articulated-body equations with the stance feet held by bilateral point contacts
(no slip, flat ground) -- a labelled stand-in, not Gazebo's contact model:

    [ M(q)  -Jc' ] [ qdd ]   [ S' tau - h(q, v)                 ]
    [ Jc     0   ] [ f   ] = [ -dJc v - (beta / dt) Jc v        ]        (Jc: rows of the stance feet)

    v+ = v + dt qdd,   q+ = q + dt v+      (semi-implicit Euler; the base coordinates are position + ZYX Euler angles, q' = v)

State in and out in the estimator's rbd layout (qm_estimation/src/StateEstimateBase.cpp:29-102), so the step chains with
WbcBase::update: [zyx; base position; joints; world angular velocity; base linear velocity; joint velocities; ee position; ee quat].
"""
import numpy as np

from . import centroidal as ce
from . import gait as G
from . import rbd
from .wbc import Wbc, euler_zyx_map


def forward_dynamics_step(model, P, rbdm, tau, mode, dt, beta=0.0):
    """-> (rbd_next[55], contact forces[12] (zero rows for swing feet), qdd[24])."""
    s = Wbc(model, P).update_measured(np.asarray(rbdm, dtype=float))
    q, v = s["q"], s["v"]
    flags = G.stance_legs(int(mode))
    rows = [3 * i + d for i in range(4) if flags[i] for d in range(3)]
    nc = len(rows)
    Jc = s["J"][rows]
    dJv = (s["dJ"] @ v)[rows]
    n = 24 + nc
    K = np.zeros((n, n))
    K[:24, :24] = s["M"]
    K[:24, 24:] = -Jc.T
    K[24:, :24] = Jc
    rhs = np.concatenate([np.concatenate([np.zeros(6), tau]) - s["nle"], -dJv - (beta / dt) * (Jc @ v)])
    sol = np.linalg.solve(K, rhs)
    qdd = sol[:24]
    f = np.zeros(12)
    f[rows] = sol[24:]
    vn = v + dt * qdd
    qn = q + dt * vn
    out = np.zeros(55)
    out[0:3], out[3:6], out[6:24] = qn[3:6], qn[0:3], qn[6:]
    out[24:27] = euler_zyx_map(qn[3:6]) @ vn[3:6]
    out[27:30] = vn[0:3]
    out[30:48] = vn[6:]
    kin = rbd.kinematics(model, qn)
    out[48:51] = rbd.frame_position(model, kin, model.ee_joint, model.ee_off)
    out[51:55] = ce.quat_from_matrix(kin["R"][model.ee_joint] @ model.ee_Roff)
    return out, f, qdd
