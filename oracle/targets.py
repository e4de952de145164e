"""TEST INFRASTRUCTURE (see oracle/__init__.py): NumPy restatement of the command -> TargetTrajectories conversion of
qm_controllers/src/QmTargetTrajectoriesPublisher_node.cpp. Parity unpinned by the reference (it has
no tests); the functions follow the file line by line."""
import numpy as np

ARM_DIST = 0.6          # qm_controllers/include/qm_controllers/StartingPosition.h:13


class TargetParams:
    """File-scope constants of the node (:18-25), filled in its main() (:268-272)."""

    def __init__(self, P, feet_height=0.0):
        self.com_height = P.com_height
        self.default_joint_state = np.asarray(P.default_joint_state, dtype=float)
        self.time_to_target = P.time_horizon
        self.target_displacement_velocity = P.target_displacement_velocity
        self.target_rotation_velocity = P.target_rotation_velocity
        self.feet_height = feet_height


def rot_zyx(e):
    """[upstream] getRotationMatrixFromZyxEulerAngles."""
    z, y, x = e
    Rz = np.array([[np.cos(z), -np.sin(z), 0], [np.sin(z), np.cos(z), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(y), 0, np.sin(y)], [0, 1, 0], [-np.sin(y), 0, np.cos(y)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(x), -np.sin(x)], [0, np.sin(x), np.cos(x)]])
    return Rz @ Ry @ Rx


def rot_quat(q):
    """Eigen::Quaterniond(w, x, y, z).toRotationMatrix() for q = (x, y, z, w); no normalisation, as in Eigen."""
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quaternion_distance(q, qref):
    """[upstream] quaternionDistance(q, qRef) for (x, y, z, w)."""
    return q[3] * qref[:3] - qref[3] * q[:3] + np.cross(q[:3], qref[:3])


def estimate_time_to_target(p, delta):
    """:40-57"""
    return max(np.linalg.norm(delta[3:6]) / p.target_rotation_velocity, np.linalg.norm(delta[0:3]) / p.target_displacement_velocity)


def target_pose_to_trajectories(p, ee_target, base_target, obs_time, obs_state, ee_current, reach_time):
    """:60-86 -> (times[2], states[2][37])"""
    base_cur = obs_state[6:12].copy()
    base_cur[2] = p.com_height + p.feet_height
    base_cur[4] = 0.0
    base_cur[5] = 0.0
    s0 = np.concatenate([np.zeros(6), base_cur, p.default_joint_state, ee_current])
    s1 = np.concatenate([np.zeros(6), base_target, p.default_joint_state, ee_target])
    return np.array([obs_time, reach_time]), np.stack([s0, s1])


def cmd_vel_to_target(p, cmd_vel, last_ee, obs_time, obs_state, ee_state):
    """:91-134; last_ee is updated in place."""
    base_cur = obs_state[6:12]
    v = rot_zyx(base_cur[3:6]) @ cmd_vel[:3]
    T = p.time_to_target
    base_target = np.array([base_cur[0] + v[0] * T, base_cur[1] + v[1] * T, p.com_height + p.feet_height, base_cur[3] + cmd_vel[3] * T, 0.0, 0.0])
    if np.linalg.norm(last_ee[:3] - ee_state[:3]) > 0.1:
        last_ee[:3] = ee_state[:3]
    ee_target = last_ee.copy()
    tt, tx = target_pose_to_trajectories(p, ee_target, base_target, obs_time, obs_state, ee_target, obs_time + T)
    tx[0, :3] = v
    tx[1, :3] = v
    return tt, tx


def ee_cmd_vel_to_target(p, cmd_vel, last_ee, obs_time, obs_state, ee_state):
    """:139-194"""
    base_cur = obs_state[6:12]
    qi = np.array([0.0, 0.0, np.sin(base_cur[3] / 2), np.cos(base_cur[3] / 2)])
    v = rot_quat(ee_state[3:7]) @ rot_quat(qi).T @ cmd_vel[:3]
    T = p.time_to_target
    ee_target = ee_state.copy()
    ee_target[0] = ee_state[0] + v[0] * T
    ee_target[1] = ee_state[1] + v[1] * T
    ee_target[2] = last_ee[2]
    ee_target[3] = last_ee[3]
    ee_target[4] = last_ee[4]
    ee_target[5] = ee_state[5] + np.sin(v[2] * T / 2)
    ee_target[6] = ee_state[6] + np.cos(v[2] * T / 2)
    yaw = np.arctan2(2.0 * (ee_target[6] * ee_target[5] + ee_target[3] * ee_target[4]),
                     1.0 - 2.0 * (ee_target[4] ** 2 + ee_target[5] ** 2))
    base_target = np.array([ee_target[0] - ARM_DIST * np.cos(base_cur[3]), ee_target[1] - ARM_DIST * np.sin(base_cur[3]),
                            p.com_height + p.feet_height, yaw, 0.0, 0.0])
    return target_pose_to_trajectories(p, ee_target, base_target, obs_time, obs_state, ee_state, obs_time + T)


def ee_goal_to_target(p, goal, last_ee, obs_time, obs_state, ee_state):
    """:201-241 and the lastEeTarget_ update of the callback (:243-257); goal = (position, quat xyzw)."""
    pos, q = goal[:3], goal[3:7]
    yaw = np.arctan2(2.0 * (q[3] * q[2] + q[0] * q[1]), 1.0 - 2.0 * (q[1] ** 2 + q[2] ** 2))
    base_target = np.array([pos[0] - ARM_DIST * np.cos(yaw), pos[1] - ARM_DIST * np.sin(yaw), p.com_height + p.feet_height, yaw, 0.0, 0.0])
    delta = np.concatenate([goal[:3] - ee_state[:3], quaternion_distance(ee_state[3:7], q)])
    out = target_pose_to_trajectories(p, goal.copy(), base_target, obs_time, obs_state, ee_state, obs_time + estimate_time_to_target(p, delta))
    last_ee[:] = goal
    return out


CONVERTERS = (cmd_vel_to_target, ee_cmd_vel_to_target, ee_goal_to_target)
