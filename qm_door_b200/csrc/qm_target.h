// Command -> TargetTrajectories conversion: the three pure functions of
// qm_controllers/src/QmTargetTrajectoriesPublisher_node.cpp that turn a base velocity command, an end-effector velocity
// command or an end-effector goal pose into the two-knot reference the MPC tracks. One thread per problem; the same source is
// the kernel body and the CPU port. Output layout = the reference's stateTrajectory: [x_ref(30); ee position(3); ee quat xyzw(4)].
#pragma once
#include "qm_mpc.h"

namespace qm {

enum { TG_BASE_CMD_VEL = 0, TG_EE_CMD_VEL = 1, TG_EE_GOAL = 2 };

// [upstream] getRotationMatrixFromZyxEulerAngles: R = Rz(z) Ry(y) Rx(x)
QM_HD void rot_from_zyx(const double* e, double* R) {
  double sz, cz, sy, cy, sx, cx;
  sincos(e[0], &sz, &cz); sincos(e[1], &sy, &cy); sincos(e[2], &sx, &cx);
  R[0] = cz * cy; R[1] = cz * sy * sx - sz * cx; R[2] = cz * sy * cx + sz * sx;
  R[3] = sz * cy; R[4] = sz * sy * sx + cz * cx; R[5] = sz * sy * cx - cz * sx;
  R[6] = -sy;     R[7] = cy * sx;                R[8] = cy * cx;
}
// Eigen::Quaternion::toRotationMatrix for (x, y, z, w)
QM_HD void rot_from_quat(const double* q, double* R) {
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// targetPoseToTargetTrajectories (QmTargetTrajectoriesPublisher_node.cpp:60-86)
QM_HDN void target_pose_to_trajectories(const qmb200_target_desc& D, const double* ee_target, const double* base_target,
                                        double obs_time, const double* obs_state, const double* ee_current, double reach_time,
                                        double* tt, double* tx) {
  tt[0] = obs_time; tt[1] = reach_time;
  double base_cur[6];
  for (int k = 0; k < 6; ++k) base_cur[k] = obs_state[6 + k];
  base_cur[2] = D.com_height + D.feet_height;
  base_cur[4] = 0.0; base_cur[5] = 0.0;
  for (int knot = 0; knot < 2; ++knot) {
    double* s = tx + QM_NTARGET * knot;
    for (int k = 0; k < 6; ++k) s[k] = 0.0;
    for (int k = 0; k < 6; ++k) s[6 + k] = knot == 0 ? base_cur[k] : base_target[k];
    for (int k = 0; k < 18; ++k) s[12 + k] = D.default_joint_state[k];
    for (int k = 0; k < 7; ++k) s[30 + k] = knot == 0 ? ee_current[k] : ee_target[k];
  }
}

// One command. cmd: (vx, vy, vz, yaw rate) for the velocity commands, (position, quat xyzw) for the goal.
// last_ee[7] is the publisher's lastEeTarget_ state (read and updated as the reference does).
QM_HDN void command_to_target(const qmb200_target_desc& D, int kind, const double* cmd, double obs_time, const double* obs_state,
                              const double* ee_state, double* last_ee, double* tt, double* tx) {
  const double* base_cur = obs_state + 6;
  if (kind == TG_BASE_CMD_VEL) {
    // cmdVelToTargetTrajectories (:91-134)
    double R[9], v[3];
    rot_from_zyx(base_cur + 3, R);
    for (int r = 0; r < 3; ++r) v[r] = R[3 * r] * cmd[0] + R[3 * r + 1] * cmd[1] + R[3 * r + 2] * cmd[2];
    const double T = D.time_to_target;
    double base_target[6] = {base_cur[0] + v[0] * T, base_cur[1] + v[1] * T, D.com_height + D.feet_height, base_cur[3] + cmd[3] * T, 0.0, 0.0};
    const double d0 = last_ee[0] - ee_state[0], d1 = last_ee[1] - ee_state[1], d2 = last_ee[2] - ee_state[2];
    if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 0.1) for (int k = 0; k < 3; ++k) last_ee[k] = ee_state[k];
    double ee_target[7];
    for (int k = 0; k < 7; ++k) ee_target[k] = last_ee[k];
    target_pose_to_trajectories(D, ee_target, base_target, obs_time, obs_state, ee_target, obs_time + T, tt, tx);
    for (int k = 0; k < 3; ++k) { tx[k] = v[k]; tx[QM_NTARGET + k] = v[k]; }
  } else if (kind == TG_EE_CMD_VEL) {
    // EeCmdVelToTargetTrajectories (:139-194)
    const double half = 0.5 * base_cur[3];
    const double qi[4] = {0.0, 0.0, sin(half), cos(half)};
    double Rq[9], Ri[9], t[3], v[3];
    rot_from_quat(ee_state + 3, Rq);
    rot_from_quat(qi, Ri);
    for (int r = 0; r < 3; ++r) t[r] = Ri[r] * cmd[0] + Ri[3 + r] * cmd[1] + Ri[6 + r] * cmd[2];      // Ri' cmd
    for (int r = 0; r < 3; ++r) v[r] = Rq[3 * r] * t[0] + Rq[3 * r + 1] * t[1] + Rq[3 * r + 2] * t[2];
    const double T = D.time_to_target;
    double ee_target[7];
    for (int k = 0; k < 7; ++k) ee_target[k] = ee_state[k];
    ee_target[0] = ee_state[0] + v[0] * T;
    ee_target[1] = ee_state[1] + v[1] * T;
    ee_target[2] = last_ee[2];
    ee_target[3] = last_ee[3];
    ee_target[4] = last_ee[4];
    ee_target[5] = ee_state[5] + sin(v[2] * T / 2);
    ee_target[6] = ee_state[6] + cos(v[2] * T / 2);
    const double siny = 2.0 * (ee_target[6] * ee_target[5] + ee_target[3] * ee_target[4]);
    const double cosy = 1.0 - 2.0 * (ee_target[4] * ee_target[4] + ee_target[5] * ee_target[5]);
    const double yaw = atan2(siny, cosy);
    double base_target[6] = {ee_target[0] - D.arm_dist * cos(base_cur[3]), ee_target[1] - D.arm_dist * sin(base_cur[3]),
                             D.com_height + D.feet_height, yaw, 0.0, 0.0};
    target_pose_to_trajectories(D, ee_target, base_target, obs_time, obs_state, ee_state, obs_time + T, tt, tx);
  } else {
    // EEgoalPoseToTargetTrajectories (:201-241) + positionCommandCallback (:243-257)
    const double* pos = cmd; const double* q = cmd + 3;     // quat xyzw
    double ee_target[7];
    for (int k = 0; k < 7; ++k) ee_target[k] = cmd[k];
    const double siny = 2.0 * (q[3] * q[2] + q[0] * q[1]);
    const double cosy = 1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]);
    const double yaw = atan2(siny, cosy);
    double base_target[6] = {pos[0] - D.arm_dist * cos(yaw), pos[1] - D.arm_dist * sin(yaw), D.com_height + D.feet_height, yaw, 0.0, 0.0};
    // estimateTimeToTarget (:40-57) of [position error; quaternionDistance(current, target)]
    double delta[6], cr[3];
    const double* qc = ee_state + 3;
    for (int k = 0; k < 3; ++k) delta[k] = ee_target[k] - ee_state[k];
    cross3(qc, q, cr);
    for (int k = 0; k < 3; ++k) delta[3 + k] = qc[3] * q[k] - q[3] * qc[k] + cr[k];
    const double disp = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]) / D.target_displacement_velocity;
    const double rot = sqrt(delta[3] * delta[3] + delta[4] * delta[4] + delta[5] * delta[5]) / D.target_rotation_velocity;
    target_pose_to_trajectories(D, ee_target, base_target, obs_time, obs_state, ee_state, obs_time + (rot > disp ? rot : disp), tt, tx);
    for (int k = 0; k < 7; ++k) last_ee[k] = ee_target[k];
  }
}

}  // namespace qm
