// POD descriptors shared by the host library, the CUDA kernels and the C-ABI (include/qmb200.h).
// Layouts follow SURVEY.md App. A: state x[30] = [normalized momentum(6); base pos(3); ZYX Euler(3); joints(18)],
// input u[30] = [contact forces LF,RF,LH,RH (12); joint velocities (18)].
#pragma once
#include <stdint.h>

#define QM_NX 30
#define QM_NU 30
#define QM_NJ 24          // one-DoF joints: 3 prismatic + 3 revolute (ZYX Euler root) + 18 actuated
#define QM_NFEET 4
#define QM_NUT 18         // max reduced input dimension after projection (14 + #stance feet)
#define QM_NCV 12         // max velocity-constraint rows (3*#stance + #swing)
#define QM_NTARGET 37     // target knot: [x_ref(30); ee position(3); ee quaternion xyzw(4)]

#ifdef __cplusplus
extern "C" {
#endif

// Rigid-body tree. Replaces the pinocchio::Model built at qm_interface/src/QMInterface.cpp:408-416.
typedef struct qmb200_model_desc {
  int32_t nj;
  int32_t parent[QM_NJ];     // -1 = world
  int32_t jtype[QM_NJ];      // 0 prismatic, 1 revolute
  int32_t depth[QM_NJ];
  uint32_t submask[QM_NJ];   // bit i set  <=> body i lies in the subtree of joint j (j included)
  uint32_t pathmask[QM_NJ];  // bit k set  <=> joint k is an ancestor-or-self of joint j
  int32_t max_depth;
  int32_t foot_joint[QM_NFEET];
  int32_t ee_joint;
  double axis[QM_NJ][3];     // joint axis, joint frame
  double Rp[QM_NJ][9];       // joint frame placement in parent body frame (row major)
  double pp[QM_NJ][3];
  double mass[QM_NJ];
  double com[QM_NJ][3];      // body frame
  double inertia[QM_NJ][9];  // about com, body frame
  double foot_off[QM_NFEET][3];
  double ee_off[3];
  double ee_Roff[9];
  double total_mass;
  double lower[QM_NJ], upper[QM_NJ], effort[QM_NJ];
  int32_t root6_standard;    // joints 0..5 are the standard floating base (prismatic x,y,z; revolute z,y,x; identity placements)
  int32_t reserved;
} qmb200_model_desc;

// Optimal-control-problem constants. Replaces what QMInterface::setupOptimalControlProblem
// (qm_interface/src/QMInterface.cpp:79-142) extracts from task.info / reference.info.
typedef struct qmb200_problem_desc {
  double Q[QM_NX * QM_NX];         // task.info:193-234
  double R[QM_NU * QM_NU];         // J'R J already applied (QMInterface.cpp:274-299)
  double mu_ee_pos, mu_ee_ori;     // "endEffector"      task.info:236-240
  double mu_fee_pos, mu_fee_ori;   // "finalEndEffector" task.info:241-246
  double fric_mu, fric_bar_mu, fric_bar_delta;  // task.info:291-298
  double fric_reg, fric_grip, fric_hess_shift;  // [upstream] FrictionConeConstraint::Config defaults
  double pos_bar_mu, pos_bar_delta, vel_bar_mu, vel_bar_delta;  // task.info:300-316
  double arm_pos_lo[6], arm_pos_hi[6], arm_vel_lo[6], arm_vel_hi[6];
  double box_offset;               // StateInputSoftBoxConstraint::initializeOffset(0,0,0) (QMInterface.cpp:257)
  double swing_liftoff_vel, swing_touchdown_vel, swing_height, swing_time_scale;  // task.info:24-31
  double gravity;                  // 9.81
} qmb200_problem_desc;

// Solver settings. task.info:76-93 (sqp) and :139-149 (mpc) + [upstream] sqp::Settings defaults.
typedef struct qmb200_solver_desc {
  double dt;            // sqp.dt
  double horizon;       // mpc.timeHorizon
  double delta_tol, g_max, g_min;
  double alpha_decay, alpha_min, gamma_c, armijo_factor;
  double weak_eps;      // ocs2::numeric_traits::weakEpsilon
  double dt_min;        // 10 * limitEpsilon
  int32_t max_nodes;    // capacity of the node axis (>= horizon/dt + 1 + 2*events in horizon)
  int32_t max_events;   // capacity of a problem's mode schedule
  int32_t max_targets;  // capacity of target knots
  int32_t sqp_iterations;  // sqp.sqpIteration (task.info:80); values < 1 mean 1
  double cost_tol;      // [upstream] sqp::Settings::costTol (1e-4): merit change below which a feasible iterate counts as converged
} qmb200_solver_desc;

// Command -> reference conversion constants. Replaces the file-scope globals of
// qm_controllers/src/QmTargetTrajectoriesPublisher_node.cpp:18-25 (filled in its main(), :268-272) and ARM_DIST of
// qm_controllers/include/qm_controllers/StartingPosition.h:13.
typedef struct qmb200_target_desc {
  double com_height;                     // reference.info comHeight
  double feet_height;                    // mean z of the feet in contact (runtime topic, :28-35); 0 until set
  double arm_dist;                       // base CoM to end-effector distance in the XY plane (0.6)
  double time_to_target;                 // task.info mpc.timeHorizon
  double target_displacement_velocity;   // reference.info targetDisplacementVelocity
  double target_rotation_velocity;       // reference.info targetRotationVelocity
  double default_joint_state[18];        // reference.info defaultJointState
} qmb200_target_desc;

// Whole-body-controller gains and limits. Defaults: qm_wbc/cfg/wbcWigeht.cfg:7-47 (delivered by WbcBase::dynamicCallback,
// qm_wbc/src/WbcBase.cpp:74-121), torque limits and friction coefficient from WbcBase::loadTasksSetting (WbcBase.cpp:597-627).
typedef struct qmb200_wbc_desc {
  double kp_swing, kd_swing;
  double kp_base_height, kd_base_height;
  double kp_base_linear, kd_base_linear;
  double kp_base_angular, kd_base_angular;
  double kp_arm_joint[6], kd_arm_joint[6];
  double kp_ee_linear[3], kd_ee_linear[3];
  double kp_ee_angular[3], kd_ee_angular[3];
  double friction_mu;          // task.info frictionConeTask.frictionCoefficient
  double tau_max[18];          // URDF effort limits in joint order
  double swing_weight;         // 100 (HierarchicalWbc.cpp:29)
  double init_time;            // 10 s: before it level 1 is the arm-joint tracking task (HierarchicalWbc.cpp:32-43)
  double gravity;
  int32_t mpc_variant;         // task stack: 0 HierarchicalWbc, 1 HierarchicalMpcWbc, 2 six-level split of the same tasks (synthetic, BASELINE config 5)
  int32_t reserved;
} qmb200_wbc_desc;

// Joint-level control law and simulated actuator with transport delay.
// Control law: QMController::updateControlLaw (qm_controllers/src/QMController.cpp:178-191): legs (joints 0..11) are commanded
// (q_des, v_des, kp 0, kd 3, tau) once the observation time exceeds 10 s, the arm (12..17) (q_des, 0, kp_arm_wbc, kd_arm_wbc, tau)
// with the gains of qm_controllers/cfg/weight.cfg:7-8. Actuator: QMHWSim::writeSim (qm_gazebo/src/QMHWSim.cpp:98-114), delay of
// qm_gazebo/config/default.yaml:2; ros::Time arithmetic is integer nanoseconds, so are the stamps here.
typedef struct qmb200_actuator_desc {
  double leg_kp, leg_kd;       // 0, 3
  double arm_kp, arm_kd;       // 0, 0.5
  double leg_enable_time;      // 10 s
  int64_t delay_ns;            // 9 000 000
} qmb200_actuator_desc;
#define QMB200_ACT_CAPACITY 32   /* commands kept per problem; more than delay / period + 2 are never alive */

#ifdef __cplusplus
}
#endif
