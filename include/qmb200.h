/* qmb200 — C-ABI of the B200-native MPC + WBC hot path of danisotelo/qm_door.
 *
 * Every entry point is extern "C", takes plain pointers and sizes, and returns an int status (0 = OK,
 * negative = call-level failure, see qmb200_last_error).  Per-problem failures are reported through the
 * `status` output arrays (bit flags QMB200_ST_*).  Citations are relative to the reference checkout.
 *
 *   qmb200_create / destroy          <- QMInterface ctor + setupOptimalControlProblem
 *                                       (qm_interface/src/QMInterface.cpp:37-74, 79-142) and the SqpMpc object built in
 *                                       QMController::setupMpc (qm_controllers/src/QMController.cpp:287-307)
 *   qmb200_mpc_cycle_batch[_dev]     <- MPC_MRT_Interface::advanceMpc() -> SqpSolver::run for B independent problems
 *                                       (QMController.cpp:316-333, :119-122)
 *   qmb200_mpc_reset                 <- coldStart / resetMpcNode semantics (task.info:143)
 *   qmb200_evaluate_policy_batch     <- MPC_MRT_Interface::evaluatePolicy (QMController.cpp:140-143)
 *   qmb200_feedback_gains[_dev], qmb200_evaluate_feedback_policy_batch <- [upstream] SqpSolver::toPrimalSolution /
 *                                       LinearController::computeInput when task.info:90 useFeedbackPolicy is true
 *   qmb200_rbd_to_state_batch[_dev]  <- CentroidalModelRbdConversions::computeCentroidalStateFromRbdModel + yaw unwrapping in
 *                                       QMController::updateStateEstimation (QMController.cpp:239-244); rbd layout of
 *                                       qm_estimation/src/StateEstimateBase.cpp:29-102
 *   qmb200_targets_batch[_dev]       <- cmdVelToTargetTrajectories / EeCmdVelToTargetTrajectories / EEgoalPoseToTargetTrajectories
 *                                       (qm_controllers/src/QmTargetTrajectoriesPublisher_node.cpp:60-257); qmb200_load_targets <- its main() (:268-272)
 *   qmb200_load_urdf                 <- centroidal_model::createPinocchioInterface(urdf, jointNames) (QMInterface.cpp:408-416)
 *   qmb200_load_problem              <- loadData / loadEigenMatrix calls on task.info and reference.info
 *                                       (QMInterface.cpp:65-73,85,152-156,199-234,291,306,395-397)
 *   qmb200_load_gait / tile_schedule <- loadModeSequenceTemplate + GaitSchedule tiling (QMInterface.cpp:455-480,
 *                                       qm_controllers/src/GaitTopicPublisher.cpp:31-44, config/gait.info)
 *
 * Array layouts (row-major): state x[30], input u[30], target knot[37] as in SURVEY.md App. A.1;
 * the node axis of every trajectory output has capacity solver.max_nodes (NMAX) per problem and n_out[b] valid entries.
 */
#ifndef QMB200_H
#define QMB200_H
#include <stdint.h>
#include "../qm_door_b200/csrc/qm_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qmb200_ctx qmb200_ctx;

enum {
  QMB200_ST_OK = 0,
  QMB200_ST_GRID_OVERFLOW = 1,
  QMB200_ST_BAD_SCHEDULE = 2,
  QMB200_ST_RANK = 4,
  QMB200_ST_CHOL = 8,
  QMB200_ST_NAN = 16,
  QMB200_ST_STEP_REJECTED = 32
};
#define QMB200_INFO_SIZE 16   /* per-problem info record: alpha, done, armijo, |dx|, |du|, base(merit,dyn,eq), new(merit,dyn,eq), line-search
                                 trials, |x0 - x[0]|^2, SQP iterations carried out (13), why the SQP loop stopped (14: 1 iteration budget,
                                 2 step size, 3 metrics, 4 primal step; [upstream] SqpSolver::Convergence) */
#define QMB200_NUM_KERNELS 13 /* schedule, init_guess, kin1, kin2, lq, solve, trial, decide, finalize, policy, proj, backtrack, step */

int qmb200_version(void);
const char* qmb200_last_error(void);
int qmb200_device_count(void);

/* ---- host-side ingestion of the reference's input files (no GPU needed) */
int qmb200_load_urdf(const char* urdf_path, qmb200_model_desc* model);
int qmb200_load_problem(const char* task_info, const char* reference_info, const qmb200_model_desc* model,
                        qmb200_problem_desc* problem, qmb200_solver_desc* solver, double* initial_state30);
int qmb200_load_gait(const char* gait_info, const char* gait_name, int32_t capacity, double* switching_times,
                     int32_t* modes, int32_t* num_modes);
int qmb200_tile_schedule(const double* switching_times, const int32_t* modes, int32_t num_modes, double t_insert,
                         double t_upper, int32_t capacity, double* events, int32_t* mode_sequence, int32_t* num_events);

/* ---- context: owns every device buffer (LQ blocks, warm start) for `batch` problems on `device` */
int qmb200_create(const qmb200_model_desc* model, const qmb200_problem_desc* problem, const qmb200_solver_desc* solver,
                  int32_t batch, int32_t device, qmb200_ctx** out);
int qmb200_destroy(qmb200_ctx* ctx);
int qmb200_mpc_reset(qmb200_ctx* ctx);
int qmb200_sync(qmb200_ctx* ctx);
/* Order the context's stream behind the work enqueued so far on `stream` (a cudaStream_t, e.g. qmb200_wbc_stream(..)): an event
 * is recorded there and waited for here; nothing blocks on the host. The _dev entry points of the two contexts run on their own
 * streams, so a device-side chain MPC -> policy -> WBC -> actuator needs one such call per hand-over. */
int qmb200_wait_stream(qmb200_ctx* ctx, void* stream);

/* One SQP cycle for every problem of the batch, HOST buffers (copies inside the call, returns when results are on the host).
 *  t0[B] x0[B][30] events[B][EMAX] modes[B][EMAX+1] nevents[B] target_t[B][KT] target_x[B][KT][37]
 *  -> t_out[B][NMAX] x_out[B][NMAX][30] u_out[B][NMAX][30] n_out[B] mode_out[B][NMAX] info[B][16] status[B]   (outputs may be NULL) */
int qmb200_mpc_cycle_batch(qmb200_ctx* ctx, const double* t0, const double* x0, const double* events, const int32_t* modes,
                           const int32_t* nevents, const double* target_t, const double* target_x, double* t_out,
                           double* x_out, double* u_out, int32_t* n_out, int32_t* mode_out, double* info, int32_t* status);
/* The same cycle as a submit / wait pair (host buffers, which must be page-locked for the copies to be asynchronous and must stay
 * untouched until the wait returns): the call enqueues the input copies, the cycle and -- on a separate copy stream -- the
 * output copies, and returns a ticket; qmb200_mpc_cycle_wait(ticket) blocks until that cycle's outputs are on the host. Up to
 * two submissions may be outstanding, so the copy-out of cycle k overlaps the computation of cycle k + 1 (the reference runs
 * its solver in a separate MPC thread for the same reason, QMController.cpp:316-333). qmb200_mpc_cycle_batch = submit + wait. */
int qmb200_mpc_cycle_batch_async(qmb200_ctx* ctx, const double* t0, const double* x0, const double* events, const int32_t* modes,
                                 const int32_t* nevents, const double* target_t, const double* target_x, double* t_out,
                                 double* x_out, double* u_out, int32_t* n_out, int32_t* mode_out, double* info, int32_t* status,
                                 int64_t* ticket);
int qmb200_mpc_cycle_wait(qmb200_ctx* ctx, int64_t ticket);
/* Same with DEVICE buffers; fully asynchronous on the context's stream (the filter line search and the SQP loop's early exit
 * run on the device: no host synchronisation inside the cycle). solver.sqp_iterations > 1 runs the multi-iteration SQP. */
int qmb200_mpc_cycle_batch_dev(qmb200_ctx* ctx, const double* t0, const double* x0, const double* events,
                               const int32_t* modes, const int32_t* nevents, const double* target_t, const double* target_x,
                               double* t_out, double* x_out, double* u_out, int32_t* n_out, int32_t* mode_out, double* info,
                               int32_t* status);

/* ---- multi-GPU (one process per GPU): independent problems are sharded over the ranks, there is no data-path collective; the
 * policy shard of every rank is all-gathered once per cycle over NCCL (NVLink / NVSwitch). NCCL is bound at run time
 * (dlopen of libnccl.so.2, or $QMB200_NCCL_LIB), so single-GPU hosts need no NCCL.
 *   qmb200_nccl_unique_id        ncclGetUniqueId (128 bytes): call on rank 0, hand to every rank with the host's own bootstrap
 *   qmb200_comm_init             ncclCommInitRank for this context (collective call); the context owns the communicator
 *   qmb200_enable_policy_buffer  without a communicator of the context's own: have k_finalize write the packed policy anyway
 *                                (for use with a caller-owned ncclComm_t)
 *   qmb200_allgather_policy      all-gather of the packed policy of the LAST cycle, [B][NMAX][61] rows (t, x*[30], u*[30]) per
 *                                rank -> gathered[world][B][NMAX][61] (device). nccl_comm: a caller-owned ncclComm_t, or NULL for
 *                                the context's own. Runs on the context's communication stream, behind the cycle and beside the
 *                                next one (two send buffers); qmb200_policy_wait_stream orders a consumer stream behind it,
 *                                qmb200_comm_sync blocks the host until it is complete. */
#define QMB200_POLICY_WIDTH 61
int qmb200_nccl_unique_id(void* id128);
int qmb200_comm_init(qmb200_ctx* ctx, const void* id128, int32_t rank, int32_t world);
int qmb200_comm_destroy(qmb200_ctx* ctx);
int qmb200_enable_policy_buffer(qmb200_ctx* ctx);
const double* qmb200_policy_buffer(qmb200_ctx* ctx);   /* device pointer of the packed policy of the last cycle (or NULL) */
int qmb200_allgather_policy(qmb200_ctx* ctx, void* nccl_comm, double* gathered);
int qmb200_policy_wait_stream(qmb200_ctx* ctx, void* stream);
int qmb200_comm_sync(qmb200_ctx* ctx);

/* Linear interpolation of the stored policy at t[B] (host buffers): x_des[B][30], u_des[B][30], mode[B]. */
int qmb200_evaluate_policy_batch(qmb200_ctx* ctx, const double* t, double* x_des, double* u_des, int32_t* mode);
/* Same with device pointers, enqueued on the context's stream (MPC -> policy -> WBC without leaving the GPU). */
int qmb200_evaluate_policy_batch_dev(qmb200_ctx* ctx, const double* t, double* x_des, double* u_des, int32_t* mode);

/* Command -> two-knot reference (target_t[n][2], target_x[n][2][37] = [x_ref(30); ee position(3); ee quat xyzw(4)]), ready to be
 * passed to qmb200_mpc_cycle_batch. kind: 0 base velocity command, 1 end-effector velocity command (cmd[n][7]: vx, vy, vz,
 * yaw rate, 3 unused), 2 end-effector goal (cmd[n][7]: position, quat xyzw). obs_time[n], obs_state[n][30]: current observation;
 * ee_state[n][7]: measured end-effector pose; last_ee_target[n][7]: the publisher's lastEeTarget_ (read and updated). */
int qmb200_load_targets(const char* task_info, const char* reference_info, qmb200_target_desc* desc);
int qmb200_targets_batch(qmb200_ctx* ctx, const qmb200_target_desc* desc, int32_t kind, int32_t n, const double* cmd,
                         const double* obs_time, const double* obs_state, const double* ee_state, double* last_ee_target,
                         double* target_t, double* target_x);
int qmb200_targets_batch_dev(qmb200_ctx* ctx, const qmb200_target_desc* desc, int32_t kind, int32_t n, const double* cmd,
                             const double* obs_time, const double* obs_state, const double* ee_state, double* last_ee_target,
                             double* target_t, double* target_x);

/* Feedback policy (useFeedbackPolicy, task.info:90; [upstream] SqpSolver::toPrimalSolution / LinearController): gains of the
 * last cycle in the original input coordinates, K_out[B][NMAX][30][30] = Pu K~ + Px per node (pre-event and final nodes repeat
 * the previous node). The context keeps them (allocated on first use); K_out may be NULL for the _dev variant.
 * qmb200_evaluate_feedback_policy_batch: u[B][30] = uff(t) + K(t) x with uff_i = u*_i - K_i x*_i (host buffers). */
int qmb200_feedback_gains(qmb200_ctx* ctx, double* K_out);
int qmb200_feedback_gains_dev(qmb200_ctx* ctx, double* K_out);
int qmb200_evaluate_feedback_policy_batch(qmb200_ctx* ctx, const double* t, const double* x, double* u_out, int32_t* mode);

/* Measured rbd state rbd[n][55] -> MPC state x_out[n][30] = [A(q) v / m; base position; zyx; joints]. yaw_last[n] (may be
 * NULL): previous yaw per state; when given, x[9] = yaw_last + shortest_angular_distance(yaw_last, yaw). n need not equal the
 * context's batch. Host buffers / device pointers (on the context's stream) respectively. */
int qmb200_rbd_to_state_batch(qmb200_ctx* ctx, int32_t n, const double* rbd, const double* yaw_last, double* x_out);
int qmb200_rbd_to_state_batch_dev(qmb200_ctx* ctx, int32_t n, const double* rbd, const double* yaw_last, double* x_out);

/* Kernel timing (CUDA events on the context's stream): accumulated ms and launch counts per kernel since the last reset. */
int qmb200_set_profiling(qmb200_ctx* ctx, int32_t enable);
int qmb200_get_kernel_times(qmb200_ctx* ctx, double* total_ms, int64_t* launches, int32_t reset);
const char* qmb200_kernel_name(int32_t index);
void* qmb200_stream(qmb200_ctx* ctx);
int64_t qmb200_device_bytes(qmb200_ctx* ctx);

/* ---- whole-body controller (second half of the hot path)
 *   qmb200_wbc_create / destroy    <- HierarchicalWbc construction + loadTasksSetting in QMController::setupWbc
 *                                     (qm_controllers/src/QMController.cpp:273-277, qm_wbc/src/WbcBase.cpp:22-72,597-627)
 *   qmb200_load_wbc                <- WbcBase::loadTasksSetting + the dynamic_reconfigure defaults (qm_wbc/cfg/wbcWigeht.cfg:7-47)
 *   qmb200_actuator_batch[_dev]    <- QMController::updateControlLaw (QMController.cpp:178-191) + QMHWSim::writeSim
 *                                       (qm_gazebo/src/QMHWSim.cpp:98-114)
 *   qmb200_wbc_batch[_dev]         <- WbcBase::update(stateDesired, inputDesired, rbdStateMeasured, mode, period, time)
 *                                     (qm_wbc/include/qm_wbc/WbcBase.h:31-32, called at QMController.cpp:147), B independent solves:
 *                                     x_des[B][30] u_des[B][30] rbd[B][55] mode[B] period[B] time[B] -> cmd[B][54] = [accelerations(24);
 *                                     contact forces(12); joint torques(18)], status[B] (QMB200_WST_* bits)
 *   qmb200_wbc_reset               <- inputLast_ = 0 (WbcBase.cpp:41): the finite-difference joint acceleration state, per solve
 * A batch larger than one wave of the single-kernel solve (4 solves per SM) runs as a sequence of kernels (task builder, level 0,
 * level products / kernel bases, active-set iteration) over a per-solve workspace image the context owns in device memory
 * (72 KB per solve of the batch); smaller batches -- a controller solving one configuration per tick -- run the single kernel,
 * which has the lower latency. Same results bit for bit; QMB200_WBC_SPLIT=0 / 1 in the environment at create time forces one. */
typedef struct qmb200_wbc_ctx qmb200_wbc_ctx;
enum { QMB200_WST_OK = 0, QMB200_WST_QP_MAX_ITER = 1, QMB200_WST_DEGENERATE = 2, QMB200_WST_NAN = 4, QMB200_WST_BAD_MODE = 8 };

int qmb200_load_wbc(const char* task_info, const qmb200_model_desc* model, qmb200_wbc_desc* wbc);
int qmb200_wbc_create(const qmb200_model_desc* model, const qmb200_wbc_desc* wbc, int32_t batch, int32_t device,
                      qmb200_wbc_ctx** out);
int qmb200_wbc_destroy(qmb200_wbc_ctx* ctx);
int qmb200_wbc_reset(qmb200_wbc_ctx* ctx);
int qmb200_wbc_set_gains(qmb200_wbc_ctx* ctx, const qmb200_wbc_desc* wbc);   /* dynamic_reconfigure callback equivalent */
int qmb200_wbc_batch(qmb200_wbc_ctx* ctx, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                     const double* period, const double* time, double* cmd, int32_t* status);
int qmb200_wbc_batch_dev(qmb200_wbc_ctx* ctx, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                         const double* period, const double* time, double* cmd, int32_t* status);
/* One solve with what the reference's HoQp objects expose per level (qm_wbc/include/qm_wbc/HoQp.h:21-36: getSolutions,
 * getStackedZMatrix, getStackedSlackSolutions): diagnostic entry (host buffers, one solve, does not touch the context's inputLast_
 * state: u_last[30] is passed in). levels[QMB200_WBC_LEVELS_SIZE]: per level p (6 records of 685 doubles): [columns n_p of the stacked
 * null-space basis | x after the level (36) | stacked Z after the level (36 x 18 row major, n_p columns valid; Eigen FullPivLU basis as
 * HoQp.cpp:129 builds it)], then the number of levels of the stack, then the level-0 slack (56). */
#define QMB200_WBC_LEVELS_SIZE (6 * (37 + 36 * 18) + 1 + 56)
int qmb200_wbc_levels(qmb200_wbc_ctx* ctx, const double* x_des, const double* u_des, const double* rbd, int32_t mode, double period,
                      double time, const double* u_last, double* cmd, int32_t* status, double* levels);
int qmb200_wbc_sync(qmb200_wbc_ctx* ctx);
int qmb200_wbc_wait_stream(qmb200_wbc_ctx* ctx, void* stream);   /* see qmb200_wait_stream */
void* qmb200_wbc_stream(qmb200_wbc_ctx* ctx);
int qmb200_wbc_kernel_time(qmb200_wbc_ctx* ctx, double* total_ms, int64_t* launches, int32_t reset);

/* Joint-level control law + simulated actuator with transport delay, the stage behind the whole-body controller in the
 * reference's simulation loop:
 *   QMController::updateControlLaw (qm_controllers/src/QMController.cpp:178-191, inputs formed at :147-157) and
 *   QMHWSim::writeSim (qm_gazebo/src/QMHWSim.cpp:98-114; delay: qm_gazebo/config/default.yaml:2).
 * One call = one control tick of every problem of the WBC context (its batch, its stream). The context keeps the delay
 * buffers (QMB200_ACT_CAPACITY commands per problem, allocated on first use) and the command each joint handle holds.
 * time_ns[B]: simulation clock in integer nanoseconds (ros::Time arithmetic); period_ns: tick length (a tick with
 * time_ns == period_ns clears the buffer, as the reference does on simulation reset); obs_time[B]: controller observation time
 * (legs are commanded only after leg_enable_time); x_des, u_des [B][30]: evaluated policy; cmd [B][54]: WBC output;
 * q, v [B][18]: measured joint positions / velocities; tau [B][18]: torque applied to the joints; status[B]: 0 or 64 (overflow). */
void qmb200_actuator_defaults(qmb200_actuator_desc* desc);
int qmb200_actuator_batch(qmb200_wbc_ctx* ctx, const qmb200_actuator_desc* desc, const int64_t* time_ns, int64_t period_ns,
                          const double* obs_time, const double* x_des, const double* u_des, const double* cmd, const double* q,
                          const double* v, double* tau, int32_t* status);
int qmb200_actuator_batch_dev(qmb200_wbc_ctx* ctx, const qmb200_actuator_desc* desc, const int64_t* time_ns, int64_t period_ns,
                              const double* obs_time, const double* x_des, const double* u_des, const double* cmd, const double* q,
                              const double* v, double* tau, int32_t* status);
int qmb200_actuator_reset(qmb200_wbc_ctx* ctx);

/* Forward-dynamics step behind the actuator. In the reference Gazebo integrates the robot
 * (qm_gazebo/src/QMHWSim.cpp:98-114 only hands over the joint efforts); for batched, device-resident closed-loop studies this is
 * ONE explicit step of the articulated-body equations with the stance feet of `mode` held by bilateral point contacts -- a
 * labelled stand-in, not Gazebo's contact model:  M qdd + h = S' tau + Jc' f,  Jc qdd = -dJc v - (beta / dt) Jc v,
 * v+ = v + dt qdd, q+ = q + dt v+.  rbd, rbd_next [B][55] in the estimator's layout (so the step chains with qmb200_wbc_batch and
 * qmb200_rbd_to_state_batch; the end-effector pose entries are the forward kinematics of the new configuration), tau [B][18] from
 * qmb200_actuator_batch, contact_forces [B][12] (zero for swing feet), status [B] (QMB200_WST_NAN). One call per simulation tick on the
 * WBC context (its batch, its stream). */
int qmb200_forward_dynamics_batch(qmb200_wbc_ctx* ctx, const double* rbd, const double* tau, const int32_t* mode, double dt, double beta,
                                  double* rbd_next, double* contact_forces, int32_t* status);
int qmb200_forward_dynamics_batch_dev(qmb200_wbc_ctx* ctx, const double* rbd, const double* tau, const int32_t* mode, double dt, double beta,
                                      double* rbd_next, double* contact_forces, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif
