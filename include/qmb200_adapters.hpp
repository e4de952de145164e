// C++ host-side mirror of the reference's plugin interfaces above the C-ABI (include/qmb200.h).
//
// The reference is a ros-control plugin whose two hot objects are created by virtual factory hooks:
//   QMController::setupMpc  (qm_controllers/include/qm_controllers/QMController.h:52, src/QMController.cpp:287-307)
//   QMController::setupWbc  (QMController.h:54, src/QMController.cpp:273-277)
// This header provides drop-in replacements for what those hooks build:
//   qmb200::HierarchicalWbcB200 : same update() signature and return value as qm::WbcBase::update
//                                 (qm_wbc/include/qm_wbc/WbcBase.h:31-32; 54 doubles, caller keeps .tail(18))
//   qmb200::SqpMpcB200          : advanceMpc()/evaluatePolicy() shaped like MPC_MRT_Interface as used at
//                                 QMController.cpp:116-120,134-143 (run one SQP cycle, then interpolate the policy)
// The classes here are written against plain std::vector<double> ("vector_t" in the reference) and need nothing but the C-ABI;
// include/qmb200_ocs2_adapters.hpp derives the reference-typed objects from them (B200SqpMpc : ocs2::MPC_BASE,
// B200HierarchicalWbc : qm::WbcBase), which is what a QMController subclass installs.
// Error behaviour follows the reference: construction problems throw std::runtime_error / std::invalid_argument
// (QMInterface.cpp:41-62); per-solve WBC failures are not thrown (the reference ignores qpOASES' return code,
// HoQp.cpp:143-146) but kept in lastStatus(); MPC failures throw std::runtime_error like the MPC thread's catch block expects
// (QMController.cpp:328-331).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "qmb200.h"

namespace qmb200 {

using vector_t = std::vector<double>;

inline void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + qmb200_last_error());
}

// Mirror of QMInterface (qm_interface/include/qm_interface/QMInterface.h:29-54): file ingestion + problem constants.
class InterfaceB200 {
 public:
  InterfaceB200(const std::string& taskFile, const std::string& urdfFile, const std::string& referenceFile) {
    if (qmb200_load_urdf(urdfFile.c_str(), &model_) != 0) throw std::invalid_argument(qmb200_last_error());
    initialState_.resize(QM_NX);
    if (qmb200_load_problem(taskFile.c_str(), referenceFile.c_str(), &model_, &problem_, &solver_, initialState_.data()) != 0)
      throw std::invalid_argument(qmb200_last_error());
    if (qmb200_load_wbc(taskFile.c_str(), &model_, &wbc_) != 0) throw std::invalid_argument(qmb200_last_error());
  }
  const qmb200_model_desc& model() const { return model_; }
  const qmb200_problem_desc& problem() const { return problem_; }
  qmb200_solver_desc& solverSettings() { return solver_; }
  qmb200_wbc_desc& wbcSettings() { return wbc_; }
  const vector_t& getInitialState() const { return initialState_; }   // QMInterface.h:45

 private:
  qmb200_model_desc model_{};
  qmb200_problem_desc problem_{};
  qmb200_solver_desc solver_{};
  qmb200_wbc_desc wbc_{};
  vector_t initialState_;
};

// Mirror of qm::HierarchicalWbc (qm_wbc/include/qm_wbc/HierarchicalWbc.h) for one robot (batch of 1) or a batch.
class HierarchicalWbcB200 {
 public:
  HierarchicalWbcB200(const InterfaceB200& interface, int batch = 1, int device = 0, bool mpcVariant = false) : batch_(batch) {
    qmb200_wbc_desc w = const_cast<InterfaceB200&>(interface).wbcSettings();
    w.mpc_variant = mpcVariant ? 1 : 0;
    check(qmb200_wbc_create(&interface.model(), &w, batch, device, &ctx_), "qmb200_wbc_create");
    status_.assign(batch, 0);
  }
  ~HierarchicalWbcB200() { qmb200_wbc_destroy(ctx_); }
  HierarchicalWbcB200(const HierarchicalWbcB200&) = delete;
  HierarchicalWbcB200& operator=(const HierarchicalWbcB200&) = delete;

  // vector_t WbcBase::update(stateDesired, inputDesired, rbdStateMeasured, mode, period, time)   (WbcBase.h:31-32)
  vector_t update(const vector_t& stateDesired, const vector_t& inputDesired, const vector_t& rbdStateMeasured, size_t mode,
                  double period, double time) {
    if (batch_ != 1) throw std::logic_error("update(): single-robot entry on a batched context");
    if (stateDesired.size() != 30 || inputDesired.size() != 30 || rbdStateMeasured.size() < 48)
      throw std::invalid_argument("HierarchicalWbcB200::update: wrong vector sizes");
    vector_t rbd(55, 0.0);
    for (size_t i = 0; i < rbdStateMeasured.size() && i < 55; ++i) rbd[i] = rbdStateMeasured[i];
    vector_t cmd(54);
    const int32_t m = static_cast<int32_t>(mode);
    check(qmb200_wbc_batch(ctx_, stateDesired.data(), inputDesired.data(), rbd.data(), &m, &period, &time, cmd.data(), status_.data()),
          "qmb200_wbc_batch");
    return cmd;
  }
  // batched form: row-major [B][..] buffers
  void updateBatch(const double* xDes, const double* uDes, const double* rbd, const int32_t* mode, const double* period,
                   const double* time, double* cmd) {
    check(qmb200_wbc_batch(ctx_, xDes, uDes, rbd, mode, period, time, cmd, status_.data()), "qmb200_wbc_batch");
  }
  void setGains(const qmb200_wbc_desc& w) { check(qmb200_wbc_set_gains(ctx_, &w), "qmb200_wbc_set_gains"); }   // dynamicCallback
  void reset() { check(qmb200_wbc_reset(ctx_), "qmb200_wbc_reset"); }
  const std::vector<int32_t>& lastStatus() const { return status_; }

 private:
  qmb200_wbc_ctx* ctx_ = nullptr;
  int batch_;
  std::vector<int32_t> status_;
};

// What SqpMpc + MPC_MRT_Interface give the controller: advance one MPC cycle, then evaluate the buffered policy.
struct ModeScheduleB200 {          // ocs2::ModeSchedule: eventTimes + modeSequence (size events + 1)
  std::vector<double> eventTimes;
  std::vector<int32_t> modeSequence;
};
struct TargetTrajectoriesB200 {    // ocs2::TargetTrajectories with the 37-dim states of QMController.cpp:107-113
  std::vector<double> timeTrajectory;
  std::vector<vector_t> stateTrajectory;
};

class SqpMpcB200 {
 public:
  SqpMpcB200(InterfaceB200& interface, int device = 0) : s_(interface.solverSettings()) {
    check(qmb200_create(&interface.model(), &interface.problem(), &s_, 1, device, &ctx_), "qmb200_create");
    t_.resize(s_.max_nodes); x_.resize(s_.max_nodes * 30); u_.resize(s_.max_nodes * 30);
  }
  ~SqpMpcB200() { qmb200_destroy(ctx_); }
  SqpMpcB200(const SqpMpcB200&) = delete;
  SqpMpcB200& operator=(const SqpMpcB200&) = delete;

  void reset() { check(qmb200_mpc_reset(ctx_), "qmb200_mpc_reset"); policyReceived_ = false; }

  // mpcMrtInterface_->setCurrentObservation(obs); advanceMpc();   (QMController.cpp:116-120, 316-333)
  void advanceMpc(double time, const vector_t& state, const ModeScheduleB200& schedule, const TargetTrajectoriesB200& target) {
    if (state.size() != 30) throw std::invalid_argument("advanceMpc: state must have 30 entries");
    const int E = s_.max_events, K = s_.max_targets;
    if ((int)schedule.eventTimes.size() > E) throw std::invalid_argument("advanceMpc: mode schedule exceeds max_events");
    if ((int)target.timeTrajectory.size() > K || target.timeTrajectory.empty()) throw std::invalid_argument("advanceMpc: bad target size");
    std::vector<double> ev(E, 1e30), tt(K), tx((size_t)K * QM_NTARGET);
    std::vector<int32_t> md(E + 1, 15);
    for (size_t i = 0; i < schedule.eventTimes.size(); ++i) ev[i] = schedule.eventTimes[i];
    for (size_t i = 0; i < schedule.modeSequence.size() && i < md.size(); ++i) md[i] = schedule.modeSequence[i];
    const int32_t nev = (int32_t)schedule.eventTimes.size();
    for (int k = 0; k < K; ++k) {     // fewer knots than capacity: repeat the last one (same interpolant)
      const size_t src = std::min<size_t>(k, target.timeTrajectory.size() - 1);
      tt[k] = target.timeTrajectory[src] + (k > (int)src ? 1e3 * (k - (int)src) : 0.0);
      for (int c = 0; c < QM_NTARGET; ++c) tx[(size_t)k * QM_NTARGET + c] = target.stateTrajectory[src][c];
    }
    int32_t status = 0;
    check(qmb200_mpc_cycle_batch(ctx_, &time, state.data(), ev.data(), md.data(), &nev, tt.data(), tx.data(), t_.data(), x_.data(),
                                 u_.data(), &n_, nullptr, info_, &status),
          "qmb200_mpc_cycle_batch");
    if (status & ~QMB200_ST_STEP_REJECTED) throw std::runtime_error("[SqpMpcB200] solver status " + std::to_string(status));
    policyReceived_ = true;
  }
  bool initialPolicyReceived() const { return policyReceived_; }

  // mpcMrtInterface_->evaluatePolicy(t, x, optimizedState, optimizedInput, plannedMode)   (QMController.cpp:140-143)
  void evaluatePolicy(double time, vector_t& optimizedState, vector_t& optimizedInput, size_t& plannedMode) {
    optimizedState.resize(30); optimizedInput.resize(30);
    int32_t mode = 15;
    check(qmb200_evaluate_policy_batch(ctx_, &time, optimizedState.data(), optimizedInput.data(), &mode), "qmb200_evaluate_policy_batch");
    plannedMode = (size_t)mode;
  }
  // PrimalSolution{timeTrajectory_, stateTrajectory_, inputTrajectory_}
  int numNodes() const { return n_; }
  const std::vector<double>& timeTrajectory() const { return t_; }
  const std::vector<double>& stateTrajectory() const { return x_; }
  const std::vector<double>& inputTrajectory() const { return u_; }
  double stepSize() const { return info_[0]; }
  const double* info() const { return info_; }          // QMB200_INFO_SIZE entries, see include/qmb200.h

 private:
  qmb200_solver_desc s_;
  qmb200_ctx* ctx_ = nullptr;
  std::vector<double> t_, x_, u_;
  double info_[QMB200_INFO_SIZE] = {0};
  int32_t n_ = 0;
  bool policyReceived_ = false;
};

}  // namespace qmb200
