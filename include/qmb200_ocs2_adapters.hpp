// Reference-typed adapters: the B200 hot path behind the exact types qm_controllers holds.
//
//   qm_controllers/include/qm_controllers/QMController.h:78   std::shared_ptr<MPC_BASE> mpc_;
//   qm_controllers/include/qm_controllers/QMController.h:80   std::shared_ptr<WbcBase> wbc_;
//
//   qmb200::B200SqpMpc          : ocs2::MPC_BASE     replaces the ocs2::SqpMpc built at QMController.cpp:288-289
//   qmb200::B200SqpSolver       : ocs2::SolverBase   replaces ocs2::SqpSolver: runImpl(t0, x0, tf) = one qmb200_mpc_cycle_batch
//                                                    (B = 1), getPrimalSolution() = the cycle's trajectories + FeedforwardController
//   qmb200::B200HierarchicalWbc : qm::WbcBase        replaces qm::HierarchicalWbc built at QMController.cpp:274-276
//                                                    (update() and loadTasksSetting() of qm_wbc/include/qm_wbc/WbcBase.h:31-34)
//
// Because mpc_ is an MPC_BASE, the reference's own MPC_MRT_Interface (QMController.cpp:311), its MPC thread (:316-333),
// starting() (:99-127) and update() (:129-157) run unchanged on top of these objects: a controller only overrides the two
// factory hooks setupMpc(ros::NodeHandle&) and setupWbc(ros::NodeHandle&, const std::string&) (QMController.h:52,54); see
// INTEGRATION.md and tests/cpp/controller_b200.cpp.
//
// Device code cannot call the virtual getValue / getQuadraticApproximation of user-defined terms, so the optimal control
// problem is fixed to the terms QMInterface::setupOptimalControlProblem registers (qm_interface/src/QMInterface.cpp:99-129);
// B200SqpMpc inspects the OptimalControlProblem it is given by those names and throws std::runtime_error on a missing or a
// foreign term, the term parameters themselves are read from the same task.info / reference.info (qmb200::InterfaceB200).
//
// Build: against the real headers in a catkin workspace (-DQMB200_WITH_OCS2), or against the mock declarations under
// tests/cpp/mock (this image has neither OCS2 nor Eigen nor ROS); only members both declare are used.
#pragma once
#if defined(QMB200_WITH_OCS2)
#include <ocs2_core/control/FeedforwardController.h>
#include <ocs2_mpc/MPC_BASE.h>
#include <ocs2_oc/oc_solver/SolverBase.h>
#include <ocs2_sqp/SqpSettings.h>
#include <qm_wbc/WbcBase.h>
#else
#include <ocs2/ocs2_mock.h>
#include <qm_wbc/WbcBase.h>
#endif
#include <algorithm>
#include <cmath>
#include <memory>
#include <string>
#include <vector>
#include "qmb200_adapters.hpp"

namespace qmb200 {

inline vector_t toStd(const ocs2::vector_t& v) { return vector_t(v.data(), v.data() + v.size()); }
inline ocs2::vector_t fromStd(const double* p, size_t n) {
  ocs2::vector_t v(static_cast<long>(n));
  std::copy(p, p + n, v.data());
  return v;
}

// The eight term names of QMInterface::setupOptimalControlProblem (QMInterface.cpp:99-129), per collection. Throws on a missing
// term and on any term outside the list (checked on a copy: every known name is erased, the collections must end up empty).
inline void checkOptimalControlProblem(const ocs2::OptimalControlProblem& ocp,
                                       const std::vector<std::string>& contactNames3DoF = {"LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"}) {
  ocs2::OptimalControlProblem p(ocp);
  auto need = [](bool erased, const std::string& name) {
    if (!erased) throw std::runtime_error("[qmb200] optimal control problem has no term \"" + name + "\" (QMInterface.cpp:99-129)");
  };
  need(p.costPtr->erase("baseTrackingCost"), "baseTrackingCost");
  need(p.stateSoftConstraintPtr->erase("endEffector"), "endEffector");
  need(p.finalSoftConstraintPtr->erase("finalEndEffector"), "finalEndEffector");
  need(p.softConstraintPtr->erase("armJointLimits"), "armJointLimits");
  for (const auto& f : contactNames3DoF) {
    need(p.softConstraintPtr->erase(f + "_frictionCone"), f + "_frictionCone");
    need(p.equalityConstraintPtr->erase(f + "_zeroForce"), f + "_zeroForce");
    need(p.equalityConstraintPtr->erase(f + "_zeroVelocity"), f + "_zeroVelocity");
    need(p.equalityConstraintPtr->erase(f + "_normalVelocity"), f + "_normalVelocity");
  }
  const bool clean = p.costPtr->empty() && p.stateCostPtr->empty() && p.finalCostPtr->empty() && p.softConstraintPtr->empty() &&
                     p.stateSoftConstraintPtr->empty() && p.finalSoftConstraintPtr->empty() && p.equalityConstraintPtr->empty();
  if (!clean)
    throw std::runtime_error("[qmb200] the optimal control problem holds a term outside the eight names of "
                             "QMInterface::setupOptimalControlProblem: the device transcription cannot evaluate it");
}

class B200SqpSolver final : public ocs2::SolverBase {
 public:
  B200SqpSolver(const ocs2::mpc::Settings& mpcSettings, const ocs2::sqp::Settings& sqp, const ocs2::OptimalControlProblem& ocp,
                InterfaceB200& files, int device, int maxTargets)
      : ocp_(ocp) {
    checkOptimalControlProblem(ocp_);
    if (!sqp.projectStateInputEqualityConstraints)
      throw std::invalid_argument("[qmb200] projectStateInputEqualityConstraints must be true (task.info:86)");
    if (sqp.useFeedbackPolicy)
      throw std::invalid_argument("[qmb200] useFeedbackPolicy true: the feedback gains are served by qmb200_feedback_gains, not by this adapter (task.info:90 sets false)");
    // the solver settings the reference parsed rule (sqp::Settings, mpc::Settings); capacities follow from them
    qmb200_solver_desc& s = files.solverSettings();
    s.dt = sqp.dt; s.horizon = mpcSettings.timeHorizon_;
    s.sqp_iterations = static_cast<int32_t>(sqp.sqpIteration);
    s.delta_tol = sqp.deltaTol; s.cost_tol = sqp.costTol; s.g_max = sqp.g_max; s.g_min = sqp.g_min;
    s.alpha_decay = sqp.alpha_decay; s.alpha_min = sqp.alpha_min; s.gamma_c = sqp.gamma_c; s.armijo_factor = sqp.armijoFactor;
    s.max_nodes = static_cast<int32_t>(s.horizon / s.dt + 0.5) + 1 + 2 * 16;
    s.max_events = 64;
    s.max_targets = maxTargets;
    horizon_ = s.horizon;
    core_.reset(new SqpMpcB200(files, device));
    capacityEvents_ = s.max_events;
  }

  void reset() override { core_->reset(); }

  void getPrimalSolution(ocs2::scalar_t /*finalTime*/, ocs2::PrimalSolution* sol) const override {
    const int n = core_->numNodes();
    sol->timeTrajectory_.assign(core_->timeTrajectory().begin(), core_->timeTrajectory().begin() + n);
    sol->stateTrajectory_.clear(); sol->inputTrajectory_.clear();
    for (int k = 0; k < n; ++k) {
      sol->stateTrajectory_.push_back(fromStd(core_->stateTrajectory().data() + 30 * k, 30));
      sol->inputTrajectory_.push_back(fromStd(core_->inputTrajectory().data() + 30 * k, 30));
    }
    sol->modeSchedule_ = lastSchedule_;
    sol->controllerPtr_.reset(new ocs2::FeedforwardController(sol->timeTrajectory_, sol->inputTrajectory_));   // task.info:90
  }
  const ocs2::PerformanceIndex& getPerformanceIndeces() const override { return performance_; }
  size_t getNumIterations() const override { return iterations_; }
  ocs2::scalar_t getFinalTime() const override { return finalTime_; }
  const ocs2::OptimalControlProblem& getOptimalControlProblem() const override { return ocp_; }
  const SqpMpcB200& core() const { return *core_; }

 private:
  // [upstream] SqpSolver::runImpl: mode schedule and target trajectories come from the reference manager, which
  // SolverBase::run has just updated (preSolverRun of the manager and of the synchronized modules, e.g. the GaitReceiver)
  void runImpl(ocs2::scalar_t initTime, const ocs2::vector_t& initState, ocs2::scalar_t finalTime) override {
    if (std::abs((finalTime - initTime) - horizon_) > 1e-9)
      throw std::invalid_argument("[qmb200] horizon differs from mpc.timeHorizon the context was created with");
    const ocs2::ModeSchedule& ms = getReferenceManager().getModeSchedule();
    const ocs2::TargetTrajectories& tg = getReferenceManager().getTargetTrajectories();
    // window of the schedule the cycle can see: from the last event a swing phase of this horizon can have started at
    // (two horizons back) to past the end of the horizon; an open-ended last mode is closed with a far-away event
    ModeScheduleB200 sched;
    const size_t ne = ms.eventTimes.size();
    size_t first = 0;
    while (first < ne && ms.eventTimes[first] < initTime - 2.0 * horizon_) ++first;
    sched.modeSequence.push_back(static_cast<int32_t>(ms.modeSequence[first]));
    for (size_t i = first; i < ne && static_cast<int>(sched.eventTimes.size()) < capacityEvents_ - 1; ++i) {
      sched.eventTimes.push_back(ms.eventTimes[i]);
      sched.modeSequence.push_back(static_cast<int32_t>(ms.modeSequence[i + 1]));
      if (ms.eventTimes[i] > finalTime + horizon_) break;
    }
    if (sched.eventTimes.empty() || sched.eventTimes.back() <= finalTime) {
      if (!sched.eventTimes.empty() && sched.eventTimes.size() + first < ne)
        throw std::runtime_error("[qmb200] mode schedule has more events inside the horizon than the context's capacity");
      sched.eventTimes.push_back(1e30);
      sched.modeSequence.push_back(sched.modeSequence.back());
    }
    TargetTrajectoriesB200 target;
    target.timeTrajectory = tg.timeTrajectory;
    for (const auto& x : tg.stateTrajectory) {
      if (x.size() != QM_NTARGET) throw std::invalid_argument("[qmb200] target states must have 37 entries (QMController.cpp:107-113)");
      target.stateTrajectory.push_back(toStd(x));
    }
    core_->advanceMpc(initTime, toStd(initState), sched, target);
    lastSchedule_.eventTimes.assign(sched.eventTimes.begin(), sched.eventTimes.end());
    lastSchedule_.modeSequence.assign(sched.modeSequence.begin(), sched.modeSequence.end());
    const double* info = core_->info();
    performance_.merit = info[8]; performance_.cost = info[8]; performance_.dynamicsViolationSSE = info[9]; performance_.equalityConstraintsSSE = info[10];
    iterations_ = static_cast<size_t>(info[13]);
    finalTime_ = finalTime;
  }

  ocs2::OptimalControlProblem ocp_;
  std::unique_ptr<SqpMpcB200> core_;
  ocs2::ModeSchedule lastSchedule_;
  ocs2::PerformanceIndex performance_;
  size_t iterations_ = 0;
  ocs2::scalar_t finalTime_ = 0.0, horizon_ = 0.0;
  int capacityEvents_ = 0;
};

// [upstream] ocs2::SqpMpc, same constructor arguments plus the ingested files and the device.
class B200SqpMpc final : public ocs2::MPC_BASE {
 public:
  B200SqpMpc(ocs2::mpc::Settings mpcSettings, ocs2::sqp::Settings settings, const ocs2::OptimalControlProblem& optimalControlProblem,
             const ocs2::Initializer& /*initializer: QMInitializer is part of the device cycle (init_guess_component)*/,
             InterfaceB200& files, int device = 0, int maxTargets = 8)
      : ocs2::MPC_BASE(mpcSettings), solverPtr_(new B200SqpSolver(mpcSettings, settings, optimalControlProblem, files, device, maxTargets)) {}
  ~B200SqpMpc() override = default;
  B200SqpSolver* getSolverPtr() override { return solverPtr_.get(); }
  const B200SqpSolver* getSolverPtr() const override { return solverPtr_.get(); }

 protected:
  void calculateController(ocs2::scalar_t initTime, const ocs2::vector_t& initState, ocs2::scalar_t finalTime) override {
    if (settings().coldStart_) solverPtr_->reset();
    solverPtr_->run(initTime, initState, finalTime);
  }

 private:
  std::unique_ptr<B200SqpSolver> solverPtr_;
};

// qm::HierarchicalWbc (or HierarchicalMpcWbc with mpcVariant) on the device, behind qm::WbcBase.
class B200HierarchicalWbc final : public qm::WbcBase {
 public:
  B200HierarchicalWbc(const ocs2::PinocchioInterface& pinocchioInterface, ocs2::CentroidalModelInfo info,
                      const ocs2::PinocchioEndEffectorKinematics& eeKinematics, const ocs2::PinocchioEndEffectorKinematics& armEeKinematics,
                      ros::NodeHandle& controller_nh, InterfaceB200& files, int device = 0, bool mpcVariant = false)
      : qm::WbcBase(pinocchioInterface, std::move(info), eeKinematics, armEeKinematics, controller_nh), files_(files),
        impl_(new HierarchicalWbcB200(files, 1, device, mpcVariant)), mpcVariant_(mpcVariant) {}

  ocs2::vector_t update(const ocs2::vector_t& stateDesired, const ocs2::vector_t& inputDesired, const ocs2::vector_t& rbdStateMeasured,
                        size_t mode, ocs2::scalar_t period, ocs2::scalar_t time) override {
    const vector_t cmd = impl_->update(toStd(stateDesired), toStd(inputDesired), toStd(rbdStateMeasured), mode, period, time);
    return fromStd(cmd.data(), cmd.size());          // [accelerations(24); contact forces(12); torques(18)], WbcBase.cpp:592
  }
  // WbcBase::loadTasksSetting (WbcBase.cpp:597-627): torque limits from the model, friction coefficient from the task file
  void loadTasksSetting(const std::string& taskFile, bool /*verbose*/) override {
    qmb200_wbc_desc w = files_.wbcSettings();
    if (qmb200_load_wbc(taskFile.c_str(), &files_.model(), &w) != 0) throw std::invalid_argument(qmb200_last_error());
    w.mpc_variant = mpcVariant_ ? 1 : 0;
    impl_->setGains(w);
  }
  int32_t lastStatus() const { return impl_->lastStatus()[0]; }

 private:
  InterfaceB200& files_;
  std::unique_ptr<HierarchicalWbcB200> impl_;
  bool mpcVariant_;
};

}  // namespace qmb200
