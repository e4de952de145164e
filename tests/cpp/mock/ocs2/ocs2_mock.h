// MOCK of the OCS2 types the reference's hot-path seams are written against (TEST INFRASTRUCTURE).
// OCS2 is not vendored in the reference checkout and not installed in this image; these declarations restate, from the published
// ocs2 sources, exactly the part of each class that qm_controllers touches (qm_controllers/src/QMController.cpp:99-157,
// 273-335) so that include/qmb200_ocs2_adapters.hpp can be compiled and driven without OCS2. With QMB200_WITH_OCS2 the adapters
// include the real headers instead; they only use members declared here.
//   ocs2_core/Types.h                      scalar_t, vector_t (Eigen::VectorXd upstream: size(), data(), operator[], (n) ctor)
//   ocs2_core/reference/ModeSchedule.h     ModeSchedule{eventTimes, modeSequence}
//   ocs2_core/reference/TargetTrajectories.h
//   ocs2_core/misc/Collection.h            add / get / erase / empty by term name
//   ocs2_oc/oc_problem/OptimalControlProblem.h
//   ocs2_oc/oc_data/PrimalSolution.h, ocs2_core/control/FeedforwardController.h
//   ocs2_oc/synchronized_module/ReferenceManagerInterface.h, SolverSynchronizedModule.h
//   ocs2_oc/oc_solver/SolverBase.h         run() = preRun + runImpl + postRun, getPrimalSolution
//   ocs2_mpc/MPC_BASE.h, MPC_Settings.h    run(t, x) -> calculateController(t0, x0, tf)
//   ocs2_mpc/MPC_MRT_Interface.h           setCurrentObservation / advanceMpc / updatePolicy / evaluatePolicy
//   ocs2_sqp/SqpSettings.h                 the fields task.info:76-93 fills
#pragma once
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace ocs2 {

using scalar_t = double;
using scalar_array_t = std::vector<scalar_t>;
using size_array_t = std::vector<size_t>;

// the subset of Eigen::VectorXd the adapters use
class vector_t {
 public:
  vector_t() = default;
  explicit vector_t(long n) : v_(static_cast<size_t>(n), 0.0) {}
  long size() const { return static_cast<long>(v_.size()); }
  scalar_t* data() { return v_.data(); }
  const scalar_t* data() const { return v_.data(); }
  scalar_t& operator[](long i) { return v_[static_cast<size_t>(i)]; }
  const scalar_t& operator[](long i) const { return v_[static_cast<size_t>(i)]; }
  vector_t tail(long n) const { vector_t o(n); for (long i = 0; i < n; ++i) o[i] = v_[v_.size() - static_cast<size_t>(n - i)]; return o; }
  vector_t& setZero(long n) { v_.assign(static_cast<size_t>(n), 0.0); return *this; }
 private:
  std::vector<scalar_t> v_;
};
using vector_array_t = std::vector<vector_t>;

struct ModeSchedule {
  scalar_array_t eventTimes;
  size_array_t modeSequence;      // eventTimes.size() + 1 entries
};

struct TargetTrajectories {
  TargetTrajectories() = default;
  TargetTrajectories(scalar_array_t t, vector_array_t x, vector_array_t u)
      : timeTrajectory(std::move(t)), stateTrajectory(std::move(x)), inputTrajectory(std::move(u)) {}
  scalar_array_t timeTrajectory;
  vector_array_t stateTrajectory;
  vector_array_t inputTrajectory;
};

struct SystemObservation {
  size_t mode = 0;
  scalar_t time = 0.0;
  vector_t state, input;
};

// ---- named term collections of the optimal control problem
struct StateInputCost { virtual ~StateInputCost() = default; };
struct StateCost { virtual ~StateCost() = default; };
struct StateInputConstraint { virtual ~StateInputConstraint() = default; };

template <typename T>
class Collection {
 public:
  void add(std::string name, std::unique_ptr<T> term) {
    if (!names_.emplace(name, terms_.size()).second) throw std::runtime_error("[Collection::add] Term with name \"" + name + "\" already exists");
    terms_.push_back(std::shared_ptr<T>(std::move(term)));
  }
  template <typename Derived = T>
  Derived& get(const std::string& name) {
    auto it = names_.find(name);
    if (it == names_.end()) throw std::out_of_range("[Collection::get] Term with name \"" + name + "\" not found");
    return dynamic_cast<Derived&>(*terms_[it->second]);
  }
  bool erase(const std::string& name) {
    auto it = names_.find(name);
    if (it == names_.end()) return false;
    terms_[it->second].reset();
    names_.erase(it);
    return true;
  }
  bool empty() const { return names_.empty(); }
 private:
  std::vector<std::shared_ptr<T>> terms_;       // shared: the mock collection is copyable like the real one (which clones)
  std::unordered_map<std::string, size_t> names_;
};
using StateInputCostCollection = Collection<StateInputCost>;
using StateCostCollection = Collection<StateCost>;
using StateInputConstraintCollection = Collection<StateInputConstraint>;

struct OptimalControlProblem {
  OptimalControlProblem()
      : costPtr(new StateInputCostCollection), stateCostPtr(new StateCostCollection), finalCostPtr(new StateCostCollection),
        softConstraintPtr(new StateInputCostCollection), stateSoftConstraintPtr(new StateCostCollection),
        finalSoftConstraintPtr(new StateCostCollection), equalityConstraintPtr(new StateInputConstraintCollection) {}
  OptimalControlProblem(const OptimalControlProblem& o)
      : costPtr(new StateInputCostCollection(*o.costPtr)), stateCostPtr(new StateCostCollection(*o.stateCostPtr)),
        finalCostPtr(new StateCostCollection(*o.finalCostPtr)), softConstraintPtr(new StateInputCostCollection(*o.softConstraintPtr)),
        stateSoftConstraintPtr(new StateCostCollection(*o.stateSoftConstraintPtr)),
        finalSoftConstraintPtr(new StateCostCollection(*o.finalSoftConstraintPtr)),
        equalityConstraintPtr(new StateInputConstraintCollection(*o.equalityConstraintPtr)) {}
  std::unique_ptr<StateInputCostCollection> costPtr;
  std::unique_ptr<StateCostCollection> stateCostPtr, finalCostPtr;
  std::unique_ptr<StateInputCostCollection> softConstraintPtr;
  std::unique_ptr<StateCostCollection> stateSoftConstraintPtr, finalSoftConstraintPtr;
  std::unique_ptr<StateInputConstraintCollection> equalityConstraintPtr;
};

struct Initializer { virtual ~Initializer() = default; };
struct RolloutBase { virtual ~RolloutBase() = default; };

// ---- controllers and primal solution
class ControllerBase {
 public:
  virtual ~ControllerBase() = default;
  virtual vector_t computeInput(scalar_t t, const vector_t& x) = 0;
  virtual ControllerBase* clone() const = 0;
};
class FeedforwardController final : public ControllerBase {
 public:
  FeedforwardController(scalar_array_t t, vector_array_t u) : timeStamp_(std::move(t)), uffArray_(std::move(u)) {}
  vector_t computeInput(scalar_t t, const vector_t&) override;      // LinearInterpolation::interpolate(t, timeStamp_, uffArray_)
  FeedforwardController* clone() const override { return new FeedforwardController(*this); }
  scalar_array_t timeStamp_;
  vector_array_t uffArray_;
};
struct PrimalSolution {
  scalar_array_t timeTrajectory_;
  vector_array_t stateTrajectory_, inputTrajectory_;
  ModeSchedule modeSchedule_;
  std::unique_ptr<ControllerBase> controllerPtr_;
};

inline void linear_segment(scalar_t t, const scalar_array_t& ts, size_t* idx, scalar_t* alpha) {    // LinearInterpolation::timeSegment
  const size_t n = ts.size();
  if (n <= 1) { *idx = 0; *alpha = 1.0; return; }
  size_t part = 0;
  while (part < n && ts[part] < t) ++part;
  const long i = (part == 0 && t == ts[0]) ? 0 : static_cast<long>(part) - 1;
  const long last = static_cast<long>(n) - 1;
  if (i >= 0) {
    if (i < last) { *idx = static_cast<size_t>(i); *alpha = (ts[i + 1] - t) / (ts[i + 1] - ts[i]); }
    else { *idx = static_cast<size_t>(last - 1 > 0 ? last - 1 : 0); *alpha = 0.0; }
  } else { *idx = 0; *alpha = 1.0; }
}
inline vector_t linear_interpolate(scalar_t t, const scalar_array_t& ts, const vector_array_t& data) {
  size_t i; scalar_t a;
  linear_segment(t, ts, &i, &a);
  const size_t j = (ts.size() > 1) ? i + 1 : i;
  vector_t out(data[i].size());
  for (long c = 0; c < out.size(); ++c) out[c] = a * data[i][c] + (1.0 - a) * data[j][c];
  return out;
}
inline vector_t FeedforwardController::computeInput(scalar_t t, const vector_t&) { return linear_interpolate(t, timeStamp_, uffArray_); }

// ---- reference manager and solver
class ReferenceManagerInterface {
 public:
  virtual ~ReferenceManagerInterface() = default;
  virtual void preSolverRun(scalar_t initTime, scalar_t finalTime, const vector_t& initState) = 0;
  virtual const ModeSchedule& getModeSchedule() const = 0;
  virtual void setModeSchedule(const ModeSchedule&) = 0;
  virtual const TargetTrajectories& getTargetTrajectories() const = 0;
  virtual void setTargetTrajectories(const TargetTrajectories&) = 0;
};
class ReferenceManager : public ReferenceManagerInterface {
 public:
  void preSolverRun(scalar_t, scalar_t, const vector_t&) override {}
  const ModeSchedule& getModeSchedule() const override { return modeSchedule_; }
  void setModeSchedule(const ModeSchedule& m) override { modeSchedule_ = m; }
  const TargetTrajectories& getTargetTrajectories() const override { return target_; }
  void setTargetTrajectories(const TargetTrajectories& t) override { target_ = t; }
 private:
  ModeSchedule modeSchedule_;
  TargetTrajectories target_;
};
class SolverSynchronizedModule {
 public:
  virtual ~SolverSynchronizedModule() = default;
  virtual void preSolverRun(scalar_t initTime, scalar_t finalTime, const vector_t& currentState, const ReferenceManagerInterface& referenceManager) = 0;
  virtual void postSolverRun(const PrimalSolution& primalSolution) = 0;
};
struct PerformanceIndex { scalar_t merit = 0, cost = 0, dynamicsViolationSSE = 0, equalityConstraintsSSE = 0; };

class SolverBase {
 public:
  virtual ~SolverBase() = default;
  virtual void reset() = 0;
  void run(scalar_t initTime, const vector_t& initState, scalar_t finalTime) {
    if (!referenceManagerPtr_) throw std::runtime_error("[SolverBase] ReferenceManager is not set");
    referenceManagerPtr_->preSolverRun(initTime, finalTime, initState);
    for (auto& m : synchronizedModules_) m->preSolverRun(initTime, finalTime, initState, *referenceManagerPtr_);
    runImpl(initTime, initState, finalTime);
    if (!synchronizedModules_.empty()) {
      PrimalSolution sol;
      getPrimalSolution(finalTime, &sol);
      for (auto& m : synchronizedModules_) m->postSolverRun(sol);
    }
  }
  void setReferenceManager(std::shared_ptr<ReferenceManagerInterface> p) { referenceManagerPtr_ = std::move(p); }
  const ReferenceManagerInterface& getReferenceManager() const { return *referenceManagerPtr_; }
  ReferenceManagerInterface& getReferenceManager() { return *referenceManagerPtr_; }
  void addSynchronizedModule(std::shared_ptr<SolverSynchronizedModule> m) { synchronizedModules_.push_back(std::move(m)); }
  virtual void getPrimalSolution(scalar_t finalTime, PrimalSolution* primalSolutionPtr) const = 0;
  virtual const PerformanceIndex& getPerformanceIndeces() const = 0;
  virtual size_t getNumIterations() const = 0;
  virtual scalar_t getFinalTime() const = 0;
  virtual const OptimalControlProblem& getOptimalControlProblem() const = 0;
 private:
  virtual void runImpl(scalar_t initTime, const vector_t& initState, scalar_t finalTime) = 0;
  std::shared_ptr<ReferenceManagerInterface> referenceManagerPtr_;
  std::vector<std::shared_ptr<SolverSynchronizedModule>> synchronizedModules_;
};

namespace mpc {
struct Settings {          // task.info:139-149
  scalar_t timeHorizon_ = 1.0;
  scalar_t solutionTimeWindow_ = -1;
  bool coldStart_ = false;
  bool debugPrint_ = false;
  scalar_t mpcDesiredFrequency_ = 100;
  scalar_t mrtDesiredFrequency_ = 400;
};
}  // namespace mpc
namespace sqp {
struct Settings {          // task.info:76-93 + defaults of ocs2_sqp/SqpSettings.h
  size_t sqpIteration = 1;
  scalar_t deltaTol = 1e-6, costTol = 1e-4;
  scalar_t alpha_decay = 0.5, alpha_min = 1e-4, gamma_c = 1e-6, g_max = 1e6, g_min = 1e-6, armijoFactor = 1e-4;
  scalar_t dt = 0.01;
  bool projectStateInputEqualityConstraints = true;
  bool useFeedbackPolicy = true;
  size_t nThreads = 4;
  int threadPriority = 50;
};
}  // namespace sqp

class MPC_BASE {
 public:
  explicit MPC_BASE(mpc::Settings s) : mpcSettings_(std::move(s)) {}
  virtual ~MPC_BASE() = default;
  virtual void reset() { getSolverPtr()->reset(); }
  virtual bool run(scalar_t currentTime, const vector_t& currentState) {
    calculateController(currentTime, currentState, currentTime + mpcSettings_.timeHorizon_);
    return true;
  }
  virtual SolverBase* getSolverPtr() = 0;
  virtual const SolverBase* getSolverPtr() const = 0;
  scalar_t getTimeHorizon() const { return mpcSettings_.timeHorizon_; }
  const mpc::Settings& settings() const { return mpcSettings_; }
 protected:
  virtual void calculateController(scalar_t initTime, const vector_t& initState, scalar_t finalTime) = 0;
 private:
  mpc::Settings mpcSettings_;
};

// The part of MPC_MRT_Interface QMController drives (QMController.cpp:116-143, 316-323): run the MPC on the last observation,
// buffer its primal solution, evaluate the buffered policy.
class MPC_MRT_Interface {
 public:
  explicit MPC_MRT_Interface(MPC_BASE& mpc) : mpc_(mpc) {}
  void initRollout(const RolloutBase*) {}
  void setCurrentObservation(const SystemObservation& o) { observation_ = o; }
  ReferenceManagerInterface& getReferenceManager() { return mpc_.getSolverPtr()->getReferenceManager(); }
  void advanceMpc() {
    mpc_.run(observation_.time, observation_.state);
    auto sol = std::make_unique<PrimalSolution>();
    mpc_.getSolverPtr()->getPrimalSolution(observation_.time + mpc_.getTimeHorizon(), sol.get());
    buffer_ = std::move(sol);
  }
  bool initialPolicyReceived() const { return buffer_ != nullptr || active_ != nullptr; }
  bool updatePolicy() { if (!buffer_) return false; active_ = std::move(buffer_); return true; }
  void evaluatePolicy(scalar_t t, const vector_t& x, vector_t& xOut, vector_t& uOut, size_t& mode) {
    if (!active_ && !updatePolicy()) throw std::runtime_error("[MPC_MRT_Interface] no policy");
    uOut = active_->controllerPtr_->computeInput(t, x);
    xOut = linear_interpolate(t, active_->timeTrajectory_, active_->stateTrajectory_);
    size_t i = 0;
    const auto& ev = active_->modeSchedule_.eventTimes;
    while (i < ev.size() && ev[i] < t) ++i;
    mode = active_->modeSchedule_.modeSequence[i];
  }
 private:
  MPC_BASE& mpc_;
  SystemObservation observation_;
  std::unique_ptr<PrimalSolution> buffer_, active_;
};

// types WbcBase / QMInterface name in their signatures (opaque here)
struct PinocchioInterface {};
struct CentroidalModelInfo { size_t stateDim = 30, inputDim = 30, generalizedCoordinatesNum = 24, actuatedDofNum = 18, numThreeDofContacts = 4; };
struct PinocchioEndEffectorKinematics {};
struct CentroidalModelRbdConversions { CentroidalModelRbdConversions(PinocchioInterface&, const CentroidalModelInfo&) {} };

}  // namespace ocs2
