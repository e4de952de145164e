// MOCK of qm_interface/include/qm_interface/QMInterface.h (TEST INFRASTRUCTURE): the accessors qm_controllers uses
// (QMInterface.h:37-54) and the registration of the eight OCP term names as setupOptimalControlProblem does it
// (qm_interface/src/QMInterface.cpp:99-129), with empty term objects.
#pragma once
#include <memory>
#include <string>
#include "../ocs2/ocs2_mock.h"

namespace qm {
using namespace ocs2;

class QMInterface {
 public:
  QMInterface(const std::string& taskFile, const std::string& urdfFile, const std::string& referenceFile)
      : taskFile_(taskFile), urdfFile_(urdfFile), referenceFile_(referenceFile), problemPtr_(new OptimalControlProblem),
        referenceManagerPtr_(new ReferenceManager), initializerPtr_(new Initializer), rolloutPtr_(new RolloutBase) {
    initialState_.setZero(30);
  }
  void setupOptimalControlProblem(const std::string&, const std::string&, const std::string&, bool) {
    struct C : StateInputCost {}; struct S : StateCost {}; struct E : StateInputConstraint {};
    problemPtr_->costPtr->add("baseTrackingCost", std::unique_ptr<StateInputCost>(new C));                     // QMInterface.cpp:99
    problemPtr_->stateSoftConstraintPtr->add("endEffector", std::unique_ptr<StateCost>(new S));               // :103
    problemPtr_->finalSoftConstraintPtr->add("finalEndEffector", std::unique_ptr<StateCost>(new S));          // :104
    problemPtr_->softConstraintPtr->add("armJointLimits", std::unique_ptr<StateInputCost>(new C));            // :108
    for (const char* foot : {"LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"}) {                                      // task.info contactNames3DoF
      const std::string f(foot);
      problemPtr_->softConstraintPtr->add(f + "_frictionCone", std::unique_ptr<StateInputCost>(new C));       // :120
      problemPtr_->equalityConstraintPtr->add(f + "_zeroForce", std::unique_ptr<StateInputConstraint>(new E));       // :123
      problemPtr_->equalityConstraintPtr->add(f + "_zeroVelocity", std::unique_ptr<StateInputConstraint>(new E));    // :126
      problemPtr_->equalityConstraintPtr->add(f + "_normalVelocity", std::unique_ptr<StateInputConstraint>(new E));  // :129
    }
  }
  OptimalControlProblem& mutableProblem() { return *problemPtr_; }          // test hook only (add a foreign term)
  const OptimalControlProblem& getOptimalControlProblem() const { return *problemPtr_; }
  const mpc::Settings& mpcSettings() const { return mpcSettings_; }
  const sqp::Settings& sqpSettings() { return sqpSettings_; }
  const vector_t& getInitialState() const { return initialState_; }
  const RolloutBase& getRollout() const { return *rolloutPtr_; }
  PinocchioInterface& getPinocchioInterface() { return pinocchio_; }
  const CentroidalModelInfo& getCentroidalModelInfo() const { return info_; }
  const Initializer& getInitializer() const { return *initializerPtr_; }
  std::shared_ptr<ReferenceManagerInterface> getReferenceManagerPtr() const { return referenceManagerPtr_; }
  // the file names the constructor was given (the real class keeps them only implicitly: QMController::init reads them from
  // the parameter server, QMController.cpp:39-49, and hands them to setupInterface)
  const std::string taskFile_, urdfFile_, referenceFile_;
  mpc::Settings mpcSettings_;
  sqp::Settings sqpSettings_;
 private:
  std::unique_ptr<OptimalControlProblem> problemPtr_;
  std::shared_ptr<ReferenceManager> referenceManagerPtr_;
  std::unique_ptr<Initializer> initializerPtr_;
  std::unique_ptr<RolloutBase> rolloutPtr_;
  PinocchioInterface pinocchio_;
  CentroidalModelInfo info_;
  vector_t initialState_;
};
}  // namespace qm
