// MOCK of qm_controllers/include/qm_controllers/QMController.h (TEST INFRASTRUCTURE): the protected virtual hooks with the
// reference's exact signatures (QMController.h:50-54), the members they fill (QMController.h:62-80) and bodies of init /
// starting / update reduced to the calls on mpc_, mpcMrtInterface_ and wbc_ that the reference makes
// (QMController.cpp:39-97 init, :99-127 starting, :129-157 update, :310-335 setupMrt). ros-control, the estimator, the
// safety checker and the publishers are left out; the MPC thread is replaced by a direct call (same call, no thread).
#pragma once
#include <memory>
#include <string>
#include "../ocs2/ocs2_mock.h"
#include "../qm_interface/QMInterface.h"
#include "../qm_wbc/WbcBase.h"
#include "../ros/ros.h"

namespace qm {
using namespace ocs2;

class QMController {
 public:
  QMController() = default;
  virtual ~QMController() = default;

  // QMController::init (QMController.cpp:39-97): file names from the parameter server, then the hooks in this order
  bool init(ros::NodeHandle& controller_nh, const std::string& taskFile, const std::string& urdfFile, const std::string& referenceFile) {
    setupInterface(taskFile, urdfFile, referenceFile, false);
    setupMpc(controller_nh);
    setupMrt();
    setupWbc(controller_nh, taskFile);
    return true;
  }
  // QMController::starting (:99-127): first observation, initial target, wait for the initial policy
  void starting(scalar_t time, const vector_t& state, const TargetTrajectories& target) {
    currentObservation_.time = time;
    currentObservation_.state = state;
    currentObservation_.input.setZero(30);
    currentObservation_.mode = 15;
    mpcMrtInterface_->setCurrentObservation(currentObservation_);
    mpcMrtInterface_->getReferenceManager().setTargetTrajectories(target);
    while (!mpcMrtInterface_->initialPolicyReceived()) mpcMrtInterface_->advanceMpc();
    mpcRunning_ = true;
  }
  // the body of the MPC thread (:316-323), on the observation update() stored last (:116-120)
  void observe(scalar_t time, const vector_t& state) {
    currentObservation_.time = time; currentObservation_.state = state;
    mpcMrtInterface_->setCurrentObservation(currentObservation_);
  }
  void mpcThreadTick() { if (mpcRunning_) mpcMrtInterface_->advanceMpc(); }
  // QMController::update (:129-157) up to the torque
  vector_t update(scalar_t time, scalar_t period, const vector_t& state, const vector_t& measuredRbdState, vector_t* optimizedStateOut = nullptr,
                  vector_t* optimizedInputOut = nullptr, size_t* plannedModeOut = nullptr) {
    currentObservation_.time = time;
    currentObservation_.state = state;
    measuredRbdState_ = measuredRbdState;
    mpcMrtInterface_->setCurrentObservation(currentObservation_);
    mpcMrtInterface_->updatePolicy();
    vector_t optimizedState, optimizedInput;
    size_t plannedMode = 0;
    mpcMrtInterface_->evaluatePolicy(currentObservation_.time, currentObservation_.state, optimizedState, optimizedInput, plannedMode);
    currentObservation_.input = optimizedInput;
    vector_t x = wbc_->update(optimizedState, optimizedInput, measuredRbdState_, plannedMode, period, currentObservation_.time);
    if (optimizedStateOut) *optimizedStateOut = optimizedState;
    if (optimizedInputOut) *optimizedInputOut = optimizedInput;
    if (plannedModeOut) *plannedModeOut = plannedMode;
    return x;          // the reference takes x.tail(18) as the joint torques
  }

 protected:
  virtual void setupInterface(const std::string& taskFile, const std::string& urdfFile, const std::string& referenceFile, bool verbose) {
    qmInterface_ = std::make_shared<QMInterface>(taskFile, urdfFile, referenceFile);                       // QMController.cpp:337-346
    qmInterface_->setupOptimalControlProblem(taskFile, urdfFile, referenceFile, verbose);
    eeKinematicsPtr_ = std::make_shared<PinocchioEndEffectorKinematics>();
    armEeKinematicsPtr_ = std::make_shared<PinocchioEndEffectorKinematics>();
  }
  virtual void setupMpc(ros::NodeHandle& controller_nh) = 0;                                               // QMController.h:52
  virtual void setupMrt() {                                                                                // QMController.h:53, .cpp:310-313
    mpcMrtInterface_ = std::make_shared<MPC_MRT_Interface>(*mpc_);
    mpcMrtInterface_->initRollout(&qmInterface_->getRollout());
  }
  virtual void setupWbc(ros::NodeHandle& controller_nh, const std::string& taskFile) = 0;                  // QMController.h:54

  // Interface (QMController.h:62-64)
  std::shared_ptr<QMInterface> qmInterface_;
  std::shared_ptr<PinocchioEndEffectorKinematics> eeKinematicsPtr_;
  std::shared_ptr<PinocchioEndEffectorKinematics> armEeKinematicsPtr_;
  // State Estimation (:70-73)
  SystemObservation currentObservation_;
  vector_t measuredRbdState_;
  std::shared_ptr<CentroidalModelRbdConversions> rbdConversions_;
  // MPC & WBC (:78-80)
  std::shared_ptr<MPC_BASE> mpc_;
  std::shared_ptr<MPC_MRT_Interface> mpcMrtInterface_;
  std::shared_ptr<WbcBase> wbc_;
  bool mpcRunning_ = false;
};
}  // namespace qm
