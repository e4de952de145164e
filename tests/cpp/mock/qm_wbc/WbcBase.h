// MOCK of qm_wbc/include/qm_wbc/WbcBase.h (TEST INFRASTRUCTURE): the public interface of qm::WbcBase exactly as the reference
// declares it (WbcBase.h:28-34: constructor, virtual update, virtual loadTasksSetting), without the Pinocchio-based body.
#pragma once
#include <string>
#include "../ocs2/ocs2_mock.h"
#include "../ros/ros.h"

namespace qm {
using namespace ocs2;

class WbcBase {
 public:
  WbcBase(const PinocchioInterface& pinocchioInterface, CentroidalModelInfo info, const PinocchioEndEffectorKinematics& eeKinematics,
          const PinocchioEndEffectorKinematics& armEeKinematics, ros::NodeHandle& controller_nh) : info_(info) {
    (void)pinocchioInterface; (void)eeKinematics; (void)armEeKinematics; (void)controller_nh;
  }
  virtual ~WbcBase() = default;
  virtual vector_t update(const vector_t& stateDesired, const vector_t& inputDesired, const vector_t& rbdStateMeasured, size_t mode,
                          scalar_t period, scalar_t time) = 0;
  virtual void loadTasksSetting(const std::string& taskFile, bool verbose) { (void)taskFile; (void)verbose; }
 protected:
  CentroidalModelInfo info_;
};
}  // namespace qm
