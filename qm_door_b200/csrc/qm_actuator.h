// Joint command of the control law and the simulated actuator with transport delay: what sits between
// the whole-body controller's torques and the joints in the reference's simulation loop. One thread group per problem, a lane
// per joint; the same source is the kernel body and the CPU port.
//   control law   QMController::updateControlLaw   qm_controllers/src/QMController.cpp:178-191 (inputs formed at :147-157)
//   actuator      QMHWSim::writeSim                qm_gazebo/src/QMHWSim.cpp:98-114 (delay: qm_gazebo/config/default.yaml:2)
#pragma once
#include "qm_core.h"

namespace qm {

enum { ACT_POS = 0, ACT_VEL, ACT_KP, ACT_KD, ACT_FF, ACT_NF };           // one buffered HybridJointCommand
enum { ST_ACT_OVERFLOW = 64 };

// State of one problem: stamp [CAP] (ns), buf [CAP][18][ACT_NF], hc = {head (newest slot), count}, last [18][ACT_NF] (the
// command each joint handle currently holds: leg handles are not written before leg_enable_time and keep their value).
// time_ns / period_ns: simulation clock of writeSim; obs_time: the controller's observation time (the `> 10` test).
// x_des, u_des: evaluated policy; cmd: the whole-body controller's output [x*(36); tau(18)]; q, v: measured joint state.
template <class G>
QM_HDN void actuator_step(G g, const qmb200_actuator_desc& D, long long time_ns, long long period_ns, double obs_time,
                          const double* x_des, const double* u_des, const double* cmd, const double* q, const double* v,
                          long long* stamp, double* buf, int* hc, double* last, double* tau, int* status) {
  const int CAP = QMB200_ACT_CAPACITY;
  // updateControlLaw: posDes = joint angles of the optimized state, velDes = joint velocities of the optimized input
  QM_PFOR(g, j, 18) {
    double* c = last + ACT_NF * j;
    if (j >= 12) {
      c[ACT_POS] = x_des[12 + j]; c[ACT_VEL] = 0.0; c[ACT_KP] = D.arm_kp; c[ACT_KD] = D.arm_kd; c[ACT_FF] = cmd[36 + j];
    } else if (obs_time > D.leg_enable_time) {
      c[ACT_POS] = x_des[12 + j]; c[ACT_VEL] = u_des[12 + j]; c[ACT_KP] = D.leg_kp; c[ACT_KD] = D.leg_kd; c[ACT_FF] = cmd[36 + j];
    }
  }
  // writeSim: every member of the group derives the new buffer shape from the old one, then the slots are written
  int head = hc[0], count = hc[1];
  if (time_ns == period_ns) count = 0;                                            // simulation reset
  while (count > 0 && stamp[(head - count + 1 + CAP) % CAP] + D.delay_ns < time_ns) --count;   // commands older than the delay
  bool overflow = false;
  if (count == CAP) { overflow = true; --count; }
  head = (head + 1) % CAP;
  ++count;
  g.sync();
  if (g.tid() == 0) {
    stamp[head] = time_ns; hc[0] = head; hc[1] = count;
    if (overflow) status_or(status, ST_ACT_OVERFLOW);
  }
  QM_PFOR(g, j, 18)
    for (int f = 0; f < ACT_NF; ++f) buf[(head * 18 + j) * ACT_NF + f] = last[ACT_NF * j + f];
  g.sync();
  const int oldest = (head - count + 1 + CAP) % CAP;
  QM_PFOR(g, j, 18) {
    const double* c = buf + (oldest * 18 + j) * ACT_NF;
    tau[j] = c[ACT_KP] * (c[ACT_POS] - q[j]) + c[ACT_KD] * (c[ACT_VEL] - v[j]) + c[ACT_FF];
  }
  g.sync();
}

}  // namespace qm
