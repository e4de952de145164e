"""Oracle (test infrastructure only): the joint-level control law and the simulated actuator with transport delay.

Restates, with the reference's own container (a deque per joint, newest command at the front):
  QMController::updateControlLaw   qm_controllers/src/QMController.cpp:178-191 (posDes / velDes / torque formed at :147-157)
  QMHWSim::writeSim                qm_gazebo/src/QMHWSim.cpp:98-114 (delay: qm_gazebo/config/default.yaml:2; gains:
                                   qm_controllers/cfg/weight.cfg:7-8)
ros::Time / ros::Duration are integer nanoseconds; so are the stamps here. Parity unpinned by the reference (it holds no test
for this path); the CUDA kernel and the CPU port are compared with this restatement.
"""
from collections import deque

import numpy as np


class HybridJointCommand:
    __slots__ = ("stamp", "pos", "vel", "kp", "kd", "ff")

    def __init__(self, stamp, pos, vel, kp, kd, ff):
        self.stamp, self.pos, self.vel, self.kp, self.kd, self.ff = stamp, pos, vel, kp, kd, ff


class Actuator:
    def __init__(self, leg_kp=0.0, leg_kd=3.0, arm_kp=0.0, arm_kd=0.5, leg_enable_time=10.0, delay_ns=9_000_000):
        self.leg_kp, self.leg_kd, self.arm_kp, self.arm_kd = leg_kp, leg_kd, arm_kp, arm_kd
        self.leg_enable_time, self.delay_ns = leg_enable_time, delay_ns
        self.handles = [[0.0, 0.0, 0.0, 0.0, 0.0] for _ in range(18)]      # what each HybridJointHandle holds
        self.buffers = [deque() for _ in range(18)]

    def update_control_law(self, obs_time, x_des, u_des, cmd):
        pos_des, vel_des, torque = x_des[12:30], u_des[12:30], cmd[36:54]
        if obs_time > self.leg_enable_time:                                 # QMController.cpp:181-186
            for j in range(12):
                self.handles[j] = [pos_des[j], vel_des[j], self.leg_kp, self.leg_kd, torque[j]]
        for j in range(12, 18):                                             # QMController.cpp:188-190
            self.handles[j] = [pos_des[j], 0.0, self.arm_kp, self.arm_kd, torque[j]]

    def write_sim(self, time_ns, period_ns, q, v):
        tau = np.zeros(18)
        for j in range(18):
            buf = self.buffers[j]
            if time_ns == period_ns:                                        # simulation reset (QMHWSim.cpp:101-103)
                buf.clear()
            while buf and buf[-1].stamp + self.delay_ns < time_ns:          # :105-107
                buf.pop()
            buf.appendleft(HybridJointCommand(time_ns, *self.handles[j]))   # :108-109
            c = buf[-1]                                                     # :111-113
            tau[j] = c.kp * (c.pos - q[j]) + c.kd * (c.vel - v[j]) + c.ff
        return tau

    def step(self, time_ns, period_ns, obs_time, x_des, u_des, cmd, q, v):
        self.update_control_law(obs_time, x_des, u_des, cmd)
        return self.write_sim(time_ns, period_ns, q, v)
