"""Mode schedule, swing-height reference and time discretisation (TEST INFRASTRUCTURE).

Restates [upstream] ocs2_legged_robot GaitSchedule / SwingTrajectoryPlanner / CubicSpline / SplineCpg and
ocs2_oc timeDiscretizationWithEvents as wired by the reference at
qm_interface/src/QMInterface.cpp:444-480 (setupReferenceManager / loadGaitSchedule) with the
parameters of qm_controllers/config/task.info:24-31 and gait templates of config/gait.info:17-255.
Mode ids are integers (4-bit stance mask LF=8 RF=4 LH=2 RH=1) and must match the CUDA path bit-exactly.
"""
import bisect

import numpy as np

EV_NONE, EV_PRE, EV_POST = 0, 1, 2
WEAK_EPS = 1e-6      # [upstream] ocs2::numeric_traits::weakEpsilon<double>()
DT_MIN = 1e-8        # [upstream] 10 * numeric_traits::limitEpsilon<double>()
STANCE = 15


def stance_legs(mode):
    """[upstream] modeNumber2StanceLeg: contact flags in contact order LF, RF, LH, RH."""
    return [(mode >> 3) & 1, (mode >> 2) & 1, (mode >> 1) & 1, mode & 1]


def tile_schedule(gait, t_insert, t_upper):
    """Mode schedule = STANCE until t_insert, then the template tiled until an event >= t_upper, then STANCE.
    ([upstream] GaitSchedule::insertModeSequenceTemplate + tileModeSequenceTemplate, steady state.)"""
    events = [float(t_insert)]
    modes = [STANCE]
    sw = gait["times"]
    while events[-1] < t_upper:
        for i, md in enumerate(gait["modes"]):
            modes.append(md)
            events.append(events[-1] + (sw[i + 1] - sw[i]))
    modes.append(STANCE)
    return np.array(events), np.array(modes, dtype=np.int32)


def mode_index(events, t):
    """[upstream] lookup::findIndexInTimeArray: number of events strictly below t (std::lower_bound)."""
    return bisect.bisect_left(list(events), t)


def time_grid(t0, tf, dt, events):
    """[upstream] timeDiscretizationWithEvents. Returns (times, event_flags)."""
    times = [t0]
    flags = [EV_NONE]
    nxt = mode_index(events, t0)
    t_next = t0
    while times[-1] < tf:
        t_next = t_next + dt
        ev = EV_NONE
        if nxt < len(events) and t_next >= events[nxt]:
            t_next = events[nxt]
            ev = EV_PRE
            nxt += 1
        if t_next >= tf:
            t_next = tf
            ev = EV_NONE
        if t_next > times[-1] + DT_MIN:
            times.append(t_next)
            flags.append(ev)
        else:                      # closer than dt_min to the last node: that node is moved (and may become the pre-event node)
            times[-1] = t_next
            flags[-1] = ev
        if ev == EV_PRE:           # the post-event node follows the pre-event node in both branches
            times.append(t_next)
            flags.append(EV_POST)
    return np.array(times), np.array(flags, dtype=np.int32)


def interval_start(t, flag):
    return t + WEAK_EPS if flag == EV_POST else t


def interval_end(t, flag):
    return t - WEAK_EPS if flag == EV_PRE else t


class _Cubic:
    """[upstream] ocs2::legged_robot::CubicSpline (Hermite, normalised time)."""

    def __init__(self, t0, p0, v0, t1, p1, v1):
        self.t0, self.dt = t0, t1 - t0
        dp, dv = p1 - p0, v1 - v0
        self.c0 = p0
        self.c1 = v0 * self.dt
        self.c2 = -(3.0 * v0 + dv) * self.dt + 3.0 * dp
        self.c3 = (2.0 * v0 + dv) * self.dt - 2.0 * dp

    def position(self, t):
        tn = (t - self.t0) / self.dt
        return self.c3 * tn ** 3 + self.c2 * tn ** 2 + self.c1 * tn + self.c0

    def velocity(self, t):
        tn = (t - self.t0) / self.dt
        return (3.0 * self.c3 * tn * tn + 2.0 * self.c2 * tn + self.c1) / self.dt


class SwingPlanner:
    """[upstream] SwingTrajectoryPlanner::update on flat terrain (height 0) + getZvelocityConstraint."""

    def __init__(self, events, modes, cfg):
        self.events = list(events)
        nph = len(modes)
        self.splines = [[None] * nph for _ in range(4)]
        for leg in range(4):
            flags = [stance_legs(int(md))[leg] for md in modes]
            for p in range(nph):
                if flags[p]:
                    self.splines[leg][p] = None  # stance: height 0, velocity 0
                    continue
                start = -1
                for ip in range(p - 1, -1, -1):
                    if flags[ip]:
                        start = ip
                        break
                final = nph - 1
                for ip in range(p + 1, nph):
                    if flags[ip]:
                        final = ip - 1
                        break
                assert start >= 0 and final < len(self.events), "swing phase not bracketed by stance phases"
                ts, tf = self.events[start], self.events[final]
                scaling = min(1.0, (tf - ts) / cfg["swingTimeScale"])
                tm = 0.5 * (ts + tf)
                left = _Cubic(ts, 0.0, scaling * cfg["liftOffVelocity"], tm, scaling * cfg["swingHeight"], 0.0)
                right = _Cubic(tm, scaling * cfg["swingHeight"], 0.0, tf, 0.0, scaling * cfg["touchDownVelocity"])
                self.splines[leg][p] = (tm, left, right)

    def z_velocity(self, leg, t):
        sp = self.splines[leg][mode_index(self.events, t)]
        if sp is None:
            return 0.0
        tm, left, right = sp
        return left.velocity(t) if t < tm else right.velocity(t)

    def z_position(self, leg, t):
        sp = self.splines[leg][mode_index(self.events, t)]
        if sp is None:
            return 0.0
        tm, left, right = sp
        return left.position(t) if t < tm else right.position(t)
