"""Fill the C-ABI descriptors from the oracle's own parse of the reference inputs (TEST INFRASTRUCTURE).
Used to (a) drive the CPU port / CUDA path with oracle-parsed data and (b) check the product's C++ loaders."""
import ctypes as C
import os
import subprocess

import numpy as np

from qm_door_b200 import _abi
from . import sqp

HERE = os.path.dirname(os.path.abspath(__file__))


def model_desc(m):
    d = _abi.ModelDesc()
    d.nj = m.nj
    depth = np.zeros(m.nj, dtype=np.int32)
    for j in range(m.nj):
        depth[j] = 0 if m.parent[j] < 0 else depth[m.parent[j]] + 1
    _abi._set(d.parent, m.parent)
    _abi._set(d.jtype, m.jtype)
    _abi._set(d.depth, depth)
    sub = [sum(1 << i for i in range(m.nj) if m.path[i, j]) for j in range(m.nj)]
    pth = [sum(1 << k for k in range(m.nj) if m.path[j, k]) for j in range(m.nj)]
    _abi._set(d.submask, np.array(sub, dtype=np.uint32))
    _abi._set(d.pathmask, np.array(pth, dtype=np.uint32))
    d.max_depth = int(depth.max())
    _abi._set(d.foot_joint, m.foot_joint)
    d.ee_joint = int(m.ee_joint)
    _abi._set(d.axis, m.axis)
    _abi._set(d.Rp, m.Rp.reshape(m.nj, 9))
    _abi._set(d.pp, m.pp)
    _abi._set(d.mass, m.mass)
    _abi._set(d.com, m.com)
    _abi._set(d.inertia, m.inertia.reshape(m.nj, 9))
    _abi._set(d.foot_off, m.foot_off)
    _abi._set(d.ee_off, m.ee_off)
    _abi._set(d.ee_Roff, np.asarray(m.ee_Roff).reshape(9))
    d.total_mass = m.total_mass
    _abi._set(d.lower, np.nan_to_num(m.lower, neginf=-1e30))
    _abi._set(d.upper, np.nan_to_num(m.upper, posinf=1e30))
    _abi._set(d.effort, m.effort)
    std = [(0, (1, 0, 0)), (0, (0, 1, 0)), (0, (0, 0, 1)), (1, (0, 0, 1)), (1, (0, 1, 0)), (1, (1, 0, 0))]
    d.root6_standard = int(all(m.parent[k] == k - 1 and m.jtype[k] == std[k][0] and np.array_equal(m.axis[k], std[k][1])
                               and np.array_equal(m.Rp[k], np.eye(3)) and not m.pp[k].any() for k in range(6)))
    return d


def problem_desc(m, P):
    d = _abi.ProblemDesc()
    _abi._set(d.Q, P.Q.reshape(-1))
    _abi._set(d.R, sqp.input_cost_weight(m, P).reshape(-1))
    for k in ("mu_ee_pos", "mu_ee_ori", "mu_fee_pos", "mu_fee_ori", "fric_mu", "fric_bar_mu", "fric_bar_delta",
              "fric_reg", "fric_grip", "fric_hess_shift", "pos_bar_mu", "pos_bar_delta", "vel_bar_mu", "vel_bar_delta"):
        setattr(d, k, float(getattr(P, k)))
    for k in ("arm_pos_lo", "arm_pos_hi", "arm_vel_lo", "arm_vel_hi"):
        _abi._set(getattr(d, k), getattr(P, k))
    z = np.zeros(6)
    d.box_offset = float(sqp.relaxed_barrier(z - P.arm_pos_lo, P.pos_bar_mu, P.pos_bar_delta)[0].sum()
                         + sqp.relaxed_barrier(P.arm_pos_hi - z, P.pos_bar_mu, P.pos_bar_delta)[0].sum()
                         + sqp.relaxed_barrier(z - P.arm_vel_lo, P.vel_bar_mu, P.vel_bar_delta)[0].sum()
                         + sqp.relaxed_barrier(P.arm_vel_hi - z, P.vel_bar_mu, P.vel_bar_delta)[0].sum())
    d.swing_liftoff_vel = P.swing["liftOffVelocity"]
    d.swing_touchdown_vel = P.swing["touchDownVelocity"]
    d.swing_height = P.swing["swingHeight"]
    d.swing_time_scale = P.swing["swingTimeScale"]
    d.gravity = 9.81
    return d


def solver_desc(P, horizon=None, dt=None, max_nodes=None, max_events=32, max_targets=2):
    d = _abi.SolverDesc()
    d.dt = P.sqp["dt"] if dt is None else dt
    d.horizon = P.time_horizon if horizon is None else horizon
    d.delta_tol, d.g_max, d.g_min = P.sqp["deltaTol"], P.sqp["g_max"], P.sqp["g_min"]
    d.alpha_decay, d.alpha_min = P.sqp["alpha_decay"], P.sqp["alpha_min"]
    d.gamma_c, d.armijo_factor = P.sqp["gamma_c"], P.sqp["armijoFactor"]
    d.weak_eps, d.dt_min = 1e-6, 1e-8
    d.sqp_iterations, d.cost_tol = P.sqp["sqpIteration"], P.sqp.get("costTol", 1e-4)
    n = int(round(d.horizon / d.dt))
    d.max_nodes = (n + 1 + 24) if max_nodes is None else max_nodes
    d.max_events = max_events
    d.max_targets = max_targets
    return d


def wbc_desc(m, P, gains=None, mpc_variant=False):
    from . import wbc
    g = dict(wbc.DEFAULT_GAINS if gains is None else gains)
    d = _abi.WbcDesc()
    d.kp_swing, d.kd_swing = g["kp_swing"], g["kd_swing"]
    d.kp_base_height, d.kd_base_height = g["baseHeightKp"], g["baseHeightKd"]
    d.kp_base_linear, d.kd_base_linear = g["kp_base_linear"], g["kd_base_linear"]
    d.kp_base_angular, d.kd_base_angular = g["kp_base_angular"], g["kd_base_angular"]
    for k in ("kp_arm_joint", "kd_arm_joint", "kp_ee_linear", "kd_ee_linear", "kp_ee_angular", "kd_ee_angular"):
        _abi._set(getattr(d, k), g[k])
    d.friction_mu = P.friction_wbc
    _abi._set(d.tau_max, m.effort[6:])
    d.swing_weight, d.init_time, d.gravity = 100.0, 10.0, 9.81
    d.mpc_variant = int(mpc_variant)
    return d


def cport_wbc(md, wd, xd, ud, rbdm, mode, period, time, u_last, threads=1):
    lib = load_cport()
    B = xd.shape[0]
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    xd, ud, rbdm = (np.ascontiguousarray(a, dtype=np.float64) for a in (xd, ud, rbdm))
    mode = np.ascontiguousarray(mode, dtype=np.int32)
    period = np.ascontiguousarray(np.broadcast_to(period, (B,)), dtype=np.float64)
    time = np.ascontiguousarray(np.broadcast_to(time, (B,)), dtype=np.float64)
    cmd = np.zeros((B, 54))
    status = np.zeros(B, dtype=np.int32)
    lib.cport_wbc_batch(C.byref(md), C.byref(wd), B, dp(xd), dp(ud), dp(rbdm), dp(mode), dp(period), dp(time), dp(u_last),
                        dp(cmd), dp(status), threads)
    return cmd, status


def cport_wbc_levels(md, wd, xd, ud, rbdm, mode, period, time, u_last):
    """One solve on the CPU port with the per-level record -> (cmd, status, levels list, level-0 slack)."""
    from qm_door_b200 import unpack_wbc_levels
    lib = load_cport()
    dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)
    cmd, st, lv = np.zeros(54), C.c_int32(), np.zeros(lib.cport_wbc_levels_size())
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (xd, ud, rbdm, u_last)]
    lib.cport_wbc_levels(C.byref(md), C.byref(wd), keep[0].ctypes.data_as(C.c_void_p), keep[1].ctypes.data_as(C.c_void_p),
                         keep[2].ctypes.data_as(C.c_void_p), int(mode), C.c_double(period), C.c_double(time),
                         keep[3].ctypes.data_as(C.c_void_p), cmd.ctypes.data_as(C.c_void_p), C.byref(st), lv.ctypes.data_as(C.c_void_p))
    return (cmd, st.value) + unpack_wbc_levels(lv)


def cport_forward_dynamics(md, gravity, rbdm, tau, mode, dt, beta=0.0, threads=1):
    lib = load_cport()
    n = rbdm.shape[0]
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    rbdm, tau = np.ascontiguousarray(rbdm, dtype=np.float64), np.ascontiguousarray(tau, dtype=np.float64)
    mode = np.ascontiguousarray(mode, dtype=np.int32)
    nxt, f, st = np.zeros((n, 55)), np.zeros((n, 12)), np.zeros(n, dtype=np.int32)
    lib.cport_forward_dynamics(C.byref(md), C.c_double(gravity), n, dp(rbdm), dp(tau), dp(mode), C.c_double(dt), C.c_double(beta),
                               dp(nxt), dp(f), dp(st), threads)
    return nxt, f, st


KW_NAMES = ["R", "P", "AX", "BODY", "COMP", "ACM", "SV", "V", "HB", "FPOS", "FVEL", "EEP", "EER", "COM", "ABINV", "VEL", "RHS", "VSIZE",
            "FJ", "EEJ", "DH", "DFV", "F", "SIZE"]


def kw_offsets():
    """Offsets of the kinematics workspace (qm_core.h KW_*), read from the compiled CPU port (never hand-copied)."""
    out = (C.c_int * len(KW_NAMES))()
    n = load_cport().cport_kw_offsets(out)
    assert n == len(KW_NAMES)
    return dict(zip(KW_NAMES, [int(v) for v in out]))


_cport = None


def load_cport():
    """Build (if stale) and load the CPU port shared library."""
    global _cport
    if _cport is not None:
        return _cport
    so = os.path.join(HERE, "cport", "libcport.so")
    src = os.path.join(HERE, "cport", "cport.cpp")
    deps = [src] + [os.path.join(HERE, "..", "qm_door_b200", "csrc", f) for f in ("qm_core.h", "qm_types.h", "qm_mpc.h", "qm_buffers.h", "qm_wbc.h", "qm_sim.h", "qm_value.h", "qm_actuator.h", "qm_target.h")]
    deps = [p for p in deps if os.path.exists(p)]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-shared", "-fPIC", "-pthread",
                               "-o", so, src])
    _cport = C.CDLL(so)
    return _cport


class CPort:
    """ctypes front-end of the CPU port (same call shape as the CUDA C-ABI, host buffers)."""

    LS = dict(alpha=0, done=1, armijo=2, dxnorm=3, dunorm=4, base_merit=5, base_dyn=6, base_eq=7,
              new_merit=8, new_dyn=9, new_eq=10, iters=11)

    def __init__(self, md, pd, sd, B, threads=1, node_threads=1):
        """threads: workers over the problems of the batch; node_threads: workers over the nodes of one problem (the
        reference's sqp.nThreads, task.info:78) -- the latency setting, used with threads = 1."""
        self.lib = load_cport()
        self.lib.cport_create.restype = C.c_void_p
        self.md, self.pd, self.sd, self.B = md, pd, sd, B
        self.NMAX, self.EMAX, self.KT = sd.max_nodes, sd.max_events, sd.max_targets
        self.ctx = C.c_void_p(self.lib.cport_create(C.byref(md), C.byref(pd), C.byref(sd), B, threads))
        self.lib.cport_set_node_threads(self.ctx, int(node_threads))

    def close(self):
        if self.ctx:
            self.lib.cport_destroy(self.ctx)
            self.ctx = None

    def reset(self):
        self.lib.cport_reset(self.ctx)

    def cycle(self, t0, x0, events, modes, nevents, target_t, target_x):
        B, N = self.B, self.NMAX
        dp = lambda a: a.ctypes.data_as(C.c_void_p)
        t0 = np.ascontiguousarray(t0, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        events = np.ascontiguousarray(events, dtype=np.float64)
        modes = np.ascontiguousarray(modes, dtype=np.int32)
        nevents = np.ascontiguousarray(nevents, dtype=np.int32)
        target_t = np.ascontiguousarray(target_t, dtype=np.float64)
        target_x = np.ascontiguousarray(target_x, dtype=np.float64)
        assert events.shape == (B, self.EMAX) and modes.shape == (B, self.EMAX + 1)
        assert target_t.shape == (B, self.KT) and target_x.shape == (B, self.KT, 37)
        out = dict(t=np.zeros((B, N)), x=np.zeros((B, N, 30)), u=np.zeros((B, N, 30)), n=np.zeros(B, dtype=np.int32),
                   mode=np.zeros((B, N), dtype=np.int32), info=np.zeros((B, 16)), status=np.zeros(B, dtype=np.int32))
        self.lib.cport_mpc_cycle(self.ctx, dp(t0), dp(x0), dp(events), dp(modes), dp(nevents), dp(target_t), dp(target_x),
                                 dp(out["t"]), dp(out["x"]), dp(out["u"]), dp(out["n"]), dp(out["mode"]), dp(out["info"]),
                                 dp(out["status"]))
        return out

    def feedback_gains(self):
        K = np.zeros((self.B, self.NMAX, 30, 30))
        self.lib.cport_feedback_gains(self.ctx, K.ctypes.data_as(C.c_void_p))
        return K


class CPortActuator:
    """CPU port of the control law + delayed actuator with caller-owned state (same call shape as the CUDA C-ABI)."""

    def __init__(self, desc, n, capacity=32):
        self.lib, self.desc, self.n = load_cport(), desc, n
        self.stamp = np.zeros((n, capacity), dtype=np.int64)
        self.buf = np.zeros((n, capacity, 18, 5))
        self.hc = np.zeros((n, 2), dtype=np.int32)
        self.last = np.zeros((n, 18, 5))

    def step(self, time_ns, period_ns, obs_time, x_des, u_des, cmd, q, v):
        dp = lambda a: a.ctypes.data_as(C.c_void_p)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        time_ns = np.ascontiguousarray(time_ns, dtype=np.int64)
        obs_time, x_des, u_des, cmd, q, v = f(obs_time), f(x_des), f(u_des), f(cmd), f(q), f(v)
        tau, status = np.zeros((self.n, 18)), np.zeros(self.n, dtype=np.int32)
        self.lib.cport_actuator(C.byref(self.desc), self.n, dp(time_ns), C.c_int64(int(period_ns)), dp(obs_time), dp(x_des), dp(u_des),
                                dp(cmd), dp(q), dp(v), dp(self.stamp), dp(self.buf), dp(self.hc), dp(self.last), dp(tau), dp(status))
        return tau, status


def pack_schedules(schedules, EMAX):
    """[(events, modes)] -> padded arrays (events padded with +1e30, modes with STANCE)."""
    B = len(schedules)
    ev = np.full((B, EMAX), 1e30)
    md = np.full((B, EMAX + 1), 15, dtype=np.int32)
    ne = np.zeros(B, dtype=np.int32)
    for b, (e, m) in enumerate(schedules):
        assert len(e) <= EMAX and len(m) == len(e) + 1
        ev[b, :len(e)] = e
        md[b, :len(m)] = m
        ne[b] = len(e)
    return ev, md, ne
