"""qm_door_b200 — B200-native MPC + WBC hot path of danisotelo/qm_door behind a C-ABI (include/qmb200.h).

This package is only the thin Python mirror of the host interface (ctypes over libqmb200.so); all numerics
run in hand-written sm_100a CUDA kernels (csrc/).  There is NO CPU fallback: compute entry points raise if the
shared library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import ModelDesc, ProblemDesc, SolverDesc  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QMB200_LIB_PATH", os.path.join(_HERE, "libqmb200.so"))   # override: development builds only
INFO_SIZE = 16
KERNEL_NAMES = ["k_schedule", "k_init_guess", "k_kin1", "k_kin2", "k_lq", "k_solve", "k_trial", "k_decide", "k_finalize", "k_policy", "k_proj"]
INFO = dict(alpha=0, done=1, armijo=2, dxnorm=3, dunorm=4, base_merit=5, base_dyn=6, base_eq=7,
            new_merit=8, new_dyn=9, new_eq=10, iters=11)

_lib = None


class Qmb200Error(RuntimeError):
    pass


def lib():
    """Load libqmb200.so (built in-tree by __graft_entry__.build() / make). Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Qmb200Error("libqmb200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.qmb200_last_error.restype = C.c_char_p
        L.qmb200_kernel_name.restype = C.c_char_p
        L.qmb200_stream.restype = C.c_void_p
        L.qmb200_device_bytes.restype = C.c_int64
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise Qmb200Error(lib().qmb200_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class MpcContext:
    """Mirror of the reference's SqpMpc + MPC_MRT_Interface pair for a batch of independent problems
    (qm_controllers/src/QMController.cpp:287-335): cycle() = advanceMpc(), evaluate_policy() = evaluatePolicy()."""

    def __init__(self, model, problem, solver, batch, device=0):
        self.L = lib()
        self.model, self.problem, self.solver = model, problem, solver
        self.B, self.NMAX, self.EMAX, self.KT = batch, solver.max_nodes, solver.max_events, solver.max_targets
        h = C.c_void_p()
        _check(self.L.qmb200_create(C.byref(model), C.byref(problem), C.byref(solver), batch, device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.qmb200_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        _check(self.L.qmb200_mpc_reset(self.h))

    def sync(self):
        _check(self.L.qmb200_sync(self.h))

    @property
    def stream(self):
        return self.L.qmb200_stream(self.h)

    @property
    def device_bytes(self):
        return self.L.qmb200_device_bytes(self.h)

    def alloc_outputs(self, pinned=False):
        B, N = self.B, self.NMAX
        shapes = dict(t=((B, N), np.float64), x=((B, N, 30), np.float64), u=((B, N, 30), np.float64),
                      n=((B,), np.int32), mode=((B, N), np.int32), info=((B, INFO_SIZE), np.float64),
                      status=((B,), np.int32))
        if pinned:
            import torch
            out = {}
            self._pinned_keepalive = []
            for k, (shp, dt) in shapes.items():
                t = torch.zeros(shp, dtype=torch.float64 if dt == np.float64 else torch.int32).pin_memory()
                self._pinned_keepalive.append(t)
                out[k] = t.numpy()
            return out
        return {k: np.zeros(shp, dtype=dt) for k, (shp, dt) in shapes.items()}

    def cycle(self, t0, x0, events, modes, nevents, target_t, target_x, out=None):
        """One SQP cycle for the whole batch; host (numpy) buffers in and out."""
        B = self.B
        t0 = np.ascontiguousarray(t0, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        events = np.ascontiguousarray(events, dtype=np.float64)
        modes = np.ascontiguousarray(modes, dtype=np.int32)
        nevents = np.ascontiguousarray(nevents, dtype=np.int32)
        target_t = np.ascontiguousarray(target_t, dtype=np.float64)
        target_x = np.ascontiguousarray(target_x, dtype=np.float64)
        if t0.shape != (B,) or x0.shape != (B, 30):
            raise ValueError("t0 must be [B], x0 [B,30]")
        if events.shape != (B, self.EMAX) or modes.shape != (B, self.EMAX + 1) or nevents.shape != (B,):
            raise ValueError("events must be [B,EMAX], modes [B,EMAX+1], nevents [B]")
        if target_t.shape != (B, self.KT) or target_x.shape != (B, self.KT, 37):
            raise ValueError("target_t must be [B,KT], target_x [B,KT,37]")
        if out is None:
            out = self.alloc_outputs()
        _check(self.L.qmb200_mpc_cycle_batch(self.h, _p(t0), _p(x0), _p(events), _p(modes), _p(nevents), _p(target_t),
                                             _p(target_x), _p(out["t"]), _p(out["x"]), _p(out["u"]), _p(out["n"]),
                                             _p(out["mode"]), _p(out["info"]), _p(out["status"])))
        return out

    def cycle_dev(self, t0, x0, events, modes, nevents, target_t, target_x, t_out=None, x_out=None, u_out=None,
                  n_out=None, mode_out=None, info=None, status=None):
        """Same with device-resident torch tensors (raw pointers passed to the C-ABI)."""
        dp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        _check(self.L.qmb200_mpc_cycle_batch_dev(self.h, dp(t0), dp(x0), dp(events), dp(modes), dp(nevents), dp(target_t),
                                                 dp(target_x), dp(t_out), dp(x_out), dp(u_out), dp(n_out), dp(mode_out),
                                                 dp(info), dp(status)))

    def evaluate_policy(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        x = np.zeros((self.B, 30))
        u = np.zeros((self.B, 30))
        mode = np.zeros(self.B, dtype=np.int32)
        _check(self.L.qmb200_evaluate_policy_batch(self.h, _p(t), _p(x), _p(u), _p(mode)))
        return x, u, mode

    def feedback_gains(self):
        """K [B][NMAX][30][30] of the last cycle in the original input coordinates (useFeedbackPolicy)."""
        K = np.zeros((self.B, self.solver.max_nodes, 30, 30))
        _check(self.L.qmb200_feedback_gains(self.h, _p(K)))
        return K

    def evaluate_feedback_policy(self, t, x):
        t = np.ascontiguousarray(t, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.zeros((self.B, 30))
        mode = np.zeros(self.B, dtype=np.int32)
        _check(self.L.qmb200_evaluate_feedback_policy_batch(self.h, _p(t), _p(x), _p(u), _p(mode)))
        return u, mode

    def evaluate_policy_dev(self, t, x_des, u_des, mode):
        """evaluatePolicy with torch CUDA tensors, enqueued on the context's stream."""
        dp = lambda a: C.c_void_p(a.data_ptr())
        _check(self.L.qmb200_evaluate_policy_batch_dev(self.h, dp(t), dp(x_des), dp(u_des), dp(mode)))

    def rbd_to_state_dev(self, rbd, x_out, yaw_last=None):
        dp = lambda a: None if a is None else C.c_void_p(a.data_ptr())
        _check(self.L.qmb200_rbd_to_state_batch_dev(self.h, int(rbd.shape[0]), dp(rbd), dp(yaw_last), dp(x_out)))

    def rbd_to_state(self, rbd, yaw_last=None):
        """Measured rbdState [n][55] -> MPC state [n][30] (QMController.cpp:239-244); yaw_last enables the yaw unwrapping."""
        rbd = np.ascontiguousarray(rbd, dtype=np.float64)
        n = rbd.shape[0]
        x = np.zeros((n, 30))
        yl = None if yaw_last is None else np.ascontiguousarray(yaw_last, dtype=np.float64)
        _check(self.L.qmb200_rbd_to_state_batch(self.h, n, _p(rbd), _p(yl), _p(x)))
        return x

    def targets(self, desc, kind, cmd, obs_time, obs_state, ee_state, last_ee_target):
        """Command -> (target_t [n][2], target_x [n][2][37]) (QmTargetTrajectoriesPublisher_node.cpp:60-257); last_ee_target
        [n][7] is updated in place. kind: 0 base cmd_vel, 1 ee cmd_vel, 2 ee goal; cmd [n][7]."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        cmd, obs_time, obs_state, ee_state = f(cmd), f(obs_time), f(obs_state), f(ee_state)
        assert last_ee_target.dtype == np.float64 and last_ee_target.flags.c_contiguous
        n = cmd.shape[0]
        tt, tx = np.zeros((n, 2)), np.zeros((n, 2, 37))
        _check(self.L.qmb200_targets_batch(self.h, C.byref(desc), int(kind), n, _p(cmd), _p(obs_time), _p(obs_state), _p(ee_state),
                                           _p(last_ee_target), _p(tt), _p(tx)))
        return tt, tx

    def set_profiling(self, on):
        _check(self.L.qmb200_set_profiling(self.h, int(bool(on))))

    def kernel_times(self, reset=False):
        ms = np.zeros(len(KERNEL_NAMES))
        cnt = np.zeros(len(KERNEL_NAMES), dtype=np.int64)
        _check(self.L.qmb200_get_kernel_times(self.h, _p(ms), _p(cnt), int(reset)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_NAMES)}


# ----------------------------------------------------------------------------- host-side loaders (no GPU needed)
DATA_DIR = os.path.join(_HERE, "data")
DEFAULT_URDF = os.path.join(DATA_DIR, "aliengo_z1.urdf")
DEFAULT_TASK = os.path.join(DATA_DIR, "aliengo_z1_task.info")
DEFAULT_REFERENCE = os.path.join(DATA_DIR, "aliengo_z1_reference.info")
DEFAULT_GAIT = os.path.join(DATA_DIR, "aliengo_z1_gait.info")


def load_model(urdf_path=DEFAULT_URDF):
    """URDF -> ModelDesc (mirror of QMInterface::setupModel, qm_interface/src/QMInterface.cpp:408-416)."""
    m = ModelDesc()
    _check(lib().qmb200_load_urdf(urdf_path.encode(), C.byref(m)))
    return m


def load_problem(model, task_info=DEFAULT_TASK, reference_info=DEFAULT_REFERENCE):
    """task.info / reference.info -> (ProblemDesc, SolverDesc, initial_state[30]) (QMInterface.cpp:37-142)."""
    p, s = ProblemDesc(), SolverDesc()
    x = np.zeros(30)
    _check(lib().qmb200_load_problem(task_info.encode(), reference_info.encode() if reference_info else None,
                                     C.byref(model), C.byref(p), C.byref(s), _p(x)))
    return p, s, x


def actuator_defaults():
    """Gains and delay of the control law / simulated actuator (QMController.cpp:181-190, weight.cfg:7-8, default.yaml:2)."""
    from ._abi import ActuatorDesc
    d = ActuatorDesc()
    lib().qmb200_actuator_defaults(C.byref(d))
    return d


def load_targets(task_info=DEFAULT_TASK, reference_info=DEFAULT_REFERENCE):
    """Constants of the command -> reference conversion (QmTargetTrajectoriesPublisher_node.cpp:268-272)."""
    from ._abi import TargetDesc
    d = TargetDesc()
    _check(lib().qmb200_load_targets(task_info.encode(), reference_info.encode(), C.byref(d)))
    return d


def load_gait(name, gait_info=DEFAULT_GAIT, capacity=32):
    """gait.info template -> (switching_times[n+1], modes[n]) (qm_controllers/src/GaitTopicPublisher.cpp:31-44)."""
    sw = np.zeros(capacity + 1)
    md = np.zeros(capacity, dtype=np.int32)
    n = C.c_int32()
    _check(lib().qmb200_load_gait(gait_info.encode(), name.encode(), capacity, _p(sw), _p(md), C.byref(n)))
    return sw[:n.value + 1].copy(), md[:n.value].copy()


def tile_schedule(switching_times, modes, t_insert, t_upper, capacity):
    """Mode schedule: STANCE until t_insert, template tiled past t_upper, STANCE. Returns padded (events, modes, n)."""
    ev = np.zeros(capacity)
    ms = np.zeros(capacity + 1, dtype=np.int32)
    n = C.c_int32()
    sw = np.ascontiguousarray(switching_times, dtype=np.float64)
    md = np.ascontiguousarray(modes, dtype=np.int32)
    _check(lib().qmb200_tile_schedule(_p(sw), _p(md), len(md), C.c_double(t_insert), C.c_double(t_upper), capacity,
                                      _p(ev), _p(ms), C.byref(n)))
    return ev, ms, n.value


# ----------------------------------------------------------------------------- whole-body controller
from ._abi import WbcDesc  # noqa: E402,F401


def load_wbc(model, task_info=DEFAULT_TASK):
    """Gains (qm_wbc/cfg/wbcWigeht.cfg defaults), torque limits and friction coefficient (WbcBase::loadTasksSetting)."""
    w = WbcDesc()
    _check(lib().qmb200_load_wbc(task_info.encode(), C.byref(model), C.byref(w)))
    return w


class WbcContext:
    """Mirror of qm::WbcBase / HierarchicalWbc for a batch of independent solves:
    update() has the argument meaning of WbcBase::update (qm_wbc/include/qm_wbc/WbcBase.h:31-32) and returns [x*(36); tau(18)]."""

    def __init__(self, model, wbc, batch, device=0):
        self.L = lib()
        self.L.qmb200_wbc_stream.restype = C.c_void_p
        self.B = batch
        h = C.c_void_p()
        _check(self.L.qmb200_wbc_create(C.byref(model), C.byref(wbc), batch, device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.qmb200_wbc_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        _check(self.L.qmb200_wbc_reset(self.h))

    def set_gains(self, wbc):
        _check(self.L.qmb200_wbc_set_gains(self.h, C.byref(wbc)))

    @property
    def stream(self):
        return self.L.qmb200_wbc_stream(self.h)

    def update(self, x_des, u_des, rbd, mode, period, time, cmd=None, status=None):
        B = self.B
        x_des = np.ascontiguousarray(x_des, dtype=np.float64)
        u_des = np.ascontiguousarray(u_des, dtype=np.float64)
        rbd = np.ascontiguousarray(rbd, dtype=np.float64)
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        period = np.ascontiguousarray(np.broadcast_to(period, (B,)), dtype=np.float64)
        time = np.ascontiguousarray(np.broadcast_to(time, (B,)), dtype=np.float64)
        if x_des.shape != (B, 30) or u_des.shape != (B, 30) or rbd.shape != (B, 55) or mode.shape != (B,):
            raise ValueError("x_des, u_des must be [B,30], rbd [B,55], mode [B]")
        cmd = np.zeros((B, 54)) if cmd is None else cmd
        status = np.zeros(B, dtype=np.int32) if status is None else status
        _check(self.L.qmb200_wbc_batch(self.h, _p(x_des), _p(u_des), _p(rbd), _p(mode), _p(period), _p(time), _p(cmd), _p(status)))
        return cmd, status

    def update_dev(self, x_des, u_des, rbd, mode, period, time, cmd, status):
        dp = lambda t: C.c_void_p(t.data_ptr())
        _check(self.L.qmb200_wbc_batch_dev(self.h, dp(x_des), dp(u_des), dp(rbd), dp(mode), dp(period), dp(time), dp(cmd), dp(status)))

    def actuator(self, desc, time_ns, period_ns, obs_time, x_des, u_des, cmd, q, v):
        """One tick of the control law + delayed actuator (QMController.cpp:178-191, QMHWSim.cpp:98-114) -> (tau [B][18], status)."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        time_ns = np.ascontiguousarray(time_ns, dtype=np.int64)
        obs_time, x_des, u_des, cmd, q, v = f(obs_time), f(x_des), f(u_des), f(cmd), f(q), f(v)
        tau, status = np.zeros((self.B, 18)), np.zeros(self.B, dtype=np.int32)
        _check(self.L.qmb200_actuator_batch(self.h, C.byref(desc), _p(time_ns), C.c_int64(int(period_ns)), _p(obs_time), _p(x_des),
                                            _p(u_des), _p(cmd), _p(q), _p(v), _p(tau), _p(status)))
        return tau, status

    def actuator_dev(self, desc, time_ns, period_ns, obs_time, x_des, u_des, cmd, q, v, tau, status):
        dp = lambda t: C.c_void_p(t.data_ptr())
        _check(self.L.qmb200_actuator_batch_dev(self.h, C.byref(desc), dp(time_ns), C.c_int64(int(period_ns)), dp(obs_time), dp(x_des),
                                                dp(u_des), dp(cmd), dp(q), dp(v), dp(tau), dp(status)))

    def actuator_reset(self):
        _check(self.L.qmb200_actuator_reset(self.h))

    def sync(self):
        _check(self.L.qmb200_wbc_sync(self.h))

    def kernel_time(self, reset=False):
        ms = C.c_double()
        n = C.c_int64()
        _check(self.L.qmb200_wbc_kernel_time(self.h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value
