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
KERNEL_NAMES = ["k_schedule", "k_init_guess", "k_kin1", "k_kin2", "k_lq", "k_solve", "k_trial", "k_decide", "k_finalize", "k_policy", "k_proj",
                "k_backtrack", "k_step"]
INFO = dict(alpha=0, done=1, armijo=2, dxnorm=3, dunorm=4, base_merit=5, base_dyn=6, base_eq=7,
            new_merit=8, new_dyn=9, new_eq=10, iters=11, dx0sq=12, sqp_iterations=13, convergence=14)

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


def _host(a, shape, dtype, name):
    """Host buffer handed to the C-ABI: contiguous, of the dtype and shape the entry point reads (an undersized buffer would be
    an out-of-bounds read in the library)."""
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.shape != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), a.shape))
    return a


def _dev(t, shape, dtype, name, device=None, optional=False):
    """Device tensor handed to the C-ABI as a raw pointer: CUDA, contiguous, of the dtype and shape the kernels index."""
    import torch
    if t is None:
        if optional:
            return None
        raise ValueError("%s is required" % name)
    want = {np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64}[dtype]
    if not t.is_cuda or t.dtype != want or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
        raise ValueError("%s must be a contiguous CUDA tensor of dtype %s and shape %s (got %s %s on %s)"
                         % (name, want, tuple(shape), t.dtype, tuple(t.shape), t.device))
    if device is not None and t.device.index != device:
        raise ValueError("%s lives on cuda:%s, the context on cuda:%s" % (name, t.device.index, device))
    return C.c_void_p(t.data_ptr())


class MpcContext:
    """Mirror of the reference's SqpMpc + MPC_MRT_Interface pair for a batch of independent problems
    (qm_controllers/src/QMController.cpp:287-335): cycle() = advanceMpc(), evaluate_policy() = evaluatePolicy()."""

    def __init__(self, model, problem, solver, batch, device=0):
        self.L = lib()
        self.model, self.problem, self.solver = model, problem, solver
        self.B, self.NMAX, self.EMAX, self.KT = batch, solver.max_nodes, solver.max_events, solver.max_targets
        self.device = device
        # arguments of the asynchronous entry points are kept alive for a few calls: the stream may still be reading them when
        # the caller drops its references (torch's caching allocator would hand the memory to the next tensor)
        import collections
        self._keep = collections.deque(maxlen=8)
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

    def wait_for(self, other):
        """Order this context's stream behind everything enqueued so far on `other`'s stream (a context or a raw cudaStream_t)."""
        st = other if isinstance(other, int) or other is None else other.stream
        _check(self.L.qmb200_wait_stream(self.h, C.c_void_p(st)))

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
            self._pinned_keepalive = getattr(self, "_pinned_keepalive", [])
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

    def cycle_async(self, t0, x0, events, modes, nevents, target_t, target_x, out):
        """Submit one cycle (host buffers; `out` from alloc_outputs(pinned=True)) and return a ticket for wait(). The inputs are
        copied asynchronously from the arrays given here: pass page-locked arrays and leave them and `out` untouched until wait().
        Two submissions may be outstanding: the copy-out of cycle k overlaps the computation of cycle k + 1."""
        B = self.B
        f8, i4 = np.float64, np.int32
        args = (_host(t0, (B,), f8, "t0"), _host(x0, (B, 30), f8, "x0"), _host(events, (B, self.EMAX), f8, "events"),
                _host(modes, (B, self.EMAX + 1), i4, "modes"), _host(nevents, (B,), i4, "nevents"),
                _host(target_t, (B, self.KT), f8, "target_t"), _host(target_x, (B, self.KT, 37), f8, "target_x"))
        ticket = C.c_int64()
        _check(self.L.qmb200_mpc_cycle_batch_async(self.h, *[_p(a) for a in args], _p(out["t"]), _p(out["x"]), _p(out["u"]), _p(out["n"]),
                                                   _p(out["mode"]), _p(out["info"]), _p(out["status"]), C.byref(ticket)))
        self._keep.append(args)
        return ticket.value

    def wait(self, ticket):
        _check(self.L.qmb200_mpc_cycle_wait(self.h, C.c_int64(ticket)))

    # ---- multi-GPU: all-gather of the packed policy (one process per GPU; see include/qmb200.h)
    def comm_init(self, unique_id, rank, world):
        """ncclCommInitRank for this context (collective: every rank calls it with rank 0's id from nccl_unique_id())."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(self.L.qmb200_comm_init(self.h, buf, int(rank), int(world)))
        self.world, self.rank = int(world), int(rank)

    def enable_policy_buffer(self):
        _check(self.L.qmb200_enable_policy_buffer(self.h))
        self.world = getattr(self, "world", 1)

    def allgather_policy(self, gathered, comm=None):
        """All-gather of the last cycle's packed policy [B][NMAX][61] into gathered [world][B][NMAX][61] (torch CUDA tensor), on the
        context's communication stream: runs beside the next cycle. comm: a caller-owned ncclComm_t value, or None."""
        world = getattr(self, "world", 1)
        self._keep.append((gathered,))
        _check(self.L.qmb200_allgather_policy(self.h, C.c_void_p(comm), _dev(gathered, (world, self.B, self.NMAX, 61), np.float64, "gathered", self.device)))

    def policy_wait_stream(self, stream):
        _check(self.L.qmb200_policy_wait_stream(self.h, C.c_void_p(stream)))

    def comm_sync(self):
        _check(self.L.qmb200_comm_sync(self.h))

    def cycle_dev(self, t0, x0, events, modes, nevents, target_t, target_x, t_out=None, x_out=None, u_out=None,
                  n_out=None, mode_out=None, info=None, status=None):
        """Same with device-resident torch tensors (raw pointers passed to the C-ABI); asynchronous on the context's stream."""
        B, N, E, K, d = self.B, self.NMAX, self.EMAX, self.KT, self.device
        f8, i4 = np.float64, np.int32
        self._keep.append((t0, x0, events, modes, nevents, target_t, target_x, t_out, x_out, u_out, n_out, mode_out, info, status))
        _check(self.L.qmb200_mpc_cycle_batch_dev(
            self.h, _dev(t0, (B,), f8, "t0", d), _dev(x0, (B, 30), f8, "x0", d), _dev(events, (B, E), f8, "events", d),
            _dev(modes, (B, E + 1), i4, "modes", d), _dev(nevents, (B,), i4, "nevents", d), _dev(target_t, (B, K), f8, "target_t", d),
            _dev(target_x, (B, K, 37), f8, "target_x", d), _dev(t_out, (B, N), f8, "t_out", d, True),
            _dev(x_out, (B, N, 30), f8, "x_out", d, True), _dev(u_out, (B, N, 30), f8, "u_out", d, True),
            _dev(n_out, (B,), i4, "n_out", d, True), _dev(mode_out, (B, N), i4, "mode_out", d, True),
            _dev(info, (B, INFO_SIZE), f8, "info", d, True), _dev(status, (B,), i4, "status", d, True)))

    def evaluate_policy(self, t):
        t = _host(t, (self.B,), np.float64, "t")
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
        t = _host(t, (self.B,), np.float64, "t")
        x = _host(x, (self.B, 30), np.float64, "x")
        u = np.zeros((self.B, 30))
        mode = np.zeros(self.B, dtype=np.int32)
        _check(self.L.qmb200_evaluate_feedback_policy_batch(self.h, _p(t), _p(x), _p(u), _p(mode)))
        return u, mode

    def evaluate_policy_dev(self, t, x_des, u_des, mode):
        """evaluatePolicy with torch CUDA tensors, enqueued on the context's stream."""
        B, d = self.B, self.device
        self._keep.append((t, x_des, u_des, mode))
        _check(self.L.qmb200_evaluate_policy_batch_dev(self.h, _dev(t, (B,), np.float64, "t", d), _dev(x_des, (B, 30), np.float64, "x_des", d),
                                                       _dev(u_des, (B, 30), np.float64, "u_des", d), _dev(mode, (B,), np.int32, "mode", d)))

    def rbd_to_state_dev(self, rbd, x_out, yaw_last=None):
        n, d = int(rbd.shape[0]), self.device
        self._keep.append((rbd, x_out, yaw_last))
        _check(self.L.qmb200_rbd_to_state_batch_dev(self.h, n, _dev(rbd, (n, 55), np.float64, "rbd", d),
                                                    _dev(yaw_last, (n,), np.float64, "yaw_last", d, True), _dev(x_out, (n, 30), np.float64, "x_out", d)))

    def rbd_to_state(self, rbd, yaw_last=None):
        """Measured rbdState [n][55] -> MPC state [n][30] (QMController.cpp:239-244); yaw_last enables the yaw unwrapping."""
        n = int(np.shape(rbd)[0])
        rbd = _host(rbd, (n, 55), np.float64, "rbd")
        x = np.zeros((n, 30))
        yl = None if yaw_last is None else _host(yaw_last, (n,), np.float64, "yaw_last")
        _check(self.L.qmb200_rbd_to_state_batch(self.h, n, _p(rbd), _p(yl), _p(x)))
        return x

    def targets(self, desc, kind, cmd, obs_time, obs_state, ee_state, last_ee_target):
        """Command -> (target_t [n][2], target_x [n][2][37]) (QmTargetTrajectoriesPublisher_node.cpp:60-257); last_ee_target
        [n][7] is updated in place. kind: 0 base cmd_vel, 1 ee cmd_vel, 2 ee goal; cmd [n][7]."""
        n = int(np.shape(cmd)[0])
        cmd, obs_time = _host(cmd, (n, 7), np.float64, "cmd"), _host(obs_time, (n,), np.float64, "obs_time")
        obs_state, ee_state = _host(obs_state, (n, 30), np.float64, "obs_state"), _host(ee_state, (n, 7), np.float64, "ee_state")
        if not (isinstance(last_ee_target, np.ndarray) and last_ee_target.dtype == np.float64 and last_ee_target.flags.c_contiguous
                and last_ee_target.shape == (n, 7)):
            raise ValueError("last_ee_target must be a contiguous float64 array of shape (n, 7): it is updated in place")
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


def nccl_unique_id():
    """ncclGetUniqueId through the library's run-time NCCL binding: 128 bytes for rank 0 to hand to every rank."""
    buf = (C.c_char * 128)()
    _check(lib().qmb200_nccl_unique_id(buf))
    return bytes(buf.raw)


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


def unpack_wbc_levels(lv):
    """levels[QMB200_WBC_LEVELS_SIZE] -> ([dict(n, x[36], Z[36, n])] for the levels of the stack, level-0 slack[56])."""
    rec = 37 + 36 * 18
    nlev = int(lv[6 * rec])
    out = []
    for p in range(nlev):
        r = lv[p * rec:(p + 1) * rec]
        n = int(r[0])
        out.append(dict(n=n, x=r[1:37].copy(), Z=r[37:].reshape(36, 18)[:, :n].copy()))
    return out, lv[6 * rec + 1:6 * rec + 57].copy()


class WbcContext:
    """Mirror of qm::WbcBase / HierarchicalWbc for a batch of independent solves:
    update() has the argument meaning of WbcBase::update (qm_wbc/include/qm_wbc/WbcBase.h:31-32) and returns [x*(36); tau(18)]."""

    def __init__(self, model, wbc, batch, device=0):
        self.L = lib()
        self.L.qmb200_wbc_stream.restype = C.c_void_p
        self.B = batch
        self.device = device
        import collections
        self._keep = collections.deque(maxlen=8)      # see MpcContext
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
        """Device tensors, enqueued on this context's stream. Inputs produced on another stream (e.g. the MPC context's policy
        evaluation) must be ordered first: call wait_for(mpc_ctx) (or synchronise) before this."""
        B, d, f8 = self.B, self.device, np.float64
        self._keep.append((x_des, u_des, rbd, mode, period, time, cmd, status))
        _check(self.L.qmb200_wbc_batch_dev(self.h, _dev(x_des, (B, 30), f8, "x_des", d), _dev(u_des, (B, 30), f8, "u_des", d),
                                           _dev(rbd, (B, 55), f8, "rbd", d), _dev(mode, (B,), np.int32, "mode", d),
                                           _dev(period, (B,), f8, "period", d), _dev(time, (B,), f8, "time", d),
                                           _dev(cmd, (B, 54), f8, "cmd", d), _dev(status, (B,), np.int32, "status", d)))

    def wait_for(self, other):
        """Order this context's stream behind everything enqueued so far on `other`'s stream (an MpcContext, a WbcContext or a
        raw cudaStream_t value): the device-side MPC -> policy -> WBC chain needs no host synchronisation."""
        st = other if isinstance(other, int) or other is None else other.stream
        _check(self.L.qmb200_wbc_wait_stream(self.h, C.c_void_p(st)))

    def actuator(self, desc, time_ns, period_ns, obs_time, x_des, u_des, cmd, q, v):
        """One tick of the control law + delayed actuator (QMController.cpp:178-191, QMHWSim.cpp:98-114) -> (tau [B][18], status)."""
        B, f8 = self.B, np.float64
        time_ns = _host(time_ns, (B,), np.int64, "time_ns")
        obs_time, x_des, u_des = _host(obs_time, (B,), f8, "obs_time"), _host(x_des, (B, 30), f8, "x_des"), _host(u_des, (B, 30), f8, "u_des")
        cmd, q, v = _host(cmd, (B, 54), f8, "cmd"), _host(q, (B, 18), f8, "q"), _host(v, (B, 18), f8, "v")
        tau, status = np.zeros((self.B, 18)), np.zeros(self.B, dtype=np.int32)
        _check(self.L.qmb200_actuator_batch(self.h, C.byref(desc), _p(time_ns), C.c_int64(int(period_ns)), _p(obs_time), _p(x_des),
                                            _p(u_des), _p(cmd), _p(q), _p(v), _p(tau), _p(status)))
        return tau, status

    def actuator_dev(self, desc, time_ns, period_ns, obs_time, x_des, u_des, cmd, q, v, tau, status):
        B, d, f8 = self.B, self.device, np.float64
        self._keep.append((time_ns, obs_time, x_des, u_des, cmd, q, v, tau, status))
        _check(self.L.qmb200_actuator_batch_dev(self.h, C.byref(desc), _dev(time_ns, (B,), np.int64, "time_ns", d), C.c_int64(int(period_ns)),
                                                _dev(obs_time, (B,), f8, "obs_time", d), _dev(x_des, (B, 30), f8, "x_des", d),
                                                _dev(u_des, (B, 30), f8, "u_des", d), _dev(cmd, (B, 54), f8, "cmd", d),
                                                _dev(q, (B, 18), f8, "q", d), _dev(v, (B, 18), f8, "v", d),
                                                _dev(tau, (B, 18), f8, "tau", d), _dev(status, (B,), np.int32, "status", d)))

    def levels(self, x_des, u_des, rbd, mode, period, time, u_last):
        """One solve with what the reference's HoQp objects expose per level (HoQp.h:21-36) -> (cmd[54], status, [dict(n, x, Z)] per
        level, level-0 slack[56]). Diagnostic entry; does not touch the per-solve inputLast_ state of the context."""
        f = lambda a, n: _host(a, (n,), np.float64, "argument")
        LV = 6 * (37 + 36 * 18) + 1 + 56
        cmd, st, lv = np.zeros(54), C.c_int32(), np.zeros(LV)
        _check(self.L.qmb200_wbc_levels(self.h, _p(f(x_des, 30)), _p(f(u_des, 30)), _p(f(rbd, 55)), int(mode), C.c_double(period), C.c_double(time),
                                        _p(f(u_last, 30)), _p(cmd), C.byref(st), _p(lv)))
        return (cmd, st.value) + unpack_wbc_levels(lv)

    def forward_dynamics(self, rbd, tau, mode, dt, beta=0.0):
        """One forward-dynamics step behind the actuator (stance feet of `mode` held by point contacts; see include/qmb200.h)
        -> (rbd_next [B][55], contact forces [B][12], status [B])."""
        B = self.B
        rbd, tau = _host(rbd, (B, 55), np.float64, "rbd"), _host(tau, (B, 18), np.float64, "tau")
        mode = _host(mode, (B,), np.int32, "mode")
        nxt, f, st = np.zeros((B, 55)), np.zeros((B, 12)), np.zeros(B, dtype=np.int32)
        _check(self.L.qmb200_forward_dynamics_batch(self.h, _p(rbd), _p(tau), _p(mode), C.c_double(dt), C.c_double(beta), _p(nxt), _p(f), _p(st)))
        return nxt, f, st

    def forward_dynamics_dev(self, rbd, tau, mode, dt, beta, rbd_next, contact_forces, status):
        B, d, f8 = self.B, self.device, np.float64
        self._keep.append((rbd, tau, mode, rbd_next, contact_forces, status))
        _check(self.L.qmb200_forward_dynamics_batch_dev(self.h, _dev(rbd, (B, 55), f8, "rbd", d), _dev(tau, (B, 18), f8, "tau", d),
                                                        _dev(mode, (B,), np.int32, "mode", d), C.c_double(dt), C.c_double(beta),
                                                        _dev(rbd_next, (B, 55), f8, "rbd_next", d), _dev(contact_forces, (B, 12), f8, "contact_forces", d),
                                                        _dev(status, (B,), np.int32, "status", d)))

    def actuator_reset(self):
        _check(self.L.qmb200_actuator_reset(self.h))

    def sync(self):
        _check(self.L.qmb200_wbc_sync(self.h))

    def kernel_time(self, reset=False):
        ms = C.c_double()
        n = C.c_int64()
        _check(self.L.qmb200_wbc_kernel_time(self.h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value
