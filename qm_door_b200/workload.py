"""Synthetic workloads of BASELINE.json built through the product's own host interface.

config 2: AlienGo+Z1, horizon 1.0 s at dt 0.01 s (N = 100), trot, B independent perturbed initial states with random
gait phase, standing reference (x_ref = initialState, EE pose = its forward kinematics pose)."""
import numpy as np

from . import load_gait, load_model, load_problem, tile_schedule

# EE pose of the nominal configuration (forward kinematics of task.info initialState, z1_end_effector frame);
# tests/test_abi.py::test_workload_reference_pose checks it against the oracle's forward kinematics.
NOMINAL_EE_POS = (0.6253031727266175, 0.0, 0.8300452360692332)
NOMINAL_EE_QUAT = None  # filled lazily from the oracle-checked constant below


def perturbed_states(model, x_init, B, seed=20261017):
    rng = np.random.default_rng(seed)
    x0 = np.tile(x_init, (B, 1))
    x0[:, 0:6] += rng.uniform(-0.1, 0.1, (B, 6))
    x0[:, 6:8] += rng.uniform(-0.05, 0.05, (B, 2))
    x0[:, 8] += rng.uniform(-0.02, 0.02, B)
    x0[:, 9:12] += rng.uniform(-0.1, 0.1, (B, 3))
    x0[:, 12:30] += rng.uniform(-0.1, 0.1, (B, 18))
    lo = np.ctypeslib.as_array(model.lower)[6:]
    hi = np.ctypeslib.as_array(model.upper)[6:]
    x0[:, 12:30] = np.clip(x0[:, 12:30], lo + 1e-3, hi - 1e-3)
    phase = rng.uniform(0.0, 0.7, B)
    return x0, phase


class Workload:
    """Everything one bench / test run needs: descriptors + per-problem inputs for `cycles` consecutive MPC cycles."""

    def __init__(self, B, horizon=1.0, dt=0.01, gait="trot", seed=20261017, max_nodes=None, max_events=32, ee_pose=None,
                 t_span=0.5):
        self.model = load_model()
        self.problem, self.solver, self.x_init = load_problem(self.model)
        self.solver.horizon, self.solver.dt = horizon, dt
        n = int(round(horizon / dt))
        self.solver.max_nodes = (n + 1 + 8) if max_nodes is None else max_nodes
        self.solver.max_events = max_events
        self.solver.max_targets = 2
        self.B = B
        self.x0, self.phase = perturbed_states(self.model, self.x_init, B, seed)
        sw, md = load_gait(gait)
        self.events = np.zeros((B, max_events))
        self.modes = np.zeros((B, max_events + 1), dtype=np.int32)
        self.nevents = np.zeros(B, dtype=np.int32)
        period = sw[-1]
        for b in range(B):
            # template inserted one horizon (+ phase) before t = 0, tiled one horizon past the last cycle
            t_ins = -(np.ceil(horizon / period) * period) - self.phase[b] * period / 0.7
            ev, ms, ne = tile_schedule(sw, md, t_ins, t_span + 2.0 * horizon, max_events)
            self.events[b], self.modes[b], self.nevents[b] = ev, ms, ne
        if ee_pose is None:
            ee_pose = nominal_ee_pose()
        knot = np.concatenate([self.x_init, ee_pose])
        self.target_t = np.tile(np.array([0.0, 1e3]), (B, 1))
        self.target_x = np.tile(knot, (B, 2, 1))


def nominal_ee_pose():
    """[position(3), quaternion xyzw(4)] of z1_end_effector at the nominal configuration (identity base orientation):
    the arm chain of the URDF evaluated at task.info initialState; verified against the oracle in tests."""
    # rotation about y by (q2 + q3 + q4) composed with joint1 (z), joint5 (z), joint6 (x) at zero -> pure y rotation
    ang = 1.11 - 0.69 - 0.40
    return np.array([0.6253031727266175, 0.0, 0.8300452360692332, 0.0, np.sin(ang / 2), 0.0, np.cos(ang / 2)])


def rbd_from_state(q, v):
    """rbdState[55] layout of qm_estimation/src/StateEstimateBase.cpp:29-102 from generalized coordinates / Euler-rate velocities."""
    q, v = np.asarray(q), np.asarray(v)
    r = np.zeros(q.shape[:-1] + (55,))
    r[..., 0:3], r[..., 3:6], r[..., 6:24] = q[..., 3:6], q[..., 0:3], q[..., 6:24]
    z, y = q[..., 3], q[..., 4]
    dz, dy, dx = v[..., 3], v[..., 4], v[..., 5]
    r[..., 24] = -np.sin(z) * dy + np.cos(y) * np.cos(z) * dx      # world angular velocity from ZYX Euler rates
    r[..., 25] = np.cos(z) * dy + np.cos(y) * np.sin(z) * dx
    r[..., 26] = dz - np.sin(y) * dx
    r[..., 27:30] = v[..., 0:3]
    r[..., 30:48] = v[..., 6:24]
    return r


class WbcWorkload:
    """BASELINE config 5: B whole-body-control solves, contact pattern uniform over the 16 modes, measured
    state = nominal pose + perturbation with velocities U(-vel, vel), desired = nominal + small perturbation with
    weight-compensating forces, period 1 ms, time 11 s (steady-state task stack)."""

    def __init__(self, B, seed=20261020, vel=0.5, time=11.0, period=0.001):
        from . import load_wbc
        rng = np.random.default_rng(seed)
        self.model = load_model()
        _, _, self.x_init = load_problem(self.model)
        self.wbc = load_wbc(self.model)
        self.B = B
        self.mode = rng.integers(0, 16, B).astype(np.int32)
        q = np.tile(self.x_init[6:], (B, 1))
        q[:, 0:2] += rng.uniform(-0.05, 0.05, (B, 2))
        q[:, 2] += rng.uniform(-0.02, 0.02, B)
        q[:, 3:6] += rng.uniform(-0.1, 0.1, (B, 3))
        q[:, 6:] += rng.uniform(-0.1, 0.1, (B, 18))
        v = rng.uniform(-vel, vel, (B, 24))
        self.rbd = rbd_from_state(q, v)
        self.x_des = np.tile(self.x_init, (B, 1))
        self.x_des[:, 0:6] += rng.uniform(-0.05, 0.05, (B, 6))
        self.x_des[:, 6:] += rng.uniform(-0.02, 0.02, (B, 24))
        self.u_des = np.zeros((B, 30))
        m, g = self.model.total_mass, 9.81
        for b in range(B):
            md = int(self.mode[b])
            ns = bin(md).count("1")
            for leg in range(4):
                if (md >> (3 - leg)) & 1:
                    self.u_des[b, 3 * leg + 2] = m * g / ns
        self.u_des[:, 12:] += rng.uniform(-0.1, 0.1, (B, 18))
        self.u_last = self.u_des + rng.uniform(-1e-3, 1e-3, (B, 30))   # previous MPC input (inputLast_), applied as a warm-up call
        self.period = np.full(B, period)
        self.time = np.full(B, time)
