"""Seeded synthetic workloads of SURVEY.md §8(d) (TEST INFRASTRUCTURE; shared by tests and bench.py)."""
import numpy as np

from . import centroidal as ce
from . import gait as G


def standing_target(model, P, t0=0.0, tf=1e3):
    """2-knot, 37-dim standing reference: x_ref = initialState, EE pose = FK(initialState) (config 1/2)."""
    nk = ce.node_kinematics(model, P.x_init)
    quat = ce.quat_from_matrix(nk["ee_rot"])
    s = np.concatenate([P.x_init, nk["ee_pos"], quat])
    return np.array([t0, tf]), np.stack([s, s])


def perturbed_states(model, P, B, seed=20261017):
    """Config-2 initial states: initialState + delta."""
    rng = np.random.default_rng(seed)
    x0 = np.tile(P.x_init, (B, 1))
    x0[:, 0:6] += rng.uniform(-0.1, 0.1, (B, 6))
    x0[:, 6:8] += rng.uniform(-0.05, 0.05, (B, 2))
    x0[:, 8] += rng.uniform(-0.02, 0.02, B)
    x0[:, 9:12] += rng.uniform(-0.1, 0.1, (B, 3))
    x0[:, 12:30] += rng.uniform(-0.1, 0.1, (B, 18))
    x0[:, 12:30] = np.clip(x0[:, 12:30], model.lower[6:] + 1e-3, model.upper[6:] - 1e-3)
    phase = rng.uniform(0.0, 0.7, B)
    return x0, phase
