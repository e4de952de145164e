"""Forward-dynamics step behind the actuator (include/qmb200.h qmb200_forward_dynamics_batch): the
labelled stand-in for the simulator that closes the reference's loop (Gazebo, qm_gazebo/src/QMHWSim.cpp:98-114). CPU port and
CUDA path against the NumPy oracle (oracle/sim.py), physical properties, and the device-resident closed loop
WBC -> control law / actuator -> forward dynamics."""
import numpy as np
import pytest

from helpers import rel_l2


def scenario(n, seed=5, vel=0.2):
    from qm_door_b200 import workload
    W = workload.WbcWorkload(n, seed=seed, vel=vel)
    W.mode[:16] = np.arange(16)
    rng = np.random.default_rng(seed)
    tau = rng.uniform(-20.0, 20.0, (n, 18))
    return W, tau


def test_cport_step_matches_oracle_and_holds_the_contacts(descs, oracle_inputs):
    from oracle import abi_fill, sim, gait as G
    from oracle.wbc import Wbc
    m, P = oracle_inputs
    W, tau = scenario(24)
    dt = 0.001
    for beta in (0.0, 0.5):
        nxt, f, st = abi_fill.cport_forward_dynamics(W.model, 9.81, W.rbd, tau, W.mode, dt, beta)
        assert (st == 0).all()
        for b in range(24):
            ref, fr, qdd = sim.forward_dynamics_step(m, P, W.rbd[b], tau[b], int(W.mode[b]), dt, beta)
            assert rel_l2(nxt[b], ref) < 1e-10 and np.abs(f[b] - fr).max() < 1e-7 * max(1.0, np.abs(fr).max()), (b, int(W.mode[b]))
            flags = G.stance_legs(int(W.mode[b]))
            for leg in range(4):
                if not flags[leg]:
                    assert not f[b, 3 * leg:3 * leg + 3].any()          # swing feet carry no force
            if beta == 0.0:
                # stance feet: J qdd + dJ v = 0 (the contact constraint at the acceleration level)
                s = Wbc(m, P).update_measured(W.rbd[b])
                acc = (s["J"] @ qdd + s["dJ"] @ s["v"]).reshape(4, 3)
                for leg in range(4):
                    if flags[leg]:
                        assert np.abs(acc[leg]).max() < 1e-8 * max(1.0, np.abs(qdd).max())


def test_free_fall_and_consistency_with_the_whole_body_controller(descs, oracle_inputs):
    """All feet in the air, zero torque, zero velocity: the base accelerates with -g. Nominal stance driven with the torques of
    the whole-body solution: the step reproduces the accelerations and contact forces the controller planned (both sides use
    the same equations of motion and contact constraints, so this ties M, h, J, dJ v of the two code paths together)."""
    from oracle import abi_fill, sqp, wbc
    from qm_door_b200 import workload
    m, P = oracle_inputs
    model = descs[0]
    x = P.x_init
    rbd0 = wbc.rbd_from_state(m, x, np.zeros(24))[None]
    dt = 1e-3
    nxt, f, _ = abi_fill.cport_forward_dynamics(model, 9.81, rbd0, np.zeros((1, 18)), np.array([0], dtype=np.int32), dt)
    assert abs(nxt[0, 29] + 9.81 * dt) < 1e-9 and not f.any()            # base linear velocity z after one step
    u = sqp.weight_compensating_input(m, 15)
    O = wbc.Wbc(m, P)
    O.input_last = u.copy()
    cmd = O.update(x, u, rbd0[0], 15, 0.002, 11.0)
    nxt, f, _ = abi_fill.cport_forward_dynamics(model, 9.81, rbd0, cmd[None, 36:], np.array([15], dtype=np.int32), dt)
    assert np.abs(nxt[0, 27:30] - dt * cmd[0:3]).max() < 1e-9 and np.abs(nxt[0, 30:48] - dt * cmd[6:24]).max() < 1e-9
    assert np.abs(f[0] - cmd[24:36]).max() < 1e-6 * np.abs(cmd[24:36]).max()
    fz = f.reshape(4, 3)[:, 2]
    assert (fz > 0).all() and abs(fz.sum() - m.total_mass * 9.81) < 0.05 * m.total_mass * 9.81


@pytest.mark.gpu
def test_cuda_step_matches_oracle_and_cpu_port(descs, oracle_inputs):
    import qm_door_b200 as q
    from oracle import abi_fill, sim
    m, P = oracle_inputs
    W, tau = scenario(64, seed=6)
    ctx = q.WbcContext(W.model, W.wbc, W.B)
    dt = 0.001
    nxt, f, st = ctx.forward_dynamics(W.rbd, tau, W.mode, dt, 0.2)
    ref, fr, _ = abi_fill.cport_forward_dynamics(W.model, 9.81, W.rbd, tau, W.mode, dt, 0.2, threads=8)
    assert (st == 0).all()
    assert rel_l2(nxt, ref) < 1e-11 and np.abs(f - fr).max() < 1e-7 * np.abs(fr).max()
    for b in range(0, 64, 7):
        o, fo, _ = sim.forward_dynamics_step(m, P, W.rbd[b], tau[b], int(W.mode[b]), dt, 0.2)
        assert rel_l2(nxt[b], o) < 1e-10
    ctx.close()


@pytest.mark.gpu
def test_device_resident_closed_loop_matches_host_calls(descs):
    """WBC -> control law / delayed actuator -> forward dynamics for several ticks with device pointers only (the state of tick k + 1
    is the output of tick k), against the same loop through the host-buffer entry points."""
    import torch
    import qm_door_b200 as q
    from qm_door_b200 import workload
    B, ticks, period = 16, 5, 0.002
    W = workload.WbcWorkload(B, seed=12, vel=0.05)
    W.mode[:] = 15
    ad = q.actuator_defaults()
    ctx = q.WbcContext(W.model, W.wbc, B)
    # host loop
    rbd = W.rbd.copy()
    hist = []
    for k in range(ticks):
        t_obs = 11.0 + k * period
        cmd, _ = ctx.update(W.x_des, W.u_des, rbd, W.mode, period, t_obs)
        tau, _ = ctx.actuator(ad, np.full(B, int((k + 2) * period * 1e9), dtype=np.int64), int(period * 1e9), np.full(B, t_obs), W.x_des, W.u_des,
                              cmd, rbd[:, 6:24], rbd[:, 30:48])
        rbd, f, st = ctx.forward_dynamics(rbd, tau, W.mode, period, 0.2)
        assert (st == 0).all()
        hist.append(rbd.copy())
    assert np.isfinite(hist[-1]).all() and np.abs(hist[-1][:, 5] - W.rbd[:, 5]).max() < 0.05      # the base height stays put over 10 ms
    # device loop
    ctx.reset(); ctx.actuator_reset()
    dev = torch.device("cuda", 0)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    f64, i32 = torch.float64, torch.int32
    xd, ud, md = T(W.x_des), T(W.u_des), T(W.mode)
    r = [T(W.rbd), torch.zeros(B, 55, dtype=f64, device=dev)]
    cmd_d, tau_d, f_d = torch.zeros(B, 54, dtype=f64, device=dev), torch.zeros(B, 18, dtype=f64, device=dev), torch.zeros(B, 12, dtype=f64, device=dev)
    st_d = torch.zeros(B, dtype=i32, device=dev)
    per_d = T(np.full(B, period))
    torch.cuda.synchronize()
    for k in range(ticks):
        t_obs = 11.0 + k * period
        cur, nxt = r[k & 1], r[(k + 1) & 1]
        tobs_d, tns_d = T(np.full(B, t_obs)), T(np.full(B, int((k + 2) * period * 1e9), dtype=np.int64))
        q_d, v_d = cur[:, 6:24].contiguous(), cur[:, 30:48].contiguous()
        torch.cuda.synchronize()
        ctx.update_dev(xd, ud, cur, md, per_d, tobs_d, cmd_d, st_d)
        ctx.actuator_dev(ad, tns_d, int(period * 1e9), tobs_d, xd, ud, cmd_d, q_d, v_d, tau_d, st_d)
        ctx.forward_dynamics_dev(cur, tau_d, md, period, 0.2, nxt, f_d, st_d)
        ctx.sync()
        assert np.array_equal(nxt.cpu().numpy(), hist[k]), k
    ctx.close()
