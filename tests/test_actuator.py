"""Control law + simulated actuator with transport delay: CPU port and CUDA kernel against the oracle's
deque restatement of QMController::updateControlLaw (QMController.cpp:178-191) and QMHWSim::writeSim (QMHWSim.cpp:98-114).
Stamps are integer nanoseconds and the buffered command that is picked must agree exactly, so the torques are compared
bit for bit up to the FMA contraction of kp * dq + kd * dv + ff (1e-13 relative)."""
import numpy as np
import pytest

from oracle import abi_fill, actuator as oa


def tick_schedule(n_ticks, seed):
    """Per-problem clocks: a 1 ms control period with jitter-free stamps, one problem crossing the 10 s leg-enable time, one
    whose tick equals the period (the reference's simulation-reset test), one with a longer period than the delay."""
    rng = np.random.default_rng(seed)
    B = 6
    period_ns = 1_000_000
    start_ns = np.array([1, 9_990, 5_000, 20_000, 1, 123_456], dtype=np.int64) * period_ns
    step_ns = np.array([1, 1, 1, 1, 12, 3], dtype=np.int64) * period_ns          # problem 4: every command is older than the delay
    ticks = []
    for k in range(n_ticks):
        time_ns = start_ns + k * step_ns
        obs = time_ns / 1e9
        ticks.append(dict(time_ns=time_ns, obs_time=obs, x=rng.normal(size=(B, 30)), u=rng.normal(size=(B, 30)),
                          cmd=10.0 * rng.normal(size=(B, 54)), q=rng.normal(size=(B, 18)), v=rng.normal(size=(B, 18))))
    return B, period_ns, ticks


def run_oracle(B, period_ns, ticks, **kw):
    acts = [oa.Actuator(**kw) for _ in range(B)]
    return [np.array([acts[b].step(int(t["time_ns"][b]), period_ns, float(t["obs_time"][b]), t["x"][b], t["u"][b], t["cmd"][b],
                                   t["q"][b], t["v"][b]) for b in range(B)]) for t in ticks]


def check(step, B, period_ns, ticks, ref):
    for t, r in zip(ticks, ref):
        tau, status = step(t["time_ns"], period_ns, t["obs_time"], t["x"], t["u"], t["cmd"], t["q"], t["v"])
        assert (status == 0).all()
        assert np.allclose(tau, r, rtol=1e-13, atol=1e-13)


def test_oracle_delay_semantics():
    """The command applied at time t is the oldest one not older than the delay: with a 1 ms period and 9 ms delay the torque
    follows the feed-forward command of 9 ticks ago; before 10 s the legs hold zero commands."""
    a = oa.Actuator()
    q = v = np.zeros(18)
    out = []
    for k in range(1, 30):
        cmd = np.zeros(54); cmd[36:] = k
        out.append(a.step(k * 1_000_000, 1_000_000, 11.0, np.zeros(30), np.zeros(30), cmd, q, v))
    out = np.array(out)
    assert np.array_equal(out[:9, 0], np.ones(9)) and np.array_equal(out[9:, 0], np.arange(1, 21))
    b = oa.Actuator()
    cmd = np.zeros(54); cmd[36:] = 5.0
    tau = b.step(1_000_000, 1_000_000, 3.0, np.zeros(30), np.zeros(30), cmd, q, v)
    assert np.array_equal(tau[:12], np.zeros(12)) and np.array_equal(tau[12:], 5.0 * np.ones(6))


def test_cport_matches_oracle():
    from qm_door_b200 import _abi
    B, period_ns, ticks = tick_schedule(60, seed=3)
    d = _abi.ActuatorDesc(0.0, 3.0, 0.0, 0.5, 10.0, 9_000_000)
    cp = abi_fill.CPortActuator(d, B)
    check(cp.step, B, period_ns, ticks, run_oracle(B, period_ns, ticks))
    # non-default gains and delay
    d2 = _abi.ActuatorDesc(20.0, 1.5, 35.0, 0.7, 9.995, 2_500_000)
    cp2 = abi_fill.CPortActuator(d2, B)
    check(cp2.step, B, period_ns, ticks,
          run_oracle(B, period_ns, ticks, leg_kp=20.0, leg_kd=1.5, arm_kp=35.0, arm_kd=0.7, leg_enable_time=9.995, delay_ns=2_500_000))


def test_cport_overflow_is_flagged():
    from qm_door_b200 import _abi
    d = _abi.ActuatorDesc(0.0, 3.0, 0.0, 0.5, 10.0, 9_000_000)
    cp = abi_fill.CPortActuator(d, 1)
    z = lambda *s: np.zeros(s)
    st = 0
    for k in range(2, 60):                                  # 0.1 ms ticks: more than 32 commands younger than 9 ms
        _, status = cp.step(np.array([k * 100_000]), 100_000, z(1), z(1, 30), z(1, 30), z(1, 54), z(1, 18), z(1, 18))
        st |= int(status[0])
    assert st == 64


@pytest.mark.gpu
def test_cuda_matches_oracle_and_reset(descs):
    import torch
    import qm_door_b200 as q
    from qm_door_b200 import workload
    B, period_ns, ticks = tick_schedule(60, seed=4)
    W = workload.WbcWorkload(B, seed=1)
    ctx = q.WbcContext(W.model, W.wbc, B)
    d = q.actuator_defaults()
    assert (d.leg_kd, d.arm_kd, d.leg_enable_time, d.delay_ns) == (3.0, 0.5, 10.0, 9_000_000)
    ref = run_oracle(B, period_ns, ticks)
    check(lambda *a: ctx.actuator(d, *a), B, period_ns, ticks, ref)
    ctx.actuator_reset()                                     # same ticks again from a clean buffer: same torques
    check(lambda *a: ctx.actuator(d, *a), B, period_ns, ticks, ref)
    # device-pointer entry on the context's stream
    ctx.actuator_reset()
    dev = torch.device("cuda", 0)
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    tau = torch.zeros(B, 18, dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    for t, r in zip(ticks, ref):
        ctx.actuator_dev(d, T(t["time_ns"]), period_ns, T(t["obs_time"]), T(t["x"]), T(t["u"]), T(t["cmd"]), T(t["q"]), T(t["v"]), tau, st)
        ctx.sync()
        assert np.allclose(tau.cpu().numpy(), r, rtol=1e-13, atol=1e-13)
    ctx.close()
