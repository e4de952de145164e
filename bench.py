#!/usr/bin/env python
"""bench.py — MPC-cycles/s of the qm_door hot path on B200 (BASELINE.json metric, config 2).

A "step" is one full SQP MPC cycle (schedule -> transcription -> Riccati -> filter line search -> policy) of the whole
batch: AlienGo+Z1, horizon 1.0 s at dt 0.01 s (N = 100 intervals + gait-event nodes), trot, 1024 independent perturbed
initial states per GPU, warm-started and advanced by 0.01 s per step.

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [...]                    the CPU implementation of the same path on the host cores
                                                            (oracle/cport: the reference binary cannot be built here)
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MPC-cycles/sec (N=100, 24-state, batch=1024) at 1/2/4/8 B200 vs CPU ref"
UNIT = "MPC-cycles/s"
BATCH_PER_GPU = 1024
HORIZON, DT = 1.0, 0.01
CYCLE_DT = 0.01


GAIT_LIBRARY = ["trot", "standing_trot", "flying_trot", "pace", "standing_pace", "dynamic_walk", "static_walk", "amble",
                "lindyhop", "skipping", "pawup"]          # the 11 moving gaits of gait.info (all but stance)
CONFIG4_SEEDS = 512


def workload_config(n_gpus, mode="config2", scaling="weak", batch=BATCH_PER_GPU):
    par = "independent problems sharded across %d GPU(s)%s" % (
        n_gpus, ", one NCCL all-gather of the policy per cycle (qmb200_allgather_policy, beside the next cycle)" if n_gpus > 1 else "")
    l2 = "working set per step (6.4 MB of LQ and kinematics blocks per problem) is far larger than the 126 MB L2; no explicit flush"
    if mode == "config4":
        return {"workload": "config 4: gait library, 11 gait schedules x 512 disturbance seeds = 5632 problems (config-2 perturbation + "
                            "base momentum kick U(-0.3, 0.3)), N=100, sharded over the GPUs by interleaved (seed, gait) index, warm start, "
                            "t0 += 0.01 s per step, 1 SQP iteration + filter line search per cycle",
                "batch_total": len(GAIT_LIBRARY) * CONFIG4_SEEDS, "batch_per_gpu": batch, "horizon_s": HORIZON, "dt_s": DT,
                "gait": "library of 11", "parallelism": par, "l2_policy": l2}
    w = "config 2: AlienGo+Z1 nx=30 nu=30, horizon 1.0 s / dt 0.01 s (N=100 + gait-event nodes), trot, "
    if scaling == "strong":
        w += "batch=1024 independent perturbed initial states IN TOTAL (%d per GPU), " % batch
    else:
        w += "batch=1024 independent perturbed initial states per GPU, "
    w += "warm start, t0 += 0.01 s per step, 1 SQP iteration + filter line search per cycle"
    return {"workload": w, "batch_per_gpu": batch, "horizon_s": HORIZON, "dt_s": DT, "gait": "trot", "parallelism": par, "l2_policy": l2}


def make_workload(args, rank, world, total_steps):
    """The rank's shard of the synthetic workload. config2 weak: 1024 problems per GPU (seed + rank);
    config2 strong: 1024 problems in total; config4: 5632 problems in total (11 gaits x 512 seeds)."""
    import qm_door_b200 as q
    from qm_door_b200 import distributed as D, workload
    t_span = CYCLE_DT * (2 * total_steps + 4)
    if args.workload == "config4":
        total = len(GAIT_LIBRARY) * CONFIG4_SEEDS
        lo, hi = D.shard_range(total, rank, world)
        W = workload.Workload(total, horizon=HORIZON, dt=DT, seed=20261019, t_span=t_span, max_events=64, max_nodes=int(round(HORIZON / DT)) + 1 + 32)
        rng = np.random.default_rng(20261019)
        W.x0[:, 0:6] += rng.uniform(-0.3, 0.3, (total, 6))                    # external base momentum kick
        gaits = [q.load_gait(g) for g in GAIT_LIBRARY]
        for b in range(lo, hi):
            sw, md = gaits[b % len(GAIT_LIBRARY)]                             # interleaved: every rank sees a mix of gaits
            W.events[b], W.modes[b], W.nevents[b] = q.tile_schedule(sw, md, -np.ceil(HORIZON / sw[-1]) * sw[-1] - W.phase[b],
                                                                    t_span + 2.0 * HORIZON, 64)
        for k in ("x0", "events", "modes", "nevents", "target_t", "target_x", "phase"):
            setattr(W, k, np.ascontiguousarray(getattr(W, k)[lo:hi]))
        W.B = hi - lo
        return W
    if args.scaling == "strong":
        lo, hi = D.shard_range(BATCH_PER_GPU, rank, world)
        W = workload.Workload(BATCH_PER_GPU, horizon=HORIZON, dt=DT, seed=20261017, t_span=t_span)
        for k in ("x0", "events", "modes", "nevents", "target_t", "target_x", "phase"):
            setattr(W, k, np.ascontiguousarray(getattr(W, k)[lo:hi]))
        W.B = hi - lo
        return W
    return workload.Workload(BATCH_PER_GPU, horizon=HORIZON, dt=DT, seed=20261017 + rank, t_span=t_span)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- algorithmic bytes
def kernel_bytes_per_node(kernel, nut):
    """Bytes a kernel must move per intermediate node with reduced input dimension nut (nv = 26 - nut velocity rows),
    as laid out in DESIGN.md section 2.
    stage = projected LQ block actually used (A 900, B 30 nut, b 30, q 30, r nut, 1 | Q 900, P 30 nut, R nut^2);
    proj  = compact projection block (pivot rows of Px / Pu, Pe, roles); gain = (K 30 nut, kff nut);
    kin1 / kin2 = products of the two kinematics evaluations; piv = pivot-block inverse and index record (k_proj)."""
    nv = 26 - nut
    fwd = 900 + 30 * nut + 30 + 30 + nut + 1
    stage = fwd + 900 + 30 * nut + nut * nut
    proj = 30 * nv + nut * nv + nv + 12 + 16
    gain = 30 * nut + nut
    kin1 = 540 + 30 + 30 + 49 * nv + 144 + 8 + 144
    kin2 = 540 + 30
    piv = nv * nv + 40
    d = {"k_kin1": 60 + kin1, "k_kin2": 60 + kin2, "k_proj": 18 * nv + piv, "k_lq": 90 + kin1 + kin2 + piv + stage + proj + 3,
         "k_solve": stage + gain + (fwd + proj + gain) + 60, "k_trial": 150 + 3}[kernel]
    return 8 * d


def cycle_bytes_per_node(nut):
    """SURVEY.md §8(d) whole-cycle figure: every intermediate block that cannot stay on-chip across the forward-transcribe /
    backward-Riccati / forward-rollout dependency is written once and read once, the trajectories (x, u) go in and out once:
    2 x 4999 + 120 = 10 118 doubles = 80 944 B per node for trot (nut = 16)."""
    blocks = (900 + 30 * nut + 30) + (900 + nut * nut + 30 * nut + 30 + nut + 1) + (30 * nut + 900 + 30) + (30 * nut + nut)
    return 8 * (2 * blocks + 120)


# FP64 peak of this part for the dense small-matrix products: measured with tools/micro/dmma_bench.cu on the gpurun B200
# (mma.sync.m8n8k4.f64 = SASS DMMA, 37.1 TFLOP/s; plain DFMA 34.0; profiles/dmma_micro_r01.txt). Not in MEASURED_PEAKS.json.
FP64_PEAK_TFLOPS = 37.1


def kernel_flops_per_node(kernel, nut):
    """Algorithmic FP64 flops (2 per multiply-add) of the dense products a kernel carries out per intermediate node; the
    symmetric products are counted in full (what the recursion defines), index / barrier / reference arithmetic is not counted.
    k_solve: S A, S B, S b | G = R + B'SB, g | G^-1 (Gauss-Jordan, 2 nut^3) | H = P + B'SA, Q + A'SA, A'sb | K = -G^-1 H, kff |
             S += H'K, s += H'kff | forward: K dx, A dx + B dut, projection rows.
    k_lq:    Dinv T | Heun sensitivities (9x9x60 products) | Gauss-Newton end-effector Hessian | change of variables
             (A + B Px, B Pu, R Px, R Pu, Q + Px'R Px, Pu'R Px, Pu'R Pu) with nv = 26 - nut pivot rows."""
    nv = 26 - nut
    if kernel == "k_solve":
        bwd = 2 * (27000 + 900 * nut + 900) + 2 * (30 * nut * nut + 30 * nut) + 2 * nut ** 3 + 2 * (900 * nut + 27000 + 900) \
            + 2 * (30 * nut * nut + nut * nut) + 2 * (900 * nut + 30 * nut)
        fwd = 2 * (30 * nut + 900 + 30 * nut + nv * (30 + nut) + 30 + nut)
        return bwd + fwd
    if kernel == "k_lq":
        nsel = nv + nut
        heun = 2 * (9 * 9 * 60) + 4 * 9 * 60
        cost = 2 * 6 * 24 * 24 + 2 * 2 * 900
        cov = 2 * nv * (900 + 30 * nsel + 30 * nut + nsel * nut + 900 + 30 * nut + nut * nut) + 2 * nv * nv * 49 + 2 * 30 * (nsel + 30)
        return heun + cost + cov
    return None


# ----------------------------------------------------------------------------- CPU arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import abi_fill
    from qm_door_b200 import workload
    cores = os.cpu_count() or 1
    sample = 128
    W = workload.Workload(sample, horizon=HORIZON, dt=DT, t_span=CYCLE_DT * (args.steps + args.warmup + 2))
    cp = abi_fill.CPort(W.model, W.problem, W.solver, sample, threads=cores)
    step = 0
    for _ in range(args.warmup):
        cp.cycle(np.full(sample, CYCLE_DT * step), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
        step += 1
    t = time.perf_counter()
    for _ in range(args.steps):
        cp.cycle(np.full(sample, CYCLE_DT * step), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
        step += 1
    el = time.perf_counter() - t
    cp.close()
    value = sample * args.steps / el
    latency = cpu_latency_legs()
    desc = "%d problems of the same workload per step (1/8 of one GPU batch), %d steps, all %d host threads" % (sample, args.steps, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "note": "CPU restatement (oracle/cport, g++ -O3, std::thread over problems) — not the reference "
                                 "binary: OCS2/Pinocchio/HPIPM are not vendored and cannot be built in this image"},
        "latency": latency,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


LATENCY_SETTINGS = (("reference setting: horizon 1.0 s / dt 0.015 s (task.info:79,141), N~67", 1.0, 0.015),
                    ("config 1: horizon 0.2 s / dt 0.01 s, N=20", 0.2, 0.01),
                    ("config 2 shape: horizon 1.0 s / dt 0.01 s, N=100", 1.0, 0.01))


def cpu_latency_legs(node_threads=3, cycles=20, warm=3):
    """Reference-like latency: ONE problem, 3 worker threads over the nodes (sqp.nThreads,
    task.info:78), warm-started cycles, trot; median and min wall time per MPC cycle of the CPU port."""
    from oracle import abi_fill
    from qm_door_b200 import workload
    legs = []
    for name, hor, dt in LATENCY_SETTINGS:
        W = workload.Workload(1, horizon=hor, dt=dt, t_span=CYCLE_DT * (cycles + warm + 2))
        cp = abi_fill.CPort(W.model, W.problem, W.solver, 1, threads=1, node_threads=node_threads)
        ts = []
        for c in range(cycles + warm):
            t = time.perf_counter()
            out = cp.cycle(np.full(1, CYCLE_DT * c), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
            ts.append(time.perf_counter() - t)
        cp.close()
        legs.append({"setting": name, "problems": 1, "node_threads": node_threads, "nodes": int(out["n"][0]),
                     "ms_per_cycle_median": 1e3 * float(np.median(ts[warm:])), "ms_per_cycle_min": 1e3 * float(min(ts[warm:])),
                     "cycles_per_s": 1.0 / float(np.median(ts[warm:]))})
    return legs


def gpu_latency_legs(q, workload, device, cycles=20, warm=3):
    """The same single-problem settings through qmb200_mpc_cycle_batch (host buffers, B = 1): what one robot would see."""
    legs = []
    for name, hor, dt in LATENCY_SETTINGS:
        W = workload.Workload(1, horizon=hor, dt=dt, t_span=CYCLE_DT * (cycles + warm + 2))
        ctx = q.MpcContext(W.model, W.problem, W.solver, 1, device=device)
        out = ctx.alloc_outputs(pinned=True)
        ts = []
        for c in range(cycles + warm):
            t = time.perf_counter()
            ctx.cycle(np.full(1, CYCLE_DT * c), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x, out=out)
            ts.append(time.perf_counter() - t)
        ctx.close()
        legs.append({"setting": name, "problems": 1, "nodes": int(out["n"][0]), "ms_per_cycle_median": 1e3 * float(np.median(ts[warm:])),
                     "ms_per_cycle_min": 1e3 * float(min(ts[warm:])), "cycles_per_s": 1.0 / float(np.median(ts[warm:]))})
    return legs


def wbc_cpu_baseline(WW, sub=4096):
    """WBC on the CPU port beside the GPU number: single-solve latency on one thread and
    throughput on all host threads over the first `sub` solves of the same batch."""
    from oracle import abi_fill
    cores = os.cpu_count() or 1
    sl = slice(0, sub)
    args = (WW.model, WW.wbc, WW.x_des[sl], WW.u_des[sl], WW.rbd[sl], WW.mode[sl], WW.period[sl], WW.time[sl])
    abi_fill.cport_wbc(*args, WW.u_last[sl].copy(), threads=cores)                      # warm-up (pages, thread start)
    t = time.perf_counter()
    abi_fill.cport_wbc(*args, WW.u_last[sl].copy(), threads=cores)
    thr = sub / (time.perf_counter() - t)
    one = tuple(a[:64] if isinstance(a, np.ndarray) else a for a in args)
    t = time.perf_counter()
    abi_fill.cport_wbc(*one, WW.u_last[:64].copy(), threads=1)
    lat = (time.perf_counter() - t) / 64
    return {"value": thr, "unit": "WBC-solves/s", "cores": cores, "kind": "port", "single_solve_latency_ms_1_thread": 1e3 * lat,
            "sample": "first %d solves of the same batch on %d threads; latency: 64 solves back to back on 1 thread (oracle/cport)" % (sub, cores)}


def cpu_baseline_sample():
    """Bounded CPU-port sample for the GPU arm's JSON line (rank 0, N=1): 128 problems x (1 warm-up + 3 cycles)."""
    from oracle import abi_fill
    from qm_door_b200 import workload
    cores = os.cpu_count() or 1
    sample, steps = 128, 3
    W = workload.Workload(sample, horizon=HORIZON, dt=DT)
    cp = abi_fill.CPort(W.model, W.problem, W.solver, sample, threads=cores)
    cp.cycle(np.zeros(sample), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
    t = time.perf_counter()
    for s in range(1, steps + 1):
        cp.cycle(np.full(sample, CYCLE_DT * s), W.x0, W.events, W.modes, W.nevents, W.target_t, W.target_x)
    el = time.perf_counter() - t
    cp.close()
    return {"value": sample * steps / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d problems x %d warm-started cycles of config 2 on %d threads (oracle/cport)" % (sample, steps, cores)}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, world, local_rank, result_fd=None):
    import torch
    import torch.distributed as dist
    import qm_door_b200 as q
    from qm_door_b200 import distributed as D, workload

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    total_steps = 3 * (args.steps + args.warmup)             # timed + profiled pass + the two end-to-end loops share the schedule
    W = make_workload(args, rank, world, total_steps)
    B = W.B
    ctx = q.MpcContext(W.model, W.problem, W.solver, B, device=local_rank)
    NMAX = W.solver.max_nodes
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    dev = torch.device("cuda", local_rank)
    f64, i32 = torch.float64, torch.int32
    d = dict(t0=torch.zeros(B, dtype=f64, device=dev), x0=torch.from_numpy(W.x0).to(dev),
             events=torch.from_numpy(W.events).to(dev), modes=torch.from_numpy(W.modes).to(dev),
             nevents=torch.from_numpy(W.nevents).to(dev), tt=torch.from_numpy(W.target_t).to(dev),
             tx=torch.from_numpy(W.target_x).to(dev))
    o = dict(t=torch.zeros(B, NMAX, dtype=f64, device=dev), x=torch.zeros(B, NMAX, 30, dtype=f64, device=dev),
             u=torch.zeros(B, NMAX, 30, dtype=f64, device=dev), n=torch.zeros(B, dtype=i32, device=dev),
             mode=torch.zeros(B, NMAX, dtype=i32, device=dev), info=torch.zeros(B, q.INFO_SIZE, dtype=f64, device=dev),
             status=torch.zeros(B, dtype=i32, device=dev))
    # multi-GPU: the library's own collective (qmb200_allgather_policy: k_finalize writes the packed send buffer, ncclAllGather on
    # the context's communication stream beside the next cycle); two receive buffers, consumed one cycle later
    gathered = [torch.zeros(world, B, NMAX, 61, dtype=f64, device=dev) for _ in range(2)] if world > 1 else None
    if world > 1:
        D.init_comm(ctx)
    torch.cuda.synchronize()

    def step_dev(k):
        with torch.cuda.stream(stream):
            d["t0"].fill_(CYCLE_DT * k)
            ctx.cycle_dev(d["t0"], d["x0"], d["events"], d["modes"], d["nevents"], d["tt"], d["tx"], o["t"], o["x"], o["u"],
                          o["n"], o["mode"], o["info"], o["status"])
            if world > 1:
                ctx.allgather_policy(gathered[k & 1])

    def barrier():
        if world > 1:
            ctx.comm_sync()                     # the last all-gather belongs to the timed region
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value): the cycle is replayed as a CUDA graph, no per-kernel events in this region
    k = 0
    for _ in range(args.warmup):
        step_dev(k)
        k += 1
    ctx.sync()
    ctx.kernel_times(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(args.steps):
        step_dev(k)
        k += 1
    if world > 1:
        ctx.policy_wait_stream(ctx.stream)      # the last all-gather ends inside the timed region
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    kt_timed = ctx.kernel_times(reset=True)                 # launch counts of the timed region (graph replays included)
    # ---- second pass of the same K steps with per-kernel CUDA events on the launching streams (direct launches: events cannot
    #      bracket the kernels of a graph replay): the per-kernel durations behind `roofline` and `kernel_ms_per_step`
    ctx.set_profiling(True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        p0.record(stream)
    for _ in range(args.steps):
        step_dev(k)
        k += 1
    if world > 1:
        ctx.policy_wait_stream(ctx.stream)
    with torch.cuda.stream(stream):
        p1.record(stream)
    barrier()
    prof_ms = p0.elapsed_time(p1)
    clocks = sampler.stop() if rank == 0 else None
    kt = ctx.kernel_times(reset=True)
    ctx.set_profiling(False)
    status = o["status"].cpu().numpy()
    bad = int(((status & ~32) != 0).sum())
    nn = o["n"].cpu().numpy()
    modes_last = o["mode"].cpu().numpy()
    alpha_mean = float(o["info"][:, 0].mean().item())

    # ---- end-to-end timing through the host-buffer C-ABI call (pinned host inputs, policy read back)
    ctx.reset()
    ctx.sync()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h = dict(x0=pin(W.x0), events=pin(W.events), modes=pin(W.modes), nevents=pin(W.nevents), tt=pin(W.target_t), tx=pin(W.target_x))
    ht0 = pin(np.zeros(B))
    hout = ctx.alloc_outputs(pinned=True)
    h2d = sum(a.nbytes for a in h.values()) + ht0.nbytes
    d2h = sum(hout[key].nbytes for key in ("t", "x", "u", "n", "mode", "info", "status"))
    hout2 = ctx.alloc_outputs(pinned=True)
    houts, ht0s = (hout, hout2), (ht0, pin(np.zeros(B)))
    k = 0
    for _ in range(args.warmup):
        ht0[:] = CYCLE_DT * k
        ctx.cycle(ht0, h["x0"], h["events"], h["modes"], h["nevents"], h["tt"], h["tx"], out=hout)
        k += 1
    # (a) one blocking call per step (qmb200_mpc_cycle_batch): copy-out fully exposed
    barrier()
    t_start = time.perf_counter()
    for _ in range(args.steps):
        ht0[:] = CYCLE_DT * k
        ctx.cycle(ht0, h["x0"], h["events"], h["modes"], h["nevents"], h["tt"], h["tx"], out=hout)
        k += 1
    torch.cuda.synchronize()
    e2e_serial_s = time.perf_counter() - t_start
    # (b) submit / wait (qmb200_mpc_cycle_batch_async + _wait), two host buffer sets: the policy of cycle k is copied out and read
    #     on the host while cycle k + 1 computes (the reference's MPC thread / policy buffer arrangement). Every step's inputs
    #     come from pinned host memory and every step's result is read on the host inside the timed region.
    e2e_bad = 0
    barrier()
    t_start = time.perf_counter()
    prev = None
    for i in range(args.steps):
        ht0s[i & 1][:] = CYCLE_DT * k
        tk = ctx.cycle_async(ht0s[i & 1], h["x0"], h["events"], h["modes"], h["nevents"], h["tt"], h["tx"], out=houts[i & 1])
        k += 1
        if prev is not None:
            ctx.wait(prev)
            e2e_bad += int(((houts[(i - 1) & 1]["status"] & ~32) != 0).sum())
        prev = tk
    ctx.wait(prev)
    e2e_bad += int(((houts[(args.steps - 1) & 1]["status"] & ~32) != 0).sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t_start

    # ---- secondary metric (BASELINE config 5): whole-body-control solves/s, B = 65 536 contact configurations, rank 0 only
    wbc_line = None
    if rank == 0 and not args.no_wbc:
        WW = workload.WbcWorkload(65536)
        wctx = q.WbcContext(WW.model, WW.wbc, WW.B, device=local_rank)
        wstream = torch.cuda.ExternalStream(wctx.stream, device=local_rank)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        wd = dict(x=T(WW.x_des), u=T(WW.u_des), ul=T(WW.u_last), r=T(WW.rbd), m=T(WW.mode), p=T(WW.period), t=T(WW.time))
        wcmd = torch.zeros(WW.B, 54, dtype=f64, device=dev)
        wst = torch.zeros(WW.B, dtype=i32, device=dev)
        torch.cuda.synchronize()
        for _ in range(2):
            wctx.update_dev(wd["x"], wd["ul"], wd["r"], wd["m"], wd["p"], wd["t"], wcmd, wst)
        wctx.sync()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        with torch.cuda.stream(wstream):
            w0.record(wstream)
            for i in range(reps):
                wctx.update_dev(wd["x"], wd["u"] if i % 2 == 0 else wd["ul"], wd["r"], wd["m"], wd["p"], wd["t"], wcmd, wst)
            w1.record(wstream)
        wctx.sync()
        torch.cuda.synchronize()
        wms = w0.elapsed_time(w1) / reps
        wbad = int(((wst.cpu().numpy() & ~2) != 0).sum())
        wbc_line = {"metric": "WBC-solves/s (HierarchicalWbc 3-level stack, batch=65536 contact configurations)",
                    "value": WW.B / (wms * 1e-3), "unit": "WBC-solves/s", "ms_per_batch": wms, "batch": WW.B,
                    "failed_solves": wbad,
                    "kernel": "k_wbc_order, k_wbc_tasks, k_wbc_level0 x 2, k_wbc_level x 3, k_wbc_gi x 2 (QMB200_WBC_SPLIT=0: the single kernel k_wbc)",
                    "algorithmic_bytes_per_solve": 1376, "achieved_gbs": WW.B * 1376 / (wms * 1e-3) / 1e9,
                    "state_bytes_per_solve": 196000,
                    "note": "one kernel per phase of the solve, solves ordered by contact pattern, workspace image of a solve in global "
                            "memory between kernels (196 KB per solve read + written, ncu dram bytes, profiles/r02_wbc.md); bound by "
                            "dependency chains at 7 to 24 resident solves per SM, not by HBM"}
        wctx.close()
        if not args.no_cpu_baseline:
            wbc_line["cpu_baseline"] = wbc_cpu_baseline(WW)
        # the six-level split of the same tasks (BASELINE config 5 names "6 task levels"; SYNTHETIC: the reference's stacks have three)
        WW.wbc.mpc_variant = 2
        wctx = q.WbcContext(WW.model, WW.wbc, WW.B, device=local_rank)
        wstream = torch.cuda.ExternalStream(wctx.stream, device=local_rank)
        for _ in range(2):
            wctx.update_dev(wd["x"], wd["ul"], wd["r"], wd["m"], wd["p"], wd["t"], wcmd, wst)
        wctx.sync()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(wstream):
            s0.record(wstream)
            for i in range(3):
                wctx.update_dev(wd["x"], wd["u"] if i % 2 == 0 else wd["ul"], wd["r"], wd["m"], wd["p"], wd["t"], wcmd, wst)
            s1.record(wstream)
        wctx.sync()
        torch.cuda.synchronize()
        sms = s0.elapsed_time(s1) / 3
        stv = wst.cpu().numpy()
        wbc_line["six_level_synthetic"] = {"value": WW.B / (sms * 1e-3), "unit": "WBC-solves/s", "ms_per_batch": sms,
                                           "failed_solves": int(((stv & ~2) != 0).sum()), "flagged_degenerate": int((stv & 2 != 0).sum()),
                                           "stack": "[EoM + limits + contact + friction] -> [base height + angular] -> [EE linear + angular] -> "
                                                    "[100 x swing] -> [base xy] -> [contact force]; labelled synthetic (SURVEY 8(d))"}
        wctx.close()
        WW.wbc.mpc_variant = 0

    times = torch.tensor([dev_ms, e2e_s * 1e3, e2e_serial_s * 1e3, prof_ms], dtype=f64, device=dev)
    counts = torch.tensor([B], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, e2e_serial_ms, prof_ms = (float(v) for v in times.tolist())
    if rank == 0:
        total = int(counts.sum().item())
        value = total * args.steps / (dev_ms * 1e-3)
        e2e = total * args.steps / (e2e_ms * 1e-3)
        # dominant kernel roofline: algorithmic bytes of the nodes actually processed / mean launch time of that kernel
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        modeled = ("k_kin1", "k_kin2", "k_lq", "k_solve", "k_trial")
        dom = max(modeled, key=lambda kn: kt[kn][0])                 # dominant kernel by total time in the timed region
        ms_dom = kt[dom][0] / max(1, kt[dom][1])
        nodes_bytes = 0
        for b in range(B):
            for kk in range(nn[b] - 1):
                md = int(modes_last[b, kk])
                nodes_bytes += kernel_bytes_per_node(dom, 14 + bin(md & 15).count("1"))
        achieved = nodes_bytes / (ms_dom * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")       # DRAM bytes per launch from the committed ncu --set full capture
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        cyc_bytes = 0.0
        flops_dom = 0.0
        for b in range(B):
            for kk in range(nn[b] - 1):
                nut_k = 14 + bin(int(modes_last[b, kk]) & 15).count("1")
                cyc_bytes += cycle_bytes_per_node(nut_k)
                fl = kernel_flops_per_node(dom, nut_k)
                flops_dom += fl if fl is not None else 0.0
        fp64 = None
        if flops_dom > 0:
            fp64 = {"kernel": dom, "algorithmic_flops_per_launch": flops_dom, "achieved": flops_dom / (ms_dom * 1e-3) / 1e12,
                    "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": flops_dom / (ms_dom * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                    "peak_source": "measured DMMA micro-benchmark (tools/micro/dmma_bench.cu, profiles/dmma_micro_r01.txt); "
                                   "not in MEASURED_PEAKS.json"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (args.scaling == "strong" or args.workload == "config4") else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, args.workload, args.scaling, B), "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps, "failed_problems": e2e_bad,
                    "api": "qmb200_mpc_cycle_batch_async + qmb200_mpc_cycle_wait, two pinned host buffer sets: copy-out of cycle k "
                           "overlaps cycle k + 1; every step's inputs come from the host and every result is read on the host",
                    "blocking_call": {"value": total * args.steps / (e2e_serial_ms * 1e-3), "ms_per_step": e2e_serial_ms / args.steps,
                                      "api": "qmb200_mpc_cycle_batch (submit + wait per step, copy-out exposed)"}},
            "gpu_launches": int(sum(v[1] for v in kt_timed.values())),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "ms_per_launch": ms_dom, "algorithmic_bytes_per_launch": int(nodes_bytes),
                         "share_of_step": kt[dom][0] / max(1e-9, prof_ms),     # of the profiled pass (k_proj overlaps k_kin2)
                         "pass": "second pass of the same K steps with per-kernel CUDA events on the launching streams "
                                 "(%.3f ms/step; the timed region replays the cycle as a CUDA graph, which events cannot bracket)" % (prof_ms / args.steps),
                         "fp64": fp64,
                         "cycle_level": {"algorithmic_bytes_per_step": int(cyc_bytes),
                                         "achieved_gbs": cyc_bytes / (dev_ms / args.steps * 1e-3) / 1e9,
                                         "frac": cyc_bytes / (dev_ms / args.steps * 1e-3) / 1e9 / peak,
                                         "note": "SURVEY.md 8(d): 80 944 B per node (trot; written once + read once) x nodes of the batch; the cycle is "
                                                 "FP64-FMA / dependency bound (Riccati, projection), not HBM bound"}},
            "kernel_ms_per_step": {kname: v[0] / args.steps for kname, v in kt.items() if v[1]},
            "failed_problems": bad, "mean_step_size": alpha_mean,
            "nodes_per_problem": [int(nn.min()), int(nn.max())],
        }
        if wbc_line is not None:
            line["secondary"] = wbc_line
        if world == 1 and not args.no_cpu_baseline and args.workload == "config2":
            line["cpu_baseline"] = cpu_baseline_sample()
        if world == 1 and not args.no_latency:
            line["latency"] = gpu_latency_legs(q, workload, local_rank)
        if result_fd is not None:
            os.write(result_fd, (json.dumps(line) + "\n").encode())     # the process's real stdout (see main)
        else:
            print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-wbc", action="store_true", help="skip the secondary WBC-solves/s measurement")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-problem latency legs")
    ap.add_argument("--workload", default="config2", choices=["config2", "config4"],
                    help="config2 (default, BASELINE metric): trot, 1024 problems; config4: 11 gaits x 512 seeds = 5632 problems in total")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="config2 only: weak = 1024 problems per GPU (default), strong = 1024 problems in total")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__), "--gpus",
               str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", args.workload, "--scaling", args.scaling] + (
                   ["--no-cpu-baseline"] if args.no_cpu_baseline else []) + (["--no-wbc"] if args.no_wbc else []) + (["--no-latency"] if args.no_latency else [])
        raise SystemExit(subprocess.call(cmd))
    # stdout carries exactly one JSON line: libraries that write to file descriptor 1 while the bench runs (NCCL prints its
    # version banner there) are sent to stderr, the descriptor is restored for the result line
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        run_gpu(args, rank, world, local_rank, result_fd=saved)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


if __name__ == "__main__":
    main()
