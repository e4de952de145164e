// sm_100a kernels + context + C-ABI of the MPC cycle (see include/qmb200.h for the reference interfaces replaced).
//
// Launch map of one cycle (B problems, NMAX node capacity):
//   k_schedule   <<<ceil(B/4), 128>>>        warp per problem: time grid (lane 0), then modes / swing references per node
//   k_init_guess <<<B, 64>>>                 thread per state (warp 0) / input (warp 1) component: warm start interpolation
//   k_kin<1>     <<<(NMAX, B), 32>>>         warp per node: kinematics + derivatives at (x,u), constraint rows, ee terms -> kin scratch
//   k_kin<2>     <<<(NMAX, B), 32>>>         warp per node: kinematics + derivatives at (x + dt f1, u)          -> kin scratch
//   k_proj       <<<(NMAX/4, B), 128>>>      warp per node: projection pivots (register Gauss-Jordan), side stream beside k_kin<2>
// The cycle can be pipelined over chunks of problems (one stream per chunk, run_cycle): B below is the chunk size.
//   k_lq         <<<(NMAX, B), 256>>>        CTA per node: cost/dynamics LQ approximation, projection -> stage/proj blocks
//   k_solve      <<<B, 128>>>                CTA per problem: Riccati backward sweep + forward rollout (serial in nodes)
//   k_trial      <<<B*NMAX/128, 128>>>       thread per (problem, node): value-only evaluation of the trial step (single-pass tree walk in registers)
//   k_decide     <<<ceil(B/4), 128>>>        warp per problem: stages the trial records, lane 0 takes the filter line-search decision
//   k_finalize   <<<B, 64>>>                 thread per component: publish primal solution + warm start
// There is no CPU fallback: without a CUDA device every compute entry point fails with an error.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/qmb200.h"
#include "qm_buffers.h"
#include "qm_target.h"

using namespace qm;

extern "C" void qmb200_set_error_(const char* msg);   // qm_host.cpp (shared error slot behind qmb200_last_error)
static int fail(const std::string& msg) { qmb200_set_error_(msg.c_str()); return -1; }
#define CUDA_OK(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)

enum { KN_SCHEDULE = 0, KN_INIT, KN_KIN1, KN_KIN2, KN_LQ, KN_SOLVE, KN_TRIAL, KN_DECIDE, KN_FINALIZE, KN_POLICY, KN_PROJ, KN_BACKTRACK, KN_STEP };
static const char* kKernelNames[QMB200_NUM_KERNELS] = {"k_schedule", "k_init_guess", "k_kin1",   "k_kin2",   "k_lq",   "k_solve",     "k_trial",
                                                       "k_decide",   "k_finalize",   "k_policy", "k_proj",   "k_backtrack", "k_step"};

// ------------------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(128) k_schedule(MpcBuffers m, const qmb200_solver_desc* S, const qmb200_problem_desc* P) {
  // warp per problem: lane 0 builds the (inherently serial) time grid, then the lanes annotate the nodes in parallel
  const int b = m.b0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= m.b0 + m.nb) return;
  const size_t o = (size_t)b * m.NMAX;
  const double* ev = m.events + (size_t)b * m.EMAX;
  const int32_t* md = m.modes + (size_t)b * (m.EMAX + 1);
  if ((threadIdx.x & 31) == 0) {
    build_grid(*S, m.t0[b], ev, m.nevents[b], m.node_t + o, m.node_flag + o, m.nn + b, m.status + b);
    m.conv[b] = CV_NONE; m.ls[(size_t)b * LS_SIZE + LS_SQP_ITERS] = 0.0; m.ls[(size_t)b * LS_SIZE + LS_CONV] = 0.0;
  }
  __syncwarp();
  annotate_schedule(WarpGroup(), *S, *P, ev, md, m.nevents[b], m.nn[b], m.node_t + o, m.node_flag + o, m.node_ts + o,
                    m.node_dt + o, m.node_mode + o, m.node_zvel + o * 4, m.status + b);
}

__global__ void __launch_bounds__(64) k_init_guess(MpcBuffers m, const qmb200_model_desc* M, const qmb200_problem_desc* P,
                                                    const qmb200_solver_desc* S) {
  // the schedule arrays and the previous solution's time grid are staged in shared memory: the per-component loop over
  // the nodes then only waits for the two warm-start loads of each node
  extern __shared__ __align__(16) double smem[];
  // warp 0: the 30 state components, warp 1: the 30 input components (the two follow different code paths)
  const int b = m.b0 + blockIdx.x, lane = threadIdx.x & 31, c = (threadIdx.x >> 5) * 30 + lane;
  const int NMAX = m.NMAX;
  const size_t o = (size_t)b * NMAX;
  double* s_t = smem; double* s_ts = smem + NMAX; double* s_pt = smem + 2 * NMAX; double* s_ra = smem + 3 * NMAX;   // s_ra: 2 per node
  int* s_flag = (int*)(smem + 5 * NMAX); int* s_mode = s_flag + NMAX; int* s_ri = s_mode + NMAX;                    // s_ri: 3 per node
  const int nn = m.nn[b], np = m.nprev[b];
  for (int i = threadIdx.x; i < NMAX; i += blockDim.x) {
    s_t[i] = (i < nn) ? m.node_t[o + i] : 0.0; s_ts[i] = (i < nn) ? m.node_ts[o + i] : 0.0;
    s_pt[i] = (i < np) ? m.prev_t[o + i] : 0.0;
    s_flag[i] = (i < nn) ? m.node_flag[o + i] : 0; s_mode[i] = (i < nn) ? m.node_mode[o + i] : 0;
  }
  __syncthreads();
  // where the warm start is read, once per node (segment search, weights) ...
  for (int k = threadIdx.x; k < nn - 1; k += blockDim.x) init_guess_node(S->weak_eps, k, s_t, s_flag, s_ts, np, s_pt, s_ri, s_ra);
  __syncthreads();
  if (lane >= 30) return;
  // ... then the trajectory of each component
  init_guess_component(*M, *P, c, m.x0 + 30 * b, nn, s_t, s_ts, s_mode, s_ri, s_ra, np, s_pt, m.prev_x + o * 30, m.prev_u + o * 30,
                       m.xs + o * 30, m.us + o * 30);
}

// ---- transcription = two kinematics kernels (warp per node: the kinematic tree is a dependency chain, so many
//      independent chains per SM) + one LQ assembly kernel (CTA per node: wide small-matrix work)
constexpr int kKinWarps = 1;
constexpr int kKinWarpDoubles = KW_SIZE + RF_SIZE + 12 + 60;
constexpr size_t kKinSmemBytes = (size_t)kKinWarps * kKinWarpDoubles * sizeof(double);

template <int EVAL>
__global__ void __launch_bounds__(32 * kKinWarps) k_kin(MpcBuffers m, const qmb200_model_desc* M, const qmb200_problem_desc* P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * kKinWarps + warp, b = m.b0 + blockIdx.y;
  const int n = m.nn[b] - 1;
  if (k >= n || m.conv[b] != CV_NONE) return;           // terminal node and padding: nothing to evaluate; SQP loop of b stopped
  const size_t o = (size_t)b * m.NMAX + k;
  if (m.node_flag[o] == EV_PRE) return;                 // event node: identity jump map
  extern __shared__ double smem[];
  double* kw = smem + (size_t)warp * kKinWarpDoubles;
  double* scr = kw + KW_SIZE;
  double* xin = scr + RF_SIZE + 12;                     // x[30], u[30]
  NodeIO io = node_io_at(m.kin + o * KS_SIZE);
  WarpGroup g;
  QM_TICK(-1);
  for (int i = lane; i < 60; i += 32) xin[i] = (i < 30) ? ((EVAL == 1) ? m.xs[o * 30 + i] : io.x2[i]) : m.us[o * 30 + i - 30];
  __syncwarp();
  if (EVAL == 1) {
    node_eval1(g, *M, *P, m.node_ts[o], m.node_dt[o], m.node_mode[o], m.node_zvel + o * 4, m.target_t + (size_t)b * m.KT,
               m.target_x + (size_t)b * m.KT * QM_NTARGET, m.KT, xin, xin + 30, kw, scr, io);
  } else {
    io.x2 = xin;
    node_eval2(g, *M, *P, xin + 30, kw, io);
  }
}

// ---- projection pivots: warp per node, register-resident Gauss-Jordan with full pivoting (runs beside k_kin<2>)
constexpr int kProjWarps = 4;
__global__ void __launch_bounds__(32 * kProjWarps) k_proj(MpcBuffers m) {
  const int k = blockIdx.x * kProjWarps + (threadIdx.x >> 5), b = m.b0 + blockIdx.y;
  if (k >= m.nn[b] - 1 || m.conv[b] != CV_NONE) return;
  const size_t o = (size_t)b * m.NMAX + k;
  if (m.node_flag[o] == EV_PRE) return;
  double* base = m.kin + o * KS_SIZE;
  projection_pivots_warp(base + KS_T, m.node_mode[o], base + KS_DINV, (int*)(base + KS_PIV));
}

constexpr int kLqSmemDoubles = TW_LQ_SIZE + 96;
constexpr size_t kLqSmemBytes = kLqSmemDoubles * sizeof(double) + TI_SIZE * sizeof(int);

#ifndef QM_LQ_THREADS
#define QM_LQ_THREADS 256
#endif
__global__ void __launch_bounds__(QM_LQ_THREADS, 4) k_lq(MpcBuffers m, const qmb200_model_desc* M, const qmb200_problem_desc* P) {
  const int k = blockIdx.x, b = m.b0 + blockIdx.y;
  const int nn = m.nn[b];
  if (k >= nn || m.conv[b] != CV_NONE) return;
  extern __shared__ __align__(16) double smem[];
  double* W = smem;
  double* xin = smem + TW_LQ_SIZE;              // x[30], u[30], xn[30]
  int* WI = (int*)(smem + kLqSmemDoubles);
  const size_t o = (size_t)b * m.NMAX + k;
  const int n = nn - 1;
  BlockGroup g;
  const bool regular = (k < n) && (m.node_flag[o] != EV_PRE);
  QM_TICK(-1);
  __shared__ uint64_t bar;
  if (regular && threadIdx.x == 0) {
    // stage the kinematics products of this node into the kinematics region of the workspace: one bulk copy (TMA)
    const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&bar), bytes = KS_SIZE * sizeof(double);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(W + TW_KIN)), "l"(m.kin + o * KS_SIZE), "r"(bytes), "r"(ba)
                 : "memory");
  }
  for (int i = threadIdx.x; i < 90; i += blockDim.x) {
    double v;
    if (i < 30) v = m.xs[o * 30 + i];
    else if (i < 60) v = m.us[o * 30 + i - 30];
    else v = (k < n) ? m.xs[(o + 1) * 30 + i - 60] : 0.0;
    xin[i] = v;
  }
  __syncthreads();
  if (regular) {
    const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(ba) : "memory");
  }
  double* sb = m.stage + o * SB_SIZE;
  double* pb = m.proj + o * PB_SIZE;
  double* pf = m.perf_base + o * PF_SIZE;
  const double* tt = m.target_t + (size_t)b * m.KT;
  const double* ts = m.target_x + (size_t)b * m.KT * QM_NTARGET;
  if (k == n) {
    terminal_node(g, *M, *P, m.node_t[o], m.node_mode[o], tt, ts, m.KT, xin, W + TW_KIN, W + TW_BPM, W + TW_BPM + 80, W + TW_BPM + 88,
                  W + TW_BPM + 100, sb, pf);
  } else if (!regular) {
    event_node(g, xin, xin + 60, sb, pb, pf);
  } else {
    node_lq(g, *M, *P, m.node_ts[o], m.node_dt[o], m.node_mode[o], tt, ts, m.KT, xin, xin + 30, xin + 60, W, WI,
            node_io_at(W + TW_KIN), sb, pb, pf, m.status + b);
  }
}

// ---- TMA staging of the per-node blocks for the serial sweeps (cp.async.bulk global -> shared, completion on an mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mbarrier helpers
__device__ __forceinline__ void mbar_init(uint64_t* bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// k_solve's fetcher: thread 0 issues the bulk copies, everybody waits on the mbarrier of the slot.
struct TmaFetch {
  double* bbuf;            // backward: stage block buffer
  double* fbuf;            // forward: two slots of FWD_SLOT_SIZE doubles (over the then idle backward buffers)
  uint64_t* bar;           // [0] backward, [1], [2] forward slots
  uint32_t pb, pf0, pf1;   // parity of the next completion per barrier

  __device__ void init(double* b, double* f, uint64_t* mbar) {
    bbuf = b; fbuf = f; bar = mbar; pb = pf0 = pf1 = 0;
    if (threadIdx.x == 0) {
      mbar_init(bar); mbar_init(bar + 1); mbar_init(bar + 2);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  __device__ __forceinline__ void bwd_request(BlockGroup, const double* stage) {
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic reads of the buffer are done
      mbar_expect(bar, SB_SIZE * 8u);
      bulk_g2s(bbuf, stage, SB_SIZE * 8u, bar);
    }
  }
  __device__ __forceinline__ const double* bwd_wait(BlockGroup) { mbar_wait(bar, pb); pb ^= 1u; return bbuf; }
  __device__ __forceinline__ void publish(BlockGroup) { asm volatile("fence.proxy.async;" ::: "memory"); }
  __device__ __forceinline__ void fwd_request(BlockGroup, int slot, const double* stage, const double* proj, const double* gain) {
    if (threadIdx.x == 0) {
      double* dst = fbuf + slot * FWD_SLOT_SIZE;
      uint64_t* br = bar + 1 + slot;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect(br, FWD_SLOT_SIZE * 8u);
      bulk_g2s(dst, stage, SB_FWD_SIZE * 8u, br);
      bulk_g2s(dst + SB_FWD_SIZE, proj, PB_SIZE * 8u, br);
      bulk_g2s(dst + SB_FWD_SIZE + PB_SIZE, gain, GB_SIZE * 8u, br);
    }
  }
  __device__ __forceinline__ void fwd_wait(BlockGroup, int slot, const double** st, const double** pbk, const double** gb) {
    if (slot == 0) { mbar_wait(bar + 1, pf0); pf0 ^= 1u; } else { mbar_wait(bar + 2, pf1); pf1 ^= 1u; }
    const double* base = fbuf + slot * FWD_SLOT_SIZE;
    *st = base; *pbk = base + SB_FWD_SIZE; *gb = base + SB_FWD_SIZE + PB_SIZE;
  }
};

// shared memory of k_solve: [stage block | Riccati workspace | 4 mbarriers]; the forward sweep stages its two slots
// over the (then idle) stage block and Riccati workspace, with its scratch behind them
static_assert(2 * FWD_SLOT_SIZE + 96 <= SB_SIZE + RW_SIZE, "forward staging must fit in the backward buffers");
static_assert(SB_FWD_SIZE % 2 == 0 && PB_SIZE % 2 == 0 && GB_SIZE % 2 == 0 && FWD_SLOT_SIZE % 2 == 0, "16-byte aligned bulk copies");
constexpr size_t kSolveSmemBytes = (size_t)(SB_SIZE + RW_SIZE + 4) * sizeof(double);

#ifndef QM_SOLVE_THREADS
#define QM_SOLVE_THREADS 128
#endif
__global__ void __launch_bounds__(QM_SOLVE_THREADS, 4) k_solve(MpcBuffers m) {
  if (m.conv[m.b0 + blockIdx.x] != CV_NONE) return;
  extern __shared__ __align__(16) double smem[];
  TmaFetch fetch;
  fetch.init(smem, smem, (uint64_t*)(smem + SB_SIZE + RW_SIZE));
  BlockGroup g;
  g.nwid = blockIdx.x & 3;                        // spread the serial chains of co-resident CTAs over the sub-partitions
  solve_problem(g, fetch, m, m.b0 + blockIdx.x, smem + SB_SIZE, smem + 2 * FWD_SLOT_SIZE);
}

// line-search evaluation: a thread per node (value-only single-pass tree walk in registers, qm_value.h)
constexpr int kTrialThreads = 128;
#ifndef QM_TRIAL_MINBLOCKS
#define QM_TRIAL_MINBLOCKS 2
#endif
// performance record (cost, dynamics defect, equality-constraint SSE) of node k of problem b at the trial step alpha
__device__ __forceinline__ void trial_node(const MpcBuffers& m, const qmb200_model_desc& M, const qmb200_problem_desc& P, int b, int k,
                                           int nn, double alpha, double* out) {
  const size_t o = (size_t)b * m.NMAX + k;
  const int n = nn - 1;
  double x[30], u[30], xn[30];
  for (int i = 0; i < 30; ++i) {
    x[i] = m.xs[o * 30 + i] + alpha * m.dxs[o * 30 + i];
    u[i] = m.us[o * 30 + i] + alpha * m.dus[o * 30 + i];
    xn[i] = (k < n) ? m.xs[(o + 1) * 30 + i] + alpha * m.dxs[(o + 1) * 30 + i] : 0.0;
  }
  double pf[PF_SIZE];
  const double* tt = m.target_t + (size_t)b * m.KT;
  const double* ts = m.target_x + (size_t)b * m.KT * QM_NTARGET;
  if (k == n) {
    perf_terminal_serial(M, P, m.node_t[o], m.node_mode[o], tt, ts, m.KT, x, pf);
  } else if (m.node_flag[o] == EV_PRE) {
    double d = 0.0;
    for (int i = 0; i < 30; ++i) d += (x[i] - xn[i]) * (x[i] - xn[i]);
    pf[PF_COST] = 0.0; pf[PF_DYN] = d; pf[PF_EQ] = 0.0;
  } else {
    perf_node_serial(M, P, m.node_ts[o], m.node_dt[o], m.node_mode[o], m.node_zvel + o * 4, tt, ts, m.KT, x, u, xn, pf);
  }
  out[PF_COST] = pf[PF_COST]; out[PF_DYN] = pf[PF_DYN]; out[PF_EQ] = pf[PF_EQ];
}

// first trial (alpha = 1) of every problem: (problem, node) pairs are laid over the threads back to back: a block per
// problem would leave its last warp with a handful of nodes (105 nodes on 128 threads)
__global__ void __launch_bounds__(kTrialThreads, QM_TRIAL_MINBLOCKS) k_trial(MpcBuffers m, const qmb200_model_desc* M, const qmb200_problem_desc* P) {
  const int gid = blockIdx.x * kTrialThreads + threadIdx.x;
  const int pi = gid / m.NMAX, k = gid - pi * m.NMAX;
  if (pi >= m.nb) return;
  const int b = m.b0 + pi;
  const double* ls = m.ls + (size_t)b * LS_SIZE;
  if (ls[LS_DONE] != 0.0) return;                 // also skips the problems whose SQP loop stopped in an earlier iteration
  const int nn = m.nn[b];
  if (k >= nn) return;
  trial_node(m, *M, *P, b, k, nn, ls[LS_ALPHA], m.perf_trial + ((size_t)b * m.NMAX + k) * PF_SIZE);
}

// warp per problem: the lanes stage the trial performance records of the nodes (coalesced), lane 0 decides (sequential sums)
constexpr int kDecideWarps = 4;
__global__ void __launch_bounds__(32 * kDecideWarps) k_decide(MpcBuffers m, const qmb200_solver_desc* S) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = m.b0 + blockIdx.x * kDecideWarps + warp;
  if (b >= m.b0 + m.nb) return;
  if (m.ls[(size_t)b * LS_SIZE + LS_DONE] != 0.0) return;
  double* pf = smem + (size_t)warp * m.NMAX * PF_SIZE;
  const double* src = m.perf_trial + (size_t)b * m.NMAX * PF_SIZE;
  const int cnt = m.nn[b] * PF_SIZE;
  for (int i = lane; i < cnt; i += 32) pf[i] = src[i];
  __syncwarp();
  if (lane != 0) return;
  decide_problem(*S, m, b, pf);
}

// Backtracking trials ([upstream] FilterLinesearch loop), entirely on the device: a CTA per problem whose full step was not
// accepted (a few percent of a batch) repeats {trial of every node at the reduced alpha, decision} until the step is accepted or
// alpha falls below alpha_min (decide_problem marks the step rejected then). Same per-node evaluation and the same sequential
// sums as k_trial / k_decide, so the result does not depend on which kernel took a trial. The host is not involved: the cycle
// is one asynchronous sequence of launches.
__global__ void __launch_bounds__(kTrialThreads, QM_TRIAL_MINBLOCKS) k_backtrack(MpcBuffers m, const qmb200_model_desc* M, const qmb200_problem_desc* P,
                                                                                   const qmb200_solver_desc* S) {
  const int b = m.b0 + blockIdx.x;
  volatile double* ls = m.ls + (size_t)b * LS_SIZE;
  if (ls[LS_DONE] != 0.0) return;                 // uniform over the CTA
  extern __shared__ __align__(16) double smem[];  // [NMAX][PF_SIZE] trial records of this problem
  const int nn = m.nn[b];
  for (int round = 0; round < 128; ++round) {     // alpha_decay < 1 is checked at creation: decide_problem ends the loop
    const double alpha = ls[LS_ALPHA];
    for (int k = threadIdx.x; k < nn; k += kTrialThreads) trial_node(m, *M, *P, b, k, nn, alpha, smem + (size_t)k * PF_SIZE);
    __syncthreads();
    if (threadIdx.x == 0) { decide_problem(*S, m, b, smem); __threadfence_block(); }
    __syncthreads();
    if (ls[LS_DONE] != 0.0) return;
  }
  if (threadIdx.x == 0) {                         // not reached with valid settings: no step is taken
    ls[LS_DONE] = 2.0; ls[LS_ALPHA] = 0.0; atomicOr(m.status + b, ST_STEP_REJECTED);
  }
}

// intermediate SQP iteration: take the accepted step in place, then the convergence test of the iteration
__global__ void __launch_bounds__(64) k_step(MpcBuffers m, const qmb200_solver_desc* S, int it, int iterations) {
  const int b = m.b0 + blockIdx.x, lane = threadIdx.x & 31, c = (threadIdx.x >> 5) * 30 + lane;
  if (m.conv[b] != CV_NONE) return;
  if (lane < 30) step_component(m, b, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int cv = check_convergence(*S, m.ls + (size_t)b * LS_SIZE, it, iterations);
    m.conv[b] = cv; m.ls[(size_t)b * LS_SIZE + LS_CONV] = (double)cv;
  }
}

__global__ void __launch_bounds__(64) k_finalize(MpcBuffers m, double* t_out, double* x_out, double* u_out, double* packed) {
  const int b = m.b0 + blockIdx.x, lane = threadIdx.x & 31, c = (threadIdx.x >> 5) * 30 + lane;   // warp 0: states, warp 1: inputs
  if (lane < 30) finalize_component(m, b, c, t_out, x_out, u_out, packed);
  if (threadIdx.x == 31) {            // an idle lane records why the SQP loop stopped (the last iteration ends with ITERATIONS)
    double* ls = m.ls + (size_t)b * LS_SIZE;
    if (m.conv[b] == CV_NONE) ls[LS_CONV] = (double)CV_ITERATIONS;
  }
}

// [upstream] MPC_MRT_Interface::evaluatePolicy with a FeedforwardController: linear interpolation of (x*, u*) at t.
__global__ void __launch_bounds__(64) k_policy(MpcBuffers m, const double* t, double* x_des, double* u_des, int32_t* mode) {
  const int b = blockIdx.x, c = threadIdx.x;
  if (c >= 60) return;
  const size_t o = (size_t)b * m.NMAX;
  const int np = m.nprev[b];
  int i; double a;
  time_segment(t[b], m.prev_t + o, np, &i, &a);
  const int i2 = (np > 1) ? i + 1 : i;
  if (c < 30) x_des[30 * b + c] = a * m.prev_x[(o + i) * 30 + c] + (1.0 - a) * m.prev_x[(o + i2) * 30 + c];
  else u_des[30 * b + c - 30] = a * m.prev_u[(o + i) * 30 + c - 30] + (1.0 - a) * m.prev_u[(o + i2) * 30 + c - 30];
  if (c == 0) mode[b] = m.modes[(size_t)b * (m.EMAX + 1) + mode_index(m.events + (size_t)b * m.EMAX, m.nevents[b], t[b])];
}

// Command -> two-knot reference (QmTargetTrajectoriesPublisher_node.cpp:60-257): thread per command.
__global__ void __launch_bounds__(128) k_targets(int n, int kind, qmb200_target_desc D, const double* cmd, const double* obs_time,
                                                 const double* obs_state, const double* ee_state, double* last_ee, double* tt, double* tx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  command_to_target(D, kind, cmd + 7 * (size_t)b, obs_time[b], obs_state + 30 * (size_t)b, ee_state + 7 * (size_t)b, last_ee + 7 * (size_t)b,
                    tt + 2 * (size_t)b, tx + 2 * QM_NTARGET * (size_t)b);
}

// [upstream] computeCentroidalStateFromRbdModel + yaw unwrapping (QMController.cpp:239-244): warp per measured state.
constexpr int kStateWarps = 4;
constexpr int kStateWarpDoubles = KW_VSIZE + 48;
__global__ void __launch_bounds__(32 * kStateWarps) k_rbd_state(int B, const qmb200_model_desc* M, const double* rbd, const double* yaw_last,
                                                                 double* x_out) {
  const int b = blockIdx.x * kStateWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  extern __shared__ __align__(16) double smem[];
  double* kw = smem + (size_t)(threadIdx.x >> 5) * kStateWarpDoubles;
  centroidal_state_from_rbd(WarpGroup(), *M, rbd + 55 * (size_t)b, yaw_last ? yaw_last[b] : 0.0, yaw_last != nullptr, kw, kw + KW_VSIZE,
                            x_out + 30 * (size_t)b);
}

// Feedback gains K[B][NMAX][30][30] of the last cycle (useFeedbackPolicy): CTA per node. Pre-event nodes and the final node
// repeat the gain of the previous node, as toPrimalSolution does for the inputs.
__global__ void __launch_bounds__(128) k_gains(MpcBuffers m, double* K) {
  const int k = blockIdx.x, b = blockIdx.y;
  const int nn = m.nn[b];
  if (k >= nn) return;
  const size_t o = (size_t)b * m.NMAX;
  double* out = K + (o + k) * 900;
  const int src = feedback_gain_source(nn, m.node_flag + o, k);
  if (src < 0) {
    for (int i = threadIdx.x; i < 900; i += blockDim.x) out[i] = 0.0;
    return;
  }
  feedback_gain_node(BlockGroup(), m.proj + (o + src) * PB_SIZE, m.gain + (o + src) * GB_SIZE, out);
}

// [upstream] LinearController::computeInput: u = uff(t) + K(t) x with uff_i = u*_i - K_i x*_i, both interpolated linearly.
__global__ void __launch_bounds__(32) k_policy_fb(MpcBuffers m, const double* K, const double* t, const double* x, double* u_out, int32_t* mode) {
  const int b = blockIdx.x, i = threadIdx.x;
  const size_t o = (size_t)b * m.NMAX;
  const int np = m.nprev[b];
  int seg; double a;
  time_segment(t[b], m.prev_t + o, np, &seg, &a);
  const int seg2 = (np > 1) ? seg + 1 : seg;
  if (i < 30) {
    double acc = 0.0;
    const int segs[2] = {seg, seg2};
    const double wts[2] = {a, 1.0 - a};
    for (int s = 0; s < 2; ++s) {
      const double* Kn = K + (o + segs[s]) * 900 + 30 * i;
      double v = m.prev_u[(o + segs[s]) * 30 + i];
      for (int j = 0; j < 30; ++j) v += Kn[j] * (x[30 * b + j] - m.prev_x[(o + segs[s]) * 30 + j]);
      acc += wts[s] * v;
    }
    u_out[30 * b + i] = acc;
  }
  if (i == 0) mode[b] = m.modes[(size_t)b * (m.EMAX + 1) + mode_index(m.events + (size_t)b * m.EMAX, m.nevents[b], t[b])];
}

// ------------------------------------------------------------------------------------------ context
struct qmb200_ctx {
  int device = 0;
  int B = 0;
  qmb200_model_desc hM;
  qmb200_problem_desc hP;
  qmb200_solver_desc hS;
  qmb200_model_desc* dM = nullptr;
  qmb200_problem_desc* dP = nullptr;
  qmb200_solver_desc* dS = nullptr;
  MpcBuffers m;          // ctx-owned device buffers
  double* fb_gains = nullptr; // [B][NMAX][900] feedback gains of the last cycle, allocated on first use
  // Asynchronous host interface (qmb200_mpc_cycle_batch_async / _wait): results leave on a copy stream while the next cycle
  // computes. The small per-cycle records the next cycle overwrites early (node counts, modes, line-search record, status) are
  // snapshotted behind k_finalize; the policy itself (prev_t / prev_x / prev_u) is only rewritten by the next k_finalize, which
  // waits for the copy of the previous one.
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_cycle[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  int64_t ticket = 0;          // number of asynchronous cycles submitted
  bool d2h_in_flight = false;  // the last submitted cycle's copies may still be running
  int32_t* snap_n = nullptr; int32_t* snap_mode = nullptr; double* snap_ls = nullptr; int32_t* snap_status = nullptr;
  // The launches of a cycle up to the line search are replayed as a CUDA graph (captured on first use, re-captured when the
  // input pointers change); per-kernel event timing (qmb200_set_profiling) launches them directly.
  bool use_graph = true;
  cudaGraphExec_t gexec = nullptr;
  const void* gkey[8] = {nullptr};
  int64_t graph_launches = 0, graph_captures = 0;
  // multi-GPU: packed policy [B][NMAX][61] of the last two cycles (send buffers of the all-gather, written by k_finalize once a
  // communicator exists), the communicator and its stream: the collective of cycle k runs beside the kernels of cycle k + 1
  double* policy[2] = {nullptr, nullptr};
  int pslot = 0;                 // send buffer the last k_finalize wrote
  void* comm = nullptr;          // ncclComm_t owned by the context (qmb200_comm_init)
  int world = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_policy = nullptr, ev_comm[2] = {nullptr, nullptr};
  bool comm_in_flight[2] = {false, false};
  bool capturing = false;
  int64_t graph_kernel_counts[QMB200_NUM_KERNELS] = {0};   // kernels per replay of the captured graph
  cudaStream_t stream = nullptr;
  // The cycle can be pipelined over chunks of problems, one stream per chunk, so that the (latency-bound, few CTAs) Riccati
  // sweeps of one chunk run beside the transcription kernels of the next ones (QMB200_CHUNKS; see qmb200_create).
  static const int kMaxChunks = 8;
  int nchunks = 1;
  cudaStream_t cs[kMaxChunks] = {nullptr};
  cudaStream_t side[kMaxChunks] = {nullptr};   // k_proj of a chunk runs here, beside its k_kin<2>
  cudaEvent_t ev_start = nullptr, ev_done[kMaxChunks] = {nullptr}, ev_fork[kMaxChunks] = {nullptr}, ev_join[kMaxChunks] = {nullptr};
  int64_t bytes = 0;
  bool profiling = false;
  double kernel_ms[QMB200_NUM_KERNELS] = {0};
  int64_t kernel_launches[QMB200_NUM_KERNELS] = {0};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
  std::vector<cudaEvent_t> event_pool;
};

static cudaEvent_t get_event(qmb200_ctx* c) {
  if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct KernelTimer {
  qmb200_ctx* c; int id; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelTimer(qmb200_ctx* c_, int id_, cudaStream_t st_ = nullptr) : c(c_), id(id_), st(st_ ? st_ : c_->stream) {
    if (c->capturing) c->graph_kernel_counts[id]++; else c->kernel_launches[id]++;
    if (c->profiling) { e0 = get_event(c); e1 = get_event(c); cudaEventRecord(e0, st); }
  }
  ~KernelTimer() {
    if (c->profiling) { cudaEventRecord(e1, st); c->pending_events.push_back({id, {e0, e1}}); }
  }
};

static void harvest_events(qmb200_ctx* c) {
  for (auto& pe : c->pending_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pe.second.first, pe.second.second) == cudaSuccess) c->kernel_ms[pe.first] += ms;
    c->event_pool.push_back(pe.second.first);
    c->event_pool.push_back(pe.second.second);
  }
  c->pending_events.clear();
}

// One MPC cycle = one asynchronous sequence of launches (no host synchronisation anywhere: the backtracking trials of the filter
// line search run in k_backtrack, the SQP loop's early exit is a per-problem flag every kernel tests).
// front: schedule ... line search of the last SQP iteration (per chunk of problems, joined on the main stream);
// back : k_finalize of the whole batch (the only writer of the stored policy) and the snapshot of the small per-cycle records.
static int enqueue_front(qmb200_ctx* c, MpcBuffers m) {
  const int B = m.B, NMAX = m.NMAX;
  const int nch = c->nchunks, cb = (B + nch - 1) / nch;
  const int iterations = c->hS.sqp_iterations < 1 ? 1 : c->hS.sqp_iterations;
  const size_t pf_bytes = (size_t)NMAX * PF_SIZE * sizeof(double);
  // inputs were enqueued on the main stream: every chunk stream starts behind them
  CUDA_OK(cudaEventRecord(c->ev_start, c->stream));
  for (int ch = 0; ch < nch; ++ch) {
    cudaStream_t st = c->cs[ch];
    m.b0 = ch * cb;
    m.nb = (m.b0 + cb <= B) ? cb : B - m.b0;
    if (m.nb <= 0) continue;
    const int nb = m.nb;
    CUDA_OK(cudaStreamWaitEvent(st, c->ev_start, 0));
    { KernelTimer kt(c, KN_SCHEDULE, st); k_schedule<<<(nb + 3) / 4, 128, 0, st>>>(m, c->dS, c->dP); }
    { KernelTimer kt(c, KN_INIT, st); k_init_guess<<<nb, 64, (size_t)NMAX * 60, st>>>(m, c->dM, c->dP, c->dS); }
    for (int it = 0; it < iterations; ++it) {
      { KernelTimer kt(c, KN_KIN1, st); k_kin<1><<<dim3((NMAX + kKinWarps - 1) / kKinWarps, nb), 32 * kKinWarps, kKinSmemBytes, st>>>(m, c->dM, c->dP); }
      // the projection pivots only need the constraint rows of k_kin<1>: they run beside k_kin<2> (FP64-issue bound warps next to
      // latency-bound ones) and join before k_lq
      CUDA_OK(cudaEventRecord(c->ev_fork[ch], st));
      CUDA_OK(cudaStreamWaitEvent(c->side[ch], c->ev_fork[ch], 0));
      { KernelTimer kt(c, KN_PROJ, c->side[ch]); k_proj<<<dim3((NMAX + kProjWarps - 1) / kProjWarps, nb), 32 * kProjWarps, 0, c->side[ch]>>>(m); }
      CUDA_OK(cudaEventRecord(c->ev_join[ch], c->side[ch]));
      { KernelTimer kt(c, KN_KIN2, st); k_kin<2><<<dim3((NMAX + kKinWarps - 1) / kKinWarps, nb), 32 * kKinWarps, kKinSmemBytes, st>>>(m, c->dM, c->dP); }
      CUDA_OK(cudaStreamWaitEvent(st, c->ev_join[ch], 0));
      { KernelTimer kt(c, KN_LQ, st); k_lq<<<dim3(NMAX, nb), QM_LQ_THREADS, kLqSmemBytes, st>>>(m, c->dM, c->dP); }
      { KernelTimer kt(c, KN_SOLVE, st); k_solve<<<nb, QM_SOLVE_THREADS, kSolveSmemBytes, st>>>(m); }
      { KernelTimer kt(c, KN_TRIAL, st); k_trial<<<(nb * NMAX + kTrialThreads - 1) / kTrialThreads, kTrialThreads, 0, st>>>(m, c->dM, c->dP); }
      { KernelTimer kt(c, KN_DECIDE, st); k_decide<<<(nb + kDecideWarps - 1) / kDecideWarps, 32 * kDecideWarps, kDecideWarps * pf_bytes, st>>>(m, c->dS); }
      { KernelTimer kt(c, KN_BACKTRACK, st); k_backtrack<<<nb, kTrialThreads, pf_bytes, st>>>(m, c->dM, c->dP, c->dS); }
      if (it + 1 < iterations) { KernelTimer kt(c, KN_STEP, st); k_step<<<nb, 64, 0, st>>>(m, c->dS, it, iterations); }
    }
    CUDA_OK(cudaEventRecord(c->ev_done[ch], st));
    CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_done[ch], 0));     // the main stream continues behind every chunk
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int run_front(qmb200_ctx* c, const MpcBuffers& m) {
  if (!c->use_graph || c->profiling) return enqueue_front(c, m);
  const void* key[8] = {m.t0, m.x0, m.events, m.modes, m.nevents, m.target_t, m.target_x, (const void*)(intptr_t)c->hS.sqp_iterations};
  if (!c->gexec || memcmp(key, c->gkey, sizeof(key)) != 0) {
    if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    for (auto& n : c->graph_kernel_counts) n = 0;
    c->capturing = true;
    const int rc = enqueue_front(c, m);
    c->capturing = false;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return -1; }
    if (e != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&c->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { c->gexec = nullptr; return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    memcpy(c->gkey, key, sizeof(key));
    c->graph_captures++;
  }
  CUDA_OK(cudaGraphLaunch(c->gexec, c->stream));
  c->graph_launches++;
  for (int i = 0; i < QMB200_NUM_KERNELS; ++i) c->kernel_launches[i] += c->graph_kernel_counts[i];
  return 0;
}

static int run_back(qmb200_ctx* c, MpcBuffers m, double* t_out, double* x_out, double* u_out) {
  // the stored policy is about to be rewritten: the asynchronous copy of the previous one must have left
  if (c->d2h_in_flight) { CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_d2h[(c->ticket - 1) & 1], 0)); c->d2h_in_flight = false; }
  m.b0 = 0; m.nb = m.B;
  double* packed = nullptr;
  if (c->policy[0]) {
    // the send buffer written two cycles ago: its all-gather must have read it
    c->pslot ^= 1;
    if (c->comm_in_flight[c->pslot]) { CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_comm[c->pslot], 0)); c->comm_in_flight[c->pslot] = false; }
    packed = c->policy[c->pslot];
  }
  { KernelTimer kt(c, KN_FINALIZE); k_finalize<<<m.B, 64, 0, c->stream>>>(m, t_out, x_out, u_out, packed); }
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int run_cycle(qmb200_ctx* c, const MpcBuffers& m, double* t_out, double* x_out, double* u_out) {
  if (run_front(c, m) != 0) return -1;
  return run_back(c, m, t_out, x_out, u_out);
}

extern "C" {

int qmb200_version(void) { return 100; }
const char* qmb200_kernel_name(int32_t i) { return (i >= 0 && i < QMB200_NUM_KERNELS) ? kKernelNames[i] : ""; }

int qmb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// CUDA_OK for the body of a create function: every failure releases what was built so far
#define CREATE_OK(call, destroy_call)                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) { destroy_call; return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); } \
  } while (0)

int qmb200_create(const qmb200_model_desc* model, const qmb200_problem_desc* problem, const qmb200_solver_desc* solver,
                  int32_t batch, int32_t device, qmb200_ctx** out) {
  if (!model || !problem || !solver || !out) return fail("qmb200_create: null argument");
  if (batch <= 0) return fail("qmb200_create: batch must be positive");
  if (model->nj != QM_NJ) return fail("qmb200_create: model must have 24 one-DoF joints (6 floating base + 18 actuated)");
  if (!value_walk_supported(*model)) return fail("qmb200_create: unsupported tree topology (serial six-joint floating base carrying all limbs expected)");
  if (solver->max_nodes < 3 || solver->max_events < 1 || solver->max_targets < 1) return fail("qmb200_create: bad capacities");
  if (!(solver->alpha_decay > 0.0 && solver->alpha_decay < 1.0) || !(solver->alpha_min > 0.0))
    return fail("qmb200_create: line search needs 0 < alpha_decay < 1 and alpha_min > 0");
  if (!(solver->dt > 0.0) || !(solver->horizon > 0.0)) return fail("qmb200_create: dt and horizon must be positive");
  // staging buffers that grow with the node capacity (long horizons): k_decide 4 x NMAX records, k_init_guess 60 B per node
  const size_t dec = (size_t)kDecideWarps * solver->max_nodes * PF_SIZE * sizeof(double), ini = (size_t)solver->max_nodes * 60;
  if (dec > 200 * 1024 || ini > 200 * 1024) return fail("qmb200_create: max_nodes too large for the staging buffers of k_decide / k_init_guess");
  if (qmb200_device_count() <= 0) return fail("qmb200_create: no CUDA device available (this library has no CPU fallback)");
  CUDA_OK(cudaSetDevice(device));
  qmb200_ctx* c = new qmb200_ctx();
  c->device = device; c->B = batch; c->hM = *model; c->hP = *problem; c->hS = *solver;
#define C_OK(call) CREATE_OK(call, qmb200_destroy(c))
  C_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // Tuning knob. Measured on B200 (config 2): 1 / 2 / 4 / 8 chunks -> 13.12 / 13.24 / 13.63 / 14.97 ms per step: the kernels
    // do overlap, but the step is bound by instruction issue across the whole GPU, so the default stays at one chunk.
    const char* env = getenv("QMB200_CHUNKS");
    int nch = env ? atoi(env) : 1;
    if (nch < 1) nch = 1;
    if (nch > qmb200_ctx::kMaxChunks) nch = qmb200_ctx::kMaxChunks;
    if (nch > batch) nch = batch;
    c->nchunks = nch;
  }
  for (int ch = 0; ch < c->nchunks; ++ch) {
    C_OK(cudaStreamCreateWithFlags(&c->cs[ch], cudaStreamNonBlocking));
    C_OK(cudaStreamCreateWithFlags(&c->side[ch], cudaStreamNonBlocking));
    C_OK(cudaEventCreateWithFlags(&c->ev_done[ch], cudaEventDisableTiming));
    C_OK(cudaEventCreateWithFlags(&c->ev_fork[ch], cudaEventDisableTiming));
    C_OK(cudaEventCreateWithFlags(&c->ev_join[ch], cudaEventDisableTiming));
  }
  C_OK(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
  C_OK(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    C_OK(cudaEventCreateWithFlags(&c->ev_cycle[i], cudaEventDisableTiming));
    C_OK(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
  }
  { const char* env = getenv("QMB200_GRAPH"); c->use_graph = !(env && atoi(env) == 0); }
  C_OK(cudaMalloc(&c->snap_n, sizeof(int32_t) * (size_t)batch));
  C_OK(cudaMalloc(&c->snap_mode, sizeof(int32_t) * (size_t)batch * solver->max_nodes));
  C_OK(cudaMalloc(&c->snap_ls, sizeof(double) * (size_t)batch * LS_SIZE));
  C_OK(cudaMalloc(&c->snap_status, sizeof(int32_t) * (size_t)batch));
  C_OK(cudaMalloc(&c->dM, sizeof(*model)));
  C_OK(cudaMalloc(&c->dP, sizeof(*problem)));
  C_OK(cudaMalloc(&c->dS, sizeof(*solver)));
  C_OK(cudaMemcpy(c->dM, model, sizeof(*model), cudaMemcpyHostToDevice));
  C_OK(cudaMemcpy(c->dP, problem, sizeof(*problem), cudaMemcpyHostToDevice));
  C_OK(cudaMemcpy(c->dS, solver, sizeof(*solver), cudaMemcpyHostToDevice));
  c->m.b0 = 0; c->m.nb = batch;
  c->m.B = batch; c->m.NMAX = solver->max_nodes; c->m.EMAX = solver->max_events; c->m.KT = solver->max_targets;
  cudaError_t err = cudaSuccess;
  int64_t total = 0;
  for_each_buffer(c->m, [&](void** p, size_t bytes) {
    if (err != cudaSuccess) { *p = nullptr; return; }
    err = cudaMalloc(p, bytes);
    if (err == cudaSuccess) { err = cudaMemset(*p, 0, bytes); total += (int64_t)bytes; } else *p = nullptr;
  });
  if (err != cudaSuccess) { qmb200_destroy(c); return fail(std::string("qmb200_create: device allocation failed: ") + cudaGetErrorString(err)); }
  c->bytes = total;
  C_OK(cudaFuncSetAttribute(k_kin<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKinSmemBytes));
  C_OK(cudaFuncSetAttribute(k_kin<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKinSmemBytes));
  C_OK(cudaFuncSetAttribute(k_lq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLqSmemBytes));
  C_OK(cudaFuncSetAttribute(k_rbd_state, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStateWarps * kStateWarpDoubles * sizeof(double))));
  C_OK(cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveSmemBytes));
  if (dec > 48 * 1024) C_OK(cudaFuncSetAttribute(k_decide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec));
  if (dec / kDecideWarps > 48 * 1024) C_OK(cudaFuncSetAttribute(k_backtrack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(dec / kDecideWarps)));
  if (ini > 48 * 1024) C_OK(cudaFuncSetAttribute(k_init_guess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ini));
#undef C_OK
  *out = c;
  return 0;
}

int qmb200_destroy(qmb200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  harvest_events(c);
  for (auto e : c->event_pool) cudaEventDestroy(e);
  for_each_buffer(c->m, [](void** p, size_t) { if (*p) cudaFree(*p); *p = nullptr; });
  if (c->dM) cudaFree(c->dM);
  if (c->dP) cudaFree(c->dP);
  if (c->dS) cudaFree(c->dS);
  if (c->fb_gains) cudaFree(c->fb_gains);
  qmb200_comm_destroy(c);
  if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
  if (c->gexec) cudaGraphExecDestroy(c->gexec);
  for (int i = 0; i < 2; ++i) { if (c->ev_cycle[i]) cudaEventDestroy(c->ev_cycle[i]); if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]); }
  for (void* p : {(void*)c->snap_n, (void*)c->snap_mode, (void*)c->snap_ls, (void*)c->snap_status}) if (p) cudaFree(p);
  for (int ch = 0; ch < qmb200_ctx::kMaxChunks; ++ch) {
    if (c->cs[ch]) { cudaStreamSynchronize(c->cs[ch]); cudaStreamDestroy(c->cs[ch]); }
    if (c->side[ch]) { cudaStreamSynchronize(c->side[ch]); cudaStreamDestroy(c->side[ch]); }
    if (c->ev_done[ch]) cudaEventDestroy(c->ev_done[ch]);
    if (c->ev_fork[ch]) cudaEventDestroy(c->ev_fork[ch]);
    if (c->ev_join[ch]) cudaEventDestroy(c->ev_join[ch]);
  }
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

// ---- multi-GPU collective. NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process already uses -- e.g.
//      the one bundled with PyTorch -- or the system one), so the library loads on hosts without NCCL and single-GPU use never
//      touches it. Minimal declarations of the NCCL C API (nccl.h) used here:
typedef struct { char internal[128]; } qm_ncclUniqueId;
typedef void* qm_ncclComm_t;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(qm_ncclUniqueId*) = nullptr;
  int (*CommInitRank)(qm_ncclComm_t*, int, qm_ncclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, qm_ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(qm_ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* env = getenv("QMB200_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm) continue;
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.GetUniqueId = (int (*)(qm_ncclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(qm_ncclComm_t*, int, qm_ncclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
      api.AllGather = (int (*)(const void*, void*, size_t, int, qm_ncclComm_t, cudaStream_t))dlsym(api.handle, "ncclAllGather");
      api.CommDestroy = (int (*)(qm_ncclComm_t))dlsym(api.handle, "ncclCommDestroy");
      api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) api.handle = nullptr;
    }
  }
  return api.handle ? &api : nullptr;
}
static int nccl_fail(NcclApi* a, const char* what, int rc) {
  return fail(std::string(what) + ": " + ((a && a->GetErrorString) ? a->GetErrorString(rc) : "NCCL error"));
}
enum { QM_NCCL_FLOAT64 = 8 };   // ncclDataType_t ncclFloat64 / ncclDouble

int qmb200_nccl_unique_id(void* id128) {
  if (!id128) return fail("qmb200_nccl_unique_id: null argument");
  NcclApi* a = nccl_api();
  if (!a) return fail("qmb200_nccl_unique_id: libnccl.so.2 not found (set QMB200_NCCL_LIB)");
  qm_ncclUniqueId id;
  const int rc = a->GetUniqueId(&id);
  if (rc != 0) return nccl_fail(a, "ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

// send buffers + stream + events of the collective (also without a communicator of our own: qmb200_allgather_policy accepts one)
static int ensure_policy_buffers(qmb200_ctx* c) {
  if (c->policy[0]) return 0;
  const size_t bytes = sizeof(double) * (size_t)c->B * c->m.NMAX * QMB200_POLICY_WIDTH;
  for (int i = 0; i < 2; ++i) {
    CUDA_OK(cudaMalloc(&c->policy[i], bytes));
    CUDA_OK(cudaMemsetAsync(c->policy[i], 0, bytes, c->stream));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_comm[i], cudaEventDisableTiming));
  }
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_policy, cudaEventDisableTiming));
  CUDA_OK(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  c->bytes += 2 * (int64_t)bytes;
  return 0;
}

int qmb200_comm_init(qmb200_ctx* c, const void* id128, int32_t rank, int32_t world) {
  if (!c || !id128) return fail("qmb200_comm_init: null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail("qmb200_comm_init: bad rank / world");
  if (c->comm) return fail("qmb200_comm_init: the context already owns a communicator");
  NcclApi* a = nccl_api();
  if (!a) return fail("qmb200_comm_init: libnccl.so.2 not found (set QMB200_NCCL_LIB)");
  CUDA_OK(cudaSetDevice(c->device));
  if (ensure_policy_buffers(c) != 0) return -1;
  qm_ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  qm_ncclComm_t comm = nullptr;
  const int rc = a->CommInitRank(&comm, world, id, rank);
  if (rc != 0) return nccl_fail(a, "ncclCommInitRank", rc);
  c->comm = comm; c->world = world; c->rank = rank;
  return 0;
}

int qmb200_comm_destroy(qmb200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (c->comm) { NcclApi* a = nccl_api(); if (a) a->CommDestroy(c->comm); c->comm = nullptr; }
  for (int i = 0; i < 2; ++i) {
    if (c->policy[i]) { cudaFree(c->policy[i]); c->policy[i] = nullptr; }
    if (c->ev_comm[i]) { cudaEventDestroy(c->ev_comm[i]); c->ev_comm[i] = nullptr; }
    c->comm_in_flight[i] = false;
  }
  if (c->ev_policy) { cudaEventDestroy(c->ev_policy); c->ev_policy = nullptr; }
  if (c->comm_stream) { cudaStreamDestroy(c->comm_stream); c->comm_stream = nullptr; }
  c->world = 1; c->rank = 0;
  return 0;
}

int qmb200_enable_policy_buffer(qmb200_ctx* c) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  return ensure_policy_buffers(c);
}

const double* qmb200_policy_buffer(qmb200_ctx* c) { return (c && c->policy[0]) ? c->policy[c->pslot] : nullptr; }

int qmb200_allgather_policy(qmb200_ctx* c, void* nccl_comm, double* gathered) {
  if (!c || !gathered) return fail("qmb200_allgather_policy: null argument");
  if (!c->policy[0]) return fail("qmb200_allgather_policy: no packed policy yet (qmb200_comm_init or qmb200_enable_policy_buffer before the cycle)");
  qm_ncclComm_t comm = nccl_comm ? nccl_comm : c->comm;
  CUDA_OK(cudaSetDevice(c->device));
  const int slot = c->pslot;
  const size_t count = (size_t)c->B * c->m.NMAX * QMB200_POLICY_WIDTH;
  // the collective starts once the cycle that wrote the send buffer is through, on its own stream
  CUDA_OK(cudaEventRecord(c->ev_policy, c->stream));
  CUDA_OK(cudaStreamWaitEvent(c->comm_stream, c->ev_policy, 0));
  if (!comm) {
    if (c->world != 1) return fail("qmb200_allgather_policy: no communicator");
    CUDA_OK(cudaMemcpyAsync(gathered, c->policy[slot], sizeof(double) * count, cudaMemcpyDeviceToDevice, c->comm_stream));   // one rank
  } else {
    NcclApi* a = nccl_api();
    if (!a) return fail("qmb200_allgather_policy: libnccl.so.2 not found");
    const int rc = a->AllGather(c->policy[slot], gathered, count, QM_NCCL_FLOAT64, comm, c->comm_stream);
    if (rc != 0) return nccl_fail(a, "ncclAllGather", rc);
  }
  CUDA_OK(cudaEventRecord(c->ev_comm[slot], c->comm_stream));
  c->comm_in_flight[slot] = true;
  return 0;
}

int qmb200_policy_wait_stream(qmb200_ctx* c, void* stream) {
  if (!c) return fail("null ctx");
  if (!c->policy[0]) return 0;
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_comm[c->pslot], 0));
  return 0;
}

int qmb200_comm_sync(qmb200_ctx* c) {
  if (!c) return fail("null ctx");
  if (!c->comm_stream) return 0;
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamSynchronize(c->comm_stream));
  return 0;
}

int qmb200_mpc_reset(qmb200_ctx* c) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaMemsetAsync(c->m.nprev, 0, sizeof(int32_t) * c->B, c->stream));
  return 0;
}

int qmb200_sync(qmb200_ctx* c) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  harvest_events(c);
  return 0;
}

static int wait_stream(int device, cudaStream_t self, void* other) {
  CUDA_OK(cudaSetDevice(device));
  cudaEvent_t e;
  CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaError_t err = cudaEventRecord(e, (cudaStream_t)other);
  if (err == cudaSuccess) err = cudaStreamWaitEvent(self, e, 0);
  cudaEventDestroy(e);                      // released once the recorded work has completed
  if (err != cudaSuccess) return fail(std::string("wait_stream: ") + cudaGetErrorString(err));
  return 0;
}

int qmb200_wait_stream(qmb200_ctx* c, void* stream) {
  if (!c) return fail("null ctx");
  return wait_stream(c->device, c->stream, stream);
}

#if defined(QM_PHASE_TIMING)
int qmb200_debug_ticks(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, qm::qm_dbg, sizeof(unsigned long long) * 64);
  if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(qm::qm_dbg, z, sizeof(z)); }
  return 0;
}
#endif

void* qmb200_stream(qmb200_ctx* c) { return c ? (void*)c->stream : nullptr; }
int64_t qmb200_device_bytes(qmb200_ctx* c) { return c ? c->bytes : 0; }

int qmb200_set_profiling(qmb200_ctx* c, int32_t enable) {
  if (!c) return fail("null ctx");
  c->profiling = enable != 0;
  return 0;
}

int qmb200_get_kernel_times(qmb200_ctx* c, double* total_ms, int64_t* launches, int32_t reset) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  harvest_events(c);
  for (int i = 0; i < QMB200_NUM_KERNELS; ++i) {
    if (total_ms) total_ms[i] = c->kernel_ms[i];
    if (launches) launches[i] = c->kernel_launches[i];
    if (reset) { c->kernel_ms[i] = 0.0; c->kernel_launches[i] = 0; }
  }
  return 0;
}

int qmb200_mpc_cycle_batch_dev(qmb200_ctx* c, const double* t0, const double* x0, const double* events, const int32_t* modes,
                               const int32_t* nevents, const double* target_t, const double* target_x, double* t_out,
                               double* x_out, double* u_out, int32_t* n_out, int32_t* mode_out, double* info, int32_t* status) {
  if (!c) return fail("null ctx");
  if (!t0 || !x0 || !events || !modes || !nevents || !target_t || !target_x) return fail("qmb200_mpc_cycle_batch_dev: null input");
  CUDA_OK(cudaSetDevice(c->device));
  MpcBuffers m = c->m;
  m.t0 = (double*)t0; m.x0 = (double*)x0; m.events = (double*)events; m.modes = (int32_t*)modes; m.nevents = (int32_t*)nevents;
  m.target_t = (double*)target_t; m.target_x = (double*)target_x;
  // keep the schedule for evaluate_policy (mode lookup)
  CUDA_OK(cudaMemcpyAsync(c->m.events, events, sizeof(double) * c->B * m.EMAX, cudaMemcpyDeviceToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->m.modes, modes, sizeof(int32_t) * c->B * (m.EMAX + 1), cudaMemcpyDeviceToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->m.nevents, nevents, sizeof(int32_t) * c->B, cudaMemcpyDeviceToDevice, c->stream));
  if (run_cycle(c, m, t_out, x_out, u_out) != 0) return -1;
  const size_t BN = (size_t)c->B * m.NMAX;
  if (n_out) CUDA_OK(cudaMemcpyAsync(n_out, m.nn, sizeof(int32_t) * c->B, cudaMemcpyDeviceToDevice, c->stream));
  if (mode_out) CUDA_OK(cudaMemcpyAsync(mode_out, m.node_mode, sizeof(int32_t) * BN, cudaMemcpyDeviceToDevice, c->stream));
  if (info) CUDA_OK(cudaMemcpyAsync(info, m.ls, sizeof(double) * c->B * LS_SIZE, cudaMemcpyDeviceToDevice, c->stream));
  if (status) CUDA_OK(cudaMemcpyAsync(status, m.status, sizeof(int32_t) * c->B, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

int qmb200_mpc_cycle_batch_async(qmb200_ctx* c, const double* t0, const double* x0, const double* events, const int32_t* modes,
                                 const int32_t* nevents, const double* target_t, const double* target_x, double* t_out, double* x_out,
                                 double* u_out, int32_t* n_out, int32_t* mode_out, double* info, int32_t* status, int64_t* ticket) {
  if (!c) return fail("null ctx");
  if (!t0 || !x0 || !events || !modes || !nevents || !target_t || !target_x) return fail("qmb200_mpc_cycle_batch_async: null input");
  CUDA_OK(cudaSetDevice(c->device));
  MpcBuffers& m = c->m;
  const size_t B = c->B, BN = B * m.NMAX;
  cudaStream_t st = c->stream;
  CUDA_OK(cudaMemcpyAsync(m.t0, t0, sizeof(double) * B, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(m.x0, x0, sizeof(double) * B * 30, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(m.events, events, sizeof(double) * B * m.EMAX, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(m.modes, modes, sizeof(int32_t) * B * (m.EMAX + 1), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(m.nevents, nevents, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(m.target_t, target_t, sizeof(double) * B * m.KT, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(m.target_x, target_x, sizeof(double) * B * m.KT * QM_NTARGET, cudaMemcpyHostToDevice, st));
  if (run_cycle(c, m, nullptr, nullptr, nullptr) != 0) return -1;
  // snapshot of the records the next cycle overwrites before its k_finalize
  CUDA_OK(cudaMemcpyAsync(c->snap_n, m.nn, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->snap_mode, m.node_mode, sizeof(int32_t) * BN, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->snap_ls, m.ls, sizeof(double) * B * LS_SIZE, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->snap_status, m.status, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
  const int slot = (int)(c->ticket & 1);
  CUDA_OK(cudaEventRecord(c->ev_cycle[slot], st));
  // results leave on the copy stream; the compute stream is free for the next cycle
  cudaStream_t cp = c->copy;
  CUDA_OK(cudaStreamWaitEvent(cp, c->ev_cycle[slot], 0));
  if (t_out) CUDA_OK(cudaMemcpyAsync(t_out, m.prev_t, sizeof(double) * BN, cudaMemcpyDeviceToHost, cp));
  if (x_out) CUDA_OK(cudaMemcpyAsync(x_out, m.prev_x, sizeof(double) * BN * 30, cudaMemcpyDeviceToHost, cp));
  if (u_out) CUDA_OK(cudaMemcpyAsync(u_out, m.prev_u, sizeof(double) * BN * 30, cudaMemcpyDeviceToHost, cp));
  if (n_out) CUDA_OK(cudaMemcpyAsync(n_out, c->snap_n, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, cp));
  if (mode_out) CUDA_OK(cudaMemcpyAsync(mode_out, c->snap_mode, sizeof(int32_t) * BN, cudaMemcpyDeviceToHost, cp));
  if (info) CUDA_OK(cudaMemcpyAsync(info, c->snap_ls, sizeof(double) * B * LS_SIZE, cudaMemcpyDeviceToHost, cp));
  if (status) CUDA_OK(cudaMemcpyAsync(status, c->snap_status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, cp));
  CUDA_OK(cudaEventRecord(c->ev_d2h[slot], cp));
  // the snapshot buffers are rewritten by the next submission: order it behind this copy (the policy buffers are guarded in run_back)
  c->d2h_in_flight = true;
  if (ticket) *ticket = c->ticket;
  c->ticket++;
  return 0;
}

int qmb200_mpc_cycle_wait(qmb200_ctx* c, int64_t ticket) {
  if (!c) return fail("null ctx");
  if (ticket < 0 || ticket >= c->ticket) return fail("qmb200_mpc_cycle_wait: unknown ticket");
  if (ticket < c->ticket - 2) return 0;      // older than the two tracked submissions: the copy stream is in order, long complete
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaEventSynchronize(c->ev_d2h[ticket & 1]));
  if (ticket == c->ticket - 1) harvest_events(c);
  return 0;
}

int qmb200_mpc_cycle_batch(qmb200_ctx* c, const double* t0, const double* x0, const double* events, const int32_t* modes,
                           const int32_t* nevents, const double* target_t, const double* target_x, double* t_out, double* x_out,
                           double* u_out, int32_t* n_out, int32_t* mode_out, double* info, int32_t* status) {
  int64_t ticket = 0;
  if (qmb200_mpc_cycle_batch_async(c, t0, x0, events, modes, nevents, target_t, target_x, t_out, x_out, u_out, n_out, mode_out, info,
                                   status, &ticket) != 0) return -1;
  if (qmb200_mpc_cycle_wait(c, ticket) != 0) return -1;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  harvest_events(c);
  return 0;
}

int qmb200_targets_batch_dev(qmb200_ctx* c, const qmb200_target_desc* desc, int32_t kind, int32_t n, const double* cmd,
                             const double* obs_time, const double* obs_state, const double* ee_state, double* last_ee_target,
                             double* target_t, double* target_x) {
  if (!c || !desc || !cmd || !obs_time || !obs_state || !ee_state || !last_ee_target || !target_t || !target_x || n <= 0)
    return fail("qmb200_targets_batch_dev: bad argument");
  if (kind < 0 || kind > 2) return fail("qmb200_targets_batch_dev: kind must be 0 (base cmd_vel), 1 (ee cmd_vel) or 2 (ee goal)");
  CUDA_OK(cudaSetDevice(c->device));
  k_targets<<<(n + 127) / 128, 128, 0, c->stream>>>(n, kind, *desc, cmd, obs_time, obs_state, ee_state, last_ee_target, target_t, target_x);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int qmb200_targets_batch(qmb200_ctx* c, const qmb200_target_desc* desc, int32_t kind, int32_t n, const double* cmd,
                         const double* obs_time, const double* obs_state, const double* ee_state, double* last_ee_target,
                         double* target_t, double* target_x) {
  if (!c || !desc || !cmd || !obs_time || !obs_state || !ee_state || !last_ee_target || !target_t || !target_x || n <= 0)
    return fail("qmb200_targets_batch: bad argument");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t N = (size_t)n;
  const size_t sizes[7] = {7 * N, N, 30 * N, 7 * N, 7 * N, 2 * N, 2 * QM_NTARGET * N};
  const double* src[5] = {cmd, obs_time, obs_state, ee_state, last_ee_target};
  double* d[7] = {nullptr};
  for (int i = 0; i < 7; ++i) CUDA_OK(cudaMallocAsync(&d[i], sizeof(double) * sizes[i], c->stream));
  for (int i = 0; i < 5; ++i) CUDA_OK(cudaMemcpyAsync(d[i], src[i], sizeof(double) * sizes[i], cudaMemcpyHostToDevice, c->stream));
  if (qmb200_targets_batch_dev(c, desc, kind, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6]) != 0) return -1;
  CUDA_OK(cudaMemcpyAsync(last_ee_target, d[4], sizeof(double) * sizes[4], cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(target_t, d[5], sizeof(double) * sizes[5], cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(target_x, d[6], sizeof(double) * sizes[6], cudaMemcpyDeviceToHost, c->stream));
  for (int i = 0; i < 7; ++i) CUDA_OK(cudaFreeAsync(d[i], c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmb200_rbd_to_state_batch_dev(qmb200_ctx* c, int32_t n, const double* rbd, const double* yaw_last, double* x_out) {
  if (!c || !rbd || !x_out || n <= 0) return fail("qmb200_rbd_to_state_batch_dev: bad argument");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t smem = (size_t)kStateWarps * kStateWarpDoubles * sizeof(double);
  k_rbd_state<<<(n + kStateWarps - 1) / kStateWarps, 32 * kStateWarps, smem, c->stream>>>(n, c->dM, rbd, yaw_last, x_out);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int qmb200_rbd_to_state_batch(qmb200_ctx* c, int32_t n, const double* rbd, const double* yaw_last, double* x_out) {
  if (!c || !rbd || !x_out || n <= 0) return fail("qmb200_rbd_to_state_batch: bad argument");
  CUDA_OK(cudaSetDevice(c->device));
  double *d_rbd = nullptr, *d_yaw = nullptr, *d_x = nullptr;
  CUDA_OK(cudaMallocAsync(&d_rbd, sizeof(double) * 55 * (size_t)n, c->stream));
  CUDA_OK(cudaMallocAsync(&d_x, sizeof(double) * 30 * (size_t)n, c->stream));
  CUDA_OK(cudaMemcpyAsync(d_rbd, rbd, sizeof(double) * 55 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  if (yaw_last) {
    CUDA_OK(cudaMallocAsync(&d_yaw, sizeof(double) * (size_t)n, c->stream));
    CUDA_OK(cudaMemcpyAsync(d_yaw, yaw_last, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  if (qmb200_rbd_to_state_batch_dev(c, n, d_rbd, d_yaw, d_x) != 0) return -1;
  CUDA_OK(cudaMemcpyAsync(x_out, d_x, sizeof(double) * 30 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaFreeAsync(d_rbd, c->stream)); CUDA_OK(cudaFreeAsync(d_x, c->stream));
  if (d_yaw) CUDA_OK(cudaFreeAsync(d_yaw, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmb200_feedback_gains_dev(qmb200_ctx* c, double* K_out) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t n = (size_t)c->B * c->m.NMAX * 900;
  if (!c->fb_gains) {
    CUDA_OK(cudaMalloc(&c->fb_gains, n * sizeof(double)));
    CUDA_OK(cudaMemsetAsync(c->fb_gains, 0, n * sizeof(double), c->stream));
  }
  k_gains<<<dim3(c->m.NMAX, c->B), 128, 0, c->stream>>>(c->m, c->fb_gains);
  CUDA_OK(cudaGetLastError());
  if (K_out) CUDA_OK(cudaMemcpyAsync(K_out, c->fb_gains, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

int qmb200_feedback_gains(qmb200_ctx* c, double* K_out) {
  if (!c || !K_out) return fail("qmb200_feedback_gains: null argument");
  if (qmb200_feedback_gains_dev(c, nullptr) != 0) return -1;
  const size_t n = (size_t)c->B * c->m.NMAX * 900;
  CUDA_OK(cudaMemcpyAsync(K_out, c->fb_gains, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmb200_evaluate_feedback_policy_batch(qmb200_ctx* c, const double* t, const double* x, double* u_out, int32_t* mode) {
  if (!c || !t || !x || !u_out || !mode) return fail("qmb200_evaluate_feedback_policy_batch: null argument");
  if (!c->fb_gains) return fail("qmb200_evaluate_feedback_policy_batch: call qmb200_feedback_gains[_dev] after the cycle first");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t B = c->B;
  double *dt = nullptr, *dx = nullptr, *du = nullptr; int32_t* dm = nullptr;
  CUDA_OK(cudaMallocAsync(&dt, sizeof(double) * B, c->stream));
  CUDA_OK(cudaMallocAsync(&dx, sizeof(double) * B * 30, c->stream));
  CUDA_OK(cudaMallocAsync(&du, sizeof(double) * B * 30, c->stream));
  CUDA_OK(cudaMallocAsync(&dm, sizeof(int32_t) * B, c->stream));
  CUDA_OK(cudaMemcpyAsync(dt, t, sizeof(double) * B, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(dx, x, sizeof(double) * B * 30, cudaMemcpyHostToDevice, c->stream));
  k_policy_fb<<<(unsigned)B, 32, 0, c->stream>>>(c->m, c->fb_gains, dt, dx, du, dm);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(u_out, du, sizeof(double) * B * 30, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(mode, dm, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaFreeAsync(dt, c->stream)); CUDA_OK(cudaFreeAsync(dx, c->stream));
  CUDA_OK(cudaFreeAsync(du, c->stream)); CUDA_OK(cudaFreeAsync(dm, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmb200_evaluate_policy_batch_dev(qmb200_ctx* c, const double* t, double* x_des, double* u_des, int32_t* mode) {
  if (!c || !t || !x_des || !u_des || !mode) return fail("qmb200_evaluate_policy_batch_dev: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  { KernelTimer kt(c, KN_POLICY); k_policy<<<(unsigned)c->B, 64, 0, c->stream>>>(c->m, t, x_des, u_des, mode); }
  CUDA_OK(cudaGetLastError());
  return 0;
}

int qmb200_evaluate_policy_batch(qmb200_ctx* c, const double* t, double* x_des, double* u_des, int32_t* mode) {
  if (!c || !t || !x_des || !u_des || !mode) return fail("qmb200_evaluate_policy_batch: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t B = c->B;
  double *dt = nullptr, *dx = nullptr, *du = nullptr; int32_t* dm = nullptr;
  CUDA_OK(cudaMallocAsync(&dt, sizeof(double) * B, c->stream));
  CUDA_OK(cudaMallocAsync(&dx, sizeof(double) * B * 30, c->stream));
  CUDA_OK(cudaMallocAsync(&du, sizeof(double) * B * 30, c->stream));
  CUDA_OK(cudaMallocAsync(&dm, sizeof(int32_t) * B, c->stream));
  CUDA_OK(cudaMemcpyAsync(dt, t, sizeof(double) * B, cudaMemcpyHostToDevice, c->stream));
  { KernelTimer kt(c, KN_POLICY); k_policy<<<(unsigned)B, 64, 0, c->stream>>>(c->m, dt, dx, du, dm); }
  CUDA_OK(cudaMemcpyAsync(x_des, dx, sizeof(double) * B * 30, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(u_des, du, sizeof(double) * B * 30, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(mode, dm, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaFreeAsync(dt, c->stream)); CUDA_OK(cudaFreeAsync(dx, c->stream));
  CUDA_OK(cudaFreeAsync(du, c->stream)); CUDA_OK(cudaFreeAsync(dm, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // extern "C"

// ============================================================================================ whole-body controller
#include "qm_wbc.h"
#include "qm_actuator.h"
#include "qm_sim.h"

constexpr int kWbcInDoubles = 30 + 30 + 56 + 30;   // xd, ud, rbd (padded), u_last
static_assert(QMB200_WBC_LEVELS_SIZE == WBL_SIZE, "include/qmb200.h and qm_wbc.h agree on the per-level record");
constexpr size_t kWbcSmemBytes = (size_t)(WW_SIZE + kWbcInDoubles) * sizeof(double) + WI_SIZE * sizeof(int);

// CTA per solve: rigid-body dynamics of both configurations, task stack, 3-level hierarchical QP, torque recovery.
#ifndef QM_WBC_THREADS
#define QM_WBC_THREADS 128
#endif
__global__ void __launch_bounds__(QM_WBC_THREADS) k_wbc(int B, const qmb200_model_desc* M, const qmb200_wbc_desc* C, const double* xd,
                                              const double* ud, const double* rbd, const int32_t* mode, const double* period,
                                              const double* time, double* u_last, double* cold, double* cmd, int32_t* status) {
  const int b = blockIdx.x;
  if (b >= B) return;
  extern __shared__ double smem[];
  double* W = smem;
  double* in = smem + WW_SIZE;
  int* WI = (int*)(smem + WW_SIZE + kWbcInDoubles);
  for (int i = threadIdx.x; i < 30; i += blockDim.x) { in[i] = xd[30 * b + i]; in[30 + i] = ud[30 * b + i]; in[116 + i] = u_last[30 * b + i]; }
  for (int i = threadIdx.x; i < 55; i += blockDim.x) in[60 + i] = rbd[55 * b + i];
  __syncthreads();
  wbc_update(BlockGroup(), *M, *C, in, in + 30, in + 60, mode[b], period[b], time[b], in + 116, W, cold + (size_t)WC_SIZE * b, WI, cmd + 54 * b, status + b);
  for (int i = threadIdx.x; i < 30; i += blockDim.x) u_last[30 * b + i] = in[30 + i];   // inputLast_ = inputDesired
}


// ---- the same solve as a sequence of kernels (the default path of a batch). k_wbc holds 56 KB of shared memory and four warps per
// solve while the active-set iterations of its levels run on one warp, and co-resident CTAs in different phases evict each
// other's instructions (k_wbc is 250 KB of code: identical solves in lockstep run 1.45x faster than a mixed batch). Split at the
// iterations, every kernel is one phase for the whole batch, and the iteration runs warp per solve with 17 KB per solve:
//   k_wbc_tasks  dynamics of both configurations, task stack                      -> D0, F0, h_j, level table, task rows
//   k_wbc_level  [first] level 0, kernel basis; [later] x += Z z, kernel basis, next stacked basis;
//                then the products and the least-squares start of the next level with rows, or the torque recovery
//   k_wbc_gi     Goldfarb-Idnani iteration of the pending level, four solves per CTA
// The state of a solve between kernels is its workspace image in global memory (ranges below; 45 KB written, 11 KB read by the
// iteration) -- a few GB per 65 536 solves, a few percent of the time it buys.
constexpr int kWbcKeepB = WS_D;                       // between levels: persistent blocks, A Z, b, Gg, J, z
// k_wbc_level keeps D0 where k_wbc_tasks left it and writes D0 Z where k_wbc_gi reads it (global memory); it holds the workspace
// from WW_F0 up to the last two blocks of the window (the iteration's triangular factor, D0 Z): 21 KB of shared memory
constexpr size_t kWbcLevelSmemBytes = (size_t)(WS_RF - WW_F0) * sizeof(double) + WI_SIZE * sizeof(int);
constexpr int kGiWarpDoubles = ((GI_MEM_DOUBLES + 1) / 2) * 2;
constexpr int kGiWarpInts = ((GI_MEM_INTS + 3) / 4) * 4;
constexpr size_t kWbcGiSmemBytes = 4 * ((size_t)kGiWarpDoubles * sizeof(double) + (size_t)kGiWarpInts * sizeof(int));

// Solves ordered by contact pattern (counting sort, 16 buckets): the CTAs that share an SM then run the same branches of the task
// builder at the same time and share the instructions they fetch (k_wbc_tasks is bound by instruction fetch: 65 % hit rate in a
// mixed batch). The order inside a bucket is whatever the atomics give; it only decides which CTA handles which solve.
__global__ void __launch_bounds__(1024) k_wbc_order(int B, const int32_t* mode, int* perm) {
  __shared__ int cnt[16], off[16];
  if (threadIdx.x < 16) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) atomicAdd(&cnt[mode[b] & 15], 1);
  __syncthreads();
  if (threadIdx.x == 0) { int a = 0; for (int m = 0; m < 16; ++m) { off[m] = a; a += cnt[m]; } }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) perm[atomicAdd(&off[mode[b] & 15], 1)] = b;
}

__device__ __forceinline__ void wbc_copy(double* dst, const double* src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// D0, F0 and h_j are written where the later kernels read them (the solve's image in global memory); only the dynamics scratch
// (WW_SCR .. WA_END) is in shared memory: 29 KB per solve, seven solves per SM.
constexpr size_t kWbcTasksSmemBytes = (size_t)(WA_END - WW_SCR + kWbcInDoubles) * sizeof(double) + WI_SIZE * sizeof(int);
#ifndef QM_WBC_TASKS_CTAS
#define QM_WBC_TASKS_CTAS 7    // measured: 6 / 7 solves per SM = 21.1 / 20.8 ms per 65 536 solves (80 / 72 registers)
#endif
__global__ void __launch_bounds__(QM_WBC_THREADS, QM_WBC_TASKS_CTAS) k_wbc_tasks(int B, const qmb200_model_desc* M, const qmb200_wbc_desc* C, const double* xd,
                                                    const double* ud, const double* rbd, const int32_t* mode, const double* period,
                                                    const double* time, double* u_last, double* cold, double* state, int* istate,
                                                    const int* perm) {
  if ((int)blockIdx.x >= B) return;
  const int b = perm[blockIdx.x];
  extern __shared__ double smem[];
  double* W = smem - WW_SCR;                         // only the scratch offsets (WA_*) are in shared memory
  double* in = smem + (WA_END - WW_SCR);
  int* WI = (int*)(in + kWbcInDoubles);
  double* S = state + (size_t)WS_END * b;
  for (int i = threadIdx.x; i < 30; i += blockDim.x) { in[i] = xd[30 * b + i]; in[30 + i] = ud[30 * b + i]; in[116 + i] = u_last[30 * b + i]; }
  for (int i = threadIdx.x; i < 55; i += blockDim.x) in[60 + i] = rbd[55 * b + i];
  __syncthreads();
  QM_TICK(-1);
  wbc_dynamics(BlockGroup(), *M, *C, in + 60, in, in + 30, in + 116, period[b], W);
  QM_TICK(33);
  wbc_tasks(BlockGroup(), *M, *C, in + 30, mode[b] & 15, time[b], W, S, S + WW_D0, cold + (size_t)WC_SIZE * b, WI);
  QM_TICK(34);
  __syncthreads();
  for (int i = threadIdx.x; i < WI_SIZE; i += blockDim.x) istate[(size_t)WI_SIZE * b + i] = WI[i];
  for (int i = threadIdx.x; i < 30; i += blockDim.x) u_last[30 * b + i] = in[30 + i];   // inputLast_ = inputDesired
}

// Level 0 (Newton iteration on the least-squares form, one Householder triangularisation per pass) with one warp per solve. D0 and
// the level-0 equality rows are read in place (global memory). The matrix has 54 fixed rows plus the ACTIVE inequality rows: a
// first launch gives every solve room for kL0NarrowActive of them (22 KB per solve, ten solves per SM; the config-5 batch never
// has more than four active), a second launch with the full-size matrix (30 KB, seven per SM) redoes the solves that needed more
// -- every other CTA of it leaves at once.
constexpr int kL0NarrowActive = 10;                 // (QMB200_WBC_L0_ACTIVE overrides it: tests force the second launch with 0 or 1)
static size_t wbc_l0_smem(int active_cap) {
  return (size_t)(((l0_mem_doubles(L0_FIXED_ROWS + active_cap) + 1) / 2) * 2) * sizeof(double) + WI_SIZE * sizeof(int);
}
__global__ void __launch_bounds__(32) k_wbc_level0(int B, int wide, int active_cap, const double* cold, double* state, int* istate,
                                                   const int* perm) {
  if ((int)blockIdx.x >= B) return;
  const int b = perm[blockIdx.x], lane = threadIdx.x;
  int* SI = istate + (size_t)WI_SIZE * b;
  if (wide && SI[WI_SC + 18] != WSS_LEVEL0_WIDE) return;
  extern __shared__ double smem[];
  const int doubles = ((l0_mem_doubles(L0_FIXED_ROWS + active_cap) + 1) / 2) * 2;
  int* WI = (int*)(smem + doubles);
  double* S = state + (size_t)WS_END * b;
  const L0Mem lm = l0_mem_compact(smem, L0_FIXED_ROWS + active_cap);
  for (int i = lane; i < 56; i += 32) smem[i] = S[WW_F0 + i];
  for (int i = lane; i < WI_SIZE; i += 32) WI[i] = SI[i];
  __syncwarp();
  if (lane == 0) WI[WI_SC + 6] = 0;
  __syncwarp();
  QM_TICK(-1);
  const bool done = wbc_level0(WarpGroup(), lm, S + WW_D0, cold + (size_t)WC_SIZE * b, WI, active_cap);
  if (!done) { if (lane == 0) SI[WI_SC + 18] = WSS_LEVEL0_WIDE; return; }
  for (int i = lane; i < 56; i += 32) S[WW_V0 + i] = lm.V0[i];
  for (int i = lane; i < 36; i += 32) S[WW_X + i] = lm.X[i];
  if (lane == 0) { SI[WI_SC + 6] = WI[WI_SC + 6]; SI[WI_SC + 18] = WSS_NONE; }
}

#ifndef QM_WBC_LEVEL_CTAS
#define QM_WBC_LEVEL_CTAS 10  // measured per 65 536 solves: 5 / 6 / 7 solves per SM = 25.7 / 24.4 / 23.2 ms with D0 Z out of shared memory,
                              // 7 / 8 = 22.0 / 21.1 with the triangular factor out as well, 8 / 9 / 10 = 20.8 / 20.4 / 20.9 with one basis
                              // buffer (128 threads: 64 / 56 / 48 registers), 10 with 96 threads (64 registers) = 20.1
#endif
#ifndef QM_WBC_LEVEL_THREADS
#define QM_WBC_LEVEL_THREADS 96
#endif
__global__ void __launch_bounds__(QM_WBC_LEVEL_THREADS, QM_WBC_LEVEL_CTAS) k_wbc_level(int B, int first, const double* cold, double* state, int* istate, double* cmd,
                                                    int32_t* status, const int* perm) {
  if ((int)blockIdx.x >= B) return;
  const int b = perm[blockIdx.x];
  int* SI = istate + (size_t)WI_SIZE * b;
  if (!first && SI[WI_SC + 18] != WSS_ITERATION) return;          // finished in an earlier round
  extern __shared__ double smem[];
  double* W = smem - WW_F0;                          // workspace offsets from WW_F0 on are in shared memory; WW_D0 is never used
  int* WI = (int*)(smem + (WS_RF - WW_F0));
  double* S = state + (size_t)WS_END * b;
  const double* D0 = S + WW_D0;
  double* GG = S + WS_GG;
  const double* Wc = cold + (size_t)WC_SIZE * b;
  wbc_copy(W + WW_F0, S + WW_F0, (first ? WW_Z0 : kWbcKeepB) - WW_F0);      // first: F0, V0, h_j, x (level 0 is done)
  for (int i = threadIdx.x; i < WI_SIZE; i += blockDim.x) WI[i] = SI[i];
  __syncthreads();
  const BlockGroup g;
  if (first) wbc_solve_begin(g, W, D0, Wc, WI, nullptr, true);
  else wbc_solve_advance(g, W, Wc, WI);
  if (wbc_solve_prepare(g, W, D0, GG, Wc, WI, nullptr, true)) {
    wbc_copy(S + WW_F0, W + WW_F0, kWbcKeepB - WW_F0);
    for (int i = threadIdx.x; i < WI_SIZE; i += blockDim.x) SI[i] = WI[i];
  } else {
    wbc_solve_finish(g, W, D0, WI, cmd + 54 * (size_t)b, status + b);
    if (threadIdx.x == 0) SI[WI_SC + 18] = WSS_DONE;
  }
}

__global__ void __launch_bounds__(128) k_wbc_gi(int B, double* state, int* istate, const int* perm) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x * 4 + w >= B) return;
  const int b = perm[blockIdx.x * 4 + w];
  int* SI = istate + (size_t)WI_SIZE * b;
  if (SI[WI_SC + 18] != WSS_ITERATION) return;
  extern __shared__ double smem[];
  double* D = smem + (size_t)w * kGiWarpDoubles;
  int* I = (int*)(smem + 4 * (size_t)kGiWarpDoubles) + (size_t)w * kGiWarpInts;
  double* S = state + (size_t)WS_END * b;
  // D0 Z stays in the solve's image: read once per candidate scan (18 independent loads per row, L2 hits) -- half the shared
  // memory per solve, twice the solves per SM
  const GiMem gm = gi_mem_compact(D, I, S + WS_GG);
  const int n = SI[WI_SC + 16], nD0 = SI[WI_SC + 9];
  for (int i = lane; i < 56; i += 32) { gm.Gg[i] = S[WS_Gg + i]; gm.ign[i] = 0; }
  for (int i = lane; i < 324; i += 32) gm.J[i] = S[WS_J + i];
  if (lane < 18) gm.z[lane] = S[WS_Z + lane];
  if (lane == 0) *gm.status = 0;
  __syncwarp();
  gi_iterate(WarpGroup(), n, nD0, gm);
  if (lane < 18) S[WS_Z + lane] = gm.z[lane];
  if (lane == 0 && *gm.status) SI[WI_SC + 6] |= *gm.status;
}

// One solve with the per-level record of the hierarchy (HoQp accessors): diagnostic entry, not on the hot path
__global__ void __launch_bounds__(128) k_wbc_levels(const qmb200_model_desc* M, const qmb200_wbc_desc* C, const double* in60_55_30, int mode,
                                                     double period, double time, double* cold, double* cmd, int32_t* status, double* levels) {
  extern __shared__ double smem[];
  double* W = smem;
  double* in = smem + WW_SIZE;
  int* WI = (int*)(smem + WW_SIZE + kWbcInDoubles);
  for (int i = threadIdx.x; i < 145; i += blockDim.x) in[i] = in60_55_30[i];          // xd(30) ud(30) rbd(55) u_last(30)
  __syncthreads();
  wbc_update(BlockGroup(), *M, *C, in, in + 30, in + 60, mode, period, time, in + 115, W, cold, WI, cmd, status, levels);
}

// ---- control law + simulated actuator with transport delay: warp per problem, lane per joint
__global__ void __launch_bounds__(128) k_actuator(int B, qmb200_actuator_desc D, const int64_t* time_ns, long long period_ns,
                                                  const double* obs_time, const double* xd, const double* ud, const double* cmd,
                                                  const double* q, const double* v, long long* stamp, double* buf, int* hc,
                                                  double* last, double* tau, int32_t* status) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  if ((threadIdx.x & 31) == 0) status[b] = 0;
  __syncwarp();
  const size_t sb = (size_t)b;
  actuator_step(WarpGroup(), D, (long long)time_ns[b], period_ns, obs_time[b], xd + 30 * sb, ud + 30 * sb, cmd + 54 * sb, q + 18 * sb,
                v + 18 * sb, stamp + QMB200_ACT_CAPACITY * sb, buf + (size_t)QMB200_ACT_CAPACITY * 18 * ACT_NF * sb, hc + 2 * sb,
                last + 18 * ACT_NF * sb, tau + 18 * sb, status + b);
}

// ---- forward-dynamics step (stand-in for the simulator behind the actuator): CTA per problem
__global__ void __launch_bounds__(128) k_fwd_dyn(int B, const qmb200_model_desc* M, double gravity, const double* rbd, const double* tau,
                                                  const int32_t* mode, double dt, double beta, double* rbd_out, double* f_out, int32_t* status) {
  const int b = blockIdx.x;
  if (b >= B) return;
  extern __shared__ double smem[];
  double* in = smem + WW_SIZE;                       // rbd (55, padded), tau (18)
  for (int i = threadIdx.x; i < 55; i += blockDim.x) in[i] = rbd[55 * (size_t)b + i];
  for (int i = threadIdx.x; i < 18; i += blockDim.x) in[56 + i] = tau[18 * (size_t)b + i];
  __syncthreads();
  fd_step(BlockGroup(), *M, gravity, in, in + 56, mode[b], dt, beta, smem, rbd_out + 55 * (size_t)b, f_out + 12 * (size_t)b, status + b);
}

struct qmb200_wbc_ctx {
  // actuator state (allocated on first use): stamps, buffered commands, {head, count}, held command per joint
  long long* act_stamp = nullptr; double* act_buf = nullptr; int* act_hc = nullptr; double* act_last = nullptr;
  int device = 0, B = 0;
  double gravity = 9.81;
  qmb200_model_desc* dM = nullptr;
  qmb200_wbc_desc* dC = nullptr;
  double *xd = nullptr, *ud = nullptr, *rbd = nullptr, *period = nullptr, *time = nullptr, *u_last = nullptr, *cmd = nullptr;
  double* cold = nullptr;                       // [B][WC_SIZE] task rows read once per level (qm_wbc.h), L2-resident per solve
  int* perm = nullptr;                          // [B] solves ordered by contact pattern
  double* state = nullptr; int* istate = nullptr;   // [B][WS_END], [B][WI_SIZE]: workspace image of a solve between the split kernels
  bool split = true;                            // sequence of kernels (large batches) or the single kernel k_wbc; see qmb200_wbc_create
  int rounds = 2;                               // levels below level 0 of the task stack (one iteration kernel each)
  int l0_active = 10;                           // active inequality rows the first level-0 launch has room for
  int32_t *mode = nullptr, *status = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double total_ms = 0.0;
  int64_t launches = 0;
  bool pending = false;
};

static void wbc_harvest(qmb200_wbc_ctx* c) {
  if (c->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->e0, c->e1) == cudaSuccess) c->total_ms += ms;
    c->pending = false;
  }
}

static int wbc_launch(qmb200_wbc_ctx* c, const double* xd, const double* ud, const double* rbd, const int32_t* mode,
                      const double* period, const double* time, double* cmd, int32_t* status) {
  wbc_harvest(c);
  CUDA_OK(cudaEventRecord(c->e0, c->stream));
  if (c->split) {
    k_wbc_order<<<1, 1024, 0, c->stream>>>(c->B, mode, c->perm);
    k_wbc_tasks<<<c->B, QM_WBC_THREADS, kWbcTasksSmemBytes, c->stream>>>(c->B, c->dM, c->dC, xd, ud, rbd, mode, period, time, c->u_last, c->cold,
                                                                    c->state, c->istate, c->perm);
    k_wbc_level0<<<c->B, 32, wbc_l0_smem(c->l0_active), c->stream>>>(c->B, 0, c->l0_active, c->cold, c->state, c->istate, c->perm);
    k_wbc_level0<<<c->B, 32, wbc_l0_smem(WB_MAXW), c->stream>>>(c->B, 1, WB_MAXW, c->cold, c->state, c->istate, c->perm);
    k_wbc_level<<<c->B, QM_WBC_LEVEL_THREADS, kWbcLevelSmemBytes, c->stream>>>(c->B, 1, c->cold, c->state, c->istate, cmd, status, c->perm);
    for (int r = 0; r < c->rounds; ++r) {
      k_wbc_gi<<<(c->B + 3) / 4, 128, kWbcGiSmemBytes, c->stream>>>(c->B, c->state, c->istate, c->perm);
      k_wbc_level<<<c->B, QM_WBC_LEVEL_THREADS, kWbcLevelSmemBytes, c->stream>>>(c->B, 0, c->cold, c->state, c->istate, cmd, status, c->perm);
    }
  } else {
    k_wbc<<<c->B, QM_WBC_THREADS, kWbcSmemBytes, c->stream>>>(c->B, c->dM, c->dC, xd, ud, rbd, mode, period, time, c->u_last, c->cold, cmd, status);
  }
  CUDA_OK(cudaEventRecord(c->e1, c->stream));
  CUDA_OK(cudaGetLastError());
  c->pending = true;
  c->launches++;
  return 0;
}

extern "C" {

int qmb200_wbc_create(const qmb200_model_desc* model, const qmb200_wbc_desc* wbc, int32_t batch, int32_t device,
                      qmb200_wbc_ctx** out) {
  if (!model || !wbc || !out) return fail("qmb200_wbc_create: null argument");
  if (batch <= 0) return fail("qmb200_wbc_create: batch must be positive");
  if (model->nj != QM_NJ) return fail("qmb200_wbc_create: model must have 24 one-DoF joints");
  if (qmb200_device_count() <= 0) return fail("qmb200_wbc_create: no CUDA device available (this library has no CPU fallback)");
  CUDA_OK(cudaSetDevice(device));
  qmb200_wbc_ctx* c = new qmb200_wbc_ctx();
  c->device = device; c->B = batch; c->gravity = wbc->gravity;
#define C_OK(call) CREATE_OK(call, qmb200_wbc_destroy(c))
  C_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  C_OK(cudaEventCreate(&c->e0));
  C_OK(cudaEventCreate(&c->e1));
  C_OK(cudaMalloc(&c->dM, sizeof(*model)));
  C_OK(cudaMalloc(&c->dC, sizeof(*wbc)));
  C_OK(cudaMemcpy(c->dM, model, sizeof(*model), cudaMemcpyHostToDevice));
  C_OK(cudaMemcpy(c->dC, wbc, sizeof(*wbc), cudaMemcpyHostToDevice));
  const size_t B = batch;
  C_OK(cudaMalloc(&c->xd, B * 30 * sizeof(double)));
  C_OK(cudaMalloc(&c->ud, B * 30 * sizeof(double)));
  C_OK(cudaMalloc(&c->rbd, B * 55 * sizeof(double)));
  C_OK(cudaMalloc(&c->period, B * sizeof(double)));
  C_OK(cudaMalloc(&c->time, B * sizeof(double)));
  C_OK(cudaMalloc(&c->u_last, B * 30 * sizeof(double)));
  C_OK(cudaMalloc(&c->cmd, B * 54 * sizeof(double)));
  C_OK(cudaMalloc(&c->cold, B * WC_SIZE * sizeof(double)));
  {
    // Kernel sequence for batches beyond one wave of the single kernel (four solves per SM), the single kernel below that: a
    // solve's latency is the same on both paths, the sequence adds seven launches and the round trips of the workspace image
    // (measured, host buffers: B = 1: 0.31 vs 0.37 ms, B = 256: 0.64 vs 0.81 ms, B = 2048: 2.24 vs 1.68 ms).
    // QMB200_WBC_SPLIT = 0 / 1 forces one or the other.
    const char* env = getenv("QMB200_WBC_SPLIT");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    c->split = env ? (atoi(env) != 0) : (batch > 4 * sms);
    c->rounds = (wbc->mpc_variant == 2) ? WB_MAXLEV - 1 : 2;
  }
  if (c->split) {
    C_OK(cudaMalloc(&c->state, B * WS_END * sizeof(double)));
    C_OK(cudaMalloc(&c->istate, B * WI_SIZE * sizeof(int)));
    C_OK(cudaMalloc(&c->perm, B * sizeof(int)));
    C_OK(cudaFuncSetAttribute(k_wbc_tasks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbcTasksSmemBytes));
    C_OK(cudaFuncSetAttribute(k_wbc_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbcLevelSmemBytes));
    C_OK(cudaFuncSetAttribute(k_wbc_gi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbcGiSmemBytes));
    C_OK(cudaFuncSetAttribute(k_wbc_level0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbc_l0_smem(WB_MAXW)));
    const char* l0env = getenv("QMB200_WBC_L0_ACTIVE");
    c->l0_active = l0env ? atoi(l0env) : kL0NarrowActive;
    if (c->l0_active < 0) c->l0_active = 0;
    if (c->l0_active > WB_MAXW) c->l0_active = WB_MAXW;
  }
  C_OK(cudaMalloc(&c->mode, B * sizeof(int32_t)));
  C_OK(cudaMalloc(&c->status, B * sizeof(int32_t)));
  C_OK(cudaMemset(c->u_last, 0, B * 30 * sizeof(double)));
  C_OK(cudaFuncSetAttribute(k_wbc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbcSmemBytes));
  C_OK(cudaFuncSetAttribute(k_fwd_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbcSmemBytes));
#undef C_OK
  *out = c;
  return 0;
}

int qmb200_wbc_destroy(qmb200_wbc_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  void* ptrs[] = {c->dM, c->dC, c->xd, c->ud, c->rbd, c->period, c->time, c->u_last, c->cmd, c->cold, c->state, c->istate, c->perm, c->mode, c->status,
                  c->act_stamp, c->act_buf, c->act_hc, c->act_last};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (c->e0) cudaEventDestroy(c->e0);
  if (c->e1) cudaEventDestroy(c->e1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int qmb200_wbc_reset(qmb200_wbc_ctx* c) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaMemsetAsync(c->u_last, 0, (size_t)c->B * 30 * sizeof(double), c->stream));
  return 0;
}

int qmb200_wbc_set_gains(qmb200_wbc_ctx* c, const qmb200_wbc_desc* wbc) {
  if (!c || !wbc) return fail("qmb200_wbc_set_gains: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamSynchronize(c->stream));     // gains are snapshotted between solves, never mid-solve
  CUDA_OK(cudaMemcpy(c->dC, wbc, sizeof(*wbc), cudaMemcpyHostToDevice));
  c->rounds = (wbc->mpc_variant == 2) ? WB_MAXLEV - 1 : 2;
  return 0;
}

int qmb200_wbc_sync(qmb200_wbc_ctx* c) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  wbc_harvest(c);
  return 0;
}

void* qmb200_wbc_stream(qmb200_wbc_ctx* c) { return c ? (void*)c->stream : nullptr; }

int qmb200_wbc_wait_stream(qmb200_wbc_ctx* c, void* stream) {
  if (!c) return fail("null ctx");
  return wait_stream(c->device, c->stream, stream);
}

int qmb200_wbc_kernel_time(qmb200_wbc_ctx* c, double* total_ms, int64_t* launches, int32_t reset) {
  if (!c) return fail("null ctx");
  if (qmb200_wbc_sync(c) != 0) return -1;
  if (total_ms) *total_ms = c->total_ms;
  if (launches) *launches = c->launches;
  if (reset) { c->total_ms = 0.0; c->launches = 0; }
  return 0;
}

int qmb200_wbc_levels(qmb200_wbc_ctx* c, const double* x_des, const double* u_des, const double* rbd, int32_t mode, double period, double time,
                      const double* u_last, double* cmd, int32_t* status, double* levels) {
  if (!c || !x_des || !u_des || !rbd || !u_last || !cmd || !status || !levels) return fail("qmb200_wbc_levels: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  double h_in[145];
  memcpy(h_in, x_des, 30 * sizeof(double)); memcpy(h_in + 30, u_des, 30 * sizeof(double));
  memcpy(h_in + 60, rbd, 55 * sizeof(double)); memcpy(h_in + 115, u_last, 30 * sizeof(double));
  double* d = nullptr; int32_t* ds = nullptr;
  const size_t nd = 145 + 54 + QMB200_WBC_LEVELS_SIZE;
  CUDA_OK(cudaMallocAsync(&d, nd * sizeof(double), c->stream));
  CUDA_OK(cudaMallocAsync(&ds, sizeof(int32_t), c->stream));
  CUDA_OK(cudaMemcpyAsync(d, h_in, sizeof(h_in), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaFuncSetAttribute(k_wbc_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWbcSmemBytes));
  k_wbc_levels<<<1, 128, kWbcSmemBytes, c->stream>>>(c->dM, c->dC, d, mode, period, time, c->cold, d + 145, ds, d + 199);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(cmd, d + 145, 54 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(levels, d + 199, QMB200_WBC_LEVELS_SIZE * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(status, ds, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaFreeAsync(d, c->stream)); CUDA_OK(cudaFreeAsync(ds, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmb200_wbc_batch_dev(qmb200_wbc_ctx* c, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                         const double* period, const double* time, double* cmd, int32_t* status) {
  if (!c || !x_des || !u_des || !rbd || !mode || !period || !time || !cmd || !status) return fail("qmb200_wbc_batch_dev: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  return wbc_launch(c, x_des, u_des, rbd, mode, period, time, cmd, status);
}

int qmb200_wbc_batch(qmb200_wbc_ctx* c, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                     const double* period, const double* time, double* cmd, int32_t* status) {
  if (!c || !x_des || !u_des || !rbd || !mode || !period || !time || !cmd) return fail("qmb200_wbc_batch: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t B = c->B;
  cudaStream_t st = c->stream;
  CUDA_OK(cudaMemcpyAsync(c->xd, x_des, B * 30 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->ud, u_des, B * 30 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->rbd, rbd, B * 55 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->mode, mode, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->period, period, B * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(c->time, time, B * sizeof(double), cudaMemcpyHostToDevice, st));
  if (wbc_launch(c, c->xd, c->ud, c->rbd, c->mode, c->period, c->time, c->cmd, c->status) != 0) return -1;
  CUDA_OK(cudaMemcpyAsync(cmd, c->cmd, B * 54 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (status) CUDA_OK(cudaMemcpyAsync(status, c->status, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  wbc_harvest(c);
  return 0;
}

static int actuator_state(qmb200_wbc_ctx* c, bool clear) {
  const size_t B = c->B, CAP = QMB200_ACT_CAPACITY;
  const bool fresh = !c->act_stamp;
  if (fresh) {
    CUDA_OK(cudaMalloc(&c->act_stamp, B * CAP * sizeof(long long)));
    CUDA_OK(cudaMalloc(&c->act_buf, B * CAP * 18 * ACT_NF * sizeof(double)));
    CUDA_OK(cudaMalloc(&c->act_hc, B * 2 * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->act_last, B * 18 * ACT_NF * sizeof(double)));
  }
  if (fresh || clear) {
    CUDA_OK(cudaMemsetAsync(c->act_stamp, 0, B * CAP * sizeof(long long), c->stream));
    CUDA_OK(cudaMemsetAsync(c->act_buf, 0, B * CAP * 18 * ACT_NF * sizeof(double), c->stream));
    CUDA_OK(cudaMemsetAsync(c->act_hc, 0, B * 2 * sizeof(int), c->stream));
    CUDA_OK(cudaMemsetAsync(c->act_last, 0, B * 18 * ACT_NF * sizeof(double), c->stream));
  }
  return 0;
}

int qmb200_forward_dynamics_batch_dev(qmb200_wbc_ctx* c, const double* rbd, const double* tau, const int32_t* mode, double dt, double beta,
                                      double* rbd_next, double* contact_forces, int32_t* status) {
  if (!c || !rbd || !tau || !mode || !rbd_next || !contact_forces || !status) return fail("qmb200_forward_dynamics_batch_dev: null argument");
  if (!(dt > 0.0) || beta < 0.0) return fail("qmb200_forward_dynamics_batch_dev: dt must be positive, beta non-negative");
  CUDA_OK(cudaSetDevice(c->device));
  k_fwd_dyn<<<c->B, 128, kWbcSmemBytes, c->stream>>>(c->B, c->dM, c->gravity, rbd, tau, mode, dt, beta, rbd_next, contact_forces, status);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int qmb200_forward_dynamics_batch(qmb200_wbc_ctx* c, const double* rbd, const double* tau, const int32_t* mode, double dt, double beta,
                                  double* rbd_next, double* contact_forces, int32_t* status) {
  if (!c || !rbd || !tau || !mode || !rbd_next || !contact_forces) return fail("qmb200_forward_dynamics_batch: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t B = c->B;
  cudaStream_t st = c->stream;
  double* d = nullptr; int32_t* di = nullptr;
  CUDA_OK(cudaMallocAsync(&d, B * (55 + 18 + 55 + 12) * sizeof(double), st));
  CUDA_OK(cudaMallocAsync(&di, B * 2 * sizeof(int32_t), st));
  double *dr = d, *dt_ = dr + 55 * B, *dn = dt_ + 18 * B, *df = dn + 55 * B;
  CUDA_OK(cudaMemcpyAsync(dr, rbd, B * 55 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(dt_, tau, B * 18 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(di, mode, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (qmb200_forward_dynamics_batch_dev(c, dr, dt_, di, dt, beta, dn, df, di + B) != 0) return -1;
  CUDA_OK(cudaMemcpyAsync(rbd_next, dn, B * 55 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(contact_forces, df, B * 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (status) CUDA_OK(cudaMemcpyAsync(status, di + B, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaFreeAsync(d, st));
  CUDA_OK(cudaFreeAsync(di, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int qmb200_actuator_reset(qmb200_wbc_ctx* c) {
  if (!c) return fail("null ctx");
  CUDA_OK(cudaSetDevice(c->device));
  return actuator_state(c, true);
}

int qmb200_actuator_batch_dev(qmb200_wbc_ctx* c, const qmb200_actuator_desc* desc, const int64_t* time_ns, int64_t period_ns,
                              const double* obs_time, const double* x_des, const double* u_des, const double* cmd, const double* q,
                              const double* v, double* tau, int32_t* status) {
  if (!c || !desc || !time_ns || !obs_time || !x_des || !u_des || !cmd || !q || !v || !tau || !status)
    return fail("qmb200_actuator_batch_dev: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  if (actuator_state(c, false) != 0) return -1;
  k_actuator<<<(c->B + 3) / 4, 128, 0, c->stream>>>(c->B, *desc, time_ns, (long long)period_ns, obs_time, x_des, u_des, cmd, q, v,
                                                    c->act_stamp, c->act_buf, c->act_hc, c->act_last, tau, status);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int qmb200_actuator_batch(qmb200_wbc_ctx* c, const qmb200_actuator_desc* desc, const int64_t* time_ns, int64_t period_ns,
                          const double* obs_time, const double* x_des, const double* u_des, const double* cmd, const double* q,
                          const double* v, double* tau, int32_t* status) {
  if (!c || !desc || !time_ns || !obs_time || !x_des || !u_des || !cmd || !q || !v || !tau) return fail("qmb200_actuator_batch: null argument");
  CUDA_OK(cudaSetDevice(c->device));
  const size_t B = c->B;
  cudaStream_t st = c->stream;
  // staging: [time_ns | obs_time | x_des | u_des | cmd | q | v | tau] doubles (time_ns as 8-byte integers), then the status words
  const size_t nd = B * (1 + 1 + 30 + 30 + 54 + 18 + 18 + 18);
  double* d = nullptr; int32_t* ds = nullptr;
  CUDA_OK(cudaMallocAsync(&d, nd * sizeof(double), st));
  CUDA_OK(cudaMallocAsync(&ds, B * sizeof(int32_t), st));
  double *dt = d, *dobs = dt + B, *dx = dobs + B, *du = dx + 30 * B, *dc = du + 30 * B, *dq = dc + 54 * B, *dv = dq + 18 * B, *dtau = dv + 18 * B;
  CUDA_OK(cudaMemcpyAsync(dt, time_ns, B * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(dobs, obs_time, B * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(dx, x_des, B * 30 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(du, u_des, B * 30 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(dc, cmd, B * 54 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(dq, q, B * 18 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(dv, v, B * 18 * sizeof(double), cudaMemcpyHostToDevice, st));
  if (qmb200_actuator_batch_dev(c, desc, (const int64_t*)dt, period_ns, dobs, dx, du, dc, dq, dv, dtau, ds) != 0) return -1;
  CUDA_OK(cudaMemcpyAsync(tau, dtau, B * 18 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (status) CUDA_OK(cudaMemcpyAsync(status, ds, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaFreeAsync(d, st));
  CUDA_OK(cudaFreeAsync(ds, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
