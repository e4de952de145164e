// Batch buffers of the MPC cycle. The same layout is used in HBM by the CUDA path and on the heap by the CPU port.
// Node axis has capacity NMAX per problem; all per-node arrays are [B][NMAX][...] row-major.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "qm_mpc.h"
#include "qm_value.h"

namespace qm {

// per-problem line-search / summary record (doubles)
enum {
  LS_ALPHA = 0,      // step size being tried / accepted (0 if rejected)
  LS_DONE = 1,       // 0 pending, 1 accepted, 2 rejected (no step)
  LS_ARMIJO = 2,
  LS_DXNORM = 3,
  LS_DUNORM = 4,
  LS_BASE_MERIT = 5, LS_BASE_DYN = 6, LS_BASE_EQ = 7,
  LS_NEW_MERIT = 8, LS_NEW_DYN = 9, LS_NEW_EQ = 10,
  LS_ITERS = 11,
  LS_DX0SQ = 12,     // |x0 - xs[0]|^2
  LS_SQP_ITERS = 13, // SQP iterations carried out in this cycle
  LS_CONV = 14,      // why the SQP loop stopped (CV_*)
  LS_SIZE = 16
};
// [upstream] SqpSolver::Convergence
enum { CV_NONE = 0, CV_ITERATIONS = 1, CV_STEPSIZE = 2, CV_METRICS = 3, CV_PRIMAL = 4 };

struct MpcBuffers {
  int B, NMAX, EMAX, KT;
  int b0, nb;          // chunk of problems [b0, b0 + nb) a kernel launch works on (the cycle is pipelined over chunks)
  // inputs (per cycle)
  double* t0;          // [B]
  double* x0;          // [B][30]
  double* events;      // [B][EMAX]
  int32_t* modes;      // [B][EMAX+1]
  int32_t* nevents;    // [B]
  double* target_t;    // [B][KT]
  double* target_x;    // [B][KT][37]
  // schedule
  double* node_t;      // [B][NMAX] annotated node time
  double* node_ts;     // [B][NMAX] interval start (post-event nodes shifted by weak_eps) = interpolation time
  double* node_dt;     // [B][NMAX]
  double* node_zvel;   // [B][NMAX][4]
  int32_t* node_flag;  // [B][NMAX]
  int32_t* node_mode;  // [B][NMAX]
  int32_t* nn;         // [B] number of nodes
  int32_t* status;     // [B]
  int32_t* conv;       // [B] CV_NONE while the problem is still iterating in this cycle, else why it stopped
  // iterate and step
  double* xs;          // [B][NMAX][30]
  double* us;          // [B][NMAX][30]
  double* dxs;         // [B][NMAX][30]
  double* dus;         // [B][NMAX][30]
  // LQ blocks
  double* stage;       // [B][NMAX][SB_SIZE]
  double* proj;        // [B][NMAX][PB_SIZE]
  double* gain;        // [B][NMAX][GB_SIZE]
  double* kin;         // [B][NMAX][KS_SIZE] kinematics products handed from k_kin1/k_kin2 to k_lq
  double* perf_base;   // [B][NMAX][PF_SIZE]
  double* perf_trial;  // [B][NMAX][PF_SIZE]
  double* ls;          // [B][LS_SIZE]
  // warm start (previous primal solution, [upstream] PrimalSolution)
  double* prev_t;      // [B][NMAX]
  double* prev_x;      // [B][NMAX][30]
  double* prev_u;      // [B][NMAX][30]
  int32_t* nprev;      // [B]
};

template <class F>
inline void for_each_buffer(MpcBuffers& m, F f) {
  const size_t B = m.B, N = m.NMAX, E = m.EMAX, K = m.KT;
  f((void**)&m.t0, B * sizeof(double));
  f((void**)&m.x0, B * 30 * sizeof(double));
  f((void**)&m.events, B * E * sizeof(double));
  f((void**)&m.modes, B * (E + 1) * sizeof(int32_t));
  f((void**)&m.nevents, B * sizeof(int32_t));
  f((void**)&m.target_t, B * K * sizeof(double));
  f((void**)&m.target_x, B * K * QM_NTARGET * sizeof(double));
  f((void**)&m.node_t, B * N * sizeof(double));
  f((void**)&m.node_ts, B * N * sizeof(double));
  f((void**)&m.node_dt, B * N * sizeof(double));
  f((void**)&m.node_zvel, B * N * 4 * sizeof(double));
  f((void**)&m.node_flag, B * N * sizeof(int32_t));
  f((void**)&m.node_mode, B * N * sizeof(int32_t));
  f((void**)&m.nn, B * sizeof(int32_t));
  f((void**)&m.status, B * sizeof(int32_t));
  f((void**)&m.conv, B * sizeof(int32_t));
  f((void**)&m.xs, B * N * 30 * sizeof(double));
  f((void**)&m.us, B * N * 30 * sizeof(double));
  f((void**)&m.dxs, B * N * 30 * sizeof(double));
  f((void**)&m.dus, B * N * 30 * sizeof(double));
  f((void**)&m.stage, B * N * SB_SIZE * sizeof(double));
  f((void**)&m.proj, B * N * PB_SIZE * sizeof(double));
  f((void**)&m.gain, B * N * GB_SIZE * sizeof(double));
  f((void**)&m.kin, B * N * KS_SIZE * sizeof(double));
  f((void**)&m.perf_base, B * N * PF_SIZE * sizeof(double));
  f((void**)&m.perf_trial, B * N * PF_SIZE * sizeof(double));
  f((void**)&m.ls, B * LS_SIZE * sizeof(double));
  f((void**)&m.prev_t, B * N * sizeof(double));
  f((void**)&m.prev_x, B * N * 30 * sizeof(double));
  f((void**)&m.prev_u, B * N * 30 * sizeof(double));
  f((void**)&m.nprev, B * sizeof(int32_t));
}

// ---- per-problem steps that are serial in the node axis (one group per problem)

// How solve_problem gets at the per-node blocks: in place (host / plain loads) or staged into shared memory by the kernel
// (k_solve: cp.async.bulk + mbarrier). Backward: one stage block in flight; forward: two slots (the blocks of stage
// k + 1 are requested before stage k is processed).
enum { FWD_SLOT_SIZE = SB_FWD_SIZE + PB_SIZE + GB_SIZE };
struct DirectFetch {
  const double* bp;
  const double* fp[2][3];
  template <class G> QM_HD void bwd_request(G, const double* stage) { bp = stage; }
  template <class G> QM_HD const double* bwd_wait(G) { return bp; }
  template <class G> QM_HD void publish(G) {}       // gains written by this group become visible to its own fetches
  template <class G> QM_HD void fwd_request(G, int slot, const double* stage, const double* proj, const double* gain) {
    fp[slot][0] = stage; fp[slot][1] = proj; fp[slot][2] = gain;
  }
  template <class G> QM_HD void fwd_wait(G, int slot, const double** st, const double** pb, const double** gb) {
    *st = fp[slot][0]; *pb = fp[slot][1]; *gb = fp[slot][2];
  }
};

// Backward Riccati sweep, forward rollout, step norms, baseline performance reduction.
// W: Riccati workspace (>= RW_SIZE doubles); R: forward scratch (>= 96 doubles; may alias W when nothing is staged over it).
template <class G, class F>
QM_HDN void solve_problem(G g, F& fetch, const MpcBuffers& m, int b, double* W, double* R) {
  const int NMAX = m.NMAX;
  const int nn = m.nn[b];
  const int n = nn - 1;
  const double* stage = m.stage + (size_t)b * NMAX * SB_SIZE;
  const double* proj = m.proj + (size_t)b * NMAX * PB_SIZE;
  double* gain = m.gain + (size_t)b * NMAX * GB_SIZE;
  double* dxs = m.dxs + (size_t)b * NMAX * 30;
  double* dus = m.dus + (size_t)b * NMAX * 30;
  const double* term = stage + (size_t)n * SB_SIZE;
  if (n > 0) fetch.bwd_request(g, stage + (size_t)(n - 1) * SB_SIZE);
  QM_PFOR(g, idx, 900) W[RW_S + idx] = term[SB_Q + idx];
  QM_PFOR(g, i, 30) W[RW_sv + i] = term[SB_q + i];
  g.sync();
  for (int k = n - 1; k >= 0; --k) {
    QM_TICK(-1);
    const double* st = fetch.bwd_wait(g);
    QM_TICK(0);
    const int nut = riccati_stage_a(g, st, W, m.status + b);
    if (k > 0) fetch.bwd_request(g, stage + (size_t)(k - 1) * SB_SIZE);      // stage buffer is free: prefetch
    riccati_stage_b(g, nut, W, gain + (size_t)k * GB_SIZE);
  }
  // forward rollout: [0:30] dx, [30:48] dut, [48:78] dx next, [80] armijo
  fetch.publish(g);
  g.sync();
  QM_TICK(-1);
  if (n > 0) fetch.fwd_request(g, 0, stage, proj, gain);
  QM_PFOR(g, i, 30) { R[i] = m.x0[30 * b + i] - m.xs[((size_t)b * NMAX) * 30 + i]; }
  if (g.tid() == 0) R[80] = 0.0;
  g.sync();
  double* cur = R;                                   // dx of the current node; the next one is written to the other buffer
  double* nxt = R + 48;
  for (int k = 0; k < n; ++k) {
    if (k + 1 < n) fetch.fwd_request(g, (k + 1) & 1, stage + (size_t)(k + 1) * SB_SIZE, proj + (size_t)(k + 1) * PB_SIZE, gain + (size_t)(k + 1) * GB_SIZE);
    QM_PFOR(g, i, 30) dxs[30 * k + i] = cur[i];
    const double *st, *pb, *gb;
    QM_TICK(10);
    fetch.fwd_wait(g, k & 1, &st, &pb, &gb);
    QM_TICK(11);
    rollout_stage(g, st, pb, gb, R, cur, nxt, dus + 30 * k);
    double* t_ = cur; cur = nxt; nxt = t_;
    QM_TICK(12);
  }
  QM_PFOR(g, i, 30) { dxs[30 * n + i] = cur[i]; dus[30 * n + i] = 0.0; }
  g.sync();
  // step norms and baseline performance: sums in a fixed order (bit-reproducible for a given group size): every thread sums
  // a strided subset of the 30 (n + 1) components, thread 0 adds the partial sums in thread order
  double* part = W + RW_SA;                         // 2 x nt partial sums (SA is idle; clear of the forward scratch R)
  {
    double px = 0.0, pu = 0.0;
    for (int idx = g.tid(); idx < 30 * (n + 1); idx += g.nt()) { px += dxs[idx] * dxs[idx]; pu += dus[idx] * dus[idx]; }
    part[2 * g.tid()] = px; part[2 * g.tid() + 1] = pu;
    // the baseline performance records of the nodes are staged by the whole group (one HBM round trip instead of one per node
    // on thread 0); their sums below stay sequential in the node index
    const double* pfg = m.perf_base + (size_t)b * NMAX * PF_SIZE;
    for (int idx = g.tid(); idx < PF_SIZE * (n + 1); idx += g.nt()) part[2 * g.nt() + idx] = pfg[idx];
  }
  g.sync();
  if (g.tid() == 0) {
    double arm = R[80];
    for (int j = 0; j < 30; ++j) arm += term[SB_q + j] * cur[j];
    double sx = 0.0, su = 0.0;
    for (int t = 0; t < g.nt(); ++t) { sx += part[2 * t]; su += part[2 * t + 1]; }
    double d0 = 0.0;
    for (int i = 0; i < 30; ++i) d0 += dxs[i] * dxs[i];
    const double* pf = part + 2 * g.nt();            // staged copy of the baseline performance records (see above)
    double c = 0.0, dy = d0, eq = 0.0;
    for (int k = 0; k <= n; ++k) { c += pf[PF_SIZE * k + PF_COST]; dy += pf[PF_SIZE * k + PF_DYN]; eq += pf[PF_SIZE * k + PF_EQ]; }
    double* ls = m.ls + (size_t)b * LS_SIZE;
    ls[LS_ALPHA] = 1.0; ls[LS_DONE] = 0.0; ls[LS_ARMIJO] = arm; ls[LS_DXNORM] = sqrt(sx); ls[LS_DUNORM] = sqrt(su);
    ls[LS_BASE_MERIT] = c; ls[LS_BASE_DYN] = dy; ls[LS_BASE_EQ] = eq; ls[LS_ITERS] = 0.0; ls[LS_DX0SQ] = d0;
    ls[LS_NEW_MERIT] = c; ls[LS_NEW_DYN] = dy; ls[LS_NEW_EQ] = eq;
    ls[LS_SQP_ITERS] += 1.0;                           // zeroed with the schedule at the start of the cycle
  }
  g.sync();
}

// Line-search decision for one problem (one thread). [upstream] SqpSolver::takeStep loop body.
// pf: the trial performance records of the problem's nodes (stride PF_SIZE): the HBM array itself or a staged copy
// (k_decide stages it with the whole warp; the sums stay sequential in the node index, so the result is the same).
QM_HDN void decide_problem(const qmb200_solver_desc& S, const MpcBuffers& m, int b, const double* pf = nullptr) {
  double* ls = m.ls + (size_t)b * LS_SIZE;
  if (ls[LS_DONE] != 0.0) return;
  const int nn = m.nn[b];
  const double alpha = ls[LS_ALPHA];
  if (pf == nullptr) pf = m.perf_trial + (size_t)b * m.NMAX * PF_SIZE;
  double c = 0.0, dy = (1.0 - alpha) * (1.0 - alpha) * ls[LS_DX0SQ], eq = 0.0;
  for (int k = 0; k < nn; ++k) { c += pf[PF_SIZE * k + PF_COST]; dy += pf[PF_SIZE * k + PF_DYN]; eq += pf[PF_SIZE * k + PF_EQ]; }
  ls[LS_ITERS] += 1.0;
  const double vb = sqrt(ls[LS_BASE_DYN] + ls[LS_BASE_EQ]), vn = sqrt(dy + eq);
  bool nan = !(c == c) || !(vn == vn);
  if (!nan && accept_step(S, ls[LS_BASE_MERIT], vb, c, vn, alpha * ls[LS_ARMIJO])) {
    ls[LS_DONE] = 1.0; ls[LS_NEW_MERIT] = c; ls[LS_NEW_DYN] = dy; ls[LS_NEW_EQ] = eq;
    return;
  }
  if (nan) m.status[b] |= ST_NAN;
  const double a2 = alpha * S.alpha_decay;
  if ((a2 * ls[LS_DXNORM] < S.delta_tol && a2 * ls[LS_DUNORM] < S.delta_tol) || a2 < S.alpha_min) {
    ls[LS_DONE] = 2.0; ls[LS_ALPHA] = 0.0; m.status[b] |= ST_STEP_REJECTED;
    return;
  }
  ls[LS_ALPHA] = a2;
}

// [upstream] SqpSolver::checkConvergence after iteration `it` (0-based) of `iterations`; ls: the problem's line-search record.
QM_HD int check_convergence(const qmb200_solver_desc& S, const double* ls, int it, int iterations) {
  if (it + 1 >= iterations) return CV_ITERATIONS;
  const double alpha = ls[LS_ALPHA];
  if (alpha < S.alpha_min) return CV_STEPSIZE;
  if (fabs(ls[LS_NEW_MERIT] - ls[LS_BASE_MERIT]) < S.cost_tol && sqrt(ls[LS_NEW_DYN] + ls[LS_NEW_EQ]) < S.g_min) return CV_METRICS;
  if (alpha * ls[LS_DXNORM] < S.delta_tol && alpha * ls[LS_DUNORM] < S.delta_tol) return CV_PRIMAL;
  return CV_NONE;
}

// Intermediate SQP iteration (sqpIteration > 1): take the accepted step in place, x += alpha dx, u += alpha du; one thread per
// (problem, component c < 60). The convergence test of the iteration is the caller's (one thread, after all components).
QM_HDN void step_component(const MpcBuffers& m, int b, int c) {
  const int NMAX = m.NMAX, nn = m.nn[b];
  const double alpha = m.ls[(size_t)b * LS_SIZE + LS_ALPHA];
  const size_t o = (size_t)b * NMAX;
  double* v = (c < 30) ? m.xs : m.us;
  const double* d = (c < 30) ? m.dxs : m.dus;
  const int cc = (c < 30) ? c : c - 30;
  for (int k = 0; k < nn; ++k) v[(o + k) * 30 + cc] += alpha * d[(o + k) * 30 + cc];
}

// Accept the step and publish the primal solution ([upstream] toPrimalSolution); one thread per (problem, component c<60).
// A problem whose SQP loop stopped in an earlier iteration has its step applied already (alpha_in = 0 is passed then).
// packed (may be null): the same policy as [B][NMAX][61] rows (t, x*[30], u*[30]), the send buffer of the multi-GPU all-gather.
QM_HDN void finalize_component(const MpcBuffers& m, int b, int c, double* t_out, double* x_out, double* u_out, double* packed = nullptr) {
  const int NMAX = m.NMAX, nn = m.nn[b];
  const double alpha = (m.conv[b] == CV_NONE) ? m.ls[(size_t)b * LS_SIZE + LS_ALPHA] : 0.0;
  const size_t o = (size_t)b * NMAX;
  constexpr int NB = 8;                      // nodes per batch: the loads of a batch are issued before its stores
  if (c < 30) {
    for (int k0 = 0; k0 < nn; k0 += NB) {
      double v[NB];
      QM_UNROLL
      for (int j = 0; j < NB; ++j) { const int k = (k0 + j < nn) ? k0 + j : nn - 1; v[j] = m.xs[(o + k) * 30 + c] + alpha * m.dxs[(o + k) * 30 + c]; }
      QM_UNROLL
      for (int j = 0; j < NB; ++j) {
        const int k = k0 + j;
        if (k < nn) {
          m.xs[(o + k) * 30 + c] = v[j]; m.prev_x[(o + k) * 30 + c] = v[j];
          if (x_out) x_out[(o + k) * 30 + c] = v[j];
          if (packed) packed[(o + k) * 61 + 1 + c] = v[j];
        }
      }
    }
    for (int k = c; k < nn; k += 30) {                                   // node times: spread over the state lanes
      const double tk = m.node_ts[o + k];
      m.prev_t[o + k] = tk;
      if (t_out) t_out[o + k] = tk;
      if (packed) packed[(o + k) * 61] = tk;
    }
    if (c == 0) m.nprev[b] = nn;
  } else {
    const int cu = c - 30;
    double last = 0.0;
    for (int k0 = 0; k0 < nn; k0 += NB) {
      double v[NB]; int fl[NB];
      QM_UNROLL
      for (int j = 0; j < NB; ++j) {
        const int k = (k0 + j < nn) ? k0 + j : nn - 1;
        v[j] = m.us[(o + k) * 30 + cu] + alpha * m.dus[(o + k) * 30 + cu]; fl[j] = m.node_flag[o + k];
      }
      QM_UNROLL
      for (int j = 0; j < NB; ++j) {
        const int k = k0 + j;
        if (k < nn) {
          double w;
          if (k == nn - 1) w = last;                                   // repeat the last input
          else if (fl[j] == EV_PRE && k > 0) w = last;                 // pre-event node repeats the previous input
          else w = v[j];
          m.us[(o + k) * 30 + cu] = w; m.prev_u[(o + k) * 30 + cu] = w;
          if (u_out) u_out[(o + k) * 30 + cu] = w;
          if (packed) packed[(o + k) * 61 + 31 + cu] = w;
          last = w;
        }
      }
    }
  }
}

}  // namespace qm
