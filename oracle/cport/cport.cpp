// CPU port of the MPC hot path (TEST INFRASTRUCTURE / CPU baseline only; see oracle/__init__.py).
// Host compilation of qm_door_b200/csrc/qm_core.h / qm_mpc.h / qm_buffers.h with the SerialGroup (every
// bulk-synchronous phase executed by one thread), threaded over independent problems with std::thread.
// Exposed to ctypes so tests can compare each routine with the NumPy oracle and bench.py can time a CPU
// baseline ("kind": "port" — not the reference binary, which cannot be built here).
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../qm_door_b200/csrc/qm_buffers.h"
#include "../../qm_door_b200/csrc/qm_target.h"

using namespace qm;

struct CportCtx {
  qmb200_model_desc M;
  qmb200_problem_desc P;
  qmb200_solver_desc S;
  MpcBuffers m;
  int threads;
  int node_threads = 1;   // worker threads over the nodes of one problem (the reference's sqp.nThreads, task.info:78)
};

template <class F>
static void parallel_for(int n, int threads, F f) {
  if (threads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) f(i); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([=]() { for (int i = t; i < n; i += threads) f(i); });
  for (auto& th : pool) th.join();
}

extern "C" {

int cport_kin_ws_size() { return KW_SIZE; }
// offsets of the kinematics workspace (qm_core.h KW_*), in the order of oracle/abi_fill.py::KW_NAMES
int cport_kw_offsets(int* out) {
  const int v[] = {KW_R, KW_P, KW_AX, KW_BODY, KW_COMP, KW_ACM, KW_SV, KW_V, KW_HB, KW_FPOS, KW_FVEL, KW_EEP, KW_EER, KW_COM, KW_ABINV,
                   KW_VEL, KW_RHS, KW_VSIZE, KW_FJ, KW_EEJ, KW_DH, KW_DFV, KW_F, KW_SIZE};
  const int n = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n; ++i) out[i] = v[i];
  return n;
}
int cport_tw_size() { return TW_SIZE; }
int cport_sizes(int* out) {
  out[0] = SB_SIZE; out[1] = PB_SIZE; out[2] = GB_SIZE; out[3] = PF_SIZE; out[4] = LS_SIZE; out[5] = TW_SIZE; out[6] = TI_SIZE;
  return 7;
}

void cport_kin_eval(const qmb200_model_desc* M, const double* x, const double* u, int deriv, double* w) {
  kin_eval(SerialGroup(), *M, x, u, deriv != 0 ? 2 : 0, w);
}

void cport_transcribe_node(const qmb200_model_desc* M, const qmb200_problem_desc* P, double t, double dt, int mode,
                           const double* zvel, const double* tt, const double* ts, int kt, const double* x, const double* u,
                           const double* xn, double* W, int* WI, double* sb, double* pb, double* perf, int* status) {
  transcribe_node(SerialGroup(), *M, *P, t, dt, mode, zvel, tt, ts, kt, x, u, xn, W, WI, sb, pb, perf, status);
}

void cport_targets(const qmb200_target_desc* D, int kind, int n, const double* cmd, const double* obs_time, const double* obs_state,
                   const double* ee_state, double* last_ee, double* tt, double* tx) {
  for (int b = 0; b < n; ++b)
    command_to_target(*D, kind, cmd + 7 * b, obs_time[b], obs_state + 30 * b, ee_state + 7 * b, last_ee + 7 * b, tt + 2 * b, tx + 2 * QM_NTARGET * b);
}

void cport_rbd_to_state(const qmb200_model_desc* M, int n, const double* rbd, const double* yaw_last, double* x_out) {
  std::vector<double> kw(KW_SIZE), qv(48);
  for (int b = 0; b < n; ++b)
    centroidal_state_from_rbd(SerialGroup(), *M, rbd + 55 * b, yaw_last ? yaw_last[b] : 0.0, yaw_last != nullptr, kw.data(), qv.data(), x_out + 30 * b);
}

CportCtx* cport_create(const qmb200_model_desc* M, const qmb200_problem_desc* P, const qmb200_solver_desc* S, int B, int threads) {
  CportCtx* c = new CportCtx();
  c->M = *M; c->P = *P; c->S = *S; c->threads = threads;
  c->m.B = B; c->m.NMAX = S->max_nodes; c->m.EMAX = S->max_events; c->m.KT = S->max_targets;
  for_each_buffer(c->m, [](void** p, size_t bytes) { *p = calloc(1, bytes); });
  return c;
}

void cport_destroy(CportCtx* c) {
  for_each_buffer(c->m, [](void** p, size_t) { free(*p); });
  delete c;
}

void cport_reset(CportCtx* c) { memset(c->m.nprev, 0, sizeof(int32_t) * c->m.B); }
void cport_set_node_threads(CportCtx* c, int n) { c->node_threads = n < 1 ? 1 : n; }

const MpcBuffers* cport_buffers(CportCtx* c) { return &c->m; }

// One MPC cycle for the whole batch. Outputs [B][NMAX][..]; info [B][LS_SIZE].
int cport_mpc_cycle(CportCtx* c, const double* t0, const double* x0, const double* events, const int32_t* modes,
                    const int32_t* nevents, const double* target_t, const double* target_x, double* t_out, double* x_out,
                    double* u_out, int32_t* n_out, int32_t* mode_out, double* info, int32_t* status) {
  MpcBuffers& m = c->m;
  const int B = m.B, NMAX = m.NMAX, E = m.EMAX, KT = m.KT;
  memcpy(m.t0, t0, sizeof(double) * B);
  memcpy(m.x0, x0, sizeof(double) * B * 30);
  memcpy(m.events, events, sizeof(double) * B * E);
  memcpy(m.modes, modes, sizeof(int32_t) * B * (E + 1));
  memcpy(m.nevents, nevents, sizeof(int32_t) * B);
  memcpy(m.target_t, target_t, sizeof(double) * B * KT);
  memcpy(m.target_x, target_x, sizeof(double) * B * KT * QM_NTARGET);
  const qmb200_model_desc& M = c->M; const qmb200_problem_desc& P = c->P; const qmb200_solver_desc& S = c->S;
  parallel_for(B, c->threads, [&](int b) {
    SerialGroup g;
    std::vector<double> W((int)TW_SIZE > (int)RW_SIZE ? (int)TW_SIZE : (int)RW_SIZE);
    std::vector<int> WI(TI_SIZE);
    const size_t o = (size_t)b * NMAX;
    build_grid(S, m.t0[b], m.events + (size_t)b * E, m.nevents[b], m.node_t + o, m.node_flag + o, m.nn + b, m.status + b);
    const int nn = m.nn[b], n = nn - 1;
    annotate_schedule(g, S, P, m.events + (size_t)b * E, m.modes + (size_t)b * (E + 1), m.nevents[b], nn, m.node_t + o,
                      m.node_flag + o, m.node_ts + o, m.node_dt + o, m.node_mode + o, m.node_zvel + o * 4, m.status + b);
    std::vector<int> gri(3 * NMAX);
    std::vector<double> gra(2 * NMAX);
    for (int k = 0; k < nn - 1; ++k)
      init_guess_node(S.weak_eps, k, m.node_t + o, m.node_flag + o, m.node_ts + o, m.nprev[b], m.prev_t + o, gri.data(), gra.data());
    for (int cc = 0; cc < 60; ++cc)
      init_guess_component(M, P, cc, m.x0 + 30 * b, nn, m.node_t + o, m.node_ts + o, m.node_mode + o, gri.data(), gra.data(), m.nprev[b],
                           m.prev_t + o, m.prev_x + o * 30, m.prev_u + o * 30, m.xs + o * 30, m.us + o * 30);
    const double* tt = m.target_t + (size_t)b * KT;
    const double* ts = m.target_x + (size_t)b * KT * QM_NTARGET;
    double* ls = m.ls + (size_t)b * LS_SIZE;
    std::vector<double> xt(30), ut(30), xnt(30);
    const int iterations = S.sqp_iterations < 1 ? 1 : S.sqp_iterations;
    m.conv[b] = CV_NONE; ls[LS_SQP_ITERS] = 0.0; ls[LS_CONV] = 0.0;
    const int nth = c->node_threads;
    std::vector<std::vector<double>> Wn(nth > 1 ? nth : 0, std::vector<double>(W.size()));
    for (int it = 0; it < iterations; ++it) {
    // transcription: independent over the nodes (the reference spreads them over sqp.nThreads workers)
    parallel_for(nth > 1 ? nth : 1, nth, [&](int tid) {
    double* Wd = (nth > 1) ? Wn[tid].data() : W.data();
    std::vector<int> WIn(TI_SIZE);
    for (int k = tid; k <= n; k += (nth > 1 ? nth : 1)) {
      double* sb = m.stage + (o + k) * SB_SIZE; double* pb = m.proj + (o + k) * PB_SIZE; double* pf = m.perf_base + (o + k) * PF_SIZE;
      const double* x = m.xs + (o + k) * 30; const double* u = m.us + (o + k) * 30; const double* xn = m.xs + (o + k + 1) * 30;
      if (k == n) terminal_node(g, M, P, m.node_t[o + k], m.node_mode[o + k], tt, ts, KT, x, Wd + TW_KIN, Wd + TW_REF,
                                 Wd + TW_E6, Wd + TW_DQ, Wd + TW_JE, sb, pf);
      else if (m.node_flag[o + k] == EV_PRE) event_node(g, x, xn, sb, pb, pf);
      else transcribe_node(g, M, P, m.node_ts[o + k], m.node_dt[o + k], m.node_mode[o + k], m.node_zvel + (o + k) * 4, tt, ts, KT,
                           x, u, xn, Wd, WIn.data(), sb, pb, pf, m.status + b);
    }
    });
    DirectFetch fetch;
    solve_problem(g, fetch, m, b, W.data(), W.data());
    // filter line search
    while (ls[LS_DONE] == 0.0) {
      const double alpha = ls[LS_ALPHA];
      parallel_for(nth > 1 ? nth : 1, nth, [&](int tid) {
      std::vector<double> xt(30), ut(30), xnt(30);
      for (int k = tid; k <= n; k += (nth > 1 ? nth : 1)) {
        double* pf = m.perf_trial + (o + k) * PF_SIZE;
        for (int i = 0; i < 30; ++i) {
          xt[i] = m.xs[(o + k) * 30 + i] + alpha * m.dxs[(o + k) * 30 + i];
          ut[i] = m.us[(o + k) * 30 + i] + alpha * m.dus[(o + k) * 30 + i];
          if (k < n) xnt[i] = m.xs[(o + k + 1) * 30 + i] + alpha * m.dxs[(o + k + 1) * 30 + i];
        }
        if (k == n) perf_terminal_serial(M, P, m.node_t[o + k], m.node_mode[o + k], tt, ts, KT, xt.data(), pf);
        else if (m.node_flag[o + k] == EV_PRE) {
          double d = 0.0;
          for (int i = 0; i < 30; ++i) d += (xt[i] - xnt[i]) * (xt[i] - xnt[i]);
          pf[PF_COST] = 0.0; pf[PF_DYN] = d; pf[PF_EQ] = 0.0;
        } else perf_node_serial(M, P, m.node_ts[o + k], m.node_dt[o + k], m.node_mode[o + k], m.node_zvel + (o + k) * 4, tt, ts, KT,
                                xt.data(), ut.data(), xnt.data(), pf);
      }
      });
      decide_problem(S, m, b);
    }
    if (it + 1 < iterations) {                         // intermediate SQP iteration: step in place, convergence test (k_step)
      for (int cc = 0; cc < 60; ++cc) step_component(m, b, cc);
      const int cv = check_convergence(S, ls, it, iterations);
      m.conv[b] = cv; ls[LS_CONV] = (double)cv;
      if (cv != CV_NONE) break;
    }
    }
    for (int cc = 0; cc < 60; ++cc) finalize_component(m, b, cc, t_out, x_out, u_out);
    if (m.conv[b] == CV_NONE) ls[LS_CONV] = (double)CV_ITERATIONS;
    if (n_out) n_out[b] = nn;
    if (mode_out) for (int k = 0; k < nn; ++k) mode_out[o + k] = m.node_mode[o + k];
    if (info) memcpy(info + (size_t)b * LS_SIZE, ls, sizeof(double) * LS_SIZE);
    if (status) status[b] = m.status[b];
  });
  return 0;
}

// Feedback gains of the last cycle, K [B][NMAX][30][30] (useFeedbackPolicy): same node functions as k_gains.
int cport_feedback_gains(CportCtx* c, double* K) {
  const MpcBuffers& m = c->m;
  parallel_for(m.B, c->threads, [&](int b) {
    const size_t o = (size_t)b * m.NMAX;
    for (int k = 0; k < m.nn[b]; ++k) {
      double* out = K + (o + k) * 900;
      const int src = feedback_gain_source(m.nn[b], m.node_flag + o, k);
      if (src < 0) memset(out, 0, sizeof(double) * 900);
      else feedback_gain_node(SerialGroup(), m.proj + (o + src) * PB_SIZE, m.gain + (o + src) * GB_SIZE, out);
    }
  });
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------- WBC
#include "../../qm_door_b200/csrc/qm_wbc.h"
#include "../../qm_door_b200/csrc/qm_actuator.h"
#include "../../qm_door_b200/csrc/qm_sim.h"

extern "C" {

int cport_wbc_ws_size() { return WW_SIZE; }

// B whole-body-control solves; u_last [B][30] is the per-solve inputLast_ state (read, then overwritten with ud).
int cport_wbc_batch(const qmb200_model_desc* M, const qmb200_wbc_desc* C, int B, const double* xd, const double* ud,
                    const double* rbd, const int32_t* mode, const double* period, const double* time, double* u_last,
                    double* cmd, int32_t* status, int threads) {
  parallel_for(B, threads, [&](int b) {
    std::vector<double> W(WW_SIZE), Wc(WC_SIZE);
    std::vector<int> WI(WI_SIZE);
    int st = 0;
    wbc_update(SerialGroup(), *M, *C, xd + 30 * b, ud + 30 * b, rbd + 55 * b, mode[b], period[b], time[b], u_last + 30 * b,
               W.data(), Wc.data(), WI.data(), cmd + 54 * b, &st);
    status[b] = st;
    memcpy(u_last + 30 * b, ud + 30 * b, sizeof(double) * 30);
  });
  return 0;
}

// Control law + simulated actuator, one tick of n problems; the caller owns the state arrays (zero-initialised):
// stamp [n][CAP] int64, buf [n][CAP][18][5], hc [n][2] int32, last [n][18][5].
void cport_actuator(const qmb200_actuator_desc* D, int n, const int64_t* time_ns, int64_t period_ns, const double* obs_time,
                    const double* xd, const double* ud, const double* cmd, const double* q, const double* v, int64_t* stamp,
                    double* buf, int32_t* hc, double* last, double* tau, int32_t* status) {
  const size_t CAP = QMB200_ACT_CAPACITY;
  for (size_t b = 0; b < (size_t)n; ++b) {
    status[b] = 0;
    actuator_step(SerialGroup(), *D, (long long)time_ns[b], (long long)period_ns, obs_time[b], xd + 30 * b, ud + 30 * b, cmd + 54 * b,
                  q + 18 * b, v + 18 * b, (long long*)stamp + CAP * b, buf + CAP * 18 * ACT_NF * b, hc + 2 * b, last + 18 * ACT_NF * b,
                  tau + 18 * b, status + b);
  }
}

// One solve with the per-level record of the hierarchy (WBL_* layout of qm_wbc.h).
int cport_wbc_levels_size() { return WBL_SIZE; }
void cport_wbc_levels(const qmb200_model_desc* M, const qmb200_wbc_desc* C, const double* xd, const double* ud, const double* rbd, int mode,
                      double period, double time, const double* u_last, double* cmd, int32_t* status, double* levels) {
  std::vector<double> W(WW_SIZE), Wc(WC_SIZE);
  std::vector<int> WI(WI_SIZE);
  int st = 0;
  wbc_update(SerialGroup(), *M, *C, xd, ud, rbd, mode, period, time, u_last, W.data(), Wc.data(), WI.data(), cmd, &st, levels);
  *status = st;
}

// Forward-dynamics step (qm_sim.h), n problems.
void cport_forward_dynamics(const qmb200_model_desc* M, double gravity, int n, const double* rbd, const double* tau, const int32_t* mode,
                            double dt, double beta, double* rbd_next, double* f, int32_t* status, int threads) {
  parallel_for(n, threads, [&](int b) {
    std::vector<double> W(WW_SIZE);
    int st = 0;
    fd_step(SerialGroup(), *M, gravity, rbd + 55 * b, tau + 18 * b, mode[b], dt, beta, W.data(), rbd_next + 55 * b, f + 12 * b, &st);
    status[b] = st;
  });
}

}  // extern "C"
