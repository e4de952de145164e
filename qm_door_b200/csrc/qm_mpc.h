// MPC half of the hot path: schedule, per-node LQ transcription with constraint projection, Riccati
// backward/forward sweep and line-search evaluation.  Same phase-structured style as qm_core.h.
// Reference call sites replaced: see the header of qm_core.h; upstream algorithms: SURVEY.md App. B.
#pragma once
#include "qm_core.h"

namespace qm {

enum { EV_NONE = 0, EV_PRE = 1, EV_POST = 2 };
enum {
  ST_OK = 0,
  ST_GRID_OVERFLOW = 1,   // node capacity exceeded
  ST_BAD_SCHEDULE = 2,    // swing phase not bracketed by stance phases / schedule does not cover the horizon
  ST_RANK = 4,            // constraint Jacobian lost rank in the projection
  ST_CHOL = 8,            // Riccati Hessian not positive definite
  ST_NAN = 16,
  ST_STEP_REJECTED = 32   // line search reached alpha_min: no step taken (informational)
};

// ---- per-node LQ blocks in HBM (offsets in doubles; each block contiguous and 16-byte aligned for bulk copies)
enum {
  SB_A = 0,                          // [30][30]
  SB_B = SB_A + 900,                 // [30][18]
  SB_b = SB_B + 30 * QM_NUT,         // [30]
  SB_q = SB_b + 30,                  // [30]
  SB_r = SB_q + 30,                  // [18]
  SB_NUT = SB_r + QM_NUT,            // reduced input dimension of this node (as double)
  SB_FWD_SIZE = SB_NUT + 2,          // the forward rollout needs [A | B | b | q | r | nut] only
  SB_Q = SB_FWD_SIZE,                // [30][30]
  SB_P = SB_Q + 900,                 // [18][30]
  SB_R = SB_P + 30 * QM_NUT,         // [18][18]
  SB_SIZE = SB_R + QM_NUT * QM_NUT
};
static_assert(SB_SIZE % 2 == 0 && SB_FWD_SIZE % 2 == 0, "bulk copies move multiples of 16 bytes");
// Projection of one node, compact: only the nv rows of the eliminated (pivot) joint velocities are stored.
//   du[12 + pivcol[p]] = PX[p] . dx + PU[p] . dut + PEC[p]      (p < nv)
//   du[fcols[a]]       = dut[a]                                  (a < nut; stance-foot forces and free joint velocities)
//   du[i]              = PEF[i]                                  (swing-foot force components: -u_i)
// role[i] (int32, i < 30): p for a pivot row, 32 + a for a free input, 64 otherwise; then nv, nut.
enum {
  PB_PX = 0,                         // [16][30]
  PB_PU = PB_PX + 16 * 30,           // [16][18]
  PB_PEC = PB_PU + 16 * QM_NUT,      // [16]
  PB_PEF = PB_PEC + 16,              // [12]
  PB_ROLE = PB_PEF + 12,             // 32 x int32
  PB_SIZE = PB_ROLE + 16
};
enum { ROLE_FREE = 32, ROLE_NONE = 64 };
enum { GB_K = 0, GB_KFF = 30 * QM_NUT, GB_SIZE = 560 };   // K [18][30], kff [18]
enum { PF_COST = 0, PF_DYN = 1, PF_EQ = 2, PF_SIZE = 4 };

// ------------------------------------------------------------------------------------------ interpolation
// [upstream] LinearInterpolation::timeSegment: value = alpha*d[i] + (1-alpha)*d[i+1]
QM_HD void time_segment(double t, const double* times, int n, int* idx, double* alpha) {
  if (n <= 1) { *idx = 0; *alpha = 1.0; return; }
  int part = 0;
  while (part < n && times[part] < t) ++part;          // findIndexInTimeArray (lower_bound)
  int i = (part == 0 && t == times[0]) ? 0 : part - 1;
  const int last = n - 1;
  if (i >= 0) {
    if (i < last) { *idx = i; *alpha = (times[i + 1] - t) / (times[i + 1] - times[i]); }
    else { *idx = (last - 1 > 0) ? last - 1 : 0; *alpha = 0.0; }
  } else { *idx = 0; *alpha = 1.0; }
}

QM_HD int mode_index(const double* events, int nev, double t) {
  int i = 0;
  while (i < nev && events[i] < t) ++i;
  return i;
}

// [upstream] RelaxedBarrierPenalty
QM_HD void relaxed_barrier(double h, double mu, double delta, double* v, double* d1, double* d2) {
  if (h > delta) {
    const double ih = 1.0 / h;
    *v = -mu * log(h); *d1 = -mu * ih; *d2 = mu * ih * ih;
  } else {
    const double z = (h - 2.0 * delta) / delta;
    *v = mu * (-log(delta) + 0.5 * z * z - 0.5); *d1 = mu * (h - 2.0 * delta) / (delta * delta); *d2 = mu / (delta * delta);
  }
}

// [upstream] CubicSpline velocity (Hermite in normalised time)
QM_HD double cubic_velocity(double t, double t0, double p0, double v0, double t1, double p1, double v1) {
  const double dt = t1 - t0, dp = p1 - p0, dv = v1 - v0;
  const double c1 = v0 * dt, c2 = -(3.0 * v0 + dv) * dt + 3.0 * dp, c3 = (2.0 * v0 + dv) * dt - 2.0 * dp;
  const double tn = (t - t0) / dt;
  return (3.0 * c3 * tn * tn + 2.0 * c2 * tn + c1) / dt;
}

// [upstream] SwingTrajectoryPlanner::getZvelocityConstraint on flat terrain
QM_HD double swing_z_velocity(const qmb200_problem_desc& P, const double* events, const int32_t* modes, int nev, int leg,
                              double t, int* status) {
  const int nph = nev + 1;
  const int p = mode_index(events, nev, t);
  const int bit = 3 - leg;
  if ((modes[p] >> bit) & 1) return 0.0;
  int start = -1, fin = -1;
  for (int ip = p - 1; ip >= 0; --ip)
    if ((modes[ip] >> bit) & 1) { start = ip; break; }
  for (int ip = p + 1; ip < nph; ++ip)
    if ((modes[ip] >> bit) & 1) { fin = ip - 1; break; }
  if (start < 0 || fin < 0) { *status |= ST_BAD_SCHEDULE; return 0.0; }
  const double ts = events[start], tf = events[fin];
  double scaling = (tf - ts) / P.swing_time_scale;
  if (scaling > 1.0) scaling = 1.0;
  const double tm = 0.5 * (ts + tf), hm = scaling * P.swing_height;
  if (t < tm) return cubic_velocity(t, ts, 0.0, scaling * P.swing_liftoff_vel, tm, hm, 0.0);
  return cubic_velocity(t, tm, hm, 0.0, tf, 0.0, scaling * P.swing_touchdown_vel);
}

// [upstream] timeDiscretizationWithEvents; one thread per problem (serial by construction).
QM_HDN void build_grid(const qmb200_solver_desc& S, double t0, const double* events, int nev, double* node_t,
                       int32_t* node_flag, int32_t* nn_out, int32_t* status) {
  const int NMAX = S.max_nodes;
  const double tf = t0 + S.horizon;
  int n = 0, st = 0;
  node_t[0] = t0; node_flag[0] = EV_NONE; n = 1;
  int nxt = mode_index(events, nev, t0);
  double tn = t0;
  while (node_t[n - 1] < tf) {
    tn = tn + S.dt;
    int ev = EV_NONE;
    if (nxt < nev && tn >= events[nxt]) { tn = events[nxt]; ev = EV_PRE; ++nxt; }
    if (tn >= tf) { tn = tf; ev = EV_NONE; }
    if (tn > node_t[n - 1] + S.dt_min) {
      if (n + 2 > NMAX) { st |= ST_GRID_OVERFLOW; node_t[n - 1] = tf; node_flag[n - 1] = EV_NONE; break; }
      node_t[n] = tn; node_flag[n] = ev; ++n;
    } else {
      node_t[n - 1] = tn; node_flag[n - 1] = ev;     // closer than dt_min to the last node: that node moves (and may become the pre-event node)
    }
    if (ev == EV_PRE) {                              // every pre-event node is followed by its post-event node, in both branches
      if (n + 1 > NMAX) { st |= ST_GRID_OVERFLOW; node_t[n - 1] = tf; node_flag[n - 1] = EV_NONE; break; }
      node_t[n] = tn; node_flag[n] = EV_POST; ++n;
    }
  }
  if (nev < 1 || events[nev - 1] < tf) st |= ST_BAD_SCHEDULE;   // schedule must extend past the horizon
  *nn_out = n;
  *status = st;
}

// Per-node annotations of the grid (interval start / duration, mode id, swing references): independent over nodes.
template <class G>
QM_HDN void annotate_schedule(G g, const qmb200_solver_desc& S, const qmb200_problem_desc& P, const double* events,
                              const int32_t* modes, int nev, int n, const double* node_t, const int32_t* node_flag,
                              double* node_ts, double* node_dt, int32_t* node_mode, double* node_zvel, int32_t* status) {
  QM_PFOR(g, i, n) {
    int st = 0;
    const double ts = node_t[i] + (node_flag[i] == EV_POST ? S.weak_eps : 0.0);
    node_ts[i] = ts;
    double dt = 0.0;
    if (i + 1 < n && node_flag[i] != EV_PRE) dt = (node_t[i + 1] - (node_flag[i + 1] == EV_PRE ? S.weak_eps : 0.0)) - ts;
    node_dt[i] = dt;
    const int md = modes[mode_index(events, nev, ts)];
    node_mode[i] = md;
    for (int leg = 0; leg < 4; ++leg) node_zvel[4 * i + leg] = swing_z_velocity(P, events, modes, nev, leg, ts, &st);
    if (st) status_or(status, st);
  }
  g.sync();
}

// [upstream] multiple_shooting::initializeStateInputTrajectories, in two steps.
// Step 1, per node k < nn - 1 (the same for every component): where the warm start is read. ri[3k] = kind: 0 pre-event node,
// 1 interpolate the previous solution, 2 beyond it (QMInitializer); ri[3k+1], ra[2k]: segment and weight of the state at the end
// of the interval; ri[3k+2], ra[2k+1]: of the input at its start ([upstream] LinearInterpolation::timeSegment: lower_bound).
QM_HD int times_below(const double* t, int n, double q) {             // number of entries strictly below q (t ascending)
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (t[mid] < q) lo = mid + 1; else hi = mid; }
  return lo;
}
QM_HD void interp_segment(int part, int nprev, const double* prev_t, double q, int* i, double* a) {
  if (part == 0) { *i = 0; *a = (q == prev_t[0]) ? (prev_t[1] - q) / (prev_t[1] - prev_t[0]) : 1.0; }
  else if (part - 1 < nprev - 1) { *i = part - 1; *a = (prev_t[part] - q) / (prev_t[part] - prev_t[part - 1]); }
  else { *i = nprev - 2; *a = 0.0; }
}
QM_HDN void init_guess_node(double weak_eps, int k, const double* node_t, const int32_t* node_flag, const double* node_ts,
                            int nprev, const double* prev_t, int* ri, double* ra) {
  const bool has_prev = nprev >= 2;
  const double till_x = has_prev ? prev_t[nprev - 1] : node_t[0];
  const double till_u = has_prev ? prev_t[nprev - 2] : node_t[0];
  int kind = 0, ix = 0, iu = 0;
  double ax = 0.0, au = 0.0;
  if (node_flag[k] != EV_PRE) {
    const double t = node_ts[k], tn = node_t[k + 1] - (node_flag[k + 1] == EV_PRE ? weak_eps : 0.0);
    if (t > till_u || tn > till_x) kind = 2;
    else {
      kind = 1;
      interp_segment(times_below(prev_t, nprev, tn), nprev, prev_t, tn, &ix, &ax);
      interp_segment(times_below(prev_t, nprev, t), nprev, prev_t, t, &iu, &au);
    }
  }
  ri[3 * k] = kind; ri[3 * k + 1] = ix; ri[3 * k + 2] = iu;
  ra[2 * k] = ax; ra[2 * k + 1] = au;
}

// Step 2, one thread per (problem, component c < 60): the trajectories of that component.
QM_HDN void init_guess_component(const qmb200_model_desc& M, const qmb200_problem_desc& P, int c, const double* x0, int nn,
                                 const double* node_t, const double* node_ts, const int32_t* node_mode, const int* ri, const double* ra,
                                 int nprev, const double* prev_t, const double* prev_x, const double* prev_u, double* xs, double* us) {
  const bool has_prev = nprev >= 2;
  const double till_x = has_prev ? prev_t[nprev - 1] : node_t[0];
  const int n = nn - 1;
  if (c < 30) {
    double xc;
    const double t_init = node_ts[0];
    if (t_init < till_x) {
      const int part = times_below(prev_t, nprev, t_init);
      const int i = (part == 0) ? 0 : part - 1;        // t_init < till_x = prev_t[last] => i < last
      const double a = (part == 0 && !(t_init == prev_t[0])) ? 1.0 : (prev_t[i + 1] - t_init) / (prev_t[i + 1] - prev_t[i]);
      xc = a * prev_x[30 * i + c] + (1.0 - a) * prev_x[30 * (i + 1) + c];
    } else xc = x0[c];
    xs[c] = xc;
    for (int k = 0; k < n; ++k) {
      if (ri[3 * k] == 1) {
        const int i = ri[3 * k + 1];
        const double a = ra[2 * k];
        xc = a * prev_x[30 * i + c] + (1.0 - a) * prev_x[30 * (i + 1) + c];
      }
      xs[30 * (k + 1) + c] = xc;                       // pre-event nodes and nodes beyond the warm start keep the last state
    }
  } else {
    const int cu = c - 30;
    for (int k = 0; k < n; ++k) {
      double uc = 0.0;
      const int kind = ri[3 * k];
      if (kind == 2) {
        // QMInitializer::compute (qm_interface/src/initialization/QMInitializer.cpp:33-41): weight compensation
        const int md = node_mode[k];
        const int ns = ((md >> 3) & 1) + ((md >> 2) & 1) + ((md >> 1) & 1) + (md & 1);
        if (cu < 12 && (cu % 3) == 2 && ((md >> (3 - cu / 3)) & 1)) uc = M.total_mass * P.gravity / ns;
      } else if (kind == 1) {
        const int i = ri[3 * k + 2];
        const double a = ra[2 * k + 1];
        uc = a * prev_u[30 * i + cu] + (1.0 - a) * prev_u[30 * (i + 1) + cu];
      }
      us[30 * k + cu] = uc;
    }
    us[30 * n + cu] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------ measured state -> MPC state
// [upstream] CentroidalModelRbdConversions::computeCentroidalStateFromRbdModel as called from
// QMController::updateStateEstimation (qm_controllers/src/QMController.cpp:239-243), followed by the yaw unwrapping of the
// same function (:244, angles::shortest_angular_distance). rbd[55] is the estimator's layout
// (qm_estimation/src/StateEstimateBase.cpp:29-102): [zyx(3); base position(3); joints(18); world angular velocity(3);
// base linear velocity(3); joint velocities(18); arm end-effector pose(7, unused here)].
// x = [A(q) v / m (6); base position; zyx; joints]. kw: kinematics workspace (value level, KW_VSIZE doubles); qv: 48 doubles.
template <class G>
QM_HDN void centroidal_state_from_rbd(G g, const qmb200_model_desc& M, const double* rbd, double yaw_last, int unwrap,
                                      double* kw, double* qv, double* x_out) {
  if (g.tid() == 0) {
    for (int k = 0; k < 3; ++k) { qv[k] = rbd[3 + k]; qv[3 + k] = rbd[k]; qv[24 + k] = rbd[27 + k]; }
    for (int k = 0; k < 18; ++k) { qv[6 + k] = rbd[6 + k]; qv[30 + k] = rbd[30 + k]; }
    // [upstream] getEulerAnglesZyxDerivativesFromGlobalAngularVelocity
    double sz, cz, sy, cy;
    sincos(qv[3], &sz, &cz); sincos(qv[4], &sy, &cy);
    const double wx = rbd[24], wy = rbd[25], wz = rbd[26];
    const double dxr = (cz * wx + sz * wy) / cy;
    qv[24 + 3] = wz + sy * dxr;
    qv[24 + 4] = -sz * wx + cz * wy;
    qv[24 + 5] = dxr;
  }
  g.sync();
  kin_positions(g, M, qv, kw, false);
  QM_PFOR(g, r, 6) {
    double acc = 0.0;
    for (int k = 0; k < QM_NJ; ++k) acc += kw[KW_ACM + r * QM_NJ + k] * qv[24 + k];
    x_out[r] = acc / M.total_mass;
  }
  QM_PFOR(g, k, QM_NJ) {
    double v = qv[k];
    if (k == 3 && unwrap) {
      // angles::shortest_angular_distance(from, to) = normalize_angle(to - from), normalize_angle -> (-pi, pi]
      const double two_pi = 6.283185307179586476925286766559, pi = 3.14159265358979323846;
      double a = fmod(fmod(v - yaw_last, two_pi) + two_pi, two_pi);
      if (a > pi) a -= two_pi;
      v = yaw_last + a;
    }
    x_out[6 + k] = v;
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------ flow map rows
// f (30) and, if Fr != nullptr, the nine non-trivial rows (f rows 3..11) of [df/dx | df/du] as Fr[9][60].
template <class G>
QM_HDN void flow_rows(G g, const qmb200_model_desc& M, double gravity, const double* w, const double* x, const double* u,
                      double* f, double* Fr, double* vb_shadow = nullptr) {
  const double m = M.total_mass;
  QM_PFOR(g, i, 30) {
    double v;
    if (i < 3) {
      v = (u[i] + u[3 + i] + u[6 + i] + u[9 + i]) / m - (i == 2 ? gravity : 0.0);
    } else if (i < 6) {
      v = 0.0;
      for (int ft = 0; ft < 4; ++ft) {
        double arm[3], t[3];
        for (int r = 0; r < 3; ++r) arm[r] = w[KW_FPOS + 3 * ft + r] - w[KW_COM + r];
        cross3(arm, u + 3 * ft, t);
        v += (i == 3) ? t[0] : ((i == 4) ? t[1] : t[2]);       // no dynamically indexed local array
      }
      v /= m;
    } else {
      v = w[KW_VEL + i - 6];
    }
    f[i] = v;
  }
  if (Fr != nullptr) {
    const double im = 1.0 / m;
    // One work item per column c of [df/dx | df/du]: the nine entries of a column share their operands (the three angular
    // rows come from one vector, the six v_b = A_b^-1 (m h - A_j v_j) rows from one column of d(A v)/dq or of A_j), and a
    // warp executes the instructions of every branch its lanes take, so per-entry items paid for all column kinds at once.
    QM_PFOR(g, c, 60) {
      double v[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (c < 6) {                                  // d v_b / d h
        for (int rr = 0; rr < 6; ++rr) v[3 + rr] = m * w[KW_ABINV + 6 * rr + c];
      } else if (c >= 30 && c < 42) {               // d/df_i: (p_i - c) x e_d / m
        const int ft = (c - 30) / 3, d = (c - 30) % 3;
        double arm[3];
        const double e[3] = {d == 0 ? 1.0 : 0.0, d == 1 ? 1.0 : 0.0, d == 2 ? 1.0 : 0.0};
        for (int rr = 0; rr < 3; ++rr) arm[rr] = w[KW_FPOS + 3 * ft + rr] - w[KW_COM + rr];
        cross3(arm, e, v);
        for (int r = 0; r < 3; ++r) v[r] *= im;
      } else {                                      // joint position (6..29) or joint velocity (42..59) columns
        const bool isq = c < 30;
        const double* X = isq ? (w + KW_DH + (c - 6)) : (w + KW_ACM + 6 + (c - 42));
        double xc[6];
        for (int cc = 0; cc < 6; ++cc) xc[cc] = X[cc * QM_NJ];
        for (int rr = 0; rr < 6; ++rr) {
          const double* Bi = w + KW_ABINV + 6 * rr;
          double a = 0.0;
          for (int cc = 0; cc < 6; ++cc) a -= Bi[cc] * xc[cc];
          v[3 + rr] = a;
        }
        if (isq) {                                  // d/dq_k sum_i (p_i - c) x f_i / m
          const int k = c - 6;
          double ac[3];
          for (int rr = 0; rr < 3; ++rr) ac[rr] = w[KW_ACM + rr * QM_NJ + k] * im;
          for (int ft = 0; ft < 4; ++ft) {
            double d[3];
            for (int rr = 0; rr < 3; ++rr) d[rr] = w[KW_FJ + (3 * ft + rr) * QM_NJ + k] - ac[rr];
            cross3_add(d, u + 3 * ft, v);
          }
          for (int r = 0; r < 3; ++r) v[r] *= im;
        }
      }
      for (int r = 0; r < 9; ++r) Fr[60 * r + c] = v[r];
      if (vb_shadow != nullptr)                     // rows of v_b kept close for the constraint rows
        for (int rr = 0; rr < 6; ++rr) vb_shadow[60 * rr + c] = v[3 + rr];
    }
  }
  g.sync(); QM_TICK(23);
}

// ------------------------------------------------------------------------------------------ quaternions (x,y,z,w)
// Eigen::Quaternion(Matrix3)
QM_HD void quat_from_matrix(const double* R, double* q) {
  const double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    double s = sqrt(t + 1.0);
    q[3] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * s;
    s = 0.5 / s;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
  }
}
// Eigen slerp(t, other)
QM_HD void quat_slerp(const double* q0, const double* q1, double t, double* out) {
  const double d = q0[0] * q1[0] + q0[1] * q1[1] + q0[2] * q1[2] + q0[3] * q1[3];
  const double ad = fabs(d);
  double s0, s1;
  if (ad >= 1.0 - 2.220446049250313e-16) { s0 = 1.0 - t; s1 = t; }
  else { const double th = acos(ad), st = sin(th); s0 = sin((1.0 - t) * th) / st; s1 = sin(t * th) / st; }
  if (d < 0.0) s1 = -s1;
  for (int k = 0; k < 4; ++k) out[k] = s0 * q0[k] + s1 * q1[k];
}

// Reference quantities of one node, computed by one thread into `ref`:
//  [0:30] x_ref  [30:60] u_nominal  [60:63] ee position ref  [63:67] ee quaternion ref
enum { RF_X = 0, RF_U = 30, RF_EEP = 60, RF_EEQ = 63, RF_SIZE = 68 };
QM_HDN void node_reference(const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, int mode, const double* tt,
                           const double* ts, int kt, double* ref) {
  int i; double a;
  time_segment(t, tt, kt, &i, &a);
  if (kt > 1) {
    const double* lhs = ts + QM_NTARGET * i;
    const double* rhs = ts + QM_NTARGET * (i + 1);
    for (int c = 0; c < 30; ++c) ref[RF_X + c] = a * lhs[c] + (1.0 - a) * rhs[c];
    for (int c = 0; c < 3; ++c) ref[RF_EEP + c] = a * lhs[30 + c] + (1.0 - a) * rhs[30 + c];
    quat_slerp(lhs + 33, rhs + 33, 1.0 - a, ref + RF_EEQ);      // EndEffectorConstraint.cpp:90-102
  } else {
    for (int c = 0; c < 30; ++c) ref[RF_X + c] = ts[c];
    for (int c = 0; c < 7; ++c) ref[RF_EEP + c] = ts[30 + c];
  }
  const int ns = ((mode >> 3) & 1) + ((mode >> 2) & 1) + ((mode >> 1) & 1) + (mode & 1);
  for (int c = 0; c < 30; ++c) ref[RF_U + c] = 0.0;
  for (int ft = 0; ft < 4; ++ft)
    if ((mode >> (3 - ft)) & 1) ref[RF_U + 3 * ft + 2] = M.total_mass * P.gravity / ns;   // [upstream] weightCompensatingInput
}

// The same over a thread group: one lane per component of x_ref / u_nominal, the end-effector reference on the last lane.
// No trailing sync: the caller's next full sync publishes `ref`.
template <class G>
QM_HDN void node_reference_group(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, int mode,
                                 const double* tt, const double* ts, int kt, double* ref) {
  int i; double a;
  time_segment(t, tt, kt, &i, &a);
  const double* lhs = ts + QM_NTARGET * i;
  const double* rhs = ts + QM_NTARGET * (i + 1);
  const int ns = ((mode >> 3) & 1) + ((mode >> 2) & 1) + ((mode >> 1) & 1) + (mode & 1);
  QM_PFOR(g, c, 32) {
    if (c < 30) {
      ref[RF_X + c] = (kt > 1) ? a * lhs[c] + (1.0 - a) * rhs[c] : ts[c];
      const bool fz = c < 12 && (c % 3) == 2 && ((mode >> (3 - c / 3)) & 1);
      ref[RF_U + c] = fz ? M.total_mass * P.gravity / ns : 0.0;                      // [upstream] weightCompensatingInput
    } else if (c == 31) {
      if (kt > 1) {
        for (int r = 0; r < 3; ++r) ref[RF_EEP + r] = a * lhs[30 + r] + (1.0 - a) * rhs[30 + r];
        quat_slerp(lhs + 33, rhs + 33, 1.0 - a, ref + RF_EEQ);      // EndEffectorConstraint.cpp:90-102
      } else {
        for (int r = 0; r < 7; ++r) ref[RF_EEP + r] = ts[30 + r];
      }
    }
  }
}

// End-effector error e[6] = [p - p_ref ; quaternionDistance(q, q_ref)] and, if JE != nullptr, de/dq [6][24]
// from the frame Jacobian in the kinematics workspace (one thread computes e and the 3x3 map, all threads JE).
template <class G>
QM_HDN void ee_terms(G g, const double* w, const double* ref, double* e, double* Dq, double* JE) {
  if (g.tid() == 0) {
    double q[4];
    quat_from_matrix(w + KW_EER, q);
    const double* qr = ref + RF_EEQ;
    double cr[3];
    cross3(q, qr, cr);
    for (int r = 0; r < 3; ++r) {
      e[r] = w[KW_EEP + r] - ref[RF_EEP + r];
      e[3 + r] = q[3] * qr[r] - qr[3] * q[r] + cr[r];        // [upstream] quaternionDistance
    }
    // d e_ori / d(world rotation vector) = -1/2 (E_w I + [e_ori]x),  E_w = qr.w q.w + qr.vec . q.vec
    const double Ew = qr[3] * q[3] + qr[0] * q[0] + qr[1] * q[1] + qr[2] * q[2];
    const double ex = e[3], ey = e[4], ez = e[5];
    Dq[0] = -0.5 * Ew; Dq[1] = 0.5 * ez;   Dq[2] = -0.5 * ey;
    Dq[3] = -0.5 * ez; Dq[4] = -0.5 * Ew;  Dq[5] = 0.5 * ex;
    Dq[6] = 0.5 * ey;  Dq[7] = -0.5 * ex;  Dq[8] = -0.5 * Ew;
  }
  g.sync();
  if (JE != nullptr) {
    QM_PFOR(g, idx, 6 * QM_NJ) {
      const int r = idx / QM_NJ, k = idx % QM_NJ;
      double v;
      if (r < 3) v = w[KW_EEJ + r * QM_NJ + k];
      else {
        const double* D = Dq + 3 * (r - 3);
        v = D[0] * w[KW_EEJ + 3 * QM_NJ + k] + D[1] * w[KW_EEJ + 4 * QM_NJ + k] + D[2] * w[KW_EEJ + 5 * QM_NJ + k];
      }
      JE[idx] = v;
    }
    g.sync();
  }
}

// Friction cone of one stance foot: [h, g0,g1,g2, H00,H01,H11, p0,p1,p2]   ([upstream] FrictionConeConstraint)
QM_HD void cone_terms(const qmb200_problem_desc& P, const double* F, double* c) {
  const double t2 = F[0] * F[0] + F[1] * F[1] + P.fric_reg;
  const double tn = sqrt(t2);
  c[0] = P.fric_mu * (F[2] + P.fric_grip) - tn;
  const double itn = 1.0 / tn;
  c[1] = -F[0] * itn; c[2] = -F[1] * itn; c[3] = P.fric_mu;
  const double ip32 = itn / t2;
  c[4] = -(F[1] * F[1] + P.fric_reg) * ip32;
  c[5] = F[0] * F[1] * ip32;
  c[6] = -(F[0] * F[0] + P.fric_reg) * ip32;
  relaxed_barrier(c[0], P.fric_bar_mu, P.fric_bar_delta, c + 7, c + 8, c + 9);
}

// Scalar running cost pieces that need no matrix products (barriers); returns value, one thread.
QM_HD double barrier_cost(const qmb200_problem_desc& P, int mode, const double* x, const double* u) {
  double L = -P.box_offset, v, d1, d2;
  for (int i = 0; i < 6; ++i) {
    relaxed_barrier(x[24 + i] - P.arm_pos_lo[i], P.pos_bar_mu, P.pos_bar_delta, &v, &d1, &d2); L += v;
    relaxed_barrier(P.arm_pos_hi[i] - x[24 + i], P.pos_bar_mu, P.pos_bar_delta, &v, &d1, &d2); L += v;
    relaxed_barrier(u[24 + i] - P.arm_vel_lo[i], P.vel_bar_mu, P.vel_bar_delta, &v, &d1, &d2); L += v;
    relaxed_barrier(P.arm_vel_hi[i] - u[24 + i], P.vel_bar_mu, P.vel_bar_delta, &v, &d1, &d2); L += v;
  }
  for (int ft = 0; ft < 4; ++ft)
    if ((mode >> (3 - ft)) & 1) { double c[10]; cone_terms(P, u + 3 * ft, c); L += c[7]; }
  return L;
}

// ------------------------------------------------------------------------------------------ transcription workspace
enum { PI_PIV = 0, PI_STATUS = 16, PI_NUT = 17, PI_NV = 18, PI_SEL = 20, PI_POS = 50, PI_SIZE = 80 };   // int32 index record
// aux layout: reference (x_ref, u_nominal, ee pose), friction-cone terms per foot, arm box gradients / Hessians / values
enum { NA_REF = 0, NA_CONE = RF_SIZE, NA_BOX = NA_CONE + 40, NA_BOXV = NA_BOX + 24, NA_SIZE = NA_BOXV + 12 };
enum { KS_FR1 = 0, KS_FR2 = 540, KS_F1 = 1080, KS_F2 = 1110, KS_X2 = 1140, KS_T = 1170, KS_JE = KS_T + 784, KS_E6 = KS_JE + 144,
       KS_AUX = KS_E6 + 8, KS_DINV = ((KS_AUX + NA_SIZE + 3) / 4) * 4, KS_PIV = KS_DINV + 256, KS_SIZE = KS_PIV + PI_SIZE / 2 };
static_assert(KS_SIZE % 2 == 0, "staged kinematics products are moved by 16-byte bulk copies");
enum {
  TW_KIN = 0,                       // kinematics workspace (fused path) / staged kinematics products (split path)
  TW_A = TW_KIN + ((int)KW_SIZE > (int)KS_SIZE ? (int)KW_SIZE : (int)KS_SIZE),     // [30][30]
  TW_BPM = TW_A + 900,              // [30][30] B with columns permuted [pivots | frees | dropped]
  TW_RPM = TW_BPM + 900,            // [<=30][30] R with rows [pivots | frees] and columns permuted
  TW_T2 = TW_RPM + 900,             // [16][49] = Dinv [Dv | C | e]: row p expresses joint velocity piv[p]
  TW_PUC = TW_T2 + 784,             // [16][18] pivot rows of Pu (reduced-input columns)
  TW_b = TW_PUC + 288,              // 30
  TW_QV = TW_b + 30,                // q
  TW_RV = TW_QV + 30,               // r
  TW_PEP = TW_RV + 30,              // Pe in permuted order (30)
  TW_TQ = TW_PEP + 30,              // Q dx
  TW_TR = TW_TQ + 30,               // R du
  TW_RP = TW_TR + 30,               // r' = r + R Pe, rows [pivots | frees]
  TW_RED = TW_RP + 30,              // 4: reduction scratch
  TW_LQ_SIZE = TW_RED + 4,          // what the split path (k_lq) needs: the rest lives in the staged kinematics products
  // RPX [<=30][30] (rows [pivots; frees] of R Px) takes the place of the flow-map rows fr1 | fr2 once A, B are assembled,
  // RPU [<=30][18] the place of the constraint rows T once T2 is formed.
  TW_JE = TW_LQ_SIZE,               // [6][24]  (fused path / terminal node: k_lq's terminal CTA places these over BPM)
  TW_REF = TW_JE + 144,             // RF_SIZE (+ 10 scratch for the ee rotation map)
  TW_E6 = TW_REF + RF_SIZE + 10,    // 6 (+2 pad)
  TW_DQ = TW_E6 + 8,                // 9 (+1)
  TW_FR1 = TW_DQ + 10,              // [9][60]   (fused path only from here on)
  TW_FR2 = TW_FR1 + 540,            // [9][60]
  TW_F1 = TW_FR2 + 540,             // 30
  TW_F2 = TW_F1 + 30,
  TW_X2 = TW_F2 + 30,
  TW_T = TW_X2 + 30,                // [16][49] = [Dv | C | e]
  TW_AUX = TW_T + 784,              // NA_SIZE
  TW_DINV = TW_AUX + 144,           // [16][16]
  TW_PIV = TW_DINV + 256,           // PI_SIZE x int32
  TW_SIZE = TW_PIV + PI_SIZE / 2
};
enum { TI_SIZE = 4 };                // integer workspace (unused by the LQ assembly: the index record comes with the pivots)

// Intermediate products of one node handed from the kinematics evaluations to the LQ assembly. In the fused (host)
// path they live in the transcription workspace; in the split CUDA path k_kin / k_proj write them to HBM (KS_* layout)
// and k_lq stages them into shared memory.
struct NodeIO {
  double* fr1; double* fr2;   // [9][60] non-trivial rows of [df/dx | df/du] at (x,u) and (x + dt f1, u)
  double* f1;  double* f2;    // [30] flow map values
  double* x2;                 // [30] x + dt f1
  double* T;                  // [nv][49] velocity-constraint rows [Dv | C | e]
  double* je;                 // [6][24] end-effector error Jacobian
  double* e6;                 // [8] end-effector error (6), [6] = sum of squared velocity-constraint values
  double* aux;                // [NA_SIZE] references and barrier terms of the node (functions of t, x, u only)
  double* dinv;               // [16][16] inverse of the pivot block of Dv (rows / columns in constraint-row order)
  int* piv;                   // PI_* index record: pivots, status bits, nut, nv, permutation sel / pos of the inputs
};
QM_HD NodeIO node_io_at(double* base) {
  NodeIO io;
  io.fr1 = base + KS_FR1; io.fr2 = base + KS_FR2; io.f1 = base + KS_F1; io.f2 = base + KS_F2;
  io.x2 = base + KS_X2; io.T = base + KS_T; io.je = base + KS_JE; io.e6 = base + KS_E6; io.aux = base + KS_AUX;
  io.dinv = base + KS_DINV; io.piv = (int*)(base + KS_PIV);
  return io;
}

QM_HD int velocity_rows(int mode) {      // velocity-constraint rows: 3 per stance foot, 1 per swing foot
  int nv = 0;
  for (int ft = 0; ft < 4; ++ft) nv += ((mode >> (3 - ft)) & 1) ? 3 : 1;
  return nv;
}

// ------------------------------------------------------------------------------------------ constraint projection pivots
// [upstream] luConstraintProjection on the joint-velocity block Dv [nv][18] of the velocity constraints: Gauss-Jordan
// with full pivoting, carried out as an in-place inversion without row exchanges. Results: piv[p] = joint-velocity
// column eliminated by constraint row p, Dinv = (Dv[:, piv])^-1 so that Dinv [Dv | C | e] has unit pivot columns.
// Pivot choice: largest magnitude among unused rows / columns; ties -> smallest row, then smallest column.
#if defined(__CUDACC__)
template <int PR>
__device__ __forceinline__ void gj_pivot_step(double (&a)[16], int nv, int pc, int lane) {
  const double p = __shfl_sync(0xffffffffu, a[PR], pc);
  const double ip = 1.0 / p;
  const double arow = a[PR] * ip;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i != PR && i < nv) {                       // nv is warp-uniform
      const double f = __shfl_sync(0xffffffffu, a[i], pc);
      a[i] = (lane == pc) ? -f * ip : a[i] - f * arow;
    }
  }
  a[PR] = (lane == pc) ? ip : arow;
}
// One warp; column c of Dv in the registers of lane c (rows = register index), multipliers broadcast with shuffles.
__device__ __forceinline__ void projection_pivots_warp(const double* T, int mode, double* Dinv, int* piv) {
  const int nv = velocity_rows(mode);
  const int lane = threadIdx.x & 31;
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (lane < 18 && i < nv) ? T[49 * i + lane] : 0.0;
  unsigned rows_used = 0;
  bool col_used = false, bad = false;
  int myrow = -1;
  for (int s = 0; s < nv; ++s) {
    double best = -1.0;
    int arg = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv && !((rows_used >> i) & 1u)) { const double v = fabs(a[i]); if (v > best) { best = v; arg = i; } }
    if (lane >= 18 || col_used) best = -1.0;
    int key = arg * 32 + lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int ok = __shfl_xor_sync(0xffffffffu, key, off);
      if (ob > best || (ob == best && ok < key)) { best = ob; key = ok; }
    }
    const int pr = key >> 5, pc = key & 31;
    if (!(best > 1e-12)) bad = true;
    switch (pr) {
      case 0: gj_pivot_step<0>(a, nv, pc, lane); break;    case 1: gj_pivot_step<1>(a, nv, pc, lane); break;
      case 2: gj_pivot_step<2>(a, nv, pc, lane); break;    case 3: gj_pivot_step<3>(a, nv, pc, lane); break;
      case 4: gj_pivot_step<4>(a, nv, pc, lane); break;    case 5: gj_pivot_step<5>(a, nv, pc, lane); break;
      case 6: gj_pivot_step<6>(a, nv, pc, lane); break;    case 7: gj_pivot_step<7>(a, nv, pc, lane); break;
      case 8: gj_pivot_step<8>(a, nv, pc, lane); break;    case 9: gj_pivot_step<9>(a, nv, pc, lane); break;
      case 10: gj_pivot_step<10>(a, nv, pc, lane); break;  case 11: gj_pivot_step<11>(a, nv, pc, lane); break;
      case 12: gj_pivot_step<12>(a, nv, pc, lane); break;  case 13: gj_pivot_step<13>(a, nv, pc, lane); break;
      case 14: gj_pivot_step<14>(a, nv, pc, lane); break;  default: gj_pivot_step<15>(a, nv, pc, lane); break;
    }
    rows_used |= 1u << pr;
    if (lane == pc) { col_used = true; myrow = pr; }
  }
  if (myrow >= 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) if (i < nv) Dinv[16 * i + myrow] = a[i];
    piv[PI_PIV + myrow] = lane;
  }
  // index record: inputs ordered [pivot joint velocities (row order) | free inputs: stance forces, free joint velocities | dropped]
  const unsigned freemask = __ballot_sync(0xffffffffu, lane < 18 && !col_used);
  const int nst = ((mode >> 3) & 1) + ((mode >> 2) & 1) + ((mode >> 1) & 1) + (mode & 1);
  const int nut = 3 * nst + __popc(freemask);
  if (lane < 18) {
    const int c = col_used ? myrow : nv + 3 * nst + __popc(freemask & ((1u << lane) - 1u));
    piv[PI_SEL + c] = 12 + lane; piv[PI_POS + 12 + lane] = c;
  }
  if (lane < 12) {
    const int ft = lane / 3;
    int before = 0;                                  // stance feet before this foot
    for (int f2 = 0; f2 < ft; ++f2) before += (mode >> (3 - f2)) & 1;
    const int c = ((mode >> (3 - ft)) & 1) ? nv + 3 * before + lane % 3 : nv + nut + 3 * (ft - before) + lane % 3;
    piv[PI_SEL + c] = lane; piv[PI_POS + lane] = c;
  }
  if (lane == 0) { piv[PI_STATUS] = bad ? ST_RANK : 0; piv[PI_NUT] = nut; piv[PI_NV] = nv; }
}
#endif
// Scalar statement of the same elimination (CPU port).
QM_HDN void projection_pivots_serial(const double* T, int mode, double* Dinv, int* piv) {
  const int nv = velocity_rows(mode);
  double a[16][18];
  for (int i = 0; i < 16; ++i) for (int c = 0; c < 18; ++c) a[i][c] = (i < nv) ? T[49 * i + c] : 0.0;
  unsigned rows_used = 0, cols_used = 0;
  int rowof[18];
  for (int c = 0; c < 18; ++c) rowof[c] = -1;
  bool bad = false;
  for (int s = 0; s < nv; ++s) {
    double best = -1.0; int pr = 0, pc = 0;
    for (int i = 0; i < nv; ++i) {
      if ((rows_used >> i) & 1u) continue;
      for (int c = 0; c < 18; ++c) {
        if ((cols_used >> c) & 1u) continue;
        const double v = fabs(a[i][c]);
        if (v > best) { best = v; pr = i; pc = c; }
      }
    }
    if (!(best > 1e-12)) bad = true;
    const double ip = 1.0 / a[pr][pc];
    double arow[18];
    for (int c = 0; c < 18; ++c) arow[c] = a[pr][c] * ip;
    for (int i = 0; i < nv; ++i) {
      if (i == pr) continue;
      const double f = a[i][pc];
      for (int c = 0; c < 18; ++c) a[i][c] = (c == pc) ? -f * ip : a[i][c] - f * arow[c];
    }
    for (int c = 0; c < 18; ++c) a[pr][c] = (c == pc) ? ip : arow[c];
    rows_used |= 1u << pr; cols_used |= 1u << pc;
    rowof[pc] = pr;
  }
  for (int c = 0; c < 18; ++c)
    if (rowof[c] >= 0) {
      for (int i = 0; i < nv; ++i) Dinv[16 * i + rowof[c]] = a[i][c];
      piv[PI_PIV + rowof[c]] = c;
    }
  int pos = nv;
  for (int p = 0; p < nv; ++p) piv[PI_SEL + p] = 12 + piv[PI_PIV + p];
  for (int ft = 0; ft < 4; ++ft)
    if ((mode >> (3 - ft)) & 1) for (int d = 0; d < 3; ++d) piv[PI_SEL + pos++] = 3 * ft + d;
  for (int c = 0; c < 18; ++c) if (rowof[c] < 0) piv[PI_SEL + pos++] = 12 + c;
  const int nut = pos - nv;
  for (int ft = 0; ft < 4; ++ft)
    if (!((mode >> (3 - ft)) & 1)) for (int d = 0; d < 3; ++d) piv[PI_SEL + pos++] = 3 * ft + d;
  for (int c = 0; c < 30; ++c) piv[PI_POS + piv[PI_SEL + c]] = c;
  piv[PI_STATUS] = bad ? ST_RANK : 0; piv[PI_NUT] = nut; piv[PI_NV] = nv;
}

// Kinematics at (x,u) with derivatives: flow map rows, end-effector terms, constraint rows (QMInterface.cpp:116-131), x2.
// kw: kinematics workspace; scr: RF_SIZE + 10 doubles of scratch.
template <class G>
QM_HDN void node_eval1(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                       const double* zvel, const double* tt, const double* ts, int kt, const double* x, const double* u,
                       double* kw, double* scr, NodeIO io) {
  node_reference_group(g, M, P, t, mode, tt, ts, kt, scr);
  // barrier terms of the node: friction cones of the stance feet, arm position / velocity boxes (independent lanes)
  QM_PFOR(g, it, 16) {
    if (it < 4) {
      if ((mode >> (3 - it)) & 1) cone_terms(P, u + 3 * it, io.aux + NA_CONE + 10 * it);
    } else {
      const int i = it - 4;
      double v1, a1, b1, v2, a2, b2;
      if (i < 6) {
        relaxed_barrier(x[24 + i] - P.arm_pos_lo[i], P.pos_bar_mu, P.pos_bar_delta, &v1, &a1, &b1);
        relaxed_barrier(P.arm_pos_hi[i] - x[24 + i], P.pos_bar_mu, P.pos_bar_delta, &v2, &a2, &b2);
      } else {
        relaxed_barrier(u[18 + i] - P.arm_vel_lo[i - 6], P.vel_bar_mu, P.vel_bar_delta, &v1, &a1, &b1);
        relaxed_barrier(P.arm_vel_hi[i - 6] - u[18 + i], P.vel_bar_mu, P.vel_bar_delta, &v2, &a2, &b2);
      }
      io.aux[NA_BOX + 2 * i] = a1 - a2;
      io.aux[NA_BOX + 2 * i + 1] = b1 + b2;
      io.aux[NA_BOXV + i] = v1 + v2;
    }
  }
  QM_TICK(24);
  // kin_eval in its parts: the end-effector terms read the end-effector Jacobian before the velocity level reuses its storage
  kin_positions(g, M, x + 6, kw, true);
  ee_terms(g, kw, scr, io.e6, scr + RF_SIZE, io.je);
  centroidal_velocity(g, M, x, u, kw);
  kin_velocities(g, M, 2, kw);
  // the six v_b rows of [df/dx | df/du] are mirrored into the (now dead) arrays P | AX | COMP of the workspace (384 doubles)
  static_assert(KW_AX == KW_P + 3 * QM_NJ && KW_COMP == KW_AX + 3 * QM_NJ, "the v_b shadow spans P | AX | COMP");
  flow_rows(g, M, P.gravity, kw, x, u, io.f1, io.fr1, kw + KW_P);
  QM_TICK(25);
  {
    // One work item per column of [Dv | C | e]: the six v_b entries of the column are loaded once, then the rows follow
    // (3 per stance foot, the normal one per swing foot).
    const double* vb = kw + KW_P;                // the six v_b rows of [df/dx | df/du], leading dimension 60
    const double zv[4] = {zvel[0], zvel[1], zvel[2], zvel[3]};
    QM_PFOR(g, c, 49) {
      double eq = 0.0;                             // sum of squares of the constraint values (last column)
      const int col = (c < 18) ? 42 + c : c - 18;
      double fc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (c < 48) for (int cc = 0; cc < 6; ++cc) fc[cc] = vb[60 * cc + col];
      int row = 0;
      for (int ft = 0; ft < 4; ++ft) {
        const bool stance = (mode >> (3 - ft)) & 1;
        for (int d = stance ? 0 : 2; d < 3; ++d, ++row) {
          const double* J = kw + KW_FJ + (3 * ft + d) * QM_NJ;
          double v;
          if (c < 18) v = J[6 + c];                                                    // d v_foot / d u_joint
          else if (c < 48) v = (col >= 6) ? kw[KW_DFV + (3 * ft + d) * QM_NJ + col - 6] : 0.0;   // d v_foot / d x
          else {
            v = kw[KW_FVEL + 3 * ft + d];
            if (!stance) v -= zv[ft];               // normal velocity: v_z - zdot_ref (QMPreComputation.cpp:56-71)
            eq += v * v;
          }
          if (c < 48) for (int cc = 0; cc < 6; ++cc) v += J[cc] * fc[cc];
          io.T[49 * row + c] = v;
        }
      }
      if (c == 48) io.e6[6] = eq;
    }
  }
  QM_PFOR(g, i, 30) io.x2[i] = x[i] + dt * io.f1[i];
  QM_PFOR(g, i, RF_SIZE) io.aux[NA_REF + i] = scr[i];
  g.sync(); QM_TICK(26);
}

// Kinematics at (x + dt f1, u) with derivatives ([upstream] RK2 sensitivity integrator = Heun): second flow map rows.
template <class G>
QM_HDN void node_eval2(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, const double* u, double* kw, NodeIO io) {
  kin_eval(g, M, io.x2, u, 1, kw);               // d(J_i v)/dq is only needed by the constraint rows of the first stage
  flow_rows(g, M, P.gravity, kw, io.x2, u, io.f2, io.fr2);
}

// LQ assembly of one intermediate node from the kinematics products: cost quadratic approximation, discrete dynamics,
// constraint projection, change of input variables; results to the HBM blocks sb / pb and perf[PF_*].
// The change of variables du = Pu dut + Px dx + Pe ([upstream] changeOfInputVariables) only involves the nv pivot rows
// of Px / Pu, so every product is a [.. x nv] x [nv x ..] tile product (mm: FP64 tensor-core tiles on the device).
// B and R are assembled directly with their input index permuted to [pivots | frees | dropped] (index record io.piv),
// Q directly in its HBM block, where the tile product then adds the projection term.
template <class G>
QM_HDN void node_lq(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                    const double* tt, const double* ts, int kt, const double* x, const double* u, const double* xn,
                    double* W, int* WI, NodeIO io, double* sb, double* pb, double* perf, int* status_out) {
  const double m = M.total_mass;
  const int nv = io.piv[PI_NV], nut = io.piv[PI_NUT], nsel = nv + nut;
  const int* SEL = io.piv + PI_SEL;
  const int* POS = io.piv + PI_POS;
  const double* T = io.T;
  const double* ref = io.aux + NA_REF;
  const double* CONE = io.aux + NA_CONE;
  const double* BOX = io.aux + NA_BOX;
  double* T2 = W + TW_T2;
  double* BPM = W + TW_BPM;
  double* RPM = W + TW_RPM;
  double* RPX = io.fr1;                                          // valid from L3 on (fr1 | fr2 are dead then)
  double* RPU = io.T;                                            // valid from L3 on (T is dead once T2 is formed)
  QM_TICK(27);
  // ---- L1: T2 = Dinv T; Q dx, R du; and everything of the assembly that does not need them (the warps that finish the small
  //         products go straight on to the wide loops instead of waiting at a barrier): cost Hessians, discrete dynamics
  mm<1, false>(g, nv, 49, nv, io.dinv, 16, T, 49, (const double*)nullptr, 0, 1.0, T2, 49);
  rows_dot(g, 60, 30, [](int) { return 0.0; },
           [&](int i, int j) {
             const double* wrow = P.Q + 30 * i;                   // P.R follows P.Q in the descriptor
             const double* xv = (i < 30) ? x : u;
             return wrow[j] * (xv[j] - ref[(i < 30 ? RF_X : RF_U) + j]);
           },
           [&](int i, double v) { W[TW_TQ + i] = v; });           // TW_TR follows TW_TQ
  {
    double shift = 0.0;
    for (int ft = 0; ft < 4; ++ft) if ((mode >> (3 - ft)) & 1) shift += CONE[10 * ft + 8] * (-P.fric_hess_shift);
    const double* JE = io.je;
    QM_PFOR(g, idx, 900) {
      const int i = idx / 30, j = idx % 30;
      double qv = P.Q[idx];
      if (i >= 6 && j >= 6) {
        double acc = 0.0;
        for (int rr = 0; rr < 6; ++rr) acc += (rr < 3 ? P.mu_ee_pos : P.mu_ee_ori) * JE[rr * QM_NJ + i - 6] * JE[rr * QM_NJ + j - 6];
        qv += acc;
      }
      double rv = P.R[idx];
      if (i == j) {
        qv += shift; rv += shift;
        if (i >= 24) { qv += BOX[2 * (i - 24) + 1]; rv += BOX[2 * (i - 18) + 1]; }
      }
      if (i < 12 && j < 12 && (i / 3) == (j / 3) && ((mode >> (3 - i / 3)) & 1)) {
        const double* c = CONE + 10 * (i / 3);
        const int a = i % 3, b = j % 3;
        double H = 0.0;
        if (a == 0 && b == 0) H = c[4];
        else if (a == 1 && b == 1) H = c[6];
        else if (a + b == 1) H = c[5];
        rv += c[9] * c[1 + a] * c[1 + b] + c[8] * H;
      }
      sb[SB_Q + idx] = dt * qv;
      const int pi = POS[i];
      if (pi < nsel) RPM[30 * pi + POS[j]] = dt * rv;
    }
    const double* F1 = io.fr1;
    const double* F2 = io.fr2;
    const double hdt = 0.5 * dt, im = 1.0 / m;
    QM_PFOR(g, idx, 900) {
      const int i = idx / 30, j = idx % 30;
      double av = (i == j) ? 1.0 : 0.0, bv = 0.0;
      if (i >= 3 && i < 12) {
        const int rr = i - 3;
        double pa = 0.0, pb2 = 0.0;
        for (int s2 = 0; s2 < 9; ++s2) {
          pa += F2[rr * 60 + 3 + s2] * F1[s2 * 60 + j];
          pb2 += F2[rr * 60 + 3 + s2] * F1[s2 * 60 + 30 + j];
        }
        if (j < 12) pb2 += F2[rr * 60 + (j % 3)] * im;
        else pb2 += F2[rr * 60 + j];
        av += hdt * (F1[rr * 60 + j] + F2[rr * 60 + j] + dt * pa);
        bv = hdt * (F1[rr * 60 + 30 + j] + F2[rr * 60 + 30 + j] + dt * pb2);
      } else if (i < 3) {
        bv = (j < 12 && (j % 3) == i) ? dt * im : 0.0;
      } else {
        bv = (j == i) ? dt : 0.0;
      }
      W[TW_A + idx] = av;
      BPM[30 * i + POS[j]] = bv;
    }
    QM_PFOR(g, i, 30) W[TW_b + i] = x[i] + hdt * (io.f1[i] + io.f2[i]) - xn[i];
    QM_PFOR(g, i, 12) pb[PB_PEF + i] = ((mode >> (3 - i / 3)) & 1) ? 0.0 : -u[i];
    int* role = (int*)(pb + PB_ROLE);
    QM_PFOR(g, i, 32) {
      int v;
      if (i == 30) v = nv;
      else if (i == 31) v = nut;
      else { const int c = POS[i]; v = (c < nv) ? c : ((c < nsel) ? ROLE_FREE + c - nv : ROLE_NONE); }
      role[i] = v;
    }
  }
  g.sync(); QM_TICK(28);
  // ---- L2: cost gradients, projection block (need T2 and Q dx, R du)
  {
    const double* JE = io.je;
    QM_PFOR(g, i, 30) {
      double qv = W[TW_TQ + i], rv = W[TW_TR + i];
      if (i >= 6) {
        const double* e = io.e6;
        for (int rr = 0; rr < 6; ++rr) qv += (rr < 3 ? P.mu_ee_pos : P.mu_ee_ori) * JE[rr * QM_NJ + i - 6] * e[rr];
      }
      if (i >= 24) { qv += BOX[2 * (i - 24)]; rv += BOX[2 * (i - 18)]; }
      if (i < 12 && ((mode >> (3 - i / 3)) & 1)) { const double* c = CONE + 10 * (i / 3); rv += c[8] * c[1 + i % 3]; }
      W[TW_QV + i] = dt * qv;
      W[TW_RV + i] = dt * rv;
    }
    // Pe (permuted): pivot joint velocities from the velocity-constraint rows, frees 0, dropped swing-foot forces -u
    QM_PFOR(g, c, 30) W[TW_PEP + c] = (c < nv) ? -T2[49 * c + 48] : ((c < nsel) ? 0.0 : -u[SEL[c]]);
    QM_PFOR2(g, p, nv, a, QM_NUT) {
      double v = 0.0;
      if (a < nut) { const int fc = SEL[nv + a]; if (fc >= 12) v = -T2[49 * p + fc - 12]; }
      W[TW_PUC + QM_NUT * p + a] = v;
      pb[PB_PU + QM_NUT * p + a] = v;
    }
    QM_PFOR2(g, p, nv, j, 30) pb[PB_PX + 30 * p + j] = -T2[49 * p + 18 + j];     // rows >= nv are never read (role)
    QM_PFOR(g, p, nv) pb[PB_PEC + p] = -T2[49 * p + 48];
  }
  g.sync(); QM_TICK(29);
  // ---- L3: r' = r + R Pe, b~ = b + B Pe; baseline performance; A~ = A + B Px, B~ = B Pu, rows [pivots; frees] of R Px, R Pu
  rows_dot(g, nsel + 30, 30, [&](int i) { return (i < nsel) ? W[TW_RV + SEL[i]] : W[TW_b + i - nsel]; },
           [&](int i, int c) { return ((i < nsel) ? RPM[30 * i + c] : BPM[30 * (i - nsel) + c]) * W[TW_PEP + c]; },
           [&](int i, double v) { if (i < nsel) W[TW_RP + i] = v; else sb[SB_b + i - nsel] = v; });
  {
    // baseline performance of this node: cost value, dynamics defect and equality-constraint SSE
    const double* e = io.e6;
#if defined(__CUDA_ARCH__)
    if (g.warp() == g.nwarps() - 1) {
      const int lane = threadIdx.x & 31;
      double c0 = 0.0, dyn = 0.0;
      if (lane < 30) {
        c0 = 0.5 * ((x[lane] - ref[RF_X + lane]) * W[TW_TQ + lane] + (u[lane] - ref[RF_U + lane]) * W[TW_TR + lane]);
        dyn = W[TW_b + lane] * W[TW_b + lane];
        if (lane < 12) c0 += io.aux[NA_BOXV + lane];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, off); dyn += __shfl_xor_sync(0xffffffffu, dyn, off); }
      if (lane == 0) { W[TW_RED] = c0; W[TW_RED + 1] = dyn; }
      __syncwarp();
    }
    if (g.tid() == g.nt() - 1) {
      double c0 = W[TW_RED] - P.box_offset, dyn = W[TW_RED + 1];
#else
    {
      double c0 = -P.box_offset, dyn = 0.0;
      for (int i = 0; i < 12; ++i) c0 += io.aux[NA_BOXV + i];
      for (int i = 0; i < 30; ++i) c0 += 0.5 * ((x[i] - ref[RF_X + i]) * W[TW_TQ + i] + (u[i] - ref[RF_U + i]) * W[TW_TR + i]);
      for (int i = 0; i < 30; ++i) dyn += W[TW_b + i] * W[TW_b + i];
#endif
      c0 += 0.5 * P.mu_ee_pos * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 0.5 * P.mu_ee_ori * (e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
      double eq = io.e6[6];
      for (int ft = 0; ft < 4; ++ft) {
        if ((mode >> (3 - ft)) & 1) c0 += CONE[10 * ft + 7];
        else eq += u[3 * ft] * u[3 * ft] + u[3 * ft + 1] * u[3 * ft + 1] + u[3 * ft + 2] * u[3 * ft + 2];
      }
      perf[PF_COST] = dt * c0;
      perf[PF_DYN] = dt * dyn;
      perf[PF_EQ] = dt * eq;
    }
  }
  const double* PXn = T2 + 18;                                   // -Px pivot rows, leading dimension 49
  mm<2, false>(g, 30, 30, nv, BPM, 30, PXn, 49, W + TW_A, 30, -1.0, sb + SB_A, 30, 0);
  mm<2, false>(g, nsel, 30, nv, RPM, 30, PXn, 49, (const double*)nullptr, 0, -1.0, RPX, 30, 0);
  if (nut > 0) {
    mm<3, false>(g, 30, nut, nv, BPM, 30, W + TW_PUC, QM_NUT, BPM + nv, 30, 1.0, sb + SB_B, QM_NUT, 6);
    mm<3, false>(g, nsel, nut, nv, RPM, 30, W + TW_PUC, QM_NUT, RPM + nv, 30, 1.0, RPU, QM_NUT, 2);
  }
  g.sync(); QM_TICK(30);
  // ---- L4: Q~ = Q + Px' R Px, P~ = Pu' R Px, R~ = Pu' R Pu, q~ = q + Px' r', r~ = Pu' r'
  mm<2, true>(g, 30, 30, nv, PXn, 49, RPX, 30, sb + SB_Q, 30, -1.0, sb + SB_Q, 30, 0);
  if (nut > 0) {
    mm<2, true>(g, nut, 30, nv, W + TW_PUC, QM_NUT, RPX, 30, RPX + 30 * nv, 30, 1.0, sb + SB_P, 30, 0);
    mm<3, true>(g, nut, nut, nv, W + TW_PUC, QM_NUT, RPU, QM_NUT, RPU + QM_NUT * nv, QM_NUT, 1.0, sb + SB_R, QM_NUT, 4);
  }
  QM_PFOR(g, j, 30) {
    double qv = W[TW_QV + j];
    for (int p = 0; p < nv; ++p) qv -= PXn[49 * p + j] * W[TW_RP + p];
    sb[SB_q + j] = qv;
  }
  QM_PFOR(g, a, nut) {
    double rv = W[TW_RP + nv + a];
    for (int p = 0; p < nv; ++p) rv += W[TW_PUC + QM_NUT * p + a] * W[TW_RP + p];
    sb[SB_r + a] = rv;
  }
  if (g.tid() == 0) {
    sb[SB_NUT] = (double)nut;
    if (io.piv[PI_STATUS]) status_or(status_out, io.piv[PI_STATUS]);
  }
  g.sync(); QM_TICK(31);
}

// Fused form (host port / tests): both kinematics evaluations, the projection pivots and the LQ assembly on one workspace.
template <class G>
QM_HDN void transcribe_node(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                            const double* zvel, const double* tt, const double* ts, int kt, const double* x, const double* u,
                            const double* xn, double* W, int* WI, double* sb, double* pb, double* perf, int* status_out) {
  NodeIO io;
  io.fr1 = W + TW_FR1; io.fr2 = W + TW_FR2; io.f1 = W + TW_F1; io.f2 = W + TW_F2; io.x2 = W + TW_X2;
  io.T = W + TW_T; io.je = W + TW_JE; io.e6 = W + TW_E6; io.aux = W + TW_AUX; io.dinv = W + TW_DINV; io.piv = (int*)(W + TW_PIV);
  node_eval1(g, M, P, t, dt, mode, zvel, tt, ts, kt, x, u, W + TW_KIN, W + TW_REF, io);
  node_eval2(g, M, P, u, W + TW_KIN, io);
  if (g.tid() == 0) projection_pivots_serial(io.T, mode, io.dinv, io.piv);
  g.sync();
  node_lq(g, M, P, t, dt, mode, tt, ts, kt, x, u, xn, W, WI, io, sb, pb, perf, status_out);
}

// Pre-event node: identity jump map, no input, no cost ([upstream] setupEventNode).
template <class G>
QM_HDN void event_node(G g, const double* x, const double* xn, double* sb, double* pb, double* perf) {
  QM_PFOR(g, idx, 900) { sb[SB_A + idx] = (idx / 30 == idx % 30) ? 1.0 : 0.0; sb[SB_Q + idx] = 0.0; }
  QM_PFOR(g, i, 30) { sb[SB_b + i] = x[i] - xn[i]; sb[SB_q + i] = 0.0; }
  QM_PFOR(g, i, 12) pb[PB_PEF + i] = 0.0;
  { int* role = (int*)(pb + PB_ROLE); QM_PFOR(g, i, 32) role[i] = (i < 30) ? ROLE_NONE : 0; }
  if (g.tid() == 0) {
    sb[SB_NUT] = 0.0;
    double d = 0.0;
    for (int i = 0; i < 30; ++i) d += (x[i] - xn[i]) * (x[i] - xn[i]);
    perf[PF_COST] = 0.0; perf[PF_DYN] = d; perf[PF_EQ] = 0.0;
  }
  g.sync();
}

// Terminal node: "finalEndEffector" soft constraint only (QMInterface.cpp:104), Gauss-Newton.
// kw: kinematics workspace, ref: RF_SIZE, e6: 8, dq: 10, JE: [6][24] or nullptr (value only), sb may be nullptr when JE is.
template <class G>
QM_HDN void terminal_node(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, int mode, const double* tt,
                          const double* ts, int kt, const double* x, double* kw, double* ref, double* e6, double* dq, double* JE,
                          double* sb, double* perf) {
  const bool deriv = JE != nullptr;
  if (g.tid() == 0) node_reference(M, P, t, mode, tt, ts, kt, ref);
  kin_eval(g, M, x, (const double*)nullptr, 0, kw, deriv);      // Jacobians only when the Gauss-Newton terms are wanted
  ee_terms(g, kw, ref, e6, dq, JE);
  if (g.tid() == 0) {
    const double* e = e6;
    perf[PF_COST] = 0.5 * P.mu_fee_pos * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 0.5 * P.mu_fee_ori * (e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
    perf[PF_DYN] = 0.0; perf[PF_EQ] = 0.0;
    if (deriv) sb[SB_NUT] = 0.0;
  }
  if (deriv) {
    const double* e = e6;
    QM_PFOR(g, idx, 900) {
      const int i = idx / 30, j = idx % 30;
      double acc = 0.0;
      if (i >= 6 && j >= 6)
        for (int r = 0; r < 6; ++r) acc += (r < 3 ? P.mu_fee_pos : P.mu_fee_ori) * JE[r * QM_NJ + i - 6] * JE[r * QM_NJ + j - 6];
      sb[SB_Q + idx] = acc;
    }
    QM_PFOR(g, i, 30) {
      double acc = 0.0;
      if (i >= 6)
        for (int r = 0; r < 6; ++r) acc += (r < 3 ? P.mu_fee_pos : P.mu_fee_ori) * JE[r * QM_NJ + i - 6] * e[r];
      sb[SB_q + i] = acc;
    }
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------ Riccati
// Workspace. S is kept as a symmetric matrix of which only the tiles on or above the diagonal are valid (MM_UP / MM_XSYM);
// K aliases SB (SB is dead once G is formed); LI holds L^-1 of the Cholesky factor (host route only).
enum { RW_S = 0, RW_SA = 900, RW_SB = 1800, RW_K = RW_SB, RW_H = RW_SB + 540, RW_G = RW_H + 540, RW_LI = RW_G + 324,
       RW_sv = RW_LI + 324, RW_sb = RW_sv + 30, RW_gv = RW_sb + 30, RW_kf = RW_gv + 18, RW_COL = RW_kf + 18, RW_SIZE = RW_COL + 40 };

#if defined(__CUDACC__)
// 1 / a for a pivot on the dependency chain: hardware seed (MUFU.RCP64H, ~2^-20) and two Newton steps, without the range
// checks and the slow path of the IEEE division (pivots of an SPD block are normal numbers; a non-positive pivot is
// flagged by the caller). Accurate to 1-2 ulp.
__device__ __forceinline__ double rcp_newton(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  return fma(y, e, y);
}

// In-place inverse of the symmetric positive definite n x n matrix Gm (shared memory, leading dimension QM_NUT, n <= 18)
// by one warp: Gauss-Jordan sweeps without pivoting, column j of the matrix in the registers of lane j. In every sweep
// the owner of the pivot column publishes the negated, scaled column through shared memory (col: 2 x 20 doubles, double
// buffered, 16-byte stores) and all lanes read it back with nine unpredicated broadcast loads and update all 18 rows
// (rows >= n carry zero multipliers): one __syncwarp per sweep on the dependency chain, no shuffles.
__device__ __forceinline__ void spd_inverse_warp(double* Gm, int n, double* col, int* status) {
  const int lane = threadIdx.x & 31;
  const bool active = lane < n;
  double r[QM_NUT];
#pragma unroll
  for (int i = 0; i < QM_NUT; ++i) r[i] = (active && i < n) ? Gm[QM_NUT * i + lane] : ((i == lane) ? 1.0 : 0.0);
  bool bad = false;
#pragma unroll
  for (int c = 0; c < QM_NUT; ++c) {
    if (c < n) {                                   // uniform across the warp
      double* cb = col + 20 * (c & 1);
      const double rc = r[c];                      // lane c: the pivot a_cc; other lanes: a_cj of their column
      double rce = rc;                             // 0 in the pivot lane: its column is final once scaled (no per-row select)
      if (lane == c) {
        if (!(rc > 0.0)) bad = true;
        const double ip = rcp_newton(rc);
#pragma unroll
        for (int i = 0; i < QM_NUT; ++i) r[i] = (i == c) ? ip : -r[i] * ip;   // -a_ic / a_cc, and 1 / a_cc in the pivot position
#pragma unroll
        for (int i = 0; i < QM_NUT; i += 2) {
          double2 v;
          v.x = r[i]; v.y = r[i + 1];
          reinterpret_cast<double2*>(cb)[i >> 1] = v;
        }
        rce = 0.0;
      }
      __syncwarp();
      double2 m2[QM_NUT / 2];
#pragma unroll
      for (int q = 0; q < QM_NUT / 2; ++q) m2[q] = reinterpret_cast<const double2*>(cb)[q];
#pragma unroll
      for (int i = 0; i < QM_NUT; ++i) {
        if (i != c) {
          const double mlt = (i & 1) ? m2[i >> 1].y : m2[i >> 1].x;
          r[i] = fma(mlt, rce, r[i]);
        }
      }
      const double p = (c & 1) ? m2[c >> 1].y : m2[c >> 1].x;
      r[c] = (lane == c) ? p : rc * p;
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < QM_NUT; ++i) if (i < n) Gm[QM_NUT * i + lane] = r[i];
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) status_or(status, ST_CHOL);
}
#endif

// One backward stage, in two halves so the caller can re-use the stage buffer (prefetch the next block) in between.
// Dense products are tile products (mm); the symmetric ones only form the tiles on or above the diagonal.
//   riccati_stage_a: P1  SA = S A, SB = S B, sb = s + S b
//                    P2  G = R + B' SB, g = r + B' sb
//                    P3  narrow: G^-1 (dependency chain, one warp)
//                        rest  : H = P + B' SA;  S <- Q + A' SA, s <- q + A' sb   (independent of the gains)
//   riccati_stage_b: P4  K = -G^-1 H, kff = -G^-1 g
//                    P5  S += H' K, s += H' kff;  gains to HBM
template <class G>
QM_HDN int riccati_stage_a(G g, const double* st, double* W, int* status) {
  const int nut = (int)st[SB_NUT];
  const double* A = st + SB_A; const double* B = st + SB_B; const double* b = st + SB_b;
  double* S = W + RW_S; double* s = W + RW_sv;
  // ---- P1
  mm<4, false, MM_XSYM>(g, 30, 30, 30, S, 30, A, 30, (const double*)nullptr, 0, 1.0, W + RW_SA, 30);
  if (nut > 0) mm<3, false, MM_XSYM>(g, 30, nut, 30, S, 30, B, QM_NUT, (const double*)nullptr, 0, 1.0, W + RW_SB, QM_NUT);
  rows_dot(g, 30, 30, [&](int i) { return s[i]; },
           [&](int i, int j) { return ((j < i) ? S[30 * j + i] : S[30 * i + j]) * b[j]; },
           [&](int i, double v) { W[RW_sb + i] = v; });
  g.sync(); QM_TICK(1);
  // ---- P2
  if (nut > 0) {
    mm<1, true>(g, nut, nut, 30, B, QM_NUT, W + RW_SB, QM_NUT, st + SB_R, QM_NUT, 1.0, W + RW_G, QM_NUT);
    rows_dot<true>(g, nut, 30, [&](int a) { return st[SB_r + a]; },
             [&](int a, int k) { return B[QM_NUT * k + a] * W[RW_sb + k]; },
             [&](int a, double v) { W[RW_gv + a] = v; });
    g.sync(); QM_TICK(2);
  }
  // ---- P3 narrow: Ginv = G^-1 in place. Device: register Gauss-Jordan on one warp; host: Cholesky route.
#if defined(__CUDA_ARCH__)
  if (g.narrow_active() && nut > 0) { spd_inverse_warp(W + RW_G, nut, W + RW_COL, status); QM_TICK(7); }
#else
  if (g.narrow_active() && nut > 0) {
    auto w0 = g.narrow();
    double* Gm = W + RW_G;
    for (int c = 0; c < nut; ++c) {
      if (w0.tid() == 0) {
        double d = Gm[QM_NUT * c + c];
        for (int k = 0; k < c; ++k) d -= Gm[QM_NUT * c + k] * Gm[QM_NUT * c + k];
        if (!(d > 0.0)) { status_or(status, ST_CHOL); d = 1e-300; }
        Gm[QM_NUT * c + c] = sqrt(d);
      }
      w0.sync();
      QM_PFOR(w0, ii, nut - c - 1) {
        const int i = c + 1 + ii;
        double v = Gm[QM_NUT * i + c];
        for (int k = 0; k < c; ++k) v -= Gm[QM_NUT * i + k] * Gm[QM_NUT * c + k];
        Gm[QM_NUT * i + c] = v / Gm[QM_NUT * c + c];
      }
      w0.sync();
    }
    QM_PFOR(w0, c, nut) {
      for (int i = 0; i < nut; ++i) {
        double v = 0.0;
        if (i >= c) {
          v = (i == c) ? 1.0 : 0.0;
          for (int k = c; k < i; ++k) v -= Gm[QM_NUT * i + k] * W[RW_LI + QM_NUT * k + c];
          v /= Gm[QM_NUT * i + i];
        }
        W[RW_LI + QM_NUT * i + c] = v;
      }
    }
    w0.sync();
    QM_PFOR(w0, idx, nut * nut) {       // Ginv = LI' LI
      const int i = idx / nut, j = idx % nut;
      double acc = 0.0;
      for (int k = (i > j ? i : j); k < nut; ++k) acc += W[RW_LI + QM_NUT * k + i] * W[RW_LI + QM_NUT * k + j];
      Gm[QM_NUT * i + j] = acc;
    }
  }
#endif
  // ---- P3 rest: S <- Q + A' SA (upper tiles); H = P + B' SA; s <- q + A' sb   (S, s were consumed in P1)
  if (g.rest_active()) {
    auto r_ = g.rest();
    mm<1, true, MM_UP>(r_, 30, 30, 30, A, 30, W + RW_SA, 30, st + SB_Q, 30, 1.0, S, 30);
    if (nut > 0) mm<1, true>(r_, nut, 30, 30, B, QM_NUT, W + RW_SA, 30, st + SB_P, 30, 1.0, W + RW_H, 30, 1);
    rows_dot<true>(r_, 30, 30, [&](int i) { return st[SB_q + i]; },
             [&](int i, int k) { return A[30 * k + i] * W[RW_sb + k]; },
             [&](int i, double v) { s[i] = v; });
  }
  g.sync(); QM_TICK(3);
  return nut;
}

template <class G>
QM_HDN void riccati_stage_b(G g, int nut, double* W, double* gb) {
  double* S = W + RW_S; double* s = W + RW_sv;
  double* Gm = W + RW_G;
  if (nut > 0) {
    // ---- P4: K = -Ginv H, kff = -Ginv g
    mm<2, false>(g, nut, 30, nut, Gm, QM_NUT, W + RW_H, 30, (const double*)nullptr, 0, -1.0, W + RW_K, 30);
    rows_dot(g, nut, nut, [](int) { return 0.0; },
             [&](int a, int k) { return -Gm[QM_NUT * a + k] * W[RW_gv + k]; },
             [&](int a, double v) { W[RW_kf + a] = v; });
    g.sync(); QM_TICK(4);
    // ---- P5: S += H' K (upper tiles), s += H' kff
    mm<1, true, MM_UP>(g, 30, 30, nut, W + RW_H, 30, W + RW_K, 30, S, 30, 1.0, S, 30);
    rows_dot<true>(g, 30, nut, [&](int i) { return s[i]; },
             [&](int i, int a) { return W[RW_H + 30 * a + i] * W[RW_kf + a]; },
             [&](int i, double v) { s[i] = v; });
  }
  QM_PFOR(g, idx, QM_NUT * 30) gb[GB_K + idx] = (idx < 30 * nut) ? W[RW_K + idx] : 0.0;
  QM_PFOR(g, a, QM_NUT) gb[GB_KFF + a] = (a < nut) ? W[RW_kf + a] : 0.0;
  g.sync(); QM_TICK(5);
}

// One forward stage: dut = K dx + kff; du from the compact projection block; dx+ = A dx + B dut + b; armijo += q.dx + r.dut
// st: forward part of the stage block. W: [30:48] dut, [80] armijo accumulator; v: dx (30), dxn: dx of the next node (30) --
// the caller alternates the two between W[0:30] and W[48:78], so no copy (and no barrier for it) separates two stages.
template <class G>
QM_HDN void rollout_stage(G g, const double* st, const double* pb, const double* gb, double* W, const double* v, double* dxn,
                          double* du_out) {
  const int nut = (int)st[SB_NUT];
  double* dut = W + 30;
  rows_dot(g, nut, 30, [&](int a) { return gb[GB_KFF + a]; },
           [&](int a, int j) { return gb[GB_K + 30 * a + j] * v[j]; },
           [&](int a, double val) { dut[a] = val; });
  g.sync();
  const int* role = (const int*)(pb + PB_ROLE);
  // 61 rows of the form init + p1 . dx + p2 . dut (a 30- and a nut-vector per row, no per-element case distinction):
  //   rows 0..29  dx+ = b + A dx + B dut
  //   rows 30..59 du_i: pivot rows Pe + Px dx + Pu dut; free inputs: dut; dropped: Pe
  //   row  60     armijo += q . dx + r . dut
  auto row = [&](int r, const double** p1, const double** p2, double* init) -> bool {
    if (r < 30) { *p1 = st + SB_A + 30 * r; *p2 = st + SB_B + QM_NUT * r; *init = st[SB_b + r]; return true; }
    if (r == 60) { *p1 = st + SB_q; *p2 = st + SB_r; *init = W[80]; return true; }
    const int i = r - 30, rl = role[i];
    if (rl < ROLE_FREE) { *p1 = pb + PB_PX + 30 * rl; *p2 = pb + PB_PU + QM_NUT * rl; *init = pb[PB_PEC + rl]; return true; }
    *init = (rl < ROLE_NONE) ? dut[rl - ROLE_FREE] : ((i < 12) ? pb[PB_PEF + i] : 0.0);
    return false;
  };
  auto put = [&](int r, double val) {
    if (r < 30) dxn[r] = val;
    else if (r == 60) W[80] = val;
    else du_out[r - 30] = (nut > 0) ? val : 0.0;
  };
#if defined(__CUDA_ARCH__)
  for (int base = 0; base < 4 * 61; base += g.nt()) {           // four lanes per row, partial sums combined with shuffles
    const int t = base + g.tid(), rl = t >> 2, part = t & 3;
    const int r = (rl & ~7) | mm_rowperm(rl & 7);               // rows two apart within a half warp (see rows_dot)
    const double *p1 = nullptr, *p2 = nullptr;
    double init = 0.0, acc = 0.0;
    if (r < 61 && row(r, &p1, &p2, &init)) {
      for (int j = part; j < 30; j += 4) acc += p1[j] * v[j];
      for (int j = part; j < nut; j += 4) acc += p2[j] * dut[j];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (r < 61 && part == 0) put(r, init + acc);
  }
#else
  QM_PFOR(g, r, 61) {
    const double *p1 = nullptr, *p2 = nullptr;
    double init = 0.0, acc = 0.0;
    if (row(r, &p1, &p2, &init)) {
      for (int j = 0; j < 30; ++j) acc += p1[j] * v[j];
      for (int j = 0; j < nut; ++j) acc += p2[j] * dut[j];
    }
    put(r, init + acc);
  }
#endif
  g.sync();
}

// Feedback gain of one node in the original input coordinates ([upstream] SqpSolver::toPrimalSolution with
// useFeedbackPolicy, task.info:90): K = Pu K~ + Px from the Riccati gain K~ [nut][30] and the compact projection block.
// Kout [30][30]; rows of dropped inputs (swing-foot forces) are zero. One work item per entry.
template <class G>
QM_HDN void feedback_gain_node(G g, const double* pb, const double* gb, double* Kout) {
  const int* role = (const int*)(pb + PB_ROLE);
  const int nut = role[31];
  QM_PFOR(g, idx, 900) {
    const int i = idx / 30, j = idx - 30 * i;
    const int rl = role[i];
    double v = 0.0;
    if (nut > 0) {
      if (rl < ROLE_FREE) {
        v = pb[PB_PX + 30 * rl + j];
        for (int a = 0; a < nut; ++a) v += pb[PB_PU + QM_NUT * rl + a] * gb[GB_K + 30 * a + j];
      } else if (rl < ROLE_NONE) {
        v = gb[GB_K + 30 * (rl - ROLE_FREE) + j];
      }
    }
    Kout[idx] = v;
  }
}

// Node whose gain node k repeats: pre-event nodes and the final node carry no input of their own and repeat the previous
// node (toPrimalSolution does the same for the inputs). Returns -1 when there is no source (zero gain).
QM_HD int feedback_gain_source(int nn, const int32_t* node_flag, int k) {
  while (k > 0 && (k == nn - 1 || node_flag[k] == EV_PRE)) --k;
  return (k == nn - 1 || node_flag[k] == EV_PRE) ? -1 : k;
}

// [upstream] FilterLinesearch::acceptStep
QM_HD bool accept_step(const qmb200_solver_desc& S, double base_merit, double base_viol, double new_merit, double new_viol, double armijo) {
  if (new_viol > S.g_max) return new_viol < (1.0 - S.gamma_c) * base_viol;
  if (new_viol < S.g_min && base_viol < S.g_min && armijo < 0.0) return new_merit < base_merit + S.armijo_factor * armijo;
  return new_merit < base_merit - S.gamma_c * base_viol || new_viol < (1.0 - S.gamma_c) * base_viol;
}

}  // namespace qm
