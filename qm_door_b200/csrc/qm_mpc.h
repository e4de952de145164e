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
  SB_Q = SB_b + 30,                  // [30][30]
  SB_P = SB_Q + 900,                 // [18][30]
  SB_R = SB_P + 30 * QM_NUT,         // [18][18]
  SB_q = SB_R + QM_NUT * QM_NUT,     // [30]
  SB_r = SB_q + 30,                  // [18]
  SB_NUT = SB_r + QM_NUT,            // reduced input dimension of this node (as double)
  SB_SIZE = 3296
};
enum {
  PB_PU = 0,                         // [30][18]
  PB_PX = PB_PU + 30 * QM_NUT,       // [30][30]
  PB_PE = PB_PX + 900,               // [30]
  PB_SIZE = 1472
};
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
    *v = -mu * log(h); *d1 = -mu / h; *d2 = mu / (h * h);
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
      if (ev == EV_PRE) { node_t[n] = tn; node_flag[n] = EV_POST; ++n; }
    } else {
      node_t[n - 1] = tn; node_flag[n - 1] = ev;
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

// [upstream] multiple_shooting::initializeStateInputTrajectories; one thread per (problem, component c<60).
QM_HDN void init_guess_component(const qmb200_model_desc& M, const qmb200_problem_desc& P, double weak_eps, int c, const double* x0, int nn,
                                 const double* node_t, const int32_t* node_flag, const double* node_ts, const double* node_dt,
                                 const int32_t* node_mode, int nprev, const double* prev_t, const double* prev_x,
                                 const double* prev_u, double* xs, double* us) {
  const bool has_prev = nprev >= 2;
  const double till_x = has_prev ? prev_t[nprev - 1] : node_t[0];
  const double till_u = has_prev ? prev_t[nprev - 2] : node_t[0];
  const int n = nn - 1;
  // query times increase with the node index, so the interpolation segment is searched monotonically:
  // `part` = number of previous-solution times strictly below the query (LinearInterpolation::timeSegment, lower_bound)
  int part = 0;
  if (c < 30) {
    double xc;
    const double t_init = node_ts[0];
    if (t_init < till_x) {
      while (part < nprev && prev_t[part] < t_init) ++part;
      const int i = (part == 0) ? 0 : part - 1;        // t_init < till_x = prev_t[last] => i < last
      const double a = (part == 0 && !(t_init == prev_t[0])) ? 1.0 : (prev_t[i + 1] - t_init) / (prev_t[i + 1] - prev_t[i]);
      xc = a * prev_x[30 * i + c] + (1.0 - a) * prev_x[30 * (i + 1) + c];
    } else xc = x0[c];
    xs[c] = xc;
    for (int k = 0; k < n; ++k) {
      if (node_flag[k] != EV_PRE) {
        const double t = node_ts[k], tn = node_t[k + 1] - (node_flag[k + 1] == EV_PRE ? weak_eps : 0.0);
        if (!(t > till_u || tn > till_x)) {
          while (part < nprev && prev_t[part] < tn) ++part;
          int i; double a;
          if (part == 0) { i = 0; a = (tn == prev_t[0]) ? (prev_t[1] - tn) / (prev_t[1] - prev_t[0]) : 1.0; }
          else if (part - 1 < nprev - 1) { i = part - 1; a = (prev_t[i + 1] - tn) / (prev_t[i + 1] - prev_t[i]); }
          else { i = nprev - 2; a = 0.0; }
          xc = a * prev_x[30 * i + c] + (1.0 - a) * prev_x[30 * (i + 1) + c];
        }
      }
      xs[30 * (k + 1) + c] = xc;
    }
  } else {
    const int cu = c - 30;
    for (int k = 0; k < n; ++k) {
      double uc = 0.0;
      if (node_flag[k] != EV_PRE) {
        const double t = node_ts[k], tn = node_t[k + 1] - (node_flag[k + 1] == EV_PRE ? weak_eps : 0.0);
        if (t > till_u || tn > till_x) {
          // QMInitializer::compute (qm_interface/src/initialization/QMInitializer.cpp:33-41): weight compensation
          const int md = node_mode[k];
          const int ns = ((md >> 3) & 1) + ((md >> 2) & 1) + ((md >> 1) & 1) + (md & 1);
          if (cu < 12 && (cu % 3) == 2 && ((md >> (3 - cu / 3)) & 1)) uc = M.total_mass * P.gravity / ns;
        } else {
          while (part < nprev && prev_t[part] < t) ++part;
          int i; double a;
          if (part == 0) { i = 0; a = (t == prev_t[0]) ? (prev_t[1] - t) / (prev_t[1] - prev_t[0]) : 1.0; }
          else if (part - 1 < nprev - 1) { i = part - 1; a = (prev_t[i + 1] - t) / (prev_t[i + 1] - prev_t[i]); }
          else { i = nprev - 2; a = 0.0; }
          uc = a * prev_u[30 * i + cu] + (1.0 - a) * prev_u[30 * (i + 1) + cu];
        }
      }
      us[30 * k + cu] = uc;
    }
    us[30 * n + cu] = 0.0;
  }
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
        v += t[i - 3];
      }
      v /= m;
    } else {
      v = w[KW_VEL + i - 6];
    }
    f[i] = v;
  }
  if (Fr != nullptr) {
    QM_PFOR(g, idx, 540) {
      const int r = idx / 60, c = idx % 60;
      double v = 0.0;
      if (r < 3) {
        if (c >= 6 && c < 30) {
          const int k = c - 6;
          for (int ft = 0; ft < 4; ++ft) {
            double d[3], t[3];
            for (int rr = 0; rr < 3; ++rr) d[rr] = w[KW_FJ + (3 * ft + rr) * QM_NJ + k] - w[KW_ACM + rr * QM_NJ + k] / m;
            cross3(d, u + 3 * ft, t);
            v += t[r];
          }
          v /= m;
        } else if (c >= 30 && c < 42) {
          const int ft = (c - 30) / 3, d = (c - 30) % 3;
          double arm[3], e[3] = {0, 0, 0}, t[3];
          e[d] = 1.0;
          for (int rr = 0; rr < 3; ++rr) arm[rr] = w[KW_FPOS + 3 * ft + rr] - w[KW_COM + rr];
          cross3(arm, e, t);
          v = t[r] / m;
        }
      } else {
        const int rr = r - 3;
        const double* Bi = w + KW_ABINV + 6 * rr;
        if (c < 6) {
          v = m * Bi[c];
        } else if (c < 30) {
          const int k = c - 6;
          for (int cc = 0; cc < 6; ++cc) v -= Bi[cc] * w[KW_DH + cc * QM_NJ + k];
        } else if (c >= 42) {
          const int l = c - 42;
          for (int cc = 0; cc < 6; ++cc) v -= Bi[cc] * w[KW_ACM + cc * QM_NJ + 6 + l];
        }
      }
      Fr[idx] = v;
      if (vb_shadow != nullptr && r >= 3) vb_shadow[idx - 180] = v;   // rows of v_b = A_b^-1(...) kept close for the constraint rows
    }
  }
  g.sync();
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
  c[1] = -F[0] / tn; c[2] = -F[1] / tn; c[3] = P.fric_mu;
  const double p32 = tn * t2;
  c[4] = -(F[1] * F[1] + P.fric_reg) / p32;
  c[5] = F[0] * F[1] / p32;
  c[6] = -(F[0] * F[0] + P.fric_reg) / p32;
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
enum {
  TW_KIN = 0,                       // kinematics workspace; later reused for RPX / RPU
  TW_FR1 = TW_KIN + KW_SIZE,           // [9][60]
  TW_FR2 = TW_FR1 + 540,            // [9][60]
  TW_RPX = TW_KIN,                  // [30][30] alias (kinematics dead)
  TW_RPU = TW_KIN + 900,            // [30][18] alias
  TW_A = TW_FR2 + 540,              // [30][30]
  TW_B = TW_A + 900,
  TW_Q = TW_B + 900,
  TW_R = TW_Q + 900,
  TW_PX = TW_FR1,                   // [30][30] alias (FR1/FR2 dead once A, B are assembled)
  TW_PU = TW_R + 900,               // [30][18]
  TW_T = TW_PU + 540,               // [12][49] = [Dv | C | e]
  TW_JE = TW_T + 588,               // [6][24]
  TW_REF = TW_JE + 144,             // RF_SIZE (+ 10 scratch for the ee rotation map)
  TW_F1 = TW_REF + RF_SIZE + 10,    // 30
  TW_F2 = TW_F1 + 30,
  TW_X2 = TW_F2 + 30,
  TW_b = TW_X2 + 30,
  TW_QV = TW_b + 30,                // q
  TW_RV = TW_QV + 30,               // r -> r'
  TW_PE = TW_RV + 30,
  TW_DX = TW_PE + 30,               // x - x_ref
  TW_DU = TW_DX + 30,
  TW_TQ = TW_DU + 30,               // Q dx
  TW_TR = TW_TQ + 30,               // R du
  TW_E6 = TW_TR + 30,               // 6 (+2 pad)
  TW_DQ = TW_E6 + 8,                // 9 (+1)
  TW_CONE = TW_DQ + 10,             // [4][10]
  TW_BOX = TW_CONE + 40,            // [12][2]: gradient, hessian of the arm boxes (6 position, 6 velocity)
  TW_BOXV = TW_BOX + 24,            // [12] their values
  TW_ROWBEST = TW_BOXV + 12,        // 12
  TW_FAC = TW_ROWBEST + 12,         // 12
  TW_SCAL = TW_FAC + 12,            // misc scalars
  TW_AUX = TW_SCAL + 8,             // NA_SIZE: references / barrier terms of the node in the fused (host) path
  TW_SIZE = TW_AUX + 144
};
enum {                               // integer workspace
  TI_ROWARG = 0,                    // 12
  TI_PIVCOL = 12,                   // 12: joint-velocity column eliminated by row p
  TI_FCOLS = 24,                    // 18: input index of reduced input a
  TI_ISPIV = 42,                    // 18
  TI_NV = 60, TI_NUT = 61, TI_PR = 62, TI_PC = 63, TI_STATUS = 64,
  TI_SIZE = 68
};

// Intermediate products of one node handed from the kinematics evaluations to the LQ assembly. In the fused (host)
// path they live in the transcription workspace; in the split CUDA path k_kin writes them to HBM (KS_* layout) and
// k_lq stages them into shared memory with one bulk copy.
struct NodeIO {
  double* fr1; double* fr2;   // [9][60] non-trivial rows of [df/dx | df/du] at (x,u) and (x + dt f1, u)
  double* f1;  double* f2;    // [30] flow map values
  double* x2;                 // [30] x + dt f1
  double* T;                  // [12][49] velocity-constraint rows [Dv | C | e]
  double* je;                 // [6][24] end-effector error Jacobian
  double* e6;                 // [8] end-effector error (6), [6] = sum of squared velocity-constraint values
  double* aux;                // [NA_SIZE] references and barrier terms of the node (functions of t, x, u only)
};
// aux layout: reference (x_ref, u_nominal, ee pose), friction-cone terms per foot, arm box gradients / Hessians / values
enum { NA_REF = 0, NA_CONE = RF_SIZE, NA_BOX = NA_CONE + 40, NA_BOXV = NA_BOX + 24, NA_SIZE = NA_BOXV + 12 };
enum { KS_FR1 = 0, KS_FR2 = 540, KS_F1 = 1080, KS_F2 = 1110, KS_X2 = 1140, KS_T = 1170, KS_JE = 1758, KS_E6 = 1902, KS_AUX = 1910,
       KS_SIZE = ((KS_AUX + NA_SIZE + 3) / 4) * 4 };
QM_HD NodeIO node_io_at(double* base) {
  NodeIO io;
  io.fr1 = base + KS_FR1; io.fr2 = base + KS_FR2; io.f1 = base + KS_F1; io.f2 = base + KS_F2;
  io.x2 = base + KS_X2; io.T = base + KS_T; io.je = base + KS_JE; io.e6 = base + KS_E6; io.aux = base + KS_AUX;
  return io;
}

// Kinematics at (x,u) with derivatives: flow map rows, end-effector terms, constraint rows (QMInterface.cpp:116-131), x2.
// kw: kinematics workspace; scr: RF_SIZE + 10 doubles of scratch.
template <class G>
QM_HDN void node_eval1(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                       const double* zvel, const double* tt, const double* ts, int kt, const double* x, const double* u,
                       double* kw, double* scr, NodeIO io) {
  int nvc = 0;
  for (int ft = 0; ft < 4; ++ft) nvc += ((mode >> (3 - ft)) & 1) ? 3 : 1;
  const int nv = nvc;                 // velocity-constraint rows: 3 per stance foot, 1 per swing foot
  if (g.tid() == 0) node_reference(M, P, t, mode, tt, ts, kt, scr);
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
  kin_eval(g, M, x, u, true, kw);
  // the six v_b rows of [df/dx | df/du] are mirrored into the (now dead) placement arrays R | P | AX of the workspace
  flow_rows(g, M, P.gravity, kw, x, u, io.f1, io.fr1, kw + KW_R);
  ee_terms(g, kw, scr, io.e6, scr + RF_SIZE, io.je);
  {
    const double* Fr1 = kw + KW_R - 180;   // Fr1[(3 + cc) * 60 + c] -> shadow[cc * 60 + c]
    QM_PFOR(g, idx, nv * 49) {
      const int row = idx / 49, c = idx % 49;
      // map row -> (foot, component)
      int ft = 0, d = 0, acc = 0;
      for (int f2 = 0; f2 < 4; ++f2) {
        const int cnt = ((mode >> (3 - f2)) & 1) ? 3 : 1;
        if (row < acc + cnt) { ft = f2; d = (cnt == 3) ? (row - acc) : 2; break; }
        acc += cnt;
      }
      const double* J = kw + KW_FJ + (3 * ft + d) * QM_NJ;
      double v;
      if (c < 18) {               // d v_foot / d u_joint
        v = J[6 + c];
        for (int cc = 0; cc < 6; ++cc) v += J[cc] * Fr1[(3 + cc) * 60 + 42 + c];
      } else if (c < 48) {        // d v_foot / d x
        const int xc = c - 18;
        v = (xc >= 6) ? kw[KW_DFV + (3 * ft + d) * QM_NJ + xc - 6] : 0.0;
        for (int cc = 0; cc < 6; ++cc) v += J[cc] * Fr1[(3 + cc) * 60 + xc];
      } else {
        v = kw[KW_FVEL + 3 * ft + d];
        if (!((mode >> (3 - ft)) & 1)) v -= zvel[ft];       // normal velocity: v_z - zdot_ref (QMPreComputation.cpp:56-71)
      }
      io.T[idx] = v;
    }
  }
  QM_PFOR(g, i, 30) io.x2[i] = x[i] + dt * io.f1[i];
  QM_PFOR(g, i, RF_SIZE) io.aux[NA_REF + i] = scr[i];
  g.sync();
  if (g.tid() == 0) {
    double eq = 0.0;
    for (int rr = 0; rr < nv; ++rr) eq += io.T[49 * rr + 48] * io.T[49 * rr + 48];
    io.e6[6] = eq;
  }
  g.sync();
}

// Kinematics at (x + dt f1, u) with derivatives ([upstream] RK2 sensitivity integrator = Heun): second flow map rows.
template <class G>
QM_HDN void node_eval2(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, const double* u, double* kw, NodeIO io) {
  kin_eval(g, M, io.x2, u, true, kw);
  flow_rows(g, M, P.gravity, kw, io.x2, u, io.f2, io.fr2);
}

// LQ assembly of one intermediate node from the kinematics products: cost quadratic approximation, discrete dynamics,
// constraint projection, change of input variables; results to the HBM blocks sb / pb and perf[PF_*].
template <class G>
QM_HDN void node_lq(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                    const double* tt, const double* ts, int kt, const double* x, const double* u, const double* xn,
                    double* W, int* WI, NodeIO io, double* sb, double* pb, double* perf, int* status_out) {
  const double m = M.total_mass;
  int nvc = 0;
  for (int ft = 0; ft < 4; ++ft) nvc += ((mode >> (3 - ft)) & 1) ? 3 : 1;
  const int nv = nvc;
  double* T = io.T;
  // ---- L1 narrow: projection by Gauss-Jordan with full pivoting on Dv ([upstream] luConstraintProjection)
  if (g.narrow_active()) {
    auto w0 = g.narrow();
    if (w0.tid() == 0) WI[TI_STATUS] = 0;
    for (int step = 0; step < nv; ++step) {
      QM_PFOR(w0, r, nv) {
        double best = -1.0; int arg = 0;
        if (r >= step) {
          for (int c = 0; c < 18; ++c) { const double a = fabs(T[49 * r + c]); if (a > best) { best = a; arg = c; } }
        }
        W[TW_ROWBEST + r] = best; WI[TI_ROWARG + r] = arg;
      }
      w0.sync();
      if (w0.tid() == 0) {
        int pr = step; double best = W[TW_ROWBEST + step];
        for (int r = step + 1; r < nv; ++r) if (W[TW_ROWBEST + r] > best) { best = W[TW_ROWBEST + r]; pr = r; }
        WI[TI_PR] = pr; WI[TI_PC] = WI[TI_ROWARG + pr]; WI[TI_PIVCOL + step] = WI[TI_ROWARG + pr];
        if (!(best > 1e-12)) WI[TI_STATUS] |= ST_RANK;
      }
      w0.sync();
      const int pr = WI[TI_PR], pc = WI[TI_PC];
      if (pr != step) {
        QM_PFOR(w0, c, 49) { const double a = T[49 * step + c]; T[49 * step + c] = T[49 * pr + c]; T[49 * pr + c] = a; }
        w0.sync();
      }
      QM_PFOR(w0, r, nv) W[TW_FAC + r] = T[49 * r + pc];
      w0.sync();
      const double ipiv = 1.0 / W[TW_FAC + step];
      QM_PFOR(w0, idx, nv * 49) {
        const int r = idx / 49, c = idx % 49;
        if (r != step) T[idx] -= W[TW_FAC + r] * ipiv * T[49 * step + c];
      }
      w0.sync();
      QM_PFOR(w0, c, 49) T[49 * step + c] *= ipiv;
      w0.sync();
    }
    if (w0.tid() == 0) {
      for (int l = 0; l < 18; ++l) WI[TI_ISPIV + l] = 0;
      for (int p = 0; p < nv; ++p) WI[TI_ISPIV + WI[TI_PIVCOL + p]] = 1;
      int a = 0;
      for (int ft = 0; ft < 4; ++ft)
        if ((mode >> (3 - ft)) & 1) { WI[TI_FCOLS + a] = 3 * ft; WI[TI_FCOLS + a + 1] = 3 * ft + 1; WI[TI_FCOLS + a + 2] = 3 * ft + 2; a += 3; }
      for (int l = 0; l < 18; ++l) if (!WI[TI_ISPIV + l]) WI[TI_FCOLS + a++] = 12 + l;
      WI[TI_NUT] = a;
      WI[TI_NV] = nv;
    }
  }
  // ---- L1 rest: cost quadratic approximation (forward Euler, * dt), discrete dynamics. References and barrier terms
  //      come with the kinematics products (io.aux), so no serial work sits in front of the wide assembly.
  if (g.rest_active()) {
    auto r = g.rest();
    const double* ref = io.aux + NA_REF;
    const double* CONE = io.aux + NA_CONE;
    const double* BOX = io.aux + NA_BOX;
    QM_PFOR(r, i, 60) {
      double acc = 0.0;
      if (i < 30) { for (int j = 0; j < 30; ++j) acc += P.Q[30 * i + j] * (x[j] - ref[RF_X + j]); W[TW_TQ + i] = acc; }
      else { const int ii = i - 30; for (int j = 0; j < 30; ++j) acc += P.R[30 * ii + j] * (u[j] - ref[RF_U + j]); W[TW_TR + ii] = acc; }
    }
    r.sync();
    double shift = 0.0;
    for (int ft = 0; ft < 4; ++ft) if ((mode >> (3 - ft)) & 1) shift += CONE[10 * ft + 8] * (-P.fric_hess_shift);
    const double* JE = io.je;

    QM_PFOR(r, idx, 900) {
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
      W[TW_Q + idx] = dt * qv;
      W[TW_R + idx] = dt * rv;
    }
    QM_PFOR(r, i, 30) {
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
    {
      const double* F1 = io.fr1;
      const double* F2 = io.fr2;
      const double hdt = 0.5 * dt;
      QM_PFOR(r, idx, 900) {
        const int i = idx / 30, j = idx % 30;
        double av = (i == j) ? 1.0 : 0.0, bv = 0.0;
        if (i >= 3 && i < 12) {
          const int rr = i - 3;
          double pa = 0.0, pb2 = 0.0;
          for (int s2 = 0; s2 < 9; ++s2) {
            pa += F2[rr * 60 + 3 + s2] * F1[s2 * 60 + j];
            pb2 += F2[rr * 60 + 3 + s2] * F1[s2 * 60 + 30 + j];
          }
          if (j < 12) pb2 += F2[rr * 60 + (j % 3)] / m;
          else pb2 += F2[rr * 60 + j];
          av += hdt * (F1[rr * 60 + j] + F2[rr * 60 + j] + dt * pa);
          bv = hdt * (F1[rr * 60 + 30 + j] + F2[rr * 60 + 30 + j] + dt * pb2);
        } else if (i < 3) {
          bv = (j < 12 && (j % 3) == i) ? dt / m : 0.0;
        } else {
          bv = (j == i) ? dt : 0.0;
        }
        W[TW_A + idx] = av;
        W[TW_B + idx] = bv;
      }
      QM_PFOR(r, i, 30) W[TW_b + i] = x[i] + hdt * (io.f1[i] + io.f2[i]) - xn[i];
    }
    r.sync();
    if (r.tid() == 0) {
      // baseline performance of this node: cost value, dynamics defect and equality-constraint SSE
      double c0 = -P.box_offset;
      for (int i = 0; i < 12; ++i) c0 += io.aux[NA_BOXV + i];
      for (int i = 0; i < 30; ++i) c0 += 0.5 * ((x[i] - ref[RF_X + i]) * W[TW_TQ + i] + (u[i] - ref[RF_U + i]) * W[TW_TR + i]);
      const double* e = io.e6;
      c0 += 0.5 * P.mu_ee_pos * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 0.5 * P.mu_ee_ori * (e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
      double eq = io.e6[6];
      for (int ft = 0; ft < 4; ++ft) {
        if ((mode >> (3 - ft)) & 1) c0 += CONE[10 * ft + 7];
        else eq += u[3 * ft] * u[3 * ft] + u[3 * ft + 1] * u[3 * ft + 1] + u[3 * ft + 2] * u[3 * ft + 2];
      }
      double dyn = 0.0;
      for (int i = 0; i < 30; ++i) dyn += W[TW_b + i] * W[TW_b + i];
      perf[PF_COST] = dt * c0;
      perf[PF_DYN] = dt * dyn;
      perf[PF_EQ] = dt * eq;
    }
    // clear the projection matrices (in the fused layout PX aliases FR1/FR2, which are dead from here on)
    QM_PFOR(r, idx, 900) W[TW_PX + idx] = 0.0;
    QM_PFOR(r, idx, 540) W[TW_PU + idx] = 0.0;
    QM_PFOR(r, i, 30) W[TW_PE + i] = (i < 12 && !((mode >> (3 - i / 3)) & 1)) ? -u[i] : 0.0;
  }
  g.sync();
  const int nut = WI[TI_NUT];
  QM_PFOR(g, idx, nv * 49) {
    const int p = idx / 49, c = idx % 49;
    const int ip = 12 + WI[TI_PIVCOL + p];
    if (c >= 18 && c < 48) W[TW_PX + 30 * ip + c - 18] = -T[idx];
    else if (c == 48) W[TW_PE + ip] = -T[idx];
  }
  QM_PFOR(g, idx, nv * QM_NUT) {
    const int p = idx / QM_NUT, a = idx % QM_NUT;
    if (a < nut) {
      const int fc = WI[TI_FCOLS + a];
      if (fc >= 12) W[TW_PU + QM_NUT * (12 + WI[TI_PIVCOL + p]) + a] = -T[49 * p + fc - 12];
    }
  }
  g.sync();
  QM_PFOR(g, a, nut) W[TW_PU + QM_NUT * WI[TI_FCOLS + a] + a] = 1.0;
  // r' = r + R Pe ;  b~ = b + B Pe
  QM_PFOR(g, i, 60) {
    double acc = 0.0;
    if (i < 30) { for (int j = 0; j < 30; ++j) acc += W[TW_R + 30 * i + j] * W[TW_PE + j]; W[TW_TR + i] = W[TW_RV + i] + acc; }
    else { const int ii = i - 30; for (int j = 0; j < 30; ++j) acc += W[TW_B + 30 * ii + j] * W[TW_PE + j]; sb[SB_b + ii] = W[TW_b + ii] + acc; }
  }
  g.sync();
  // ---- change of input variables  du = Pu dut + Px dx + Pe   ([upstream] changeOfInputVariables), sparse in the pivot rows
  QM_PFOR(g, idx, 900) {
    const int i = idx / 30, j = idx % 30;
    double a = W[TW_A + idx], rp = 0.0;
    for (int p = 0; p < nv; ++p) {
      const int ip = 12 + WI[TI_PIVCOL + p];
      const double px = W[TW_PX + 30 * ip + j];
      a += W[TW_B + 30 * i + ip] * px;
      rp += W[TW_R + 30 * i + ip] * px;
    }
    sb[SB_A + idx] = a;
    W[TW_RPX + idx] = rp;
    pb[PB_PX + idx] = W[TW_PX + idx];
  }
  QM_PFOR(g, idx, 30 * QM_NUT) {
    const int i = idx / QM_NUT, a = idx % QM_NUT;
    double bv = 0.0, rv = 0.0;
    if (a < nut) {
      const int fc = WI[TI_FCOLS + a];
      bv = W[TW_B + 30 * i + fc];
      rv = W[TW_R + 30 * i + fc];
      if (fc >= 12)
        for (int p = 0; p < nv; ++p) {
          const int ip = 12 + WI[TI_PIVCOL + p];
          const double pu = W[TW_PU + QM_NUT * ip + a];
          bv += W[TW_B + 30 * i + ip] * pu;
          rv += W[TW_R + 30 * i + ip] * pu;
        }
    }
    sb[SB_B + idx] = bv;
    W[TW_RPU + idx] = rv;
    pb[PB_PU + idx] = W[TW_PU + idx];
  }
  QM_PFOR(g, i, 30) pb[PB_PE + i] = W[TW_PE + i];
  g.sync();
  QM_PFOR(g, idx, 900) {
    const int i = idx / 30, j = idx % 30;
    double qv = W[TW_Q + idx];
    for (int p = 0; p < nv; ++p) {
      const int ip = 12 + WI[TI_PIVCOL + p];
      qv += W[TW_PX + 30 * ip + i] * W[TW_RPX + 30 * ip + j];
    }
    sb[SB_Q + idx] = qv;
  }
  QM_PFOR(g, idx, QM_NUT * 30) {
    const int a = idx / 30, j = idx % 30;
    double pv = 0.0;
    if (a < nut) {
      const int fc = WI[TI_FCOLS + a];
      pv = W[TW_RPX + 30 * fc + j];
      if (fc >= 12)
        for (int p = 0; p < nv; ++p) { const int ip = 12 + WI[TI_PIVCOL + p]; pv += W[TW_PU + QM_NUT * ip + a] * W[TW_RPX + 30 * ip + j]; }
    }
    sb[SB_P + idx] = pv;
  }
  QM_PFOR(g, idx, QM_NUT * QM_NUT) {
    const int a = idx / QM_NUT, b = idx % QM_NUT;
    double rv = 0.0;
    if (a < nut && b < nut) {
      const int fc = WI[TI_FCOLS + a];
      rv = W[TW_RPU + QM_NUT * fc + b];
      if (fc >= 12)
        for (int p = 0; p < nv; ++p) { const int ip = 12 + WI[TI_PIVCOL + p]; rv += W[TW_PU + QM_NUT * ip + a] * W[TW_RPU + QM_NUT * ip + b]; }
    }
    sb[SB_R + idx] = rv;
  }
  QM_PFOR(g, j, 30) {
    double qv = W[TW_QV + j];
    for (int p = 0; p < nv; ++p) { const int ip = 12 + WI[TI_PIVCOL + p]; qv += W[TW_PX + 30 * ip + j] * W[TW_TR + ip]; }
    sb[SB_q + j] = qv;
  }
  QM_PFOR(g, a, QM_NUT) {
    double rv = 0.0;
    if (a < nut) {
      const int fc = WI[TI_FCOLS + a];
      rv = W[TW_TR + fc];
      if (fc >= 12)
        for (int p = 0; p < nv; ++p) { const int ip = 12 + WI[TI_PIVCOL + p]; rv += W[TW_PU + QM_NUT * ip + a] * W[TW_TR + ip]; }
    }
    sb[SB_r + a] = rv;
  }
  if (g.tid() == 0) {
    sb[SB_NUT] = (double)nut;
    if (WI[TI_STATUS]) status_or(status_out, WI[TI_STATUS]);
  }
  g.sync();
}

// Fused form (host port / tests): both kinematics evaluations and the LQ assembly on one workspace.
template <class G>
QM_HDN void transcribe_node(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                            const double* zvel, const double* tt, const double* ts, int kt, const double* x, const double* u,
                            const double* xn, double* W, int* WI, double* sb, double* pb, double* perf, int* status_out) {
  NodeIO io;
  io.fr1 = W + TW_FR1; io.fr2 = W + TW_FR2; io.f1 = W + TW_F1; io.f2 = W + TW_F2; io.x2 = W + TW_X2;
  io.T = W + TW_T; io.je = W + TW_JE; io.e6 = W + TW_E6; io.aux = W + TW_AUX;
  node_eval1(g, M, P, t, dt, mode, zvel, tt, ts, kt, x, u, W + TW_KIN, W + TW_REF, io);
  node_eval2(g, M, P, u, W + TW_KIN, io);
  node_lq(g, M, P, t, dt, mode, tt, ts, kt, x, u, xn, W, WI, io, sb, pb, perf, status_out);
}

// Pre-event node: identity jump map, no input, no cost ([upstream] setupEventNode).
template <class G>
QM_HDN void event_node(G g, const double* x, const double* xn, double* sb, double* pb, double* perf) {
  QM_PFOR(g, idx, 900) { sb[SB_A + idx] = (idx / 30 == idx % 30) ? 1.0 : 0.0; sb[SB_Q + idx] = 0.0; pb[PB_PX + idx] = 0.0; }
  QM_PFOR(g, i, 30) { sb[SB_b + i] = x[i] - xn[i]; sb[SB_q + i] = 0.0; pb[PB_PE + i] = 0.0; }
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
  kin_eval(g, M, x, (const double*)nullptr, false, kw, deriv);      // Jacobians only when the Gauss-Newton terms are wanted
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

// ------------------------------------------------------------------------------------------ line-search node evaluation
// Value-only evaluation of one intermediate node ([upstream] computeIntermediatePerformance): needs W of size PW_SIZE.
enum { PW_KIN = 0, PW_REF = KW_VSIZE, PW_F1 = PW_REF + RF_SIZE, PW_F2 = PW_F1 + 30, PW_X2 = PW_F2 + 30, PW_E6 = PW_X2 + 30,
       PW_DQ = PW_E6 + 8, PW_DX = PW_DQ + 10, PW_DU = PW_DX + 30, PW_TQ = PW_DU + 30, PW_TR = PW_TQ + 30, PW_SCAL = PW_TR + 30,
       PW_SIZE = PW_SCAL + 4 };
template <class G>
QM_HDN void perf_node(G g, const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                      const double* zvel, const double* tt, const double* ts, int kt, const double* x, const double* u,
                      const double* xn, double* W, double* perf) {
  double* kw = W + PW_KIN;
  if (g.tid() == 0) node_reference(M, P, t, mode, tt, ts, kt, W + PW_REF);
  kin_eval(g, M, x, u, false, kw, false);
  flow_rows(g, M, P.gravity, kw, x, u, W + PW_F1, (double*)nullptr);
  ee_terms(g, kw, W + PW_REF, W + PW_E6, W + PW_DQ, (double*)nullptr);
  QM_PFOR(g, i, 30) {
    W[PW_DX + i] = x[i] - W[PW_REF + RF_X + i];
    W[PW_DU + i] = u[i] - W[PW_REF + RF_U + i];
    W[PW_X2 + i] = x[i] + dt * W[PW_F1 + i];
  }
  g.sync();
  QM_PFOR(g, i, 60) {
    double acc = 0.0;
    if (i < 30) { for (int j = 0; j < 30; ++j) acc += P.Q[30 * i + j] * W[PW_DX + j]; W[PW_TQ + i] = acc; }
    else { const int ii = i - 30; for (int j = 0; j < 30; ++j) acc += P.R[30 * ii + j] * W[PW_DU + j]; W[PW_TR + ii] = acc; }
  }
  g.sync();
  if (g.tid() == 0) {
    double c0 = barrier_cost(P, mode, x, u);
    for (int i = 0; i < 30; ++i) c0 += 0.5 * (W[PW_DX + i] * W[PW_TQ + i] + W[PW_DU + i] * W[PW_TR + i]);
    const double* e = W + PW_E6;
    c0 += 0.5 * P.mu_ee_pos * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 0.5 * P.mu_ee_ori * (e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
    double eq = 0.0;
    for (int ft = 0; ft < 4; ++ft) {
      const double* v = kw + KW_FVEL + 3 * ft;
      if ((mode >> (3 - ft)) & 1) eq += v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      else {
        const double d = v[2] - zvel[ft];
        eq += d * d + u[3 * ft] * u[3 * ft] + u[3 * ft + 1] * u[3 * ft + 1] + u[3 * ft + 2] * u[3 * ft + 2];
      }
    }
    W[PW_SCAL + 0] = c0; W[PW_SCAL + 1] = eq;
  }
  g.sync();
  kin_eval(g, M, W + PW_X2, u, false, kw, false);
  flow_rows(g, M, P.gravity, kw, W + PW_X2, u, W + PW_F2, (double*)nullptr);
  if (g.tid() == 0) {
    double dyn = 0.0;
    for (int i = 0; i < 30; ++i) {
      const double d = x[i] + 0.5 * dt * (W[PW_F1 + i] + W[PW_F2 + i]) - xn[i];
      dyn += d * d;
    }
    perf[PF_COST] = dt * W[PW_SCAL + 0];
    perf[PF_DYN] = dt * dyn;
    perf[PF_EQ] = dt * W[PW_SCAL + 1];
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------ Riccati
// Workspace. K aliases SB (SB is dead once G, H are formed); LI holds L^-1 of the Cholesky factor.
enum { RW_S = 0, RW_SA = 900, RW_SB = 1800, RW_K = RW_SB, RW_H = RW_SB + 540, RW_G = RW_H + 540, RW_LI = RW_G + 324,
       RW_sv = RW_LI + 324, RW_sb = RW_sv + 30, RW_gv = RW_sb + 30, RW_kf = RW_gv + 18, RW_SIZE = RW_kf + 18 };

// 3x3 register tile of C = X Y (X: m x kd row-major ldx, Y: kd x n row-major ldy), accumulated into acc[9]
#define QM_TILE3(acc, XP, ldx, YP, ldy, i0, j0, kd)                                                        \
  do {                                                                                                     \
    for (int k_ = 0; k_ < (kd); ++k_) {                                                                    \
      const double x0_ = (XP)[((i0) + 0) * (ldx) + k_], x1_ = (XP)[((i0) + 1) * (ldx) + k_], x2_ = (XP)[((i0) + 2) * (ldx) + k_]; \
      const double y0_ = (YP)[k_ * (ldy) + (j0)], y1_ = (YP)[k_ * (ldy) + (j0) + 1], y2_ = (YP)[k_ * (ldy) + (j0) + 2];           \
      acc[0] += x0_ * y0_; acc[1] += x0_ * y1_; acc[2] += x0_ * y2_;                                       \
      acc[3] += x1_ * y0_; acc[4] += x1_ * y1_; acc[5] += x1_ * y2_;                                       \
      acc[6] += x2_ * y0_; acc[7] += x2_ * y1_; acc[8] += x2_ * y2_;                                       \
    }                                                                                                      \
  } while (0)
// same with X used transposed: C = X' Y (X: kd x m row-major ldx)
#define QM_TILE3_T(acc, XP, ldx, YP, ldy, i0, j0, kd)                                                      \
  do {                                                                                                     \
    for (int k_ = 0; k_ < (kd); ++k_) {                                                                    \
      const double x0_ = (XP)[k_ * (ldx) + (i0)], x1_ = (XP)[k_ * (ldx) + (i0) + 1], x2_ = (XP)[k_ * (ldx) + (i0) + 2]; \
      const double y0_ = (YP)[k_ * (ldy) + (j0)], y1_ = (YP)[k_ * (ldy) + (j0) + 1], y2_ = (YP)[k_ * (ldy) + (j0) + 2]; \
      acc[0] += x0_ * y0_; acc[1] += x0_ * y1_; acc[2] += x0_ * y2_;                                       \
      acc[3] += x1_ * y0_; acc[4] += x1_ * y1_; acc[5] += x1_ * y2_;                                       \
      acc[6] += x2_ * y0_; acc[7] += x2_ * y1_; acc[8] += x2_ * y2_;                                       \
    }                                                                                                      \
  } while (0)

#if defined(__CUDACC__)
// In-place inverse of the symmetric positive definite n x n matrix Gm (shared memory, leading dimension QM_NUT, n <= 18)
// by one warp: Gauss-Jordan elimination without pivoting, column j of the matrix held in the registers of lane j, the
// multipliers of each pivot step broadcast with warp shuffles (no shared-memory round trips on the dependency chain).
__device__ __forceinline__ void spd_inverse_warp(double* Gm, int n, int* status) {
  const int lane = threadIdx.x & 31;
  const bool active = lane < n;
  double r[QM_NUT];
#pragma unroll
  for (int i = 0; i < QM_NUT; ++i) r[i] = (active && i < n) ? Gm[QM_NUT * i + lane] : ((i == lane) ? 1.0 : 0.0);
  bool bad = false;
#pragma unroll
  for (int c = 0; c < QM_NUT; ++c) {
    if (c < n) {                                   // uniform across the warp
      const double acc = __shfl_sync(0xffffffffu, r[c], c);      // pivot a_cc lives in lane c, register c
      if (!(acc > 0.0)) bad = true;
      const double p = 1.0 / acc;
      const double rc = r[c];                      // a_cj of this lane's column
#pragma unroll
      for (int i = 0; i < QM_NUT; ++i) {
        if (i != c) {
          const double m = __shfl_sync(0xffffffffu, r[i], c) * p;   // multiplier a_ic / a_cc from lane c
          r[i] = (lane == c) ? -m : r[i] - m * rc;
        }
      }
      r[c] = (lane == c) ? p : rc * p;
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < QM_NUT; ++i) if (i < n) Gm[QM_NUT * i + lane] = r[i];
  }
  if (bad && lane == 0) status_or(status, ST_CHOL);
}
#endif

// One backward stage, in two halves so the caller can re-use the stage buffer (prefetch the next block) in between.
// All dense products use 3x3 register tiles (6 loads per 9 FMAs); padded input columns (a >= nut) are zero in the block.
//   riccati_stage_a: SA = S A, SB = S B, sb;  G, H, g;  then concurrently
//                    narrow: Cholesky G = L L', L^-1, G^-1 = L^-T L^-1      (dependency chain, one warp)
//                    rest  : S <- Q + A' SA (symmetric tiles), s <- q + A' sb (largest product, independent of the gains)
//   riccati_stage_b: K = -G^-1 H, kff = -G^-1 g;  S += H' K (symmetrised), s += H' kff;  gains to HBM
template <class G>
QM_HDN int riccati_stage_a(G g, const double* st, double* W, int* status) {
  const int nut = (int)st[SB_NUT];
  const double* A = st + SB_A; const double* B = st + SB_B; const double* b = st + SB_b;
  double* S = W + RW_S; double* s = W + RW_sv;
  // ---- P1: SA = S A, SB = S B, sb = s + S b
  mm<4, false>(g, 30, 30, 30, S, 30, A, 30, (const double*)nullptr, 0, 1.0, W + RW_SA, 30);
  if (nut > 0) mm<3, false>(g, 30, nut, 30, S, 30, B, QM_NUT, (const double*)nullptr, 0, 1.0, W + RW_SB, QM_NUT);
  QM_PFOR(g, i, 30) {
    double acc = s[i];
    for (int j = 0; j < 30; ++j) acc += S[30 * i + j] * b[j];
    W[RW_sb + i] = acc;
  }
  g.sync();
  // ---- P2: G = R + B' SB, H = P + B' SA, g = r + B' sb
  if (nut > 0) {
    mm<3, true>(g, nut, nut, 30, B, QM_NUT, W + RW_SB, QM_NUT, st + SB_R, QM_NUT, 1.0, W + RW_G, QM_NUT);
    mm<4, true>(g, nut, 30, 30, B, QM_NUT, W + RW_SA, 30, st + SB_P, 30, 1.0, W + RW_H, 30);
    QM_PFOR(g, a, nut) {
      double acc = st[SB_r + a];
      for (int k = 0; k < 30; ++k) acc += B[QM_NUT * k + a] * W[RW_sb + k];
      W[RW_gv + a] = acc;
    }
    g.sync();
  }
  // ---- P3 narrow: Ginv = G^-1 in place. Device: register/shuffle Gauss-Jordan on one warp; host: Cholesky route.
#if defined(__CUDA_ARCH__)
  if (g.narrow_active() && nut > 0) spd_inverse_warp(W + RW_G, nut, status);
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
  // ---- P3 rest: S <- Q + A' SA ; s <- q + A' sb   (S, s were consumed in P1; symmetrised at the end of the stage)
  if (g.rest_active()) {
    auto r_ = g.rest();
    mm<2, true>(r_, 30, 30, 30, A, 30, W + RW_SA, 30, st + SB_Q, 30, 1.0, S, 30);
    QM_PFOR(r_, i, 30) {
      double acc = st[SB_q + i];
      for (int k = 0; k < 30; ++k) acc += A[30 * k + i] * W[RW_sb + k];
      s[i] = acc;
    }
  }
  g.sync();
  return nut;
}

template <class G>
QM_HDN void riccati_stage_b(G g, int nut, double* W, double* gb) {
  double* S = W + RW_S; double* s = W + RW_sv;
  double* Gm = W + RW_G;
  if (nut > 0) {
    // ---- P4: K = -Ginv H, kff = -Ginv g
    mm<4, false>(g, nut, 30, nut, Gm, QM_NUT, W + RW_H, 30, (const double*)nullptr, 0, -1.0, W + RW_K, 30);
    QM_PFOR(g, a, nut) {
      double acc = 0.0;
      for (int k = 0; k < nut; ++k) acc += Gm[QM_NUT * a + k] * W[RW_gv + k];
      W[RW_kf + a] = -acc;
    }
    g.sync();
    // ---- P5: S += H' K, s += H' kff
    mm<4, true>(g, 30, 30, nut, W + RW_H, 30, W + RW_K, 30, S, 30, 1.0, S, 30);
    QM_PFOR(g, i, 30) {
      double acc = s[i];
      for (int a = 0; a < nut; ++a) acc += W[RW_H + 30 * a + i] * W[RW_kf + a];
      s[i] = acc;
    }
  }
  QM_PFOR(g, idx, QM_NUT * 30) gb[GB_K + idx] = (idx / 30 < nut) ? W[RW_K + idx] : 0.0;
  QM_PFOR(g, a, QM_NUT) gb[GB_KFF + a] = (a < nut) ? W[RW_kf + a] : 0.0;
  g.sync();
  QM_PFOR(g, idx, 435) {     // symmetrise: pairs i < j
    int i = 0, r = idx;
    while (r >= 29 - i) { r -= 29 - i; ++i; }
    const int j = i + 1 + r;
    const double v = 0.5 * (S[30 * i + j] + S[30 * j + i]);
    S[30 * i + j] = v; S[30 * j + i] = v;
  }
  g.sync();
}

// One forward stage: dut = K dx + kff; du = Pu dut + Px dx + Pe; dx+ = A dx + B dut + b; armijo += q.dx + r.dut
// W: [0:30] dx, [30:60] dx next, [60:78] dut, [80] armijo accumulator
template <class G>
QM_HDN void rollout_stage(G g, const double* st, const double* pb, const double* gb, double* W, double* du_out) {
  const int nut = (int)st[SB_NUT];
  double* dx = W; double* dxn = W + 30; double* dut = W + 60;
  QM_PFOR(g, a, QM_NUT) {
    double acc = 0.0;
    if (a < nut) { acc = gb[GB_KFF + a]; for (int j = 0; j < 30; ++j) acc += gb[GB_K + 30 * a + j] * dx[j]; }
    dut[a] = acc;
  }
  g.sync();
  QM_PFOR(g, i, 60) {
    if (i < 30) {
      double acc = pb[PB_PE + i];
      for (int j = 0; j < 30; ++j) acc += pb[PB_PX + 30 * i + j] * dx[j];
      for (int a = 0; a < nut; ++a) acc += pb[PB_PU + QM_NUT * i + a] * dut[a];
      du_out[i] = (nut > 0) ? acc : 0.0;
    } else {
      const int ii = i - 30;
      double acc = st[SB_b + ii];
      for (int j = 0; j < 30; ++j) acc += st[SB_A + 30 * ii + j] * dx[j];
      for (int a = 0; a < nut; ++a) acc += st[SB_B + QM_NUT * ii + a] * dut[a];
      dxn[ii] = acc;
    }
  }
  if (g.tid() == 0) {
    double acc = 0.0;
    for (int j = 0; j < 30; ++j) acc += st[SB_q + j] * dx[j];
    for (int a = 0; a < nut; ++a) acc += st[SB_r + a] * dut[a];
    W[80] += acc;
  }
  g.sync();
}

// [upstream] FilterLinesearch::acceptStep
QM_HD bool accept_step(const qmb200_solver_desc& S, double base_merit, double base_viol, double new_merit, double new_viol, double armijo) {
  if (new_viol > S.g_max) return new_viol < (1.0 - S.gamma_c) * base_viol;
  if (new_viol < S.g_min && base_viol < S.g_min && armijo < 0.0) return new_merit < base_merit + S.armijo_factor * armijo;
  return new_merit < base_merit - S.gamma_c * base_viol || new_viol < (1.0 - S.gamma_c) * base_viol;
}

}  // namespace qm
