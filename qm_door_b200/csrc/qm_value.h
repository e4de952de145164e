// Value-only evaluation of one node by ONE thread ([upstream] computeIntermediatePerformance / computeTerminalPerformance
// of the filter line search). The rigid-body tree is walked once, joint by joint, with the running chain transform in
// registers: no workspace, no synchronisation, model constants read at warp-uniform addresses. The line-search kernel
// runs a thread per node (all 32 lanes of a warp busy); the CPU port calls the same functions.
// Same formulas as the phase-structured evaluation in qm_core.h (kin_positions / centroidal_velocity / kin_velocities)
// restricted to the value level; reference computations replaced: see the header of qm_core.h.
#pragma once
#include "qm_mpc.h"

namespace qm {

enum { VK_SLOTS = 2 };     // saved transforms of joints with children other than the next joint (the floating-base link)

struct ValueKin {
  double fpos[QM_NFEET][3];
  double fvel[QM_NFEET][3];
  double eep[3], eer[9];
  double com[3];
  double vel[QM_NJ];
};

QM_HD void mat3_mul(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
QM_HD void mat3_vec(const double* A, const double* v, double* o) {
  for (int r = 0; r < 3; ++r) o[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
}

// Tree topologies the single-pass walk supports: parent[j] < j, a serial six-joint floating base carrying all other joints,
// and at most VK_SLOTS joints with a child that is not the next joint. (The phase-structured evaluation of qm_core.h has no such restriction.)
QM_HDN bool value_walk_supported(const qmb200_model_desc& M) {
  uint32_t branch = 0;
  for (int j = 0; j < QM_NJ; ++j) {
    if (M.parent[j] >= j) return false;
    if (M.parent[j] >= 0 && M.parent[j] != j - 1) branch |= 1u << M.parent[j];
    // the six base joints form a serial chain that carries everything else (composite inertias of the base columns)
    if (j < 6 && M.parent[j] != j - 1) return false;
    if (j >= 6 && !((M.pathmask[j] >> 5) & 1u)) return false;
  }
  int nb = 0;
  for (int j = 0; j < QM_NJ; ++j) nb += (branch >> j) & 1u;
  return nb <= VK_SLOTS;
}

// q = x[6:30]; u == nullptr: position level only (frames, com).
QM_HDN void kin_value_serial(const qmb200_model_desc& M, const double* x, const double* u, ValueKin& o) {
  const double* q = x + 6;
  uint32_t branch = 0;
  for (int j = 0; j < QM_NJ; ++j)
    if (M.parent[j] >= 0 && M.parent[j] != j - 1) branch |= 1u << M.parent[j];
  double Rc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pc[3] = {0, 0, 0}, Vc[6] = {0, 0, 0, 0, 0, 0};   // transform / joint-only velocity of joint j-1
  double Rs[VK_SLOTS][9], ps[VK_SLOTS][3], Vs[VK_SLOTS][6];
  int slot_joint[VK_SLOTS];
  for (int s = 0; s < VK_SLOTS; ++s) slot_joint[s] = -1;
  int nslot = 0;
  double total[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // composite inertia of the whole tree about the world origin
  double base_body[5][10];                              // bodies of the first five base joints (their subtrees exclude earlier ones)
  double base_ax[6][3], base_p[6][3];
  double hj[6] = {0, 0, 0, 0, 0, 0};                     // momentum (L0, p) due to the actuated joint velocities
  double fV[QM_NFEET][6];
  double sz = 0, cz = 1, sy = 0, cy = 1, sx = 0, cx = 1;
  if (M.root6_standard) { sincos(q[3], &sz, &cz); sincos(q[4], &sy, &cy); sincos(q[5], &sx, &cx); }
  for (int j = 0; j < QM_NJ; ++j) {
    double Rj[9], pj[3], aw[3], Vj[6];
    if (M.root6_standard && j < 6) {
      pj[0] = q[0]; pj[1] = (j >= 1) ? q[1] : 0.0; pj[2] = (j >= 2) ? q[2] : 0.0;
      if (j <= 2) {
        for (int k = 0; k < 9; ++k) Rj[k] = (k % 4 == 0) ? 1.0 : 0.0;
        aw[0] = (j == 0); aw[1] = (j == 1); aw[2] = (j == 2);
      } else if (j == 3) {
        Rj[0] = cz; Rj[1] = -sz; Rj[2] = 0; Rj[3] = sz; Rj[4] = cz; Rj[5] = 0; Rj[6] = 0; Rj[7] = 0; Rj[8] = 1;
        aw[0] = 0; aw[1] = 0; aw[2] = 1;
      } else if (j == 4) {
        Rj[0] = cz * cy; Rj[1] = -sz; Rj[2] = cz * sy; Rj[3] = sz * cy; Rj[4] = cz; Rj[5] = sz * sy; Rj[6] = -sy; Rj[7] = 0; Rj[8] = cy;
        aw[0] = -sz; aw[1] = cz; aw[2] = 0;
      } else {
        Rj[0] = cz * cy; Rj[1] = cz * sy * sx - sz * cx; Rj[2] = cz * sy * cx + sz * sx;
        Rj[3] = sz * cy; Rj[4] = sz * sy * sx + cz * cx; Rj[5] = sz * sy * cx - cz * sx;
        Rj[6] = -sy;     Rj[7] = cy * sx;                Rj[8] = cy * cx;
        aw[0] = cz * cy; aw[1] = sz * cy; aw[2] = -sy;
      }
      for (int k = 0; k < 6; ++k) Vj[k] = 0.0;
    } else {
      // parent transform: the previous joint, a saved branch joint, or the world
      double Rpar[9], ppar[3], Vpar[6];
      const int par = M.parent[j];
      if (par == j - 1 && par >= 0) {
        for (int k = 0; k < 9; ++k) Rpar[k] = Rc[k];
        for (int k = 0; k < 3; ++k) ppar[k] = pc[k];
        for (int k = 0; k < 6; ++k) Vpar[k] = Vc[k];
      } else {
        for (int k = 0; k < 9; ++k) Rpar[k] = (k % 4 == 0) ? 1.0 : 0.0;
        for (int k = 0; k < 3; ++k) ppar[k] = 0.0;
        for (int k = 0; k < 6; ++k) Vpar[k] = 0.0;
        for (int s = 0; s < VK_SLOTS; ++s)
          if (par >= 0 && slot_joint[s] == par) {
            for (int k = 0; k < 9; ++k) Rpar[k] = Rs[s][k];
            for (int k = 0; k < 3; ++k) ppar[k] = ps[s][k];
            for (int k = 0; k < 6; ++k) Vpar[k] = Vs[s][k];
          }
      }
      double R0[9], p0[3];
      mat3_mul(Rpar, M.Rp[j], R0);
      mat3_vec(Rpar, M.pp[j], p0);
      for (int r = 0; r < 3; ++r) p0[r] += ppar[r];
      const double ax = M.axis[j][0], ay = M.axis[j][1], az = M.axis[j][2];
      for (int r = 0; r < 3; ++r) aw[r] = R0[3 * r] * ax + R0[3 * r + 1] * ay + R0[3 * r + 2] * az;
      if (M.jtype[j] == 1) {
        double s, c;
        sincos(q[j], &s, &c);
        const double v = 1.0 - c;
        const double Rq[9] = {c + v * ax * ax,      v * ax * ay - s * az, v * ax * az + s * ay,
                              v * ax * ay + s * az, c + v * ay * ay,      v * ay * az - s * ax,
                              v * ax * az - s * ay, v * ay * az + s * ax, c + v * az * az};
        mat3_mul(R0, Rq, Rj);
        for (int r = 0; r < 3; ++r) pj[r] = p0[r];
      } else {
        for (int k = 0; k < 9; ++k) Rj[k] = R0[k];
        for (int r = 0; r < 3; ++r) pj[r] = p0[r] + aw[r] * q[j];
      }
      // velocity of this body due to the actuated joints alone (base at rest): V_parent + S_j v_j
      for (int k = 0; k < 6; ++k) Vj[k] = Vpar[k];
      if (u != nullptr && j >= 6) {
        const double vj = u[12 + j - 6];
        if (M.jtype[j] == 1) {
          double t[3];
          cross3(pj, aw, t);
          for (int k = 0; k < 3; ++k) { Vj[k] += aw[k] * vj; Vj[3 + k] += t[k] * vj; }
        } else {
          for (int k = 0; k < 3; ++k) Vj[3 + k] += aw[k] * vj;
        }
      }
    }
    if (j < 6) for (int k = 0; k < 3; ++k) { base_ax[j][k] = aw[k]; base_p[j][k] = pj[k]; }
    // body inertial quantities about the world origin: m, m c, I0
    {
      const double m = M.mass[j];
      double c[3], RI[9], Iw[9];
      mat3_vec(Rj, M.com[j], c);
      for (int r = 0; r < 3; ++r) c[r] += pj[r];
      mat3_mul(Rj, M.inertia[j], RI);
      for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) Iw[3 * r + cc] = RI[3 * r] * Rj[3 * cc] + RI[3 * r + 1] * Rj[3 * cc + 1] + RI[3 * r + 2] * Rj[3 * cc + 2];
      const double cc2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      double b[10];
      b[0] = m; b[1] = m * c[0]; b[2] = m * c[1]; b[3] = m * c[2];
      b[4] = Iw[0] + m * (cc2 - c[0] * c[0]);
      b[5] = Iw[1] - m * c[0] * c[1];
      b[6] = Iw[2] - m * c[0] * c[2];
      b[7] = Iw[4] + m * (cc2 - c[1] * c[1]);
      b[8] = Iw[5] - m * c[1] * c[2];
      b[9] = Iw[8] + m * (cc2 - c[2] * c[2]);
      for (int k = 0; k < 10; ++k) total[k] += b[k];
      if (j < 5) for (int k = 0; k < 10; ++k) base_body[j][k] = b[k];
      if (u != nullptr && j >= 6) {
        double h[6];
        inertia_mul(b, Vj, h);
        for (int k = 0; k < 6; ++k) hj[k] += h[k];
      }
    }
    // frames attached to this joint's body
    for (int f = 0; f < QM_NFEET; ++f)
      if (M.foot_joint[f] == j) {
        double t[3];
        mat3_vec(Rj, M.foot_off[f], t);
        for (int r = 0; r < 3; ++r) o.fpos[f][r] = pj[r] + t[r];
        for (int k = 0; k < 6; ++k) fV[f][k] = Vj[k];
      }
    if (M.ee_joint == j) {
      double t[3];
      mat3_vec(Rj, M.ee_off, t);
      for (int r = 0; r < 3; ++r) o.eep[r] = pj[r] + t[r];
      mat3_mul(Rj, M.ee_Roff, o.eer);
    }
    if ((branch >> j) & 1u) {
      if (nslot < VK_SLOTS) {
        for (int k = 0; k < 9; ++k) Rs[nslot][k] = Rj[k];
        for (int k = 0; k < 3; ++k) ps[nslot][k] = pj[k];
        for (int k = 0; k < 6; ++k) Vs[nslot][k] = Vj[k];
        slot_joint[nslot] = j;
      }
      ++nslot;
    }
    for (int k = 0; k < 9; ++k) Rc[k] = Rj[k];
    for (int k = 0; k < 3; ++k) pc[k] = pj[k];
    for (int k = 0; k < 6; ++k) Vc[k] = Vj[k];
  }
  const double itot = 1.0 / total[0];
  for (int r = 0; r < 3; ++r) o.com[r] = total[1 + r] * itot;
  if (u == nullptr) return;
  // floating-base block of the centroidal momentum matrix: columns k < 6 = I^c_k S_k, I^c_k = whole tree minus bodies < k
  double Ab[6][6];
  {
    double comp[10];
    for (int k = 0; k < 10; ++k) comp[k] = total[k];
    for (int k = 0; k < 6; ++k) {
      if (k > 0) for (int i = 0; i < 10; ++i) comp[i] -= base_body[k - 1][i];
      double S[6], h[6], t[3];
      if (M.jtype[k] == 1) {
        S[0] = base_ax[k][0]; S[1] = base_ax[k][1]; S[2] = base_ax[k][2];
        cross3(base_p[k], base_ax[k], S + 3);
      } else {
        S[0] = S[1] = S[2] = 0.0;
        S[3] = base_ax[k][0]; S[4] = base_ax[k][1]; S[5] = base_ax[k][2];
      }
      inertia_mul(comp, S, h);
      cross3(o.com, h + 3, t);
      Ab[0][k] = h[3]; Ab[1][k] = h[4]; Ab[2][k] = h[5];
      Ab[3][k] = h[0] - t[0]; Ab[4][k] = h[1] - t[1]; Ab[5][k] = h[2] - t[2];
    }
  }
  // v_b = Ab^-1 (m h_normalized - A_j v_j)   ([upstream] getPinocchioJointVelocity with the block inverse of Ab)
  {
    double t[3], rhs[6];
    cross3(o.com, hj + 3, t);
    for (int r = 0; r < 3; ++r) {
      rhs[r] = M.total_mass * x[r] - hj[3 + r];
      rhs[3 + r] = M.total_mass * x[3 + r] - (hj[r] - t[r]);
    }
    const double mass = Ab[0][0];
    const double a = Ab[3][3], b = Ab[3][4], c = Ab[3][5];
    const double d = Ab[4][3], e = Ab[4][4], f = Ab[4][5];
    const double gg = Ab[5][3], hh = Ab[5][4], ii = Ab[5][5];
    const double det = a * (e * ii - f * hh) - b * (d * ii - f * gg) + c * (d * hh - e * gg);
    const double id = 1.0 / det;
    const double inv[9] = {(e * ii - f * hh) * id, (c * hh - b * ii) * id, (b * f - c * e) * id,
                           (f * gg - d * ii) * id, (a * ii - c * gg) * id, (c * d - a * f) * id,
                           (d * hh - e * gg) * id, (b * gg - a * hh) * id, (a * e - b * d) * id};
    double wv[3];
    for (int r = 0; r < 3; ++r) wv[r] = inv[3 * r] * rhs[3] + inv[3 * r + 1] * rhs[4] + inv[3 * r + 2] * rhs[5];
    const double im = 1.0 / mass;
    for (int r = 0; r < 3; ++r) {
      // Bi[r][0:3] = I / mass, Bi[r][3:6] = -(Ab12 inv)[r] / mass
      double acc = rhs[r] * im;
      for (int cc = 0; cc < 3; ++cc) {
        double ai = 0.0;
        for (int k = 0; k < 3; ++k) ai += Ab[r][3 + k] * inv[3 * k + cc];
        acc += -ai * im * rhs[3 + cc];
      }
      o.vel[r] = acc;
      o.vel[3 + r] = wv[r];
    }
  }
  for (int l = 0; l < 18; ++l) o.vel[6 + l] = u[12 + l];
  // foot velocities: (V_base + V_joints) at the foot position
  double Vb[6] = {0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 6; ++k) {
    const double vk = o.vel[k];
    if (M.jtype[k] == 1) {
      double t[3];
      cross3(base_p[k], base_ax[k], t);
      for (int r = 0; r < 3; ++r) { Vb[r] += base_ax[k][r] * vk; Vb[3 + r] += t[r] * vk; }
    } else {
      for (int r = 0; r < 3; ++r) Vb[3 + r] += base_ax[k][r] * vk;
    }
  }
  for (int f = 0; f < QM_NFEET; ++f) {
    double V[6], t[3];
    for (int k = 0; k < 6; ++k) V[k] = Vb[k] + fV[f][k];
    cross3(V, o.fpos[f], t);
    for (int r = 0; r < 3; ++r) o.fvel[f][r] = V[3 + r] + t[r];
  }
}

// flow map value f(x,u) from the value-level kinematics (QMDynamicsAD::computeFlowMap)
QM_HDN void flow_value_serial(const qmb200_model_desc& M, double gravity, const ValueKin& k, const double* u, double* f) {
  const double im = 1.0 / M.total_mass;
  for (int i = 0; i < 3; ++i) f[i] = (u[i] + u[3 + i] + u[6 + i] + u[9 + i]) * im - (i == 2 ? gravity : 0.0);
  double acc[3] = {0, 0, 0};
  for (int ft = 0; ft < 4; ++ft) {
    double arm[3], t[3];
    for (int r = 0; r < 3; ++r) arm[r] = k.fpos[ft][r] - k.com[r];
    cross3(arm, u + 3 * ft, t);
    for (int r = 0; r < 3; ++r) acc[r] += t[r];
  }
  for (int i = 0; i < 3; ++i) f[3 + i] = acc[i] * im;
  for (int i = 0; i < QM_NJ; ++i) f[6 + i] = k.vel[i];
}

// end-effector error e[6] = [p - p_ref ; quaternionDistance(q, q_ref)]
QM_HD void ee_error_serial(const ValueKin& k, const double* ref, double* e) {
  double q[4], cr[3];
  quat_from_matrix(k.eer, q);
  const double* qr = ref + RF_EEQ;
  cross3(q, qr, cr);
  for (int r = 0; r < 3; ++r) {
    e[r] = k.eep[r] - ref[RF_EEP + r];
    e[3 + r] = q[3] * qr[r] - qr[3] * q[r] + cr[r];
  }
}

// Intermediate node: cost, dynamics defect (Heun) and equality-constraint violation at (x, u), next state xn.
QM_HDN void perf_node_serial(const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, double dt, int mode,
                             const double* zvel, const double* tt, const double* ts, int kt, const double* x, const double* u,
                             const double* xn, double* perf) {
  double ref[RF_SIZE];
  node_reference(M, P, t, mode, tt, ts, kt, ref);
  ValueKin k;
  kin_value_serial(M, x, u, k);
  double f1[30], x2[30], e[6];
  flow_value_serial(M, P.gravity, k, u, f1);
  ee_error_serial(k, ref, e);
  // tracking cost 1/2 dx'Q dx + 1/2 du'R du (zero weights skipped: the weights are warp-uniform)
  double c0 = barrier_cost(P, mode, x, u);
  {
    double dx[30], du[30];
    for (int i = 0; i < 30; ++i) { dx[i] = x[i] - ref[RF_X + i]; du[i] = u[i] - ref[RF_U + i]; }
    double acc = 0.0;
    for (int i = 0; i < 30; ++i) {
      double rq = 0.0, rr = 0.0;
      for (int j = 0; j < 30; ++j) {
        const double wq = P.Q[30 * i + j], wr = P.R[30 * i + j];
        if (wq != 0.0) rq += wq * dx[j];
        if (wr != 0.0) rr += wr * du[j];
      }
      acc += dx[i] * rq + du[i] * rr;
    }
    c0 += 0.5 * acc;
  }
  c0 += 0.5 * P.mu_ee_pos * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 0.5 * P.mu_ee_ori * (e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
  double eq = 0.0;
  for (int ft = 0; ft < 4; ++ft) {
    const double* v = k.fvel[ft];
    if ((mode >> (3 - ft)) & 1) eq += v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    else {
      const double d = v[2] - zvel[ft];
      eq += d * d + u[3 * ft] * u[3 * ft] + u[3 * ft + 1] * u[3 * ft + 1] + u[3 * ft + 2] * u[3 * ft + 2];
    }
  }
  for (int i = 0; i < 30; ++i) x2[i] = x[i] + dt * f1[i];
  kin_value_serial(M, x2, u, k);
  double f2[30];
  flow_value_serial(M, P.gravity, k, u, f2);
  double dyn = 0.0;
  for (int i = 0; i < 30; ++i) {
    const double d = x[i] + 0.5 * dt * (f1[i] + f2[i]) - xn[i];
    dyn += d * d;
  }
  perf[PF_COST] = dt * c0;
  perf[PF_DYN] = dt * dyn;
  perf[PF_EQ] = dt * eq;
}

// Terminal node: "finalEndEffector" soft constraint only (QMInterface.cpp:104).
QM_HDN void perf_terminal_serial(const qmb200_model_desc& M, const qmb200_problem_desc& P, double t, int mode, const double* tt,
                                 const double* ts, int kt, const double* x, double* perf) {
  double ref[RF_SIZE], e[6];
  node_reference(M, P, t, mode, tt, ts, kt, ref);
  ValueKin k;
  kin_value_serial(M, x, (const double*)nullptr, k);
  ee_error_serial(k, ref, e);
  perf[PF_COST] = 0.5 * P.mu_fee_pos * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + 0.5 * P.mu_fee_ori * (e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
  perf[PF_DYN] = 0.0;
  perf[PF_EQ] = 0.0;
}

}  // namespace qm
