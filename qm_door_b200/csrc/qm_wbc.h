// WBC half of the hot path: rigid-body dynamics of the measured / desired configurations, the task stack and the
// hierarchical QP, in the same phase-structured style as qm_core.h (one thread group per solve).
// Reference computations replaced (relative to /root/reference):
//   wbc_dynamics      WbcBase::updateMeasured / updateDesired          qm_wbc/src/WbcBase.cpp:146-238   (Pinocchio crba, nle, J, dJ, dccrba)
//   wbc_tasks         the formulate* task builders + Task operators    qm_wbc/src/WbcBase.cpp:240-578, qm_wbc/include/qm_wbc/Task.h:17-66
//   wbc_solve         HoQp (3 nested levels) + qpOASES                 qm_wbc/src/HoQp.cpp:12-158, HierarchicalWbc.cpp:18-44, HierarchicalMpcWbc.cpp:18-34
//   torque recovery   WbcBase::updateCmd                               qm_wbc/src/WbcBase.cpp:580-595
// Every level's QP is strictly convex (HoQp.cpp:66,72-75), so its solution is unique; it is computed here by
//   level 0 : Newton iteration on the slack-eliminated piecewise-quadratic  1/2|A z - b|^2 + eps/2 |z|^2 + 1/2 |(D z - f)+|^2
//             with Householder least squares (stable for the 1e-12 regularisation of HoQp.cpp:66)
//   level 1,2: dual active-set method (Goldfarb & Idnani) in the null-space coordinates of the levels above.
#pragma once
#include "qm_core.h"

namespace qm {

enum { WST_OK = 0, WST_QP_MAX_ITER = 1, WST_DEGENERATE = 2, WST_NAN = 4, WST_BAD_MODE = 8 };

enum { WB_NX = 36, WB_ND0 = 56, WB_POOL = 36, WB_MAXLEV = 6, WB_MAXW = 36, WB_QR_ROWS = 92, WB_QR_LD = 37 };
#define QM_WBC_EPS 1e-12        // HoQp.cpp:66

// ---- workspace (doubles)
// The equality rows of all levels are written once by the task builder and then only streamed (least-squares build, A Z
// products): they live in a separate "cold" block Wc -- global memory on the device (16 KB per solve, L2 traffic of a few hundred
// bytes per phase), so that four solves fit in the shared memory of an SM; on the host it is just another array.
enum {
  WC_A0 = 0,                         // [18][36] level-0 equality rows
  WC_B0 = WC_A0 + 18 * 36,           // [18]
  WC_AP = WC_B0 + 18,                // [36][36] equality rows of the levels below level 0, level after level (see WI_LV)
  WC_BP = WC_AP + WB_POOL * 36,      // [36]
  WC_SIZE = ((WC_BP + WB_POOL + 3) / 4) * 4
};
enum {
  WW_D0 = 0,                         // [56][36]
  WW_F0 = WW_D0 + 56 * 36,           // [56]
  WW_V0 = WW_F0 + 56,                // [56]  level-0 slack solution
  WW_HJ = WW_V0 + 56,                // [18]
  WW_X = WW_HJ + 18,                 // [36]
  WW_Z0 = WW_X + 36,                 // [36][18] null-space basis of the levels solved so far (updated in place, row by row)
  WW_SCR = ((WW_Z0 + 36 * 18 + 3) / 4) * 4,
  // scratch, dynamics phase
  WA_KIN = WW_SCR,
  WA_ACC = WA_KIN + KW_SIZE,         // [24][6] bias spatial acceleration of each body (qdd = 0, no gravity)
  WA_FB = WA_ACC + 144,              // [24][6] per-joint terms, then body forces (n, f)
  WA_M = WA_FB + 144,                // [24][24]
  WA_NLE = WA_M + 576,               // [24]
  WA_JF = WA_NLE + 24,               // [12][24]
  WA_DJV = WA_JF + 288,              // [12]
  WA_JBA = WA_DJV + 12,              // [3][24]
  WA_DJBV = WA_JBA + 72,             // [3] (+1)
  WA_JEE = WA_DJBV + 4,              // [6][24]
  WA_DJEE = WA_JEE + 144,            // [6] (+2)
  WA_MEAS = WA_DJEE + 8,             // q[24] v[24] fpos[12] fvel[12] eep[3] eev[6] eeR[9] = 90
  WA_DES = WA_MEAS + 92,             // q[24] v[24] bacc[6] fpos[12] fvel[12] eep[3] eev[3] eeR[9] = 93
  WA_END = WA_DES + 96,
  // scratch, solver phase (aliases the dynamics phase). Lifetimes: level 0 needs the big least-squares matrix QR and nothing of
  // the levels below it; the levels below need GA .. GG. QR is therefore laid over that whole window. The order of the blocks
  // below is the order in which the kernel sequence can leave them out of shared memory (see k_wbc_level).
  WS_GA = WW_SCR,                    // [22][18]  A Z of the current level
  WS_GB = WS_GA + 22 * 18,           // [22]
  WS_Gg = WS_GB + 22,                // [56]
  WS_J = WS_Gg + 56,                 // [18][18]  active-set factor; after the level's solve: its kernel basis N
  WS_Z = WS_J + 324,                 // [18]
  WS_D = WS_Z + 18,                  // [18]
  WS_RR = WS_D + 18,                 // [18]
  WS_ZD = WS_RR + 18,                // [18]
  WS_NP = WS_ZD + 18,                // [18]
  WS_KCN = WS_D,                     // [38] column maxima of a kernel basis (d .. np are only live inside the iteration)
  WS_STAGE_ROWS = 28,                // rows of 36 that fit in J .. LS (324 + 90 + 760 doubles): staging of A_p / D0, see wbc_solve_prepare
  WS_LS = WS_NP + 18,                // [40][19] least-squares matrix | rhs of a level below level 0; then scratch of its kernel basis
  WS_RF = WS_LS + 40 * 19,           // [18][18]  triangular factor of the active set: only the iteration uses it, so the kernel
                                     //           sequence keeps it out of k_wbc_level's shared memory (which ends here)
  WS_GG = WS_RF + 324,               // [18][56]  (D0 Z)' -- column c of row i at 56 c + i, so the lanes that own rows read
                                     //           consecutive words. Last block of the window: the kernel sequence keeps it in the
                                     //           solve's global-memory image only
  WS_LOWEND = WS_GG + 56 * 18,       // end of the blocks of the levels below level 0
  WS_QR = WW_SCR,                    // level 0 and its kernel basis only: [92][37] stacked least-squares matrix | rhs
  WS_OVEND = (WS_LOWEND > WS_QR + WB_QR_ROWS * WB_QR_LD) ? WS_LOWEND : WS_QR + WB_QR_ROWS * WB_QR_LD,
  WS_RES = WS_OVEND,                 // [56] constraint residuals
  WS_VH = WS_RES + 56,               // [92] Householder vector / violation scores
  WS_WJ = WS_VH + 92,                // [37]
  WS_HP = WS_WJ + 40,                // [4][56] partial sums of the Householder steps (4 row chunks per column)
  WS_U = WS_HP,                      // [60] multipliers of the active-set iteration   } never live during a Householder
  WS_CN = WS_HP + 64,                // [40] column maxima of the kernel basis / scalars of the iteration   } triangularisation
  WS_SCL = WS_HP + 104,              // [56] row scales of the violation test of the active-set iteration
  WS_END = WS_HP + 4 * 56,
  WW_SIZE = (WA_END > WS_END ? WA_END : WS_END)
};
static_assert(WB_QR_ROWS * WB_QR_LD <= WS_OVEND - WS_QR, "the level-0 least-squares matrix fits in the window it is laid over");
static_assert(WS_STAGE_ROWS * 36 <= WS_RF - WS_J && 22 * 36 <= WS_RF - WS_J, "staged rows fit in the blocks J .. LS");
// per-level record of a solve (wbc_update's `levels` output): level p at WBL_LEVEL * p: [number of null-space columns n_p | x after
// the level (36) | stacked Z after the level (36 x 18, n_p columns valid)]; after the WB_MAXLEV records: number of levels, then the
// level-0 slack (56).
enum { WBL_N = 0, WBL_X = 1, WBL_Z = 37, WBL_LEVEL = 37 + 36 * 18, WBL_NLEV = WB_MAXLEV * WBL_LEVEL, WBL_V0 = WBL_NLEV + 1, WBL_SIZE = WBL_V0 + 56 };
enum { WI_INW = 0, WI_PERM = 56, WI_ACT = 100, WI_IGN = 136, WI_SC = 192, WI_LV = 216, WI_SIZE = 232 };   // PERM may run into ACT (kernel basis, never during the iteration)
// WI_LV: [0] number of levels (level 0 included), then per level p >= 1: [2 p] first row in the pool, [2 p + 1] number of rows
// WI_SC scalars: [0] nW changed flag, [1] rank, [2] n1, [3] n2, [4] iq, [5] ip, [6] status, [7] r1, [8] r2, [9] nD0, [10] done

QM_HDO void rot_zyx(const double* e, double* R) {
  double sz, cz, sy, cy, sx, cx;
  sincos(e[0], &sz, &cz); sincos(e[1], &sy, &cy); sincos(e[2], &sx, &cx);
  R[0] = cz * cy; R[1] = cz * sy * sx - sz * cx; R[2] = cz * sy * cx + sz * sx;
  R[3] = sz * cy; R[4] = sz * sy * sx + cz * cx; R[5] = sz * sy * cx - cz * sx;
  R[6] = -sy;     R[7] = cy * sx;                R[8] = cy * cx;
}
// [upstream] rotationErrorInWorld(Rref, Rcur): rotation vector of Rref Rcur^T
QM_HDO void rotation_error_world(const double* Rr, const double* Rc, double* e) {
  double R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = Rr[3 * i] * Rc[3 * j] + Rr[3 * i + 1] * Rc[3 * j + 1] + Rr[3 * i + 2] * Rc[3 * j + 2];
  const double sk[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double th = acos(c);
  const double k = (th < 1e-8) ? 0.5 : th / (2.0 * sin(th));
  e[0] = k * sk[0]; e[1] = k * sk[1]; e[2] = k * sk[2];
}
// spatial motion cross product  X = A x B
QM_HD void crm(const double* A, const double* B, double* X) {
  cross3(A, B, X);
  cross3(A, B + 3, X + 3);
  cross3_add(A + 3, B, X + 3);
}

// single copies of the kinematic passes for the two configurations (measured, desired) of one solve
template <class G>
QM_HDO void wbc_kin_positions(G g, const qmb200_model_desc& M, const double* q, double* kw) { kin_positions(g, M, q, kw); }
template <class G>
QM_HDO void wbc_kin_velocities(G g, const qmb200_model_desc& M, double* kw) { kin_velocities(g, M, 0, kw); }

// Bias accelerations (qdd = 0, no gravity) and body forces of the configuration whose position/velocity level is in kw.
// FB_i = I_i (acc_i + a_g) + V_i x* (I_i V_i)   with a_g = (0; 0, 0, grav) (grav = 9.81 for RNEA, 0 for momentum rates)
template <class G>
QM_HDO void bias_forces(G g, const qmb200_model_desc& M, const double* kw, double grav, double* acc, double* fb) {
  QM_PFOR(g, k, QM_NJ) {
    double S[6], X[6];
    joint_S(M, kw, k, S);
    crm(kw + KW_V + 6 * k, S, X);
    const double vk = kw[KW_VEL + k];
    for (int c = 0; c < 6; ++c) fb[6 * k + c] = X[c] * vk;
  }
  g.sync();
  QM_PFOR(g, i, QM_NJ) {
    double a[6] = {0, 0, 0, 0, 0, 0};
    const uint32_t mask = M.pathmask[i];
    for (int k = 0; k < QM_NJ; ++k)
      if ((mask >> k) & 1u)
        for (int c = 0; c < 6; ++c) a[c] += fb[6 * k + c];
    for (int c = 0; c < 6; ++c) acc[6 * i + c] = a[c];
  }
  g.sync();
  QM_PFOR(g, i, QM_NJ) {
    double a[6], h[6];
    for (int c = 0; c < 6; ++c) a[c] = acc[6 * i + c];
    a[5] += grav;
    inertia_mul(kw + KW_BODY + 10 * i, a, h);
    const double* V = kw + KW_V + 6 * i;
    const double* hb = kw + KW_HB + 6 * i;
    cross3_add(V, hb, h);            // w x L0
    cross3_add(V + 3, hb + 3, h);    // + vO x p
    cross3_add(V, hb + 3, h + 3);    // w x p
    for (int c = 0; c < 6; ++c) fb[6 * i + c] = h[c];
  }
  g.sync();
}

// classical acceleration (qdd = 0) of a point p moving with body b: aO + wd x p + w x v_p
QM_HD void point_bias_accel(const double* acc_b, const double* V_b, const double* p, double* out) {
  double vp[3], t[3];
  cross3(V_b, p, vp);
  vp[0] += V_b[3]; vp[1] += V_b[4]; vp[2] += V_b[5];
  cross3(acc_b, p, t);
  cross3_add(V_b, vp, t);
  out[0] = acc_b[3] + t[0]; out[1] = acc_b[4] + t[1]; out[2] = acc_b[5] + t[2];
}

// updateMeasured (WbcBase.cpp:146-203): mass matrix, nonlinear effects, frame Jacobians and their bias accelerations at the
// measured state rbd[55]; q, v and the frame kinematics to W + WA_MEAS.
template <class G>
QM_HDN void wbc_dynamics_measured(G g, const qmb200_model_desc& M, double gravity, const double* rbd, double* W) {
  double* kw = W + WA_KIN;
  double* ms = W + WA_MEAS;
  // ---- measured (WbcBase.cpp:146-203)
  if (g.tid() == 0) {
    for (int k = 0; k < 3; ++k) { ms[k] = rbd[3 + k]; ms[3 + k] = rbd[k]; ms[24 + k] = rbd[27 + k]; }
    for (int k = 0; k < 18; ++k) { ms[6 + k] = rbd[6 + k]; ms[30 + k] = rbd[30 + k]; }
    // [upstream] getEulerAnglesZyxDerivativesFromGlobalAngularVelocity
    double sz, cz, sy, cy;
    sincos(ms[3], &sz, &cz); sincos(ms[4], &sy, &cy);
    const double wx = rbd[24], wy = rbd[25], wz = rbd[26];
    const double dxr = (cz * wx + sz * wy) / cy;
    ms[24 + 3] = wz + sy * dxr;
    ms[24 + 4] = -sz * wx + cz * wy;
    ms[24 + 5] = dxr;
  }
  g.sync();
  wbc_kin_positions(g, M, ms, kw);
  QM_PFOR(g, k, QM_NJ) kw[KW_VEL + k] = ms[24 + k];
  QM_PFOR(g, idx, 144) W[WA_JEE + idx] = kw[KW_EEJ + idx];     // before the velocity level reuses the storage (KW_HB)
  g.sync();
  wbc_kin_velocities(g, M, kw);
  bias_forces(g, M, kw, gravity, W + WA_ACC, W + WA_FB);
  QM_PFOR(g, j, QM_NJ) {     // nle = S_j . sum of subtree forces (RNEA backward pass)
    double f[6] = {0, 0, 0, 0, 0, 0}, S[6];
    const uint32_t mask = M.submask[j];
    for (int i = 0; i < QM_NJ; ++i)
      if ((mask >> i) & 1u)
        for (int c = 0; c < 6; ++c) f[c] += W[WA_FB + 6 * i + c];
    joint_S(M, kw, j, S);
    W[WA_NLE + j] = S[0] * f[0] + S[1] * f[1] + S[2] * f[2] + S[3] * f[3] + S[4] * f[4] + S[5] * f[5];
  }
  QM_PFOR(g, idx, 576) {     // CRBA: M_ij = S_j . (I^c_i S_i) for j on the path of i
    const int i = idx / 24, j = idx % 24;
    double v = 0.0, S[6];
    if ((M.pathmask[i] >> j) & 1u) {
      joint_S(M, kw, j, S);
      const double* F = kw + KW_F + 6 * i;
      v = S[0] * F[0] + S[1] * F[1] + S[2] * F[2] + S[3] * F[3] + S[4] * F[4] + S[5] * F[5];
    } else if ((M.pathmask[j] >> i) & 1u) {
      joint_S(M, kw, i, S);
      const double* F = kw + KW_F + 6 * j;
      v = S[0] * F[0] + S[1] * F[1] + S[2] * F[2] + S[3] * F[3] + S[4] * F[4] + S[5] * F[5];
    }
    W[WA_M + idx] = v;
  }
  QM_PFOR(g, idx, 288) W[WA_JF + idx] = kw[KW_FJ + idx];
  QM_PFOR(g, idx, 72) {      // angular Jacobian of the base frame (joint 5)
    const int r = idx / 24, k = idx % 24;
    W[WA_JBA + idx] = (((M.pathmask[5] >> k) & 1u) && M.jtype[k] == 1) ? kw[KW_AX + 3 * k + r] : 0.0;
  }
  QM_PFOR(g, f, 6) {
    if (f < 4) {
      const int b = M.foot_joint[f];
      point_bias_accel(W + WA_ACC + 6 * b, kw + KW_V + 6 * b, kw + KW_FPOS + 3 * f, W + WA_DJV + 3 * f);
      for (int r = 0; r < 3; ++r) { ms[48 + 3 * f + r] = kw[KW_FPOS + 3 * f + r]; ms[60 + 3 * f + r] = kw[KW_FVEL + 3 * f + r]; }
    } else if (f == 4) {
      for (int r = 0; r < 3; ++r) W[WA_DJBV + r] = W[WA_ACC + 6 * 5 + r];
    } else {
      const int b = M.ee_joint;
      const double* Vb = kw + KW_V + 6 * b;
      const double* p = kw + KW_EEP;
      point_bias_accel(W + WA_ACC + 6 * b, Vb, p, W + WA_DJEE);
      // angular part with the base-orientation columns 3..5 of dJ removed (WbcBase.cpp:553-557)
      double wd[3] = {W[WA_ACC + 6 * b], W[WA_ACC + 6 * b + 1], W[WA_ACC + 6 * b + 2]};
      for (int k = 3; k < 6; ++k) {
        double t[3];
        cross3(kw + KW_V + 6 * k, kw + KW_AX + 3 * k, t);
        for (int r = 0; r < 3; ++r) wd[r] -= t[r] * kw[KW_VEL + k];
      }
      double vp[3];
      cross3(Vb, p, vp);
      for (int r = 0; r < 3; ++r) {
        W[WA_DJEE + 3 + r] = wd[r];
        ms[72 + r] = p[r];
        ms[75 + r] = Vb[3 + r] + vp[r];
        ms[78 + r] = Vb[r];
      }
      for (int r = 0; r < 9; ++r) ms[81 + r] = kw[KW_EER + r];
    }
  }
  g.sync();
}

// updateMeasured + updateDesired. rbd[55] measured state, xd/ud[30] MPC policy sample, u_last[30] (stateful inputLast_).
template <class G>
QM_HDN void wbc_dynamics(G g, const qmb200_model_desc& M, const qmb200_wbc_desc& C, const double* rbd, const double* xd,
                         const double* ud, const double* u_last, double period, double* W) {
  wbc_dynamics_measured(g, M, C.gravity, rbd, W);
  double* kw = W + WA_KIN;
  double* ds = W + WA_DES;
  // ---- desired (WbcBase.cpp:205-238)
  wbc_kin_positions(g, M, xd + 6, kw);
  centroidal_velocity(g, M, xd, ud, kw);
  wbc_kin_velocities(g, M, kw);
  bias_forces(g, M, kw, 0.0, W + WA_ACC, W + WA_FB);
  if (g.tid() == 0) {
    // Adot v = d/dt(A) v: total momentum rate with qdd = 0, moved to the centre of mass
    double L[3] = {0, 0, 0}, p[3] = {0, 0, 0}, t[3];
    for (int i = 0; i < QM_NJ; ++i)
      for (int r = 0; r < 3; ++r) { L[r] += W[WA_FB + 6 * i + r]; p[r] += W[WA_FB + 6 * i + 3 + r]; }
    cross3(kw + KW_COM, p, t);
    double rhs[6];
    // m * getNormalizedCentroidalMomentumRate(input)
    double fs[3] = {0, 0, 0}, ts[3] = {0, 0, 0};
    for (int ft = 0; ft < 4; ++ft) {
      double arm[3];
      for (int r = 0; r < 3; ++r) { arm[r] = kw[KW_FPOS + 3 * ft + r] - kw[KW_COM + r]; fs[r] += ud[3 * ft + r]; }
      cross3_add(arm, ud + 3 * ft, ts);
    }
    fs[2] -= M.total_mass * C.gravity;
    for (int r = 0; r < 3; ++r) { rhs[r] = fs[r] - p[r]; rhs[3 + r] = ts[r] - (L[r] - t[r]); }
    for (int l = 0; l < 18; ++l) {
      const double ja = (ud[12 + l] - u_last[12 + l]) / period;      // stateful inputLast_ (WbcBase.cpp:224-225)
      for (int r = 0; r < 6; ++r) rhs[r] -= kw[KW_ACM + r * QM_NJ + 6 + l] * ja;
    }
    for (int r = 0; r < 6; ++r) {
      double a = 0.0;
      for (int c = 0; c < 6; ++c) a += kw[KW_ABINV + 6 * r + c] * rhs[c];
      ds[48 + r] = a;
    }
    const int b = M.ee_joint;
    const double* Vb = kw + KW_V + 6 * b;
    double vp[3];
    cross3(Vb, kw + KW_EEP, vp);
    for (int r = 0; r < 3; ++r) { ds[78 + r] = kw[KW_EEP + r]; ds[81 + r] = Vb[3 + r] + vp[r]; }
    for (int r = 0; r < 9; ++r) ds[84 + r] = kw[KW_EER + r];
  }
  QM_PFOR(g, k, QM_NJ) { ds[k] = xd[6 + k]; ds[24 + k] = kw[KW_VEL + k]; }
  QM_PFOR(g, idx, 12) { ds[54 + idx] = kw[KW_FPOS + idx]; ds[66 + idx] = kw[KW_FVEL + idx]; }
  g.sync();
}

// The task stack (WbcBase.cpp:240-578 and the stacks of HierarchicalWbc.cpp:23-43 / HierarchicalMpcWbc.cpp:23-31).
// C.mpc_variant selects the stack below level 0:
//   0  HierarchicalWbc    : [arm joints (t < init_time) | base height + base angular + ee linear + ee angular + 100 swing] -> [contact force + base xy]
//   1  HierarchicalMpcWbc : [base height + base angular + base xy + 100 swing] -> [contact force]
//   2  six-level split of the same tasks (BASELINE config 5's "6 task levels"; SYNTHETIC, the reference has no such stack):
//      [base height + base angular] -> [ee linear + ee angular] -> [100 swing] -> [base xy] -> [contact force]
// nD0 to WI_SC[9]; the level table to WI_LV.
template <class G>
QM_HDN void wbc_tasks(G g, const qmb200_model_desc& M, const qmb200_wbc_desc& C, const double* ud, int mode, double time,
                      double* W, double* P, double* D0, double* Wc, int* WI) {
  // W: dynamics scratch (WA_*); P: base of the blocks WW_F0, WW_V0, WW_HJ and D0: [56][36] -- all three write only here (the
  // single-kernel solve passes its workspace for them, the kernel sequence the solve's image in global memory)
  const double* ms = W + WA_MEAS;
  const double* ds = W + WA_DES;
  const int nc = ((mode >> 3) & 1) + ((mode >> 2) & 1) + ((mode >> 1) & 1) + (mode & 1);
  const int nsw = 4 - nc;
  const bool init_stack = (C.mpc_variant == 0) && (time < C.init_time);
  const int nD0 = 36 + 5 * nc + 3 * nsw;
  if (g.tid() == 0) WI[WI_SC + 9] = nD0;
  QM_PFOR(g, idx, 18 * 36) Wc[WC_A0 + idx] = 0.0;
  QM_PFOR(g, idx, 56 * 36) D0[idx] = 0.0;
  QM_PFOR(g, idx, WB_POOL * 36) Wc[WC_AP + idx] = 0.0;
  QM_PFOR(g, idx, 56) { P[WW_F0 + idx] = 0.0; P[WW_V0 + idx] = 0.0; }
  g.sync();
  // level 0, equality rows: floating-base EoM (6), no contact motion (3 nc), zero swing force (3 nsw)
  QM_PFOR(g, idx, 18 * 36) {
    const int r = idx / 36, c = idx % 36;
    double v = 0.0;
    if (r < 6) {
      v = (c < 24) ? W[WA_M + 24 * r + c] : -W[WA_JF + (c - 24) * 24 + r];
    } else {
      // map row -> (foot, component)
      int rr = r - 6;
      if (rr < 3 * nc) {
        int j = rr / 3, d = rr % 3, ft = 0, cnt = 0;
        for (int f2 = 0; f2 < 4; ++f2) if ((mode >> (3 - f2)) & 1) { if (cnt == j) { ft = f2; break; } ++cnt; }
        if (c < 24) v = W[WA_JF + (3 * ft + d) * 24 + c];
      } else {
        rr -= 3 * nc;
        int j = rr / 3, d = rr % 3, ft = 0, cnt = 0;
        for (int f2 = 0; f2 < 4; ++f2) if (!((mode >> (3 - f2)) & 1)) { if (cnt == j) { ft = f2; break; } ++cnt; }
        if (c == 24 + 3 * ft + d) v = 1.0;
      }
    }
    Wc[WC_A0 + idx] = v;
  }
  QM_PFOR(g, r, 18) {
    double v = 0.0;
    if (r < 6) v = -W[WA_NLE + r];
    else if (r - 6 < 3 * nc) {
      int j = (r - 6) / 3, d = (r - 6) % 3, ft = 0, cnt = 0;
      for (int f2 = 0; f2 < 4; ++f2) if ((mode >> (3 - f2)) & 1) { if (cnt == j) { ft = f2; break; } ++cnt; }
      v = -W[WA_DJV + 3 * ft + d];
    }
    Wc[WC_B0 + r] = v;
    P[WW_HJ + r] = W[WA_NLE + 6 + r];
  }
  // level 0, inequality rows: torque limits (36), friction pyramid (5 nc), 3 nsw all-zero rows (WbcBase.cpp:458)
  QM_PFOR(g, idx, 18 * 36) {
    const int l = idx / 36, c = idx % 36;
    const double v = (c < 24) ? W[WA_M + 24 * (6 + l) + c] : -W[WA_JF + (c - 24) * 24 + 6 + l];
    D0[idx] = v;
    D0[18 * 36 + idx] = -v;
  }
  QM_PFOR(g, l, 18) {
    P[WW_F0 + l] = C.tau_max[l] - W[WA_NLE + 6 + l];
    P[WW_F0 + 18 + l] = C.tau_max[l] + W[WA_NLE + 6 + l];
  }
  QM_PFOR(g, idx, 5 * 4) {
    const int j = idx / 5, k = idx % 5;
    if (j < nc) {
      int ft = 0, cnt = 0;
      for (int f2 = 0; f2 < 4; ++f2) if ((mode >> (3 - f2)) & 1) { if (cnt == j) { ft = f2; break; } ++cnt; }
      double* row = D0 + (36 + 5 * j + k) * 36 + 24 + 3 * ft;
      const double mu = C.friction_mu;
      if (k == 0) { row[2] = -1.0; }
      else if (k == 1) { row[0] = 1.0; row[2] = -mu; }
      else if (k == 2) { row[0] = -1.0; row[2] = -mu; }
      else if (k == 3) { row[1] = 1.0; row[2] = -mu; }
      else { row[1] = -1.0; row[2] = -mu; }
    }
  }
  // rows of the levels below level 0 (one thread: scalar right-hand sides; the matrices are copies of Jacobian rows)
  if (g.tid() == 0) {
    const double* qm = ms; const double* vm = ms + 24;
    const double* qd = ds; const double* vd = ds + 24;
    const double* bacc = ds + 48;
    double* AP = Wc + WC_AP; double* bp = Wc + WC_BP;
    int row = 0, lev = 1;
    auto level_begin = [&]() { WI[WI_LV + 2 * lev] = row; };
    auto level_end = [&]() { WI[WI_LV + 2 * lev + 1] = row - WI[WI_LV + 2 * lev]; ++lev; };
    auto arm_joints = [&]() {               // formulateArmJointNomalTrackingTask
      for (int i = 0; i < 6; ++i) {
        AP[(row + i) * 36 + 18 + i] = 1.0;
        bp[row + i] = C.kp_arm_joint[i] * (qd[18 + i] - qm[18 + i]) + C.kd_arm_joint[i] * (vd[18 + i] - vm[18 + i]);
      }
      row += 6;
    };
    auto base_height = [&]() {
      AP[row * 36 + 2] = 1.0;
      bp[row] = bacc[2] + C.kp_base_height * (qd[2] - qm[2]) + C.kd_base_height * (vd[2] - vm[2]);
      ++row;
    };
    auto base_angular = [&]() {
      double sz, cz, sy, cy;
      sincos(qm[3], &sz, &cz); sincos(qm[4], &sy, &cy);
      const double T[9] = {0, -sz, cy * cz, 0, cz, cy * sz, 1, 0, -sy};
      const double dz = vd[3], dy = vd[4];
      const double Td[9] = {0, -cz * dz, -sy * cz * dy - cy * sz * dz, 0, -sz * dz, -sy * sz * dy + cy * cz * dz, 0, 0, -cy * dy};
      double Rm[9], Rd[9], err[3];
      rot_zyx(qm + 3, Rm);
      rot_zyx(qd + 3, Rd);
      rotation_error_world(Rd, Rm, err);
      for (int r = 0; r < 3; ++r) {
        double wm = 0, wdv = 0, acc = 0;
        for (int c = 0; c < 3; ++c) {
          wm += T[3 * r + c] * vm[3 + c];
          wdv += T[3 * r + c] * vd[3 + c];
          acc += T[3 * r + c] * bacc[3 + c] + Td[3 * r + c] * vd[3 + c];
        }
        for (int c = 0; c < 24; ++c) AP[(row + r) * 36 + c] = W[WA_JBA + 24 * r + c];
        bp[row + r] = acc + C.kp_base_angular * err[r] + C.kd_base_angular * (wdv - wm) - W[WA_DJBV + r];
      }
      row += 3;
    };
    auto base_linear = [&]() {
      for (int r = 0; r < 2; ++r) {
        AP[(row + r) * 36 + r] = 1.0;
        bp[row + r] = bacc[r] + C.kp_base_linear * (qd[r] - qm[r]) + C.kd_base_linear * (vd[r] - vm[r]);
      }
      row += 2;
    };
    auto ee_linear = [&]() {
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 24; ++c) AP[(row + r) * 36 + c] = W[WA_JEE + 24 * r + c];
        bp[row + r] = C.kp_ee_linear[r] * (ds[78 + r] - ms[72 + r]) + C.kd_ee_linear[r] * (ds[81 + r] - ms[75 + r]) - W[WA_DJEE + r];
      }
      row += 3;
    };
    auto ee_angular = [&]() {               // base-orientation columns zeroed (WbcBase.cpp:553,557)
      double err[3];
      rotation_error_world(ds + 84, ms + 81, err);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 24; ++c) AP[(row + r) * 36 + c] = (c >= 3 && c < 6) ? 0.0 : W[WA_JEE + 24 * (3 + r) + c];
        bp[row + r] = C.kp_ee_angular[r] * err[r] - C.kd_ee_angular[r] * ms[78 + r] - W[WA_DJEE + 3 + r];
      }
      row += 3;
    };
    auto swing_legs = [&]() {               // weighted (Task::operator*, HierarchicalWbc.cpp:29)
      for (int ft = 0; ft < 4; ++ft) {
        if ((mode >> (3 - ft)) & 1) continue;
        for (int d = 0; d < 3; ++d) {
          const double acc = C.kp_swing * (ds[54 + 3 * ft + d] - ms[48 + 3 * ft + d]) + C.kd_swing * (ds[66 + 3 * ft + d] - ms[60 + 3 * ft + d]);
          for (int c = 0; c < 24; ++c) AP[(row + d) * 36 + c] = C.swing_weight * W[WA_JF + (3 * ft + d) * 24 + c];
          bp[row + d] = C.swing_weight * (acc - W[WA_DJV + 3 * ft + d]);
        }
        row += 3;
      }
    };
    auto contact_force = [&]() {
      for (int i = 0; i < 12; ++i) { AP[(row + i) * 36 + 24 + i] = 1.0; bp[row + i] = ud[i]; }
      row += 12;
    };
    if (C.mpc_variant == 2) {
      level_begin(); base_height(); base_angular(); level_end();
      level_begin(); ee_linear(); ee_angular(); level_end();
      level_begin(); swing_legs(); level_end();
      level_begin(); base_linear(); level_end();
      level_begin(); contact_force(); level_end();
    } else if (C.mpc_variant == 1) {
      level_begin(); base_height(); base_angular(); base_linear(); swing_legs(); level_end();
      level_begin(); contact_force(); level_end();
    } else {
      level_begin();
      if (init_stack) arm_joints();
      else { base_height(); base_angular(); ee_linear(); ee_angular(); swing_legs(); }
      level_end();
      level_begin(); contact_force(); base_linear(); level_end();
    }
    WI[WI_LV] = lev;
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------ dense helpers
// Householder triangularisation of the m x (n+1) matrix A (row major, ld): columns 0..n-1 are reduced, column n (rhs)
// is transformed along. On exit the upper triangle holds R and A[0:n][n] = Q'rhs. vh: m doubles, wj: ld + 2, hp: 4 x 56 scratch.
// Every step is three short phases over the whole group: (a) inner products of column k with the columns k..n, each split
// over four row chunks; (b) the reflector's scalars and w = beta v'A per column, the reflector itself copied aside;
// (c) the rank-one update. No phase has a serial loop longer than a quarter of a column.
enum { HH_PARTS = 4, HH_LD = 56 };
// dense: the rows below `dense` are sqrt(eps) e_c' (the regularisation block of HoQp.cpp:66, zero right-hand side): row
// dense + c is untouched until step c, so step k only involves the rows k .. dense + k (the others hold exact zeros in the
// columns k..n and their reflector entry is zero).
template <class G>
QM_HDO void householder_ls(G g, double* A, int m, int n, int ld, double* vh, double* wj, double* hp, int dense) {
  const int steps = (m - 1 < n) ? m - 1 : n;
  for (int k = 0; k < steps; ++k) {
    const int mk = (dense + k + 1 < m) ? dense + k + 1 : m;   // one past the last row step k touches
    const int rows = mk - k, cols = n - k + 1;                // columns k..n (rhs included); column k gives the norm
    const int chunk = (rows + HH_PARTS - 1) / HH_PARTS;
    QM_PFOR2(g, part, HH_PARTS, jj, cols) {
      const int j = k + jj;
      const int i0 = k + part * chunk, i1 = (i0 + chunk < mk) ? i0 + chunk : mk;
      double sp = 0.0;
      const double* ak = A + i0 * ld + k;
      const double* aj = A + i0 * ld + j;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
      for (int i = i0; i < i1; ++i, ak += ld, aj += ld) sp += *ak * *aj;
      hp[part * HH_LD + jj] = sp;
    }
    g.sync();
    {
      const double s = (hp[0] + hp[HH_LD]) + (hp[2 * HH_LD] + hp[3 * HH_LD]);      // |A[k:, k]|^2
      const double nrm = sqrt(s);
      const double akk = A[k * ld + k];
      const double alpha = (akk >= 0.0) ? -nrm : nrm;
      const double vk = akk - alpha;
      const double vv = s - akk * akk + vk * vk;               // v'v with v = A[k:, k] - alpha e_k
      const double beta = (vv > 0.0) ? 2.0 / vv : 0.0;
      QM_PFOR(g, jj, n - k) {                                  // columns k+1..n: v'A_j = A_k'A_j - alpha A[k][j]
        const int j = k + 1 + jj;
        const double c = (hp[jj + 1] + hp[HH_LD + jj + 1]) + (hp[2 * HH_LD + jj + 1] + hp[3 * HH_LD + jj + 1]);
        wj[j] = beta * (c - alpha * A[k * ld + j]);
      }
      QM_PFOR(g, i, rows) vh[k + i] = (i == 0) ? vk : A[(k + i) * ld + k];
      if (g.tid() == 0) wj[ld + 1] = alpha;
    }
    g.sync();
    // rank-one update, rows over the warps and columns over the lanes (no index division); a lane keeps its w_j in a register
    for (int jj = g.tid() & 31; jj < n - k; jj += (g.nt() < 32 ? g.nt() : 32)) {
      const double w = wj[k + 1 + jj];
      const int r0 = g.tid() >> 5, rs = (g.nt() + 31) >> 5;
      double* QM_RESTRICT a = A + (k + r0) * ld + k + 1 + jj;
      const double* QM_RESTRICT v = vh + k + r0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
      for (int ii = r0; ii < rows; ii += rs, a += rs * ld, v += rs) *a -= *v * w;
    }
    QM_PFOR(g, i, rows) A[(k + i) * ld + k] = (i == 0) ? wj[ld + 1] : 0.0;
    g.sync();
  }
}

// y[i * ldy] -= x[i * ldx] * w for i < count, four rows at a time with all loads ahead of the first store (the compiler keeps a
// load of x behind the preceding store to y otherwise: a shared-memory round trip per row on the chain)
QM_HD void hh_axpy(double* y, int ldy, const double* x, int ldx, int count, double w) {
  int i = 0;
  for (; i + 4 <= count; i += 4, y += 4 * ldy, x += 4 * ldx) {
    const double x0 = x[0], x1 = x[ldx], x2 = x[2 * ldx], x3 = x[3 * ldx];
    const double y0 = y[0], y1 = y[ldy], y2 = y[2 * ldy], y3 = y[3 * ldy];
    y[0] = y0 - x0 * w; y[ldy] = y1 - x1 * w; y[2 * ldy] = y2 - x2 * w; y[3 * ldy] = y3 - x3 * w;
  }
  for (; i < count; ++i, y += ldy, x += ldx) *y -= *x * w;
}

// The same triangularisation on one narrow group (a warp on the device), lane per column: a step is one pass of inner products
// (each lane its own columns against column k, rows in order), the reflector's scalars formed by every lane, and one pass of
// updates with w_j in a register -- no partial sums to combine, no scalar work repeated by four warps, two warp-level syncs per
// step instead of three block barriers. A third of the instructions of householder_ls and a third of its latency for the
// matrices of a solve (<= 92 x 37): the triangularisation is what bounded k_wbc_level by instruction issue.
// hp: 2 doubles of scratch. Same `dense` convention as householder_ls.
QM_HD double hh_dot(const double* QM_RESTRICT ak, const double* QM_RESTRICT aj, int ld, int count) {
  // two interleaved partial sums (even / odd rows): half the dependent multiply-add chain
  double s0 = 0.0, s1 = 0.0;
  int i = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
  for (; i + 2 <= count; i += 2, ak += 2 * ld, aj += 2 * ld) { s0 += ak[0] * aj[0]; s1 += ak[ld] * aj[ld]; }
  if (i < count) s0 += *ak * *aj;
  return s0 + s1;
}
template <class G>
QM_HDO void householder_ls_narrow(G w0, double* A, int m, int n, int ld, double* hp, int dense) {
  const int steps = (m - 1 < n) ? m - 1 : n;
  for (int k = 0; k < steps; ++k) {
    const int mk = (dense + k + 1 < m) ? dense + k + 1 : m;   // one past the last row step k touches
    const int rows = mk - k, cols = n - k + 1;
    double* colk = A + k * ld + k;
#if defined(__CUDA_ARCH__)
    // a lane owns column k + lane and, in the first steps of a matrix wider than the warp, column k + lane + 32
    (void)hp;
    const int lane = w0.tid();
    const bool has0 = lane < cols, has1 = lane + 32 < cols;
    const double d0 = has0 ? hh_dot(colk, colk + lane, ld, rows) : 0.0;
    const double d1 = has1 ? hh_dot(colk, colk + lane + 32, ld, rows) : 0.0;
    const double s = __shfl_sync(0xffffffffu, d0, 0);            // |A[k:, k]|^2
#else
    double dot[WB_QR_LD + 3];
    for (int jj = 0; jj < cols; ++jj) dot[jj] = hh_dot(colk, colk + jj, ld, rows);
    const double s = dot[0];
#endif
    const double nrm = sqrt(s);
    const double akk = *colk;
    const double alpha = (akk >= 0.0) ? -nrm : nrm;
    const double vk = akk - alpha;
    const double vv = s - akk * akk + vk * vk;                 // v'v with v = A[k:, k] - alpha e_k
    const double beta = (vv > 0.0) ? 2.0 / vv : 0.0;
    // column j > k: w = beta v'A_j with v'A_j = A_k'A_j - alpha A[k][j]; A_j -= v w
    auto update = [&](int jj, double d) {
      double* aj = colk + jj;
      const double w = beta * (d - alpha * *aj);
      *aj -= vk * w;
      hh_axpy(aj + ld, ld, colk + ld, ld, rows - 1, w);
    };
#if defined(__CUDA_ARCH__)
    if (has0 && lane > 0) update(lane, d0);
    if (has1) update(lane + 32, d1);
#else
    for (int jj = 1; jj < cols; ++jj) update(jj, dot[jj]);
#endif
    w0.sync();
    // R[k][k]; the entries below it are never read again (the later steps and the back substitution stay right of column k),
    // so they are left as they are and the next step starts without another hand-over
    if (w0.tid() == 0) colk[0] = alpha;
  }
  w0.sync();
}

// back substitution R z = c (R n x n upper in A, n <= 64, c = A[:, n]) on one narrow group, column sweep from the last unknown:
// a lane keeps the right-hand sides of its rows (lane, lane + 32) in registers, the diagonal enters through its reciprocal
// (one division per lane instead of one per unknown on the chain), z_i reaches the other lanes by a shuffle.
template <class G>
QM_HDO void back_substitute(G w0, const double* A, int n, int ld, double* z) {
#if defined(__CUDA_ARCH__)
  const int j0 = w0.tid(), j1 = j0 + 32;
  double c0 = (j0 < n) ? A[j0 * ld + n] : 0.0, c1 = (j1 < n) ? A[j1 * ld + n] : 0.0;
  const double d0 = (j0 < n) ? A[j0 * ld + j0] : 0.0, d1 = (j1 < n) ? A[j1 * ld + j1] : 0.0;
  const double inv0 = (d0 != 0.0) ? 1.0 / d0 : 0.0, inv1 = (d1 != 0.0) ? 1.0 / d1 : 0.0;
  for (int i = n - 1; i > 0; --i) {
    const double zi = (i >= 32) ? __shfl_sync(0xffffffffu, c1 * inv1, i - 32) : __shfl_sync(0xffffffffu, c0 * inv0, i);
    if (j0 < i) c0 -= A[j0 * ld + i] * zi;
    if (j1 < i) c1 -= A[j1 * ld + i] * zi;
  }
  if (j0 < n) z[j0] = c0 * inv0;
  if (j1 < n) z[j1] = c1 * inv1;
  w0.sync();
#else
  double c[64], inv[64];
  for (int j = 0; j < n; ++j) { c[j] = A[j * ld + n]; const double d = A[j * ld + j]; inv[j] = (d != 0.0) ? 1.0 / d : 0.0; }
  for (int i = n - 1; i >= 0; --i) {
    z[i] = c[i] * inv[i];
    for (int j = 0; j < i; ++j) c[j] -= A[j * ld + i] * z[i];
  }
#endif
}

// ---- cross-lane helpers of the pivot search and of the active-set iteration (one warp on the device, one thread on the host). The sums run in the same
// butterfly order on both sides, so host and device agree on them bit for bit.
template <class G, class F>
QM_HD double gi_sum(G w0, int n, F f) {                       // sum of f(i), i < n <= 32; every lane gets the result
#if defined(__CUDA_ARCH__)
  const int lane = w0.tid();
  double v = (lane < n) ? f(lane) : 0.0;
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
#else
  double v[32], t[32];
  for (int i = 0; i < 32; ++i) v[i] = (i < n) ? f(i) : 0.0;
  for (int s = 16; s > 0; s >>= 1) {
    for (int i = 0; i < 32; ++i) t[i] = v[i] + v[i ^ s];
    for (int i = 0; i < 32; ++i) v[i] = t[i];
  }
  return v[0];
#endif
}
#if defined(__CUDA_ARCH__)
// Warp arg-min of (value, index) with the hardware warp reductions: the doubles are mapped to 64-bit keys of the same order
// (sign flip; -0.0 folded into +0.0 first, NaN never wins as in the serial scan's `v < best`), the minimum key is found as the
// minimum high word, then the minimum low word among its holders, then the lowest index among those -- three redux.sync
// instead of five shuffle rounds of three shuffles. bi < 0 marks a lane without a candidate (its value is the bound).
__device__ __forceinline__ int warp_argmin(double best, int bi, double* vmin) {
  const double c = (best == best) ? best + 0.0 : __longlong_as_double(0x7ff0000000000000LL);
  unsigned long long k = (unsigned long long)__double_as_longlong(c);
  k = (k >> 63) ? ~k : (k | 0x8000000000000000ULL);
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const bool c1 = hi == mh;
  const unsigned ml = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
  const bool c2 = c1 && lo == ml;
  const unsigned idx = __reduce_min_sync(0xffffffffu, (c2 && bi >= 0) ? (unsigned)bi : 0x7fffffffu);
  unsigned long long km = ((unsigned long long)mh << 32) | ml;
  km = (km >> 63) ? (km & 0x7fffffffffffffffULL) : ~km;
  *vmin = __longlong_as_double((long long)km);
  return (idx == 0x7fffffffu) ? -1 : (int)idx;
}
#endif
// index of the smallest f(i) below `bound` over i < n (the lowest index on ties), -1 if there is none; *vmin: that value or bound
template <class G, class F>
QM_HD int gi_argmin(G w0, int n, double bound, F f, double* vmin) {
  double best = bound;
  int bi = -1;
#if defined(__CUDA_ARCH__)
  for (int i = w0.tid(); i < n; i += 32) { const double v = f(i); if (v < best) { best = v; bi = i; } }
  bi = warp_argmin(best, bi, &best);
#else
  for (int i = 0; i < n; ++i) { const double v = f(i); if (v < best) { best = v; bi = i; } }
#endif
  *vmin = (bi >= 0) ? best : bound;
  return bi;
}
// gi_argmin over i < nmin (<= 32) and two gi_sums over i < nsum in one pass (the three butterflies interleave on the device)
template <class G, class FV, class F1, class F2>
QM_HD int gi_argmin_sum2(G w0, int nmin, double bound, FV fv, double* vmin, int nsum, F1 f1, F2 f2, double* s1, double* s2) {
#if defined(__CUDA_ARCH__)
  const int lane = w0.tid();
  double best = bound, a = 0.0, b = 0.0;
  int bi = -1;
  if (lane < nmin) { const double v = fv(lane); if (v < best) { best = v; bi = lane; } }
  if (lane < nsum) { a = f1(lane); b = f2(lane); }
  for (int s = 16; s > 0; s >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, s);
    b += __shfl_xor_sync(0xffffffffu, b, s);
  }
  bi = warp_argmin(best, bi, &best);
  *vmin = (bi >= 0) ? best : bound; *s1 = a; *s2 = b;
  return bi;
#else
  *s1 = gi_sum(w0, nsum, f1);
  *s2 = gi_sum(w0, nsum, f2);
  return gi_argmin(w0, nmin, bound, fv, vmin);
#endif
}
// Kernel basis of Abar (r x n, row major ld_a) exactly as the reference obtains it (HoQp.cpp:129: (A Zprev).fullPivLu().kernel(),
// [upstream] Eigen::FullPivLU<MatrixXd>): elimination with full pivoting -- the pivot of step k is the entry of largest magnitude
// of the remaining corner, the first one in column-major order on ties --, rank = pivots above eps * min(r, n) * max|pivot|,
// kernel = Q [ -U11^-1 U12 ; I ] with Q the accumulated column permutation. The basis is NOT orthonormal: HoQp's 1e-12 |z|^2
// regularisation acts on the coordinates in this basis, which is what selects the solution of a rank-deficient level.
// T: r x n scratch; N: n x ldn output (columns 0 .. min(n - rank, maxcols) - 1); cn: n + 2 doubles; perm: 2 n + 4 ints. *rank_out = rank.
// Eigen returns a trivial kernel as one zero column (a dummy variable that moves nothing); here rank = n means "no freedom left".
// Runs on one narrow group (a warp on the device; the caller's other warps wait): lane per column in the pivot search (the
// winner by a warp arg-max with the serial scan's tie rule), the swaps and the elimination, whose rows are loaded four ahead of
// the stores -- no block barrier in the eighteen dependent steps.
template <class G>
QM_HDO void kernel_basis_lu(G g, const double* Abar, int r, int n, int ld_a, double* T, double* N, int ldn, int maxcols, double* cn,
                            int* perm, int* rank_out, int* status) {
  int* qidx = perm;                 // [n] column permutation
  int* brow = perm + n;             // [n] row of the largest entry per column
  if (ld_a == n) { QM_PFOR(g, idx, r * n) T[idx] = Abar[idx]; }       // (the level-0 rows: a straight copy out of global memory)
  else { QM_PFOR(g, idx, r * n) { const int i = idx / n, j = idx - i * n; T[idx] = Abar[i * ld_a + j]; } }
  QM_PFOR(g, j, n) qidx[j] = j;
  g.sync();
  const int size = (r < n) ? r : n;
  int nonzero = size;
  double maxpivot = 0.0;
  for (int k = 0; k < size; ++k) {
    // largest magnitude of every remaining column over the rows k..r-1 (first row on ties), then over the columns in order
    // (strictly greater wins: the first maximum in column-major order)
    double negbest;
    const int bjj = gi_argmin(g, n - k, 1e300, [&](int jj) {
      const int j = k + jj;
      double best = fabs(T[k * n + j]);
      int bi = k;
      for (int i = k + 1; i < r; ++i) { const double v = fabs(T[i * n + j]); if (v > best) { best = v; bi = i; } }
      brow[j] = bi;
      return -best;
    }, &negbest);
    g.sync();
    const double best = (bjj >= 0) ? -negbest : 0.0;
    if (best == 0.0) { nonzero = k; break; }
    if (best > maxpivot) maxpivot = best;
    const int bj = k + bjj, bi = brow[bj];
    if (bi != k) QM_PFOR(g, j, n) { const double t = T[k * n + j]; T[k * n + j] = T[bi * n + j]; T[bi * n + j] = t; }
    g.sync();
    if (bj != k) {
      QM_PFOR(g, i, r) { const double t = T[i * n + k]; T[i * n + k] = T[i * n + bj]; T[i * n + bj] = t; }
      if (g.tid() == 0) { const int t = qidx[k]; qidx[k] = qidx[bj]; qidx[bj] = t; }
    }
    g.sync();
    const double piv = T[k * n + k];
    QM_PFOR(g, ii, r - k - 1) T[(k + 1 + ii) * n + k] /= piv;
    g.sync();
    if (k < size - 1) {
      QM_PFOR(g, jj, n - k - 1) {
        const int j = k + 1 + jj;
        hh_axpy(T + (k + 1) * n + j, n, T + (k + 1) * n + k, n, r - k - 1, T[k * n + j]);
      }
      g.sync();
    }
  }
  // rank: pivots above the threshold (FullPivLU::threshold() default). They form a prefix of the diagonal with full pivoting up
  // to rounding noise; anything else is flagged and the prefix is used.
  const double thr = maxpivot * 2.220446049250313e-16 * (double)size;
  int rank = 0, above = 0;
  for (int i = 0; i < nonzero; ++i) if (fabs(T[i * n + i]) > thr) { ++above; if (rank == i) rank = i + 1; }
  if (above != rank && g.tid() == 0) status_or(status, WST_DEGENERATE);
  const int dimker = n - rank;
  // X = U11^-1 U12 in place (one right-hand side column per work item)
  // (the reciprocals of the pivots are formed side by side first: one division per lane instead of one per row on the chain)
  QM_PFOR(g, i, rank) cn[i] = 1.0 / T[i * n + i];
  g.sync();
  QM_PFOR(g, cc, dimker) {
    const int c = rank + cc;
    for (int i = rank - 1; i >= 0; --i) {
      double sacc = T[i * n + c];
      for (int j = i + 1; j < rank; ++j) sacc -= T[i * n + j] * T[j * n + c];
      T[i * n + c] = sacc * cn[i];
    }
  }
  g.sync();
  QM_PFOR2(g, i, n, kk, (dimker < maxcols ? dimker : maxcols)) {      // at most maxcols kernel columns are kept (the caller flags more)
    double v = 0.0;
    if (i < rank) v = -T[i * n + rank + kk];
    else if (i - rank == kk) v = 1.0;
    N[qidx[i] * ldn + kk] = v;
  }
  if (g.tid() == 0) *rank_out = rank;
  g.sync();
}

// ------------------------------------------------------------------------------------------ level 0
// min 1/2|A0 z - b0|^2 + eps/2 |z|^2 + 1/2 |(D0 z - f0)+|^2  by Newton iteration on the active set of violated rows.
// Storage of level 0: blocks of the solve's workspace (l0_mem_of) or the compact per-warp block of the stand-alone kernel.
struct L0Mem {
  const double* F0;                // [56] right-hand sides of the inequality rows
  double *V0, *X;                  // [56] slack solution, [36] solution
  double *QR, *res, *hp;           // [92][37] least-squares matrix | rhs, [56] row residuals, [4] scratch
};
// doubles of a compact block whose matrix holds `rows` rows (18 equality + 36 regularisation rows + the active inequality rows)
QM_HD constexpr int l0_mem_doubles(int rows) { return 56 + 56 + 36 + rows * WB_QR_LD + 56 + 4; }
enum { L0_MEM_DOUBLES = 56 + 56 + 36 + WB_QR_ROWS * WB_QR_LD + 56 + 4, L0_FIXED_ROWS = 18 + 36 };
QM_HD L0Mem l0_mem_of(double* W) {
  L0Mem m;
  m.F0 = W + WW_F0; m.V0 = W + WW_V0; m.X = W + WW_X; m.QR = W + WS_QR; m.res = W + WS_RES; m.hp = W + WS_HP;
  return m;
}
QM_HD L0Mem l0_mem_compact(double* D, int rows = WB_QR_ROWS) {   // l0_mem_doubles(rows) doubles; F0 is filled by the caller
  L0Mem m;
  m.F0 = D; m.V0 = D + 56; m.X = m.V0 + 56; m.QR = m.X + 36; m.res = m.QR + rows * WB_QR_LD; m.hp = m.res + 56;
  return m;
}
// max_nw: active inequality rows the matrix of `lm` has room for; false (nothing usable written) if a pass needs more.
template <class G>
QM_HDN bool wbc_level0(G g, const L0Mem& lm, const double* D0, const double* Wc, int* WI, int max_nw = WB_MAXW) {
  const double* F0 = lm.F0;
  double* X = lm.X;
  double* RES = lm.res;
  const int nD0 = WI[WI_SC + 9];
  const int ld = WB_QR_LD;
  // The matrix of a pass is [active inequality rows; 18 equality rows; 36 rows sqrt(eps) I]. Its first row is kept max_nw rows
  // below the start of the storage: when a later pass only ADDS active rows (the usual second pass: none -> a few), the
  // triangular factor R | Q'rhs of the previous pass is that pass's whole problem, so the new rows go on top of it and the
  // triangularisation runs on [new rows; R] -- a window of (new rows + 1) rows per step instead of (active + 19). A pass that
  // drops a row, or runs out of head room, rebuilds the matrix.
  // WI_IGN: row is in the current factor; WI_SC [2] pass is an update, [3] first row of the pass's matrix, [5] new rows,
  // [7] rows of head room used.
  int* inr = WI + WI_IGN;
  QM_PFOR(g, i, 56) { WI[WI_INW + i] = 0; inr[i] = 0; }
  if (g.tid() == 0) { WI[WI_SC + 10] = 0; WI[WI_SC + 7] = -1; }
  g.sync();
  for (int iter = 0; iter < 40; ++iter) {
    if (g.tid() == 0) {                    // active row list (all active rows, or the ones to add to the factor)
      int nw = 0, nadd = 0, nrem = 0;
      for (int i = 0; i < nD0; ++i) {
        const int in = WI[WI_INW + i];
        nw += in; nadd += (in && !inr[i]); nrem += (!in && inr[i]);
      }
      const int used = WI[WI_SC + 7];
      const int upd = (used >= 0 && nrem == 0 && used + nadd <= max_nw);
      int k = 0;
      if (upd) {
        for (int i = 0; i < nD0; ++i) if (WI[WI_INW + i] && !inr[i]) { WI[WI_PERM + k++] = i; inr[i] = 1; }
        WI[WI_SC + 7] = used + nadd;
        WI[WI_SC + 3] = max_nw - used - nadd;
      } else {
        for (int i = 0; i < nD0 && k < WB_MAXW; ++i) if (WI[WI_INW + i]) WI[WI_PERM + k++] = i;
        for (int i = 0; i < nD0; ++i) inr[i] = WI[WI_INW + i];
        nw = k;
        WI[WI_SC + 7] = nw;
        WI[WI_SC + 3] = max_nw - nw;
      }
      WI[WI_SC + 0] = nw; WI[WI_SC + 2] = upd; WI[WI_SC + 5] = k;
    }
    g.sync();
    const int nw = WI[WI_SC + 0], upd = WI[WI_SC + 2], nnew = WI[WI_SC + 5];
    if (nw > max_nw) return false;
    double* QR = lm.QR + WI[WI_SC + 3] * ld;           // first row of this pass's matrix
    const int m = upd ? nnew + 36 : nw + 18 + 36;
    if (upd) {
      QM_PFOR(g, idx, nnew * ld) {
        const int r = idx / ld, c = idx % ld, i = WI[WI_PERM + r];
        QR[idx] = (c < 36) ? D0[36 * i + c] : F0[i];
      }
    } else {
      // rows: active inequality rows, equality rows, sqrt(eps) I (large rows first: stable for the tiny regularisation)
      QM_PFOR(g, idx, m * ld) {
        const int r = idx / ld, c = idx % ld;
        double v;
        if (r < nw) { const int i = WI[WI_PERM + r]; v = (c < 36) ? D0[36 * i + c] : F0[i]; }
        else if (r < nw + 18) { const int i = r - nw; v = (c < 36) ? Wc[WC_A0 + 36 * i + c] : Wc[WC_B0 + i]; }
        else { const int i = r - nw - 18; v = (c == i) ? 1e-6 : 0.0; }
        QR[idx] = v;
      }
    }
    g.sync(); QM_TICK(35);
    if (g.narrow_active()) {
      householder_ls_narrow(g.narrow(), QR, m, 36, ld, lm.hp, upd ? nnew : nw + 18);
      QM_TICK(36);
      back_substitute(g.narrow(), QR, 36, ld, X);
    }
    g.sync();
    QM_PFOR(g, i, nD0) {
      double s = -F0[i];
      for (int c = 0; c < 36; ++c) s += D0[36 * i + c] * X[c];
      RES[i] = s;
    }
    g.sync();
    if (g.tid() == 0) {
      int changed = 0;
      for (int i = 0; i < nD0; ++i) {
        const double tol = 1e-9 * (1.0 + fabs(F0[i]));
        const int in = WI[WI_INW + i];
        const int nw_in = in ? (RES[i] > -tol) : (RES[i] > tol);
        if (nw_in != in) { WI[WI_INW + i] = nw_in; ++changed; }
      }
      WI[WI_SC + 10] = (changed == 0);
      if (changed && iter == 39) WI[WI_SC + 6] |= WST_QP_MAX_ITER;
    }
    g.sync(); QM_TICK(37);
    if (WI[WI_SC + 10]) break;
  }
  QM_PFOR(g, i, 56) lm.V0[i] = (i < nD0 && RES[i] > 0.0) ? RES[i] : 0.0;
  g.sync();
  return true;
}

// Goldfarb-Idnani iteration on a group (see wbc_gi). State: z (WS_Z), J (WS_J), RF (WS_RF), multipliers u (WS_U),
// active list (WI_ACT), iq (WI_SC+4). Scratch scalars in WS_CN: [0] t, [1] cip, [20..38] cs, [40..58]... see below.
enum { GI_T = 0, GI_CIP = 1, GI_CS = 2, GI_SN = 20 };          // offsets into WS_CN (40 doubles): t, c_ip, cs[18], sn[18]
enum { GI_ACTION = 11, GI_L = 12, GI_NROT = 13 };              // offsets into WI_SC: 0 add, 1 drop, 2 skip/ignore, 3 done
// rr = R^-1 d for the leading q x q upper triangle of RF (ld 18), column by column from the last one: lane i owns row i, the
// diagonal enters through its reciprocal (one division per lane, off the chain)
template <class G>
QM_HD void gi_backsub(G w0, const double* RF, const double* d, int q, double* rr) {
#if defined(__CUDA_ARCH__)
  const int lane = w0.tid();
  double s = (lane < q) ? d[lane] : 0.0;
  const double inv = (lane < q) ? 1.0 / RF[lane * 18 + lane] : 0.0;
  for (int j = q - 1; j > 0; --j) {
    const double rj = __shfl_sync(0xffffffffu, s * inv, j);
    if (lane < j) s -= RF[lane * 18 + j] * rj;
  }
  if (lane < q) rr[lane] = s * inv;
#else
  double s[18];
  for (int i = 0; i < q; ++i) s[i] = d[i];
  for (int j = q - 1; j >= 0; --j) {
    rr[j] = s[j] * (1.0 / RF[j * 18 + j]);
    for (int i = 0; i < j; ++i) s[i] -= RF[i * 18 + j] * rr[j];
  }
#endif
}

// Goldfarb-Idnani dual active-set iteration on  min 1/2 |z - z0|^2_H  s.t.  GG z <= Gg  in the factored form J = L^-T Q, R
// ([upstream] the QP of one HoQp level, HoQp.cpp:129-158, is handed to qpOASES; this is the dense dual method restated).
// One warp: vector operations over the lanes, scalar decisions formed redundantly by every lane from warp-wide reductions
// (no lane-0 sections on the chain except the bookkeeping stores).
// Storage of the iteration: the solve's workspace (gi_mem_of) or the compact per-warp block of the stand-alone kernel (k_wbc_gi).
struct GiMem {
  const double* GG;                // [18][56] inequality rows by column (read only; may live in global memory)
  double* Gg;                      // [56] right-hand sides (read only)
  double *J, *RF;                  // [18][18] each
  double *z, *d, *rr, *zd, *np;    // [18] each
  double *res, *viol, *scl;        // [56] each: constraint values, scaled violations, row scales 1 / (1 + |Gg_i|)
  double *u, *sc;                  // [60] multipliers, [40] scalars / rotation coefficients
  int *act, *ina, *ign;            // [36] active list, [56] row is active, [56] row is ignored (dependent and marginally violated)
  int *status;                     // WST_* flags of the solve (or-ed into)
};
enum { GI_MEM_DOUBLES = 56 + 2 * 324 + 5 * 20 + 3 * 56 + 60 + 40, GI_MEM_INTS = 36 + 56 + 56 + 4 };   // compact block, GG elsewhere
QM_HD GiMem gi_mem_of(double* W, int* WI) {
  GiMem m;
  m.GG = W + WS_GG; m.Gg = W + WS_Gg; m.J = W + WS_J; m.RF = W + WS_RF;
  m.z = W + WS_Z; m.d = W + WS_D; m.rr = W + WS_RR; m.zd = W + WS_ZD; m.np = W + WS_NP;
  m.res = W + WS_RES; m.viol = W + WS_VH;      // (the Householder vector storage is idle during the iteration)
  m.scl = W + WS_SCL; m.u = W + WS_U; m.sc = W + WS_CN;
  m.act = WI + WI_ACT; m.ina = WI + WI_INW;    // (the level-0 flags of this name are dead by now)
  m.ign = WI + WI_IGN; m.status = WI + WI_SC + 6;
  return m;
}
QM_HD GiMem gi_mem_compact(double* D, int* I, const double* GG) {   // GI_MEM_DOUBLES doubles, GI_MEM_INTS ints
  GiMem m;
  m.GG = GG; m.Gg = D; m.J = m.Gg + 56; m.RF = m.J + 324;
  m.z = m.RF + 324; m.d = m.z + 20; m.rr = m.d + 20; m.zd = m.rr + 20; m.np = m.zd + 20;
  m.res = m.np + 20; m.viol = m.res + 56; m.scl = m.viol + 56; m.u = m.scl + 56; m.sc = m.u + 60;
  m.act = I; m.ina = I + 36; m.ign = m.ina + 56; m.status = m.ign + 56;
  return m;
}

template <class G>
QM_HDN void gi_iterate(G w0, int n, int nD0, const GiMem& gm) {
  double* J = gm.J;
  double* RF = gm.RF;
  int* act = gm.act;
  int* ina = gm.ina;
  double* u = gm.u;
  double* z = gm.z;
  double* d = gm.d;
  double* rr = gm.rr;
  double* zd = gm.zd;
  double* np = gm.np;
  double* sc = gm.sc;
  double* viol = gm.viol;
  double* scl = gm.scl;
  const double* GG = gm.GG;
  const double* Gg = gm.Gg;
  double* res = gm.res;
  int total = 0;
  int iq = 0;                      // size of the active set (uniform over the lanes)
  QM_PFOR(w0, i, 56) { ina[i] = 0; scl[i] = (i < nD0) ? 1.0 / (1.0 + fabs(Gg[i])) : 0.0; }
  w0.sync();
  for (int outer = 0; outer < 200; ++outer) {
    // constraint values c_i = gg_i - Gg_i z  (>= 0 feasible) and the scaled violation of every row that may enter
    QM_PFOR(w0, i, nD0) {
      double s = Gg[i];
      for (int c = 0; c < n; ++c) s -= GG[56 * c + i] * z[c];
      res[i] = s;
      const double v = s * scl[i];
      viol[i] = (gm.ign[i] || ina[i] || !(v < -1e-9)) ? 0.0 : v;
    }
    w0.sync();
    double worst = 0.0;
    const int ip = gi_argmin(w0, nD0, 0.0, [&](int i) { return viol[i]; }, &worst);   // most violated row; first one on ties
    (void)worst;
    QM_TICK(46);
    if (ip < 0) break;
    double cip = res[ip];
    if (w0.tid() == 0) u[iq] = 0.0;
    QM_PFOR(w0, c, n) np[c] = -GG[56 * c + ip];
    w0.sync();
    for (int inner = 0; inner < 200; ++inner) {
      // d = J' np
      QM_PFOR(w0, c, n) { double s = 0.0; for (int i = 0; i < n; ++i) s += J[i * 18 + c] * np[i]; d[c] = s; }
      w0.sync(); QM_TICK(47);
      // zd = J[:, iq:] d[iq:],  rr = RF^-1 d[:iq]
      QM_PFOR(w0, i, n) { double s = 0.0; for (int c = iq; c < n; ++c) s += J[i * 18 + c] * d[c]; zd[i] = s; }
      gi_backsub(w0, RF, d, iq, rr);
      w0.sync(); QM_TICK(48);
      // step lengths: t1 = largest dual step that keeps the multipliers non-negative (and the constraint l it blocks on),
      //               t2 = full primal step onto the new constraint
      double t1, dall, dn2;
      const int l = gi_argmin_sum2(w0, iq, 1e300,
                                   [&](int k) { return (rr[k] > 1e-14 * (1.0 + fabs(u[k]))) ? u[k] / rr[k] : 1e300; }, &t1, n,
                                   [&](int c) { return d[c] * d[c]; }, [&](int c) { return (c >= iq) ? d[c] * d[c] : 0.0; },
                                   &dall, &dn2);
      double t2 = 1e300;
      if (dn2 > 1e-26 * dall) t2 = -cip / dn2;          // z' np = |d2|^2 in the J-scaled metric
      const double t = (t1 < t2) ? t1 : t2;
      QM_TICK(49);
      if (t >= 1e300) {
        // dependent normal and nothing to drop: infeasible up to rounding -> ignore a marginally violated row
        if (w0.tid() == 0) {
          if (!(fabs(cip) < 1e-6 * (1.0 + fabs(Gg[ip])))) *gm.status |= WST_DEGENERATE;
          gm.ign[ip] = 1;
        }
        w0.sync();
        break;
      }
      if (t2 < 1e300) QM_PFOR(w0, i, n) z[i] += t * zd[i];
      QM_PFOR(w0, k, iq) u[k] -= t * rr[k];
      if (w0.tid() == 0) u[iq] += t;
      w0.sync(); QM_TICK(50);
      if (t == t2) {
        // add constraint ip: Givens rotations that zero d[iq+1..n-1] bottom up. The value a rotation leaves in d[j-1] is the norm
        // of the tail d[j-1..], so all coefficients follow from the suffix sums of squares, formed by the lanes in parallel
        // (the sequential form is a chain of n - iq - 1 hypot calls on one lane).
        QM_PFOR(w0, j, n) {
          if (j >= iq) { double sfx = 0.0; for (int c = n - 1; c >= j; --c) sfx += d[c] * d[c]; rr[j] = sfx; }
        }
        w0.sync();
        QM_PFOR(w0, j, n) {
          if (j > iq) {
            // value at position j when its rotation is formed: the tail norm, or d[j] itself if nothing below it was non-zero
            const double b2 = (j + 1 < n && rr[j + 1] != 0.0) ? sqrt(rr[j]) : d[j];
            double cs = 1.0, sn = 0.0;
            if (b2 != 0.0) { const double h = sqrt(rr[j - 1]); cs = d[j - 1] / h; sn = b2 / h; }
            sc[GI_CS + j] = cs; sc[GI_SN + j] = sn;
          }
        }
        w0.sync();
        QM_PFOR(w0, i, n) {             // row i of J: columns n-1 .. iq, the running right-hand element stays in a register
          double x2 = J[i * 18 + n - 1];
          for (int j = n - 1; j > iq; --j) {
            const double cs = sc[GI_CS + j], sn = sc[GI_SN + j];
            const double x1 = J[i * 18 + j - 1];
            J[i * 18 + j] = -sn * x1 + cs * x2;
            x2 = cs * x1 + sn * x2;
          }
          J[i * 18 + iq] = x2;
        }
        QM_PFOR(w0, i, iq) RF[i * 18 + iq] = d[i];
        if (w0.tid() == 0) {
          RF[iq * 18 + iq] = (iq + 1 < n && rr[iq + 1] != 0.0) ? sqrt(rr[iq]) : d[iq];
          act[iq] = ip; ina[ip] = 1;
        }
        ++iq;
        w0.sync(); QM_TICK(51);
        break;
      }
      // drop constraint l and continue with the same ip: column l leaves R, Givens rotations of rows (k, k+1) restore the triangle
      const int q2 = iq - 1;
#if defined(__CUDA_ARCH__)
      {
        const int lane = w0.tid();
        // bookkeeping: read, then write (lanes k >= l take the entry of k + 1)
        const int a_n = (lane >= l && lane < q2) ? act[lane + 1] : 0;
        const double u_n = (lane >= l && lane < iq) ? u[lane + 1] : 0.0;
        if (lane == 0) ina[act[l]] = 0;
        // rows shift left by one from column l on (lane = row)
        if (lane < iq) for (int k = l; k < q2; ++k) RF[lane * 18 + k] = RF[lane * 18 + k + 1];
        __syncwarp();
        if (lane >= l && lane < q2) act[lane] = a_n;
        if (lane >= l && lane < iq) u[lane] = u_n;
        // rotations (lane = column): the coefficients of step k come from column k, whose running element is in that lane
        const bool mine = (lane >= l && lane < q2);
        double carry = mine ? RF[l * 18 + lane] : 0.0;
        for (int k = l; k < q2; ++k) {
          const bool on = mine && lane >= k;
          const double x2 = on ? RF[(k + 1) * 18 + lane] : 0.0;
          double cs = 1.0, sn = 0.0;
          if (lane == k && x2 != 0.0) { const double h = hypot(carry, x2); cs = carry / h; sn = x2 / h; }
          cs = __shfl_sync(0xffffffffu, cs, k);
          sn = __shfl_sync(0xffffffffu, sn, k);
          if (on) {                                       // (cs, sn) = (1, 0) where the host skips the rotation: same values
            RF[k * 18 + lane] = cs * carry + sn * x2;
            carry = -sn * carry + cs * x2;
            if (lane == k) RF[(k + 1) * 18 + lane] = carry;
          }
          if (lane == k) { sc[GI_CS + k] = cs; sc[GI_SN + k] = sn; }
        }
        __syncwarp();
      }
#else
      {
        ina[act[l]] = 0;
        for (int k = l; k < iq - 1; ++k) {
          act[k] = act[k + 1]; u[k] = u[k + 1];
          for (int i = 0; i <= k + 1; ++i) RF[i * 18 + k] = RF[i * 18 + k + 1];
        }
        u[iq - 1] = u[iq];
        for (int k = l; k < q2; ++k) {   // restore the triangle: rotate rows k, k+1 of RF
          const double a = RF[k * 18 + k], b2 = RF[(k + 1) * 18 + k];
          double cs = 1.0, sn = 0.0;
          if (b2 != 0.0) {
            const double h = hypot(a, b2);
            cs = a / h; sn = b2 / h;
            for (int c = k; c < q2; ++c) {
              const double x1 = RF[k * 18 + c], x2 = RF[(k + 1) * 18 + c];
              RF[k * 18 + c] = cs * x1 + sn * x2;
              RF[(k + 1) * 18 + c] = -sn * x1 + cs * x2;
            }
          }
          sc[GI_CS + k] = cs; sc[GI_SN + k] = sn;
        }
      }
#endif
      iq = q2;
      QM_PFOR(w0, i, n) {               // same rotations on the columns k, k+1 of J, row-parallel, running element in a register
        double x1 = J[i * 18 + l];
        for (int k = l; k < q2; ++k) {
          const double cs = sc[GI_CS + k], sn = sc[GI_SN + k];
          const double x2 = J[i * 18 + k + 1];
          J[i * 18 + k] = cs * x1 + sn * x2;
          x1 = -sn * x1 + cs * x2;
        }
        J[i * 18 + q2] = x1;
      }
      cip = Gg[ip];
      for (int c = 0; c < n; ++c) cip -= GG[56 * c + ip] * z[c];
      w0.sync(); QM_TICK(52);
      if (++total > 400) { if (w0.tid() == 0) *gm.status |= WST_QP_MAX_ITER; break; }
    }
    if (total > 400) break;
  }
  w0.sync();
}

template <class G>
QM_HDN void wbc_gi_prepare(G g, int n, int r, double* W, int* WI) {
  const int ld = n + 1;
  double* QR = W + WS_LS;
  double* J = W + WS_J;
  const int m = r + n;
  QM_PFOR(g, idx, m * ld) {
    const int i = idx / ld, c = idx % ld;
    double v;
    if (i < r) v = (c < n) ? W[WS_GA + 18 * i + c] : W[WS_GB + i];
    else v = (c == i - r) ? 1e-6 : 0.0;
    QR[idx] = v;
  }
  g.sync();
  if (g.narrow_active()) {
    householder_ls_narrow(g.narrow(), QR, m, n, ld, W + WS_HP, r);
    back_substitute(g.narrow(), QR, n, ld, W + WS_Z);
  }
  // J = R^-1 (upper triangular), column by column
  QM_PFOR(g, c, n) {
    for (int i = n - 1; i >= 0; --i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int j = i + 1; j <= c; ++j) s -= QR[i * ld + j] * J[j * 18 + c];
      J[i * 18 + c] = (i <= c) ? s / QR[i * ld + i] : 0.0;
    }
  }
  g.sync();
  if (g.tid() == 0) { WI[WI_SC + 4] = 0; WI[WI_SC + 10] = 0; }
  QM_PFOR(g, i, 56) WI[WI_IGN + i] = 0;
  g.sync();
  QM_TICK(40);
}

// HierarchicalWbc::update after the task stack is in W, as four resumable pieces around the active-set iteration of a level, so
// that the iteration (a dependency chain on one warp) can also run as its own kernel with many solves per SM:
//   wbc_solve_begin   level 0 and its kernel basis
//   wbc_solve_prepare skips empty levels; for the next level with rows: A Z, D0 Z, least-squares start, J  -> true (iteration due)
//   [gi_iterate]
//   wbc_solve_advance x += Z z, kernel basis of the level, next stacked basis
//   wbc_solve_finish  torque recovery, cmd[54], status
// D0 (the inequality rows of level 0, [56][36]) and GG (their products with the current basis, written by wbc_solve_prepare for the
// iteration) are reached through their own pointers: the workspace block WW_D0 in the single-kernel
// solve and on the host, the solve's image in global memory in the kernel sequence (the largest block, read a few times).
// Loop state of a solve in WI_SC: [15] level, [16] columns of the current basis, [18] WSS_* (what the solve waits for).
enum { WSS_NONE = 0, WSS_ITERATION = 1, WSS_DONE = 2, WSS_LEVEL0_WIDE = 3 };   // WIDE: level 0 needs the full-size matrix
template <class G>
QM_HDN void wbc_solve_begin(G g, double* W, const double* D0, const double* Wc, int* WI, double* levels = nullptr,
                            bool level0_done = false) {
  // ---- level 0 (level0_done: the stand-alone kernel k_wbc_level0 has left x, the slack and the status word)
  if (!level0_done) {
    if (g.tid() == 0) WI[WI_SC + 6] = 0;
    g.sync();
    wbc_level0(g, l0_mem_of(W), D0, Wc, WI);
  }
  QM_TICK(-1);
  // Z0 = kernel(A0) in the reference's own (FullPivLU) basis; it has 36 - rank(A0) columns, at most 18 are kept (rank(A0) = 18
  // unless the contact Jacobians are degenerate, which is flagged)
  QM_PFOR(g, idx, 36 * 18) W[WW_Z0 + idx] = 0.0;
  g.sync();
  if (g.narrow_active())
    kernel_basis_lu(g.narrow(), Wc + WC_A0, 18, 36, 36, W + WS_QR, W + WW_Z0, 18, 18, W + WS_KCN, WI + WI_PERM, WI + WI_SC + 1, WI + WI_SC + 6);
  g.sync();
  int n1 = 36 - WI[WI_SC + 1];
  if (n1 > 18) { n1 = 18; if (g.tid() == 0) WI[WI_SC + 6] |= WST_DEGENERATE; }
  QM_TICK(38);
  if (levels != nullptr) {
    QM_PFOR(g, i, WBL_SIZE) levels[i] = 0.0;
    g.sync();
    if (g.tid() == 0) { levels[WBL_N] = (double)n1; levels[WBL_NLEV] = (double)WI[WI_LV]; }
    QM_PFOR(g, i, 36) levels[WBL_X + i] = W[WW_X + i];
    QM_PFOR(g, i, 36 * 18) levels[WBL_Z + i] = W[WW_Z0 + i];
    QM_PFOR(g, i, 56) levels[WBL_V0 + i] = W[WW_V0 + i];
  }
  g.sync();
  if (g.tid() == 0) { WI[WI_SC + 15] = 1; WI[WI_SC + 16] = n1; WI[WI_SC + 17] = 0; WI[WI_SC + 18] = WSS_NONE; }
  g.sync();
}

// ---- levels 1 .. nlev - 1, each in the coordinates x = x_prev + Z z of the null space the levels above leave
//      (HoQp.cpp:12-158: the same construction for every level; an empty level -- the swing level of the six-level stack in
//      full stance -- changes nothing)
template <class G>
QM_HDN bool wbc_solve_prepare(G g, double* W, const double* D0, double* GG, const double* Wc, int* WI, double* levels = nullptr,
                              bool stage = false) {
  // stage: A_p and D0 live in global memory (kernel sequence): they pass through the blocks J .. LS of the workspace, which are
  // dead until the least-squares start below, in pieces of at most WS_STAGE_ROWS rows -- the same products from shared memory
  const int nD0 = WI[WI_SC + 9];
  const int nlev = WI[WI_LV];
  const int n = WI[WI_SC + 16];
  const double* Zc = W + WW_Z0;             // current basis [36][18]
  int p = WI[WI_SC + 15];
  g.sync();                                   // everybody has read the loop state before it is rewritten
  for (; p < nlev; ++p) {
    const int off = WI[WI_LV + 2 * p], r = WI[WI_LV + 2 * p + 1];
    if (r == 0 || n == 0) {                  // empty level, or no freedom left: x and the basis stay as they are
      if (levels != nullptr) {
        double* L = levels + WBL_LEVEL * p;
        const int nz = (p + 1 < nlev) ? n : 0;
        if (g.tid() == 0) L[WBL_N] = (double)nz;
        QM_PFOR(g, i, 36) L[WBL_X + i] = W[WW_X + i];
        QM_PFOR(g, idx, 36 * 18) L[WBL_Z + idx] = (idx % 18 < nz) ? Zc[idx] : 0.0;
        g.sync();
      }
      continue;
    }
    const double* Ap = Wc + WC_AP + 36 * off;
    const double* bpv = Wc + WC_BP + off;
    // A Z and D0 Z as tile products (FP64 tensor-core tiles on the device); columns >= n are never read
    if (stage) {
      double* st = W + WS_J;
      QM_PFOR(g, idx, r * 36) st[idx] = Ap[idx];
      g.sync();
      mm<3, false>(g, r, n, 36, st, 36, Zc, 18, (const double*)nullptr, 0, 1.0, W + WS_GA, 18);
      QM_PFOR(g, i, r) {
        double s = bpv[i];
        for (int k = 0; k < 36; ++k) s -= st[36 * i + k] * W[WW_X + k];
        W[WS_GB + i] = s;
      }
      g.sync();
      for (int r0 = 0; r0 < nD0; r0 += WS_STAGE_ROWS) {
        const int rows = (nD0 - r0 < WS_STAGE_ROWS) ? nD0 - r0 : WS_STAGE_ROWS;
        QM_PFOR(g, idx, rows * 36) st[idx] = D0[36 * r0 + idx];
        g.sync();
        mm<3, false>(g, rows, n, 36, st, 36, Zc, 18, (const double*)nullptr, 0, 1.0, GG + 18 * r0, 18);
        QM_PFOR(g, ii, rows) {
          const int i = r0 + ii;
          double s = W[WW_F0 + i] + W[WW_V0 + i];
          for (int k = 0; k < 36; ++k) s -= st[36 * ii + k] * W[WW_X + k];
          W[WS_Gg + i] = s;
        }
        g.sync();
      }
    } else {
      mm<3, false>(g, r, n, 36, Ap, 36, Zc, 18, (const double*)nullptr, 0, 1.0, W + WS_GA, 18);
      mm<3, false>(g, nD0, n, 36, D0, 36, Zc, 18, (const double*)nullptr, 0, 1.0, GG, 18);
      QM_PFOR(g, i, r) {
        double s = bpv[i];
        for (int k = 0; k < 36; ++k) s -= Ap[36 * i + k] * W[WW_X + k];
        W[WS_GB + i] = s;
      }
      QM_PFOR(g, i, nD0) {
        double s = W[WW_F0 + i] + W[WW_V0 + i];
        for (int k = 0; k < 36; ++k) s -= D0[36 * i + k] * W[WW_X + k];
        W[WS_Gg + i] = s;
      }
      g.sync();
    }
    {
      // D0 Z from [row][18] to [column][56] in place (the iteration reads it by column): every element is held in a register
      // across the barrier
#if defined(__CUDA_ARCH__)
      double v[16];                              // covers groups of 64 threads and more
      QM_UNROLL
      for (int q = 0; q < 16; ++q) { const int idx = g.tid() + q * g.nt(); v[q] = (idx < 56 * 18) ? GG[idx] : 0.0; }
      g.sync();
      QM_UNROLL
      for (int q = 0; q < 16; ++q) {
        const int idx = g.tid() + q * g.nt();
        if (idx < 56 * 18) { const int i = idx / 18, c = idx - 18 * i; GG[56 * c + i] = v[q]; }
      }
#else
      double v[56 * 18];
      for (int idx = 0; idx < 56 * 18; ++idx) v[idx] = GG[idx];
      for (int idx = 0; idx < 56 * 18; ++idx) GG[56 * (idx % 18) + idx / 18] = v[idx];
#endif
    }
    g.sync(); QM_TICK(39);
    wbc_gi_prepare(g, n, r, W, WI);
    break;
  }
  const bool due = p < nlev;
  if (g.tid() == 0) { WI[WI_SC + 15] = p; WI[WI_SC + 18] = due ? WSS_ITERATION : WSS_NONE; }
  g.sync();
  return due;
}

template <class G>
QM_HDN void wbc_solve_advance(G g, double* W, const double* Wc, int* WI, double* levels = nullptr) {
  const int nlev = WI[WI_LV];
  const int p = WI[WI_SC + 15];
  int n = WI[WI_SC + 16];
  const int r = WI[WI_LV + 2 * p + 1];
  double* Zc = W + WW_Z0;
  g.sync();
  QM_PFOR(g, k, 36) {
    double s = 0.0;
    for (int c = 0; c < n; ++c) s += Zc[18 * k + c] * W[WS_Z + c];
    W[WW_X + k] += s;
  }
  g.sync();
  if (p + 1 < nlev) {
    // kernel of A_p Z (FullPivLU basis) -> Z <- Z N, in place: a row of the product only needs the same row of Z, which its
    // thread holds in registers
    if (g.narrow_active())
      kernel_basis_lu(g.narrow(), W + WS_GA, r, n, 18, W + WS_LS, W + WS_J, 18, 18, W + WS_KCN, WI + WI_PERM, WI + WI_SC + 1, WI + WI_SC + 6);
    g.sync();
    const int nn = n - WI[WI_SC + 1];
    QM_PFOR(g, i, 36) {
      double zr[18];
      QM_UNROLL
      for (int k = 0; k < 18; ++k) zr[k] = (k < n) ? Zc[18 * i + k] : 0.0;
      for (int c = 0; c < 18; ++c) {
        double s = 0.0;
        if (c < nn) {
          QM_UNROLL
          for (int k = 0; k < 18; ++k) if (k < n) s += zr[k] * W[WS_J + 18 * k + c];
        }
        Zc[18 * i + c] = s;
      }
    }
    g.sync(); QM_TICK(42);
    n = nn;
  }
  if (levels != nullptr) {
    double* L = levels + WBL_LEVEL * p;
    const int nz = (p + 1 < nlev) ? n : 0;           // the last level's null space is never formed
    if (g.tid() == 0) L[WBL_N] = (double)nz;
    QM_PFOR(g, i, 36) L[WBL_X + i] = W[WW_X + i];
    QM_PFOR(g, idx, 36 * 18) L[WBL_Z + idx] = (idx % 18 < nz) ? Zc[idx] : 0.0;
  }
  g.sync();
  if (g.tid() == 0) { WI[WI_SC + 15] = p + 1; WI[WI_SC + 16] = n; WI[WI_SC + 18] = WSS_NONE; }
  g.sync();
}

// ---- torque recovery: tau = [M_j, -J_j'] x + h_j   (rows 0..17 of D0)
template <class G>
QM_HDN void wbc_solve_finish(G g, double* W, const double* D0, int* WI, double* cmd, int* status) {
  QM_PFOR(g, i, 54) {
    double v;
    if (i < 36) v = W[WW_X + i];
    else {
      const int l = i - 36;
      v = W[WW_HJ + l];
      for (int c = 0; c < 36; ++c) v += D0[36 * l + c] * W[WW_X + c];
    }
    cmd[i] = v;
  }
  if (g.tid() == 0) {
    int st = WI[WI_SC + 6];
    for (int i = 0; i < 36; ++i) if (!(W[WW_X + i] == W[WW_X + i])) st |= WST_NAN;
    *status = st;
    WI[WI_SC + 18] = WSS_DONE;
  }
  g.sync();
}

// The pieces in sequence on one group. The active-set iteration is a dependency chain: it runs on the narrow group (one warp)
// with lane-parallel vector operations and warp-wide decisions; the rest of the CTA waits.
template <class G>
QM_HDN void wbc_solve(G g, double* W, const double* Wc, int* WI, double* cmd, int* status, double* levels = nullptr) {
  const double* D0 = W + WW_D0;
  wbc_solve_begin(g, W, D0, Wc, WI, levels);
  while (wbc_solve_prepare(g, W, D0, W + WS_GG, Wc, WI, levels)) {
    if (g.narrow_active()) gi_iterate(g.narrow(), WI[WI_SC + 16], WI[WI_SC + 9], gi_mem_of(W, WI));
    g.sync(); QM_TICK(41);
    wbc_solve_advance(g, W, Wc, WI, levels);
  }
  wbc_solve_finish(g, W, D0, WI, cmd, status);
}

// One whole-body-control solve: WbcBase::update + HierarchicalWbc::update.
// levels (optional, WBL_SIZE doubles): what the reference's HoQp objects expose per level (qm_wbc/include/qm_wbc/HoQp.h:21-36):
// getSolutions() and getStackedZMatrix() after every level, getStackedSlackSolutions() of level 0 (the only level with
// inequality rows in the reference's stacks). See WBL_* for the layout.
template <class G>
QM_HDN void wbc_update(G g, const qmb200_model_desc& M, const qmb200_wbc_desc& C, const double* xd, const double* ud,
                       const double* rbd, int mode, double period, double time, const double* u_last, double* W, double* Wc, int* WI,
                       double* cmd, int* status, double* levels = nullptr) {
  QM_TICK(-1);
  wbc_dynamics(g, M, C, rbd, xd, ud, u_last, period, W);
  QM_TICK(33);
  wbc_tasks(g, M, C, ud, mode & 15, time, W, W, W + WW_D0, Wc, WI);
  QM_TICK(34);
  wbc_solve(g, W, Wc, WI, cmd, status, levels);
  QM_TICK(45);
}

}  // namespace qm
