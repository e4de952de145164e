// Rigid-body forward-dynamics step behind the simulated actuator. The reference closes the loop
// controller -> actuator -> robot with Gazebo (qm_gazebo/src/QMHWSim.cpp:98-114 hands the joint efforts to the physics engine);
// for batched, device-resident disturbance studies this is one explicit step of the articulated-body equations with the stance
// feet held by bilateral point contacts -- a labelled stand-in, not Gazebo's contact model (see oracle/sim.py for the equations):
//   [ M  -Jc' ] [ qdd ]   [ S' tau - h              ]         v+ = v + dt qdd,  q+ = q + dt v+
//   [ Jc   0  ] [ f   ] = [ -dJc v - (beta/dt) Jc v ]
// Same phase-structured style and the same rigid-body routines as the whole-body controller (wbc_dynamics_measured); the KKT
// system (24 + 3 nc unknowns) is solved with the Householder routine of the controller. One thread group per problem.
#pragma once
#include "qm_mpc.h"
#include "qm_wbc.h"

namespace qm {

enum { FD_LD = 37 };                         // leading dimension of the KKT matrix | right-hand side (at most 36 unknowns)
enum { FDW_K = WW_D0, FDW_SOL = WW_Z0, FDW_QN = WW_Z0 + 40 };   // storage in the (unused) task / basis area of the controller's workspace
static_assert(36 * FD_LD <= 56 * 36, "KKT matrix fits in the D0 block");

// rbd[55] state, tau[18] joint torques, mode (stance mask), dt, beta (velocity stabilisation of the contact constraint, 0..1)
// -> rbd_out[55] next state (end-effector pose from the forward kinematics of the new configuration), f_out[12] contact forces
// (zero for swing feet), status: WST_NAN if the solution is not finite.
template <class G>
QM_HDN void fd_step(G g, const qmb200_model_desc& M, double gravity, const double* rbd, const double* tau, int mode, double dt,
                    double beta, double* W, double* rbd_out, double* f_out, int* status) {
  wbc_dynamics_measured(g, M, gravity, rbd, W);
  const double* ms = W + WA_MEAS;            // q[24], v[24]
  double* K = W + FDW_K;
  double* sol = W + FDW_SOL;
  double* qn = W + FDW_QN;                   // q+[24], v+[24]
  mode &= 15;
  const int nc = ((mode >> 3) & 1) + ((mode >> 2) & 1) + ((mode >> 1) & 1) + (mode & 1);
  const int n = 24 + 3 * nc;
  // stance rows -> (foot, component)
  auto stance_row = [&](int rr) {
    int j = rr / 3, d = rr - 3 * j, ft = 0, cnt = 0;
    for (int f2 = 0; f2 < 4; ++f2) if ((mode >> (3 - f2)) & 1) { if (cnt == j) { ft = f2; break; } ++cnt; }
    return 3 * ft + d;
  };
  QM_PFOR2(g, r, n, c, n + 1) {
    double v = 0.0;
    if (r < 24) {
      if (c < 24) v = W[WA_M + 24 * r + c];
      else if (c < n) v = -W[WA_JF + stance_row(c - 24) * 24 + r];
      else v = ((r >= 6) ? tau[r - 6] : 0.0) - W[WA_NLE + r];
    } else {
      const int fr = stance_row(r - 24);
      if (c < 24) v = W[WA_JF + fr * 24 + c];
      else if (c == n) {
        double jv = 0.0;
        for (int k = 0; k < 24; ++k) jv += W[WA_JF + fr * 24 + k] * ms[24 + k];
        v = -W[WA_DJV + fr] - (beta / dt) * jv;
      }
    }
    K[r * FD_LD + c] = v;
  }
  g.sync();
  householder_ls(g, K, n, n, FD_LD, W + WS_VH, W + WS_WJ, W + WS_HP, n);
  if (g.narrow_active()) back_substitute(g.narrow(), K, n, FD_LD, sol);
  g.sync();
  QM_PFOR(g, k, 24) {
    const double vn = ms[24 + k] + dt * sol[k];
    qn[24 + k] = vn;
    qn[k] = ms[k] + dt * vn;
  }
  QM_PFOR(g, i, 12) {
    double v = 0.0;
    const int ft = i / 3;
    if ((mode >> (3 - ft)) & 1) {
      int before = 0;
      for (int f2 = 0; f2 < ft; ++f2) before += (mode >> (3 - f2)) & 1;
      v = sol[24 + 3 * before + i % 3];
    }
    f_out[i] = v;
  }
  g.sync();
  kin_positions(g, M, qn, W + WA_KIN, false);          // frames of the new configuration (end-effector pose of the estimator layout)
  if (g.tid() == 0) {
    const double* kw = W + WA_KIN;
    for (int k = 0; k < 3; ++k) { rbd_out[k] = qn[3 + k]; rbd_out[3 + k] = qn[k]; rbd_out[27 + k] = qn[24 + k]; }
    double sz, cz, sy, cy;
    sincos(qn[3], &sz, &cz); sincos(qn[4], &sy, &cy);
    const double dz = qn[24 + 3], dy = qn[24 + 4], dx = qn[24 + 5];
    rbd_out[24] = -sz * dy + cy * cz * dx;             // world angular velocity = T(zyx) d/dt[yaw, pitch, roll]
    rbd_out[25] = cz * dy + cy * sz * dx;
    rbd_out[26] = dz - sy * dx;
    for (int r = 0; r < 3; ++r) rbd_out[48 + r] = kw[KW_EEP + r];
    quat_from_matrix(kw + KW_EER, rbd_out + 51);
    int st = 0;
    for (int k = 0; k < n; ++k) if (!(sol[k] == sol[k]) || fabs(sol[k]) > 1e300) st = WST_NAN;
    *status = st;
  }
  QM_PFOR(g, k, 18) { rbd_out[6 + k] = qn[6 + k]; rbd_out[30 + k] = qn[24 + 6 + k]; }
  g.sync();
}

}  // namespace qm
