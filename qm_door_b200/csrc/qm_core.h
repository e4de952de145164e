// Per-node numerics of the MPC hot path, written once as bulk-synchronous phases over a thread group.
//   BlockGroup : all threads of a CTA, phases separated by __syncthreads()   (transcription, Riccati)
//   WarpGroup  : one warp, phases separated by __syncwarp()                  (line-search evaluation)
//   SerialGroup: one host thread executing every phase's iterations in order (CPU port / unit tests)
// Communication between phases goes through the workspace arrays only (shared memory on the device),
// so the same source is the CUDA kernel body and its scalar CPU restatement.
//
// Reference computations replaced (relative to /root/reference):
//   kin_eval            QMPreComputation::request                qm_interface/src/QMPreComputation.cpp:73-88
//   flow_rows           QMDynamicsAD::linearApproximation        qm_interface/src/dynamics/QMDynamicsAD.cpp:22-33
//   constraint rows     NormalVelocityConstraintCppAd / zero velocity / zero force
//                       qm_interface/src/constraint/NormalVelocityConstraintCppAd.cpp:37-66, QMInterface.cpp:116-131,324-339
//   ee terms            EndEffectorConstraint                    qm_interface/src/constraint/EndEffectorConstraint.cpp:36-113
//   cost                LeggedRobotStateInputQuadraticCost       qm_interface/include/qm_interface/cost/LeggedRobotQuadraticTrackingCost.h:34-40
//   barriers            QMInterface.cpp:177-259 (arm limits), :344-358 (friction cone)
//   RK2 / projection / Riccati / line search: [upstream] ocs2_sqp, driven from QMController.cpp:288-289,323
#pragma once
#include <math.h>
#include "qm_types.h"

#if defined(__CUDACC__)
#define QM_HD __host__ __device__ __forceinline__
#define QM_HDN __host__ __device__
// one copy per kernel of the large routines that are called from several places (rigid-body passes, Householder triangularisation,
// kernel basis): k_wbc was 340 KB of straight-line code with a 79 % instruction-cache hit rate
#define QM_HDO __host__ __device__ __noinline__
#else
#define QM_HD inline
#define QM_HDN inline
#define QM_HDO inline
#endif

namespace qm {

// Development aid (-DQM_PHASE_TIMING): thread 0 of CTA 0 accumulates the cycles between consecutive ticks per phase id.
#if defined(QM_PHASE_TIMING) && defined(__CUDACC__)
__device__ unsigned long long qm_dbg[64];
#endif
#if defined(QM_PHASE_TIMING) && defined(__CUDA_ARCH__)
// the previous stamp lives in shared memory and the accumulation is a fire-and-forget reduction, so a tick costs one shared-memory
// round trip (a global read-modify-write would put an L2 latency on the chain being measured)
__device__ __forceinline__ unsigned long long* qm_tick_last() { __shared__ unsigned long long last; return &last; }
#define QM_TICK(id)                                                                    \
  do {                                                                                 \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {                      \
      const unsigned long long t_ = clock64();                                         \
      if ((id) >= 0) atomicAdd(&qm_dbg[(id)], t_ - *qm_tick_last());                   \
      *qm_tick_last() = t_;                                                            \
    }                                                                                  \
  } while (0)
#else
#define QM_TICK(id) do {} while (0)
#endif

// ------------------------------------------------------------------------------------------ groups
// A group can split into a `narrow` part (dependency-chain work: kinematics tree, pivoting) and the `rest` (wide
// independent work) that run concurrently between two full-group syncs.
struct SerialGroup {
  QM_HD int tid() const { return 0; }
  QM_HD int nt() const { return 1; }
  QM_HD void sync() const {}
  QM_HD bool narrow_active() const { return true; }
  QM_HD bool rest_active() const { return true; }
  QM_HD SerialGroup narrow() const { return *this; }
  QM_HD SerialGroup rest() const { return *this; }
  QM_HD void rest_signal() const {}      // rest -> narrow hand-over inside a split phase (see BlockGroup)
  QM_HD void narrow_wait() const {}
};
#if defined(__CUDACC__)
struct WarpGroup {
  __device__ __forceinline__ int tid() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int nt() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ bool narrow_active() const { return true; }
  __device__ __forceinline__ bool rest_active() const { return true; }
  __device__ __forceinline__ WarpGroup narrow() const { return *this; }
  __device__ __forceinline__ WarpGroup rest() const { return *this; }
  __device__ __forceinline__ void rest_signal() const { __syncwarp(); }
  __device__ __forceinline__ void narrow_wait() const {}
  __device__ __forceinline__ int warp() const { return 0; }
  __device__ __forceinline__ int nwarps() const { return 1; }
};
// all warps of the CTA but the narrow one, synchronised with named barrier 1
struct RestGroup {
  int nwid;                                      // the narrow warp (excluded)
  __device__ __forceinline__ int warp() const { const int w = threadIdx.x >> 5; return w - (w > nwid ? 1 : 0); }
  __device__ __forceinline__ int tid() const { return warp() * 32 + (threadIdx.x & 31); }
  __device__ __forceinline__ int nt() const { return blockDim.x - 32; }
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x - 32) : "memory"); }
  __device__ __forceinline__ int nwarps() const { return (blockDim.x >> 5) - 1; }
};
// The narrow warp can be chosen per CTA (k_solve rotates it with the CTA index: warp w issues on sub-partition w % 4,
// so the serial chains of co-resident CTAs do not all land on the same FP64 pipe).
struct BlockGroup {
  int nwid = 0;
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int nt() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ bool narrow_active() const { return (threadIdx.x >> 5) == nwid; }
  __device__ __forceinline__ bool rest_active() const { return (threadIdx.x >> 5) != nwid; }
  __device__ __forceinline__ WarpGroup narrow() const { return WarpGroup(); }
  __device__ __forceinline__ RestGroup rest() const { return RestGroup{nwid}; }
  // hand-over inside a split phase: every rest thread signals once its part is written, the narrow warp waits (named barrier 2)
  __device__ __forceinline__ void rest_signal() const { asm volatile("bar.arrive 2, %0;" ::"r"(blockDim.x) : "memory"); }
  __device__ __forceinline__ void narrow_wait() const { asm volatile("bar.sync 2, %0;" ::"r"(blockDim.x) : "memory"); }
  __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
  __device__ __forceinline__ int nwarps() const { return blockDim.x >> 5; }
};
#endif
#if defined(__CUDA_ARCH__)
#define QM_UNROLL _Pragma("unroll")
#else
#define QM_UNROLL
#endif
#define QM_RESTRICT __restrict__
#define QM_PFOR(g, i, n) for (int i = (g).tid(); i < (n); i += (g).nt())
// two-level loop without index division: rows over the warps of the group, columns over the lanes
#define QM_PFOR2(g, i, ni, c, nc)                                              \
  for (int i = (g).tid() >> 5; i < (ni); i += ((g).nt() + 31) >> 5)            \
    for (int c = (g).tid() & 31; c < (nc); c += ((g).nt() < 32 ? (g).nt() : 32))

// several CTAs (nodes) of one problem may flag the same status word
QM_HD void status_or(int* p, int v) {
#if defined(__CUDA_ARCH__)
  atomicOr(p, v);
#else
  *p |= v;
#endif
}

QM_HD int mm_rowperm(int r) { return ((r & 3) << 1) | (r >> 2); }     // {0,2,4,6,1,3,5,7}: see the bank-conflict note at mm

// out(r, init(r) + sum_{j < len} term(r, j)) for r < nrows. Device: four lanes per row, partial sums combined with
// shuffles (the group size is a multiple of 32); host: one loop.
// TR: term(r, j) reads a matrix by columns (element j * ld + r, ld = 30 or 18): the four lanes of a row then take j in pairs
// {2 part, 2 part + 1} + 8 s and the rows stay consecutive, which puts the lanes of a half warp on 16 different bank pairs.
template <bool TR = false, class G, class FI, class FT, class FO>
QM_HDN void rows_dot(G g, int nrows, int len, FI init, FT term, FO out) {
#if defined(__CUDA_ARCH__)
  for (int base = 0; base < 4 * nrows; base += g.nt()) {
    const int t = base + g.tid(), rl = t >> 2, part = t & 3;
    const int r = TR ? rl : ((rl & ~7) | mm_rowperm(rl & 7));   // rows two apart within a half warp: no bank conflicts at row stride 30 / 18
    const bool valid = r < nrows;
    double acc = 0.0;
    if (TR) {
      if (valid) for (int j = 2 * part; j < len; j += 8) { acc += term(r, j); if (j + 1 < len) acc += term(r, j + 1); }
    } else if (valid) for (int j = part; j < len; j += 4) acc += term(r, j);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (valid && part == 0) out(r, init(r) + acc);
  }
#else
  QM_PFOR(g, r, nrows) {
    double acc = 0.0;
    for (int j = 0; j < len; ++j) acc += term(r, j);
    out(r, init(r) + acc);
  }
#endif
}

// ------------------------------------------------------------------------------------------ small vector helpers
QM_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
QM_HD void cross3_add(const double* a, const double* b, double* c) {
  c[0] += a[1] * b[2] - a[2] * b[1];
  c[1] += a[2] * b[0] - a[0] * b[2];
  c[2] += a[0] * b[1] - a[1] * b[0];
}
// symmetric 3x3 stored as (xx, xy, xz, yy, yz, zz) times vector
QM_HD void sym3_mul(const double* s, const double* v, double* o) {
  o[0] = s[0] * v[0] + s[1] * v[1] + s[2] * v[2];
  o[1] = s[1] * v[0] + s[3] * v[1] + s[4] * v[2];
  o[2] = s[2] * v[0] + s[4] * v[1] + s[5] * v[2];
}

// ------------------------------------------------------------------------------------------ small dense products
// C[m x n] = C0 + alpha * op(X) Y      op(X) = X (m x k, row-major ldx) or X' (X stored k x m) ; Y: k x n (ldy).
// Device: FP64 tensor-core tiles (mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4, 256 FMA per warp instruction) with edge
// predication instead of padding; each warp of the group owns (tile row, up to TJ tile columns) units so an X fragment is
// loaded once per k-step. Host: plain loops. C0 may alias C (each element is read and written by the same thread).
#if defined(__CUDACC__)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif
// Flags: MM_UP  : C is symmetric (xT products, m == n); only the tiles needed to cover its upper triangle are formed:
//                  tile (ti, tj) unless column group tj / 2 lies left of row group ti / 2 (16 x 16 groups, see below);
//        MM_XSYM: X is a symmetric matrix that is read from its upper triangle only (not with xT).
// Bank conflicts: the operands live in shared memory with leading dimensions 30 or 18 doubles (bank stride -4 / +4 per
// row), so a tile over consecutive rows / columns makes the 16 lanes of a half warp collide two-way on every fragment
// load. A tile therefore covers a permuted set of rows / columns (the product does not care which eight it gets):
//   * rows of a non-transposed X tile: i0 + {0,2,4,6,1,3,5,7}  (rows two apart are eight banks apart);
//   * columns of Y (and of the output), and the rows of a transposed X tile: groups of 16, two tiles per group:
//     16 g + {0,1,8,9,2,3,10,11} and 16 g + {4,5,12,13,6,7,14,15}  (the four k-rows of a step then interleave cleanly).
// Each lane still writes two adjacent output elements.
enum { MM_UP = 1, MM_XSYM = 2 };
QM_HD int mm_col16(int tile, int c) {      // matrix column of tile-column c (0..7) of tile `tile` (two tiles per 16 columns)
  const int p = c >> 1;
  return ((tile >> 1) << 4) + (((p & 1) << 3) | ((p >> 1) << 1) | ((tile & 1) << 2)) + (c & 1);
}
QM_HD int mm_tiles16(int n) {              // number of tiles with at least one column < n
  const int g = (n - 1) >> 4;
  return 2 * g + (((n - 1) - 16 * g >= 4) ? 2 : 1);
}
template <int TJ, bool xT, int FLAGS = 0, class G>
QM_HDN void mm(G g, int m, int n, int k, const double* X, int ldx, const double* Y, int ldy, const double* C0, int ld0,
               double alpha, double* C, int ldc, int rot = 0) {
#if defined(__CUDA_ARCH__)
  const int lane = threadIdx.x & 31;
  const int tm = xT ? mm_tiles16(m) : (m + 7) >> 3, tn = mm_tiles16(n);
  const int gj = (tn + TJ - 1) / TJ;
  const int r = lane >> 2, q = lane & 3;
  // `rot` (0 <= rot < nwarps) rotates the unit -> warp assignment so that products issued back to back spread over all warps
  const int nw = g.nwarps();
  int w0 = g.warp() - rot;
  if (w0 < 0) w0 += nw;
  int nunits = tm * gj;
  // the two adjacent output elements of a lane go out (and the C0 pair comes in) as one 16-byte access where the layout allows
  const bool pairc = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  const bool pair0 = C0 && ((ld0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(C0) & 15) == 0);
  if (FLAGS & MM_UP) { nunits = 0; for (int ti = 0; ti < tm; ++ti) nunits += tn - ((ti >> 1) << 1); }
  for (int unit = w0; unit < nunits; unit += nw) {
    int ti = 0, tj = unit;                                        // unit -> (tile row, tile column group) without divisions
    if (FLAGS & MM_UP) { while (tj >= tn - ((ti >> 1) << 1)) { tj -= tn - ((ti >> 1) << 1); ++ti; } tj += (ti >> 1) << 1; }
    else if (gj == 1) { ti = unit; tj = 0; }
    else if (gj == 2) { ti = unit >> 1; tj = unit & 1; }
    else if (gj == 4) { ti = unit >> 2; tj = unit & 3; }
    else { while (tj >= gj) { tj -= gj; ++ti; } }
    const int t0 = tj * TJ;                                       // first tile column of the unit
    double acc[TJ][2], c0v[TJ][2];
    const int i = xT ? mm_col16(ti, r) : (ti << 3) + mm_rowperm(r);
    // the C0 operands are requested before the product loop so that their latency (HBM blocks) overlaps the tiles
#pragma unroll
    for (int t = 0; t < TJ; ++t) {
      acc[t][0] = 0.0; acc[t][1] = 0.0;
      const int j = mm_col16(t0 + t, 2 * q);
      if (pair0 && i < m && j + 1 < n) {          // j is even: one 16-byte access (all 32 banks instead of every other pair)
        const double2 v = *reinterpret_cast<const double2*>(C0 + i * ld0 + j);
        c0v[t][0] = v.x; c0v[t][1] = v.y;
      } else {
        c0v[t][0] = (C0 && i < m && j < n) ? C0[i * ld0 + j] : 0.0;
        c0v[t][1] = (C0 && i < m && j + 1 < n) ? C0[i * ld0 + j + 1] : 0.0;
      }
    }
    const bool iok = i < m;
    const double* xp = xT ? (X + i + q * ldx) : (X + i * ldx + q);     // advances by 4 rows (xT) / 4 columns per k-step
    const int xstep = xT ? 4 * ldx : 4;
    int jb[TJ];                                                   // this lane's column of Y per tile
#pragma unroll
    for (int t = 0; t < TJ; ++t) jb[t] = mm_col16(t0 + t, r);
    const double* yp = Y + q * ldy;
    for (int k0 = 0; k0 < k; k0 += 4, xp += xstep, yp += 4 * ldy) {
      const bool kok = (k0 + q) < k;
      double a;
      if ((FLAGS & MM_XSYM) && !xT) { const int kk = k0 + q; a = (iok && kok) ? ((kk < i) ? X[kk * ldx + i] : X[i * ldx + kk]) : 0.0; }
      else a = (iok && kok) ? *xp : 0.0;
#pragma unroll
      for (int t = 0; t < TJ; ++t) {
        if (t0 + t < tn) {                        // warp-uniform
          const double b = (kok && jb[t] < n) ? yp[jb[t]] : 0.0;
          dmma884(acc[t][0], acc[t][1], a, b);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TJ; ++t) {
      const int j = mm_col16(t0 + t, 2 * q);
      if (i < m && t0 + t < tn) {
        if (pairc && j + 1 < n) {
          double2 v;
          v.x = c0v[t][0] + alpha * acc[t][0]; v.y = c0v[t][1] + alpha * acc[t][1];
          *reinterpret_cast<double2*>(C + i * ldc + j) = v;
        } else {
          if (j < n) C[i * ldc + j] = c0v[t][0] + alpha * acc[t][0];
          if (j + 1 < n) C[i * ldc + j + 1] = c0v[t][1] + alpha * acc[t][1];
        }
      }
    }
  }
#else
  QM_PFOR(g, idx, m * n) {
    const int i = idx / n, j = idx % n;
    if ((FLAGS & MM_UP) && (j >> 4) < (i >> 4)) continue;
    double acc = 0.0;
    for (int kk = 0; kk < k; ++kk) {
      double xv;
      if (xT) xv = X[kk * ldx + i];
      else if ((FLAGS & MM_XSYM) && kk < i) xv = X[kk * ldx + i];
      else xv = X[i * ldx + kk];
      acc += xv * Y[kk * ldy + j];
    }
    C[i * ldc + j] = (C0 ? C0[i * ld0 + j] : 0.0) + alpha * acc;
  }
#endif
}

// ------------------------------------------------------------------------------------------ kinematics workspace
// Offsets (in doubles) into the kinematics workspace. kin_velocities needs the full workspace (KW_SIZE).
// Regions whose lifetimes do not overlap share storage: the q-derivatives DH | DFV (written by the last phase of
// kin_velocities when deriv is requested) take the place of R | BODY (placements: dead after the body / frame phases; body
// inertias: dead once the body momenta HB are formed), and so does F (CRBA columns, read by the whole-body controller only,
// which never requests the q-derivatives).
enum {
  KW_R = 0,                          // [24][9]  world rotation of joint frames
  KW_BODY = KW_R + QM_NJ * 9,        // [24][10] per body: m, m*c[3], I0[6] (rot. inertia about world origin)
  KW_P = KW_BODY + QM_NJ * 10,       // [24][3]  world origin of joint frames
  KW_AX = KW_P + QM_NJ * 3,          // [24][3]  world joint axis
  KW_COMP = KW_AX + QM_NJ * 3,       // [24][10] same as BODY, summed over the subtree of each joint
  KW_ACM = KW_COMP + QM_NJ * 10,     // [6][24]  centroidal momentum matrix
  KW_SV = KW_ACM + 6 * QM_NJ,        // [24][6]  S_j v_j  -> reused for subtree momenta
  KW_V = KW_SV + QM_NJ * 6,          // [24][6]  spatial velocity (w, vO) of each body, world origin
  KW_FPOS = KW_V + QM_NJ * 6,        // [4][3]
  KW_FVEL = KW_FPOS + 12,            // [4][3]
  KW_EEP = KW_FVEL + 12,             // [3]
  KW_EER = KW_EEP + 3,               // [9]
  KW_COM = KW_EER + 9,               // [3]
  KW_ABINV = KW_COM + 3,             // [36]
  KW_VEL = KW_ABINV + 36,            // [24] generalized velocity
  KW_RHS = KW_VEL + QM_NJ,           // [6]
  KW_VSIZE = ((KW_RHS + 6 + 3) / 4) * 4,
  // ---- Jacobian level
  KW_FJ = KW_VSIZE,                  // [4][3][24] foot linear Jacobians
  KW_EEJ = KW_FJ + 12 * QM_NJ,       // [6][24] ee Jacobian [linear; angular]
  KW_SIZE = ((KW_EEJ + 6 * QM_NJ + 3) / 4) * 4,
  // ---- aliases (see above)
  KW_F = KW_R,                       // [24][6] composite momentum per unit joint rate I^c_j S_j = (L0, p)  (CRBA columns)
  KW_DH = KW_R,                      // [6][24]  d(A v)/dq at fixed v (centroidal)
  KW_DFV = KW_DH + 6 * QM_NJ,        // [4][3][24] d(J_i v)/dq at fixed v
  KW_HB = KW_EEJ                     // [24][6]  body momentum (L0, p): lives from kin_velocities' second phase to its subtree sums,
                                     //          in the place of the end-effector Jacobian, whose readers (ee_terms, the whole-body
                                     //          controller's copy) run between kin_positions and kin_velocities
};
static_assert(KW_DFV + 12 * QM_NJ <= KW_P, "DH | DFV must fit in R | BODY");

// The subtree of joint j consists of joints j .. j + 5 only (true for the limb joints of a tree numbered parents first): the sums
// over a subtree then take six predicated steps instead of a sweep over all joints, in the same order.
QM_HD int count_trailing_zeros(uint32_t v) {      // v != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
QM_HD bool subtree_is_window(uint32_t submask, int j) {
  return (submask & ((1u << j) - 1u)) == 0u && (submask >> j) < 64u;
}

// spatial motion vector of joint j (world coordinates, reference point = world origin): (w, vO)
QM_HD void joint_S(const qmb200_model_desc& M, const double* w, int j, double* S) {
  const double* a = w + KW_AX + 3 * j;
  if (M.jtype[j] == 1) {
    S[0] = a[0]; S[1] = a[1]; S[2] = a[2];
    cross3(w + KW_P + 3 * j, a, S + 3);
  } else {
    S[0] = S[1] = S[2] = 0.0;
    S[3] = a[0]; S[4] = a[1]; S[5] = a[2];
  }
}
// momentum (L0, p) of inertia (m, h=m*c, I0) moving with (w, vO)
QM_HD void inertia_mul(const double* I, const double* V, double* h) {
  double t[3];
  sym3_mul(I + 4, V, h);              // I0 w
  cross3_add(I + 1, V + 3, h);        // + h x vO
  cross3(V, I + 1, t);                // w x h
  h[3] = I[0] * V[3] + t[0];
  h[4] = I[0] * V[4] + t[1];
  h[5] = I[0] * V[5] + t[2];
}

// Position level: placements, composite inertias, centroidal momentum matrix, frame positions and Jacobians. q[24].
template <class G>
QM_HDN void kin_positions(G g, const qmb200_model_desc& M, const double* q, double* w, bool jac = true) {
  // P1: placements. A standard floating base (joints 0..5) is placed in closed form by its six lanes at once
  //     (R = Rz(yaw) Ry(pitch) Rx(roll), p = base position); the remaining joints level by level.
  // One sincos per lane for all joints at once (the three Euler angles of a standard floating base included), kept in the
  // velocity-level array SV, which is not live before kin_velocities.
  double* sc = w + KW_SV;
  QM_PFOR(g, j, QM_NJ) {
    double s = 0.0, c = 1.0;
    if (M.jtype[j] == 1) sincos(q[j], &s, &c);
    sc[2 * j] = s; sc[2 * j + 1] = c;
  }
  g.sync(); QM_TICK(13);
  const int d0 = M.root6_standard ? 6 : 0;
  // local transforms of all joints at once (Rodrigues, Rp Rq), stored in the joint's own R | P | AX slots; the level-by-level
  // sweep then only composes them with the parent placement. A standard floating base (joints 0..5) is placed in closed
  // form by its six lanes: R = Rz(yaw) Ry(pitch) Rx(roll), p = base position.
  QM_PFOR(g, j, QM_NJ) {
    if (M.depth[j] < d0) {
      const double sz = sc[6], cz = sc[7], sy = sc[8], cy = sc[9], sx = sc[10], cx = sc[11];
      double* Rj = w + KW_R + 9 * j;
      double* pj = w + KW_P + 3 * j;
      double* aj = w + KW_AX + 3 * j;
      pj[0] = q[0]; pj[1] = (j >= 1) ? q[1] : 0.0; pj[2] = (j >= 2) ? q[2] : 0.0;
      if (j <= 2) {
        for (int k = 0; k < 9; ++k) Rj[k] = (k % 4 == 0) ? 1.0 : 0.0;
        aj[0] = (j == 0); aj[1] = (j == 1); aj[2] = (j == 2);
      } else if (j == 3) {
        Rj[0] = cz; Rj[1] = -sz; Rj[2] = 0; Rj[3] = sz; Rj[4] = cz; Rj[5] = 0; Rj[6] = 0; Rj[7] = 0; Rj[8] = 1;
        aj[0] = 0; aj[1] = 0; aj[2] = 1;
      } else if (j == 4) {
        Rj[0] = cz * cy; Rj[1] = -sz; Rj[2] = cz * sy; Rj[3] = sz * cy; Rj[4] = cz; Rj[5] = sz * sy; Rj[6] = -sy; Rj[7] = 0; Rj[8] = cy;
        aj[0] = -sz; aj[1] = cz; aj[2] = 0;
      } else {
        Rj[0] = cz * cy; Rj[1] = cz * sy * sx - sz * cx; Rj[2] = cz * sy * cx + sz * sx;
        Rj[3] = sz * cy; Rj[4] = sz * sy * sx + cz * cx; Rj[5] = sz * sy * cx - cz * sx;
        Rj[6] = -sy;     Rj[7] = cy * sx;                Rj[8] = cy * cx;
        aj[0] = cz * cy; aj[1] = sz * cy; aj[2] = -sy;
      }
      continue;
    }
    const double ax = M.axis[j][0], ay = M.axis[j][1], az = M.axis[j][2];
    double* Lj = w + KW_R + 9 * j;
    double* lp = w + KW_P + 3 * j;
    double* la = w + KW_AX + 3 * j;
    for (int r = 0; r < 3; ++r) la[r] = M.Rp[j][3 * r] * ax + M.Rp[j][3 * r + 1] * ay + M.Rp[j][3 * r + 2] * az;
    if (M.jtype[j] == 1) {
      const double s = sc[2 * j], c = sc[2 * j + 1];
      const double v = 1.0 - c;
      // Rodrigues: I + s K + (1-c) K^2, K = skew(axis)
      const double Rq[9] = {c + v * ax * ax,      v * ax * ay - s * az, v * ax * az + s * ay,
                            v * ax * ay + s * az, c + v * ay * ay,      v * ay * az - s * ax,
                            v * ax * az - s * ay, v * ay * az + s * ax, c + v * az * az};
      for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc)
          Lj[3 * r + cc] = M.Rp[j][3 * r] * Rq[cc] + M.Rp[j][3 * r + 1] * Rq[3 + cc] + M.Rp[j][3 * r + 2] * Rq[6 + cc];
      for (int r = 0; r < 3; ++r) lp[r] = M.pp[j][r];
    } else {
      for (int k = 0; k < 9; ++k) Lj[k] = M.Rp[j][k];
      for (int r = 0; r < 3; ++r) lp[r] = M.pp[j][r] + la[r] * q[j];
    }
  }
  g.sync(); QM_TICK(14);
  for (int d = d0; d <= M.max_depth; ++d) {
    QM_PFOR(g, j, QM_NJ) {
      if (M.depth[j] != d) continue;
      const int par = M.parent[j];
      if (par < 0) continue;                     // a root joint: its local transform is its placement
      double Rpar[9], ppar[3], L[9], l[3], a[3];
      for (int k = 0; k < 9; ++k) { Rpar[k] = w[KW_R + 9 * par + k]; L[k] = w[KW_R + 9 * j + k]; }
      for (int k = 0; k < 3; ++k) { ppar[k] = w[KW_P + 3 * par + k]; l[k] = w[KW_P + 3 * j + k]; a[k] = w[KW_AX + 3 * j + k]; }
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) w[KW_R + 9 * j + 3 * r + c] = Rpar[3 * r] * L[c] + Rpar[3 * r + 1] * L[3 + c] + Rpar[3 * r + 2] * L[6 + c];
        w[KW_P + 3 * j + r] = ppar[r] + Rpar[3 * r] * l[0] + Rpar[3 * r + 1] * l[1] + Rpar[3 * r + 2] * l[2];
        w[KW_AX + 3 * j + r] = Rpar[3 * r] * a[0] + Rpar[3 * r + 1] * a[1] + Rpar[3 * r + 2] * a[2];
      }
    }
    g.sync();
  }
  QM_TICK(15);
  // P2: body inertial quantities about the world origin
  QM_PFOR(g, j, QM_NJ) {
    const double* R = w + KW_R + 9 * j;
    const double m = M.mass[j];
    double c[3];
    for (int r = 0; r < 3; ++r)
      c[r] = w[KW_P + 3 * j + r] + R[3 * r] * M.com[j][0] + R[3 * r + 1] * M.com[j][1] + R[3 * r + 2] * M.com[j][2];
    double RI[9];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc)
        RI[3 * r + cc] = R[3 * r] * M.inertia[j][cc] + R[3 * r + 1] * M.inertia[j][3 + cc] + R[3 * r + 2] * M.inertia[j][6 + cc];
    double Iw[9];
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc)
        Iw[3 * r + cc] = RI[3 * r] * R[3 * cc] + RI[3 * r + 1] * R[3 * cc + 1] + RI[3 * r + 2] * R[3 * cc + 2];
    const double cc2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    double* b = w + KW_BODY + 10 * j;
    b[0] = m; b[1] = m * c[0]; b[2] = m * c[1]; b[3] = m * c[2];
    b[4] = Iw[0] + m * (cc2 - c[0] * c[0]);
    b[5] = Iw[1] - m * c[0] * c[1];
    b[6] = Iw[2] - m * c[0] * c[2];
    b[7] = Iw[4] + m * (cc2 - c[1] * c[1]);
    b[8] = Iw[5] - m * c[1] * c[2];
    b[9] = Iw[8] + m * (cc2 - c[2] * c[2]);
  }
  // frame positions (independent of the body phase)
  QM_PFOR(g, f, QM_NFEET + 1) {
    if (f < QM_NFEET) {
      const int b = M.foot_joint[f];
      const double* R = w + KW_R + 9 * b;
      for (int r = 0; r < 3; ++r)
        w[KW_FPOS + 3 * f + r] = w[KW_P + 3 * b + r] + R[3 * r] * M.foot_off[f][0] + R[3 * r + 1] * M.foot_off[f][1] + R[3 * r + 2] * M.foot_off[f][2];
    } else {
      const int b = M.ee_joint;
      const double* R = w + KW_R + 9 * b;
      for (int r = 0; r < 3; ++r) {
        w[KW_EEP + r] = w[KW_P + 3 * b + r] + R[3 * r] * M.ee_off[0] + R[3 * r + 1] * M.ee_off[1] + R[3 * r + 2] * M.ee_off[2];
        for (int c = 0; c < 3; ++c)
          w[KW_EER + 3 * r + c] = R[3 * r] * M.ee_Roff[c] + R[3 * r + 1] * M.ee_Roff[3 + c] + R[3 * r + 2] * M.ee_Roff[6 + c];
      }
    }
  }
  g.sync(); QM_TICK(16);
  // P3: composite (subtree) inertias. One (joint, component) pair per work item: the base joints carry the whole tree,
  //     so a lane per joint would leave the phase waiting for six lanes that sum 24 bodies each.
  QM_PFOR(g, idx, QM_NJ * 10) {
    const int j = idx / 10, k = idx - 10 * j;
    double acc = 0.0;
    const uint32_t mask = M.submask[j];
    if (subtree_is_window(mask, j)) {              // limb joints: at most six bodies, joints j .. j + 5 (same summation order)
      QM_UNROLL
      for (int d = 0; d < 6; ++d)
        if ((mask >> (j + d)) & 1u) acc += w[KW_BODY + 10 * (j + d) + k];
    } else {
      for (int i = 0; i < QM_NJ; ++i)
        if ((mask >> i) & 1u) acc += w[KW_BODY + 10 * i + k];
    }
    w[KW_COMP + idx] = acc;
  }
  g.sync();
  QM_PFOR(g, r, 3) w[KW_COM + r] = w[KW_COMP + 1 + r] / w[KW_COMP];
  g.sync(); QM_TICK(17);
  // P4: centroidal momentum matrix columns, frame Jacobian columns
  QM_PFOR(g, j, QM_NJ) {
    double S[6], h[6], t[3];
    joint_S(M, w, j, S);
    inertia_mul(w + KW_COMP + 10 * j, S, h);
    cross3(w + KW_COM, h + 3, t);
    w[KW_ACM + 0 * QM_NJ + j] = h[3];
    w[KW_ACM + 1 * QM_NJ + j] = h[4];
    w[KW_ACM + 2 * QM_NJ + j] = h[5];
    w[KW_ACM + 3 * QM_NJ + j] = h[0] - t[0];
    w[KW_ACM + 4 * QM_NJ + j] = h[1] - t[1];
    w[KW_ACM + 5 * QM_NJ + j] = h[2] - t[2];
    if (!jac) continue;                 // value-only evaluation: the workspace ends at KW_VSIZE
    for (int c = 0; c < 6; ++c) w[KW_F + 6 * j + c] = h[c];
    for (int f = 0; f < QM_NFEET + 1; ++f) {
      const int b = (f < QM_NFEET) ? M.foot_joint[f] : M.ee_joint;
      const double* pos = (f < QM_NFEET) ? (w + KW_FPOS + 3 * f) : (w + KW_EEP);
      double col[3] = {0, 0, 0}, ang[3] = {0, 0, 0};
      if ((M.pathmask[b] >> j) & 1u) {
        // velocity of the frame origin per unit joint rate: vO + w x pos
        cross3(S, pos, col);
        col[0] += S[3]; col[1] += S[4]; col[2] += S[5];
        ang[0] = S[0]; ang[1] = S[1]; ang[2] = S[2];
      }
      if (f < QM_NFEET) {
        for (int r = 0; r < 3; ++r) w[KW_FJ + (3 * f + r) * QM_NJ + j] = col[r];
      } else {
        for (int r = 0; r < 3; ++r) {
          w[KW_EEJ + r * QM_NJ + j] = col[r];
          w[KW_EEJ + (3 + r) * QM_NJ + j] = ang[r];
        }
      }
    }
  }
  g.sync(); QM_TICK(18);
}

// Generalized velocity implied by the centroidal state/input: v = [Ab^-1 (m h_n - A_j v_j); v_j]
// ([upstream] CentroidalModelPinocchioMapping::getPinocchioJointVelocity). Needs kin_positions first.
template <class G>
QM_HDN void centroidal_velocity(G g, const qmb200_model_desc& M, const double* x, const double* u, double* w) {
  QM_PFOR(g, r, 6) {
    double acc = M.total_mass * x[r];
    for (int l = 0; l < 18; ++l) acc -= w[KW_ACM + r * QM_NJ + 6 + l] * u[12 + l];
    w[KW_RHS + r] = acc;
  }
  QM_PFOR(g, l, 18) w[KW_VEL + 6 + l] = u[12 + l];
  if (g.tid() == 0) {
    // [upstream] computeFloatingBaseCentroidalMomentumMatrixInverse: Ab = [m I, Ab12; 0, Ab22]
    const double* A = w + KW_ACM;
    const double mass = A[0];
    const double a = A[3 * QM_NJ + 3], b = A[3 * QM_NJ + 4], c = A[3 * QM_NJ + 5];
    const double d = A[4 * QM_NJ + 3], e = A[4 * QM_NJ + 4], f = A[4 * QM_NJ + 5];
    const double gg = A[5 * QM_NJ + 3], hh = A[5 * QM_NJ + 4], ii = A[5 * QM_NJ + 5];
    const double det = a * (e * ii - f * hh) - b * (d * ii - f * gg) + c * (d * hh - e * gg);
    const double id = 1.0 / det;
    const double inv[9] = {(e * ii - f * hh) * id, (c * hh - b * ii) * id, (b * f - c * e) * id,
                           (f * gg - d * ii) * id, (a * ii - c * gg) * id, (c * d - a * f) * id,
                           (d * hh - e * gg) * id, (b * gg - a * hh) * id, (a * e - b * d) * id};
    double* Bi = w + KW_ABINV;
    for (int k = 0; k < 36; ++k) Bi[k] = 0.0;
    const double im = 1.0 / mass;
    for (int r = 0; r < 3; ++r) {
      Bi[6 * r + r] = im;
      for (int cc = 0; cc < 3; ++cc) {
        Bi[6 * (3 + r) + 3 + cc] = inv[3 * r + cc];
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) acc += A[r * QM_NJ + 3 + k] * inv[3 * k + cc];
        Bi[6 * r + 3 + cc] = -acc * im;
      }
    }
  }
  g.sync();
  QM_PFOR(g, r, 6) {
    double acc = 0.0;
    for (int c = 0; c < 6; ++c) acc += w[KW_ABINV + 6 * r + c] * w[KW_RHS + c];
    w[KW_VEL + r] = acc;
  }
  g.sync(); QM_TICK(19);
}

// Velocity level (KW_VEL given): spatial velocities, body momenta, foot velocities and the q-derivatives at fixed
// generalized velocity: deriv >= 1: d(A v)/dq (KW_DH); deriv == 2: d(J_i v)/dq (KW_DFV) as well.
template <class G>
QM_HDN void kin_velocities(G g, const qmb200_model_desc& M, int deriv, double* w) {
  // P7: spatial velocities and body momenta
  QM_PFOR(g, j, QM_NJ) {
    double S[6];
    joint_S(M, w, j, S);
    const double vj = w[KW_VEL + j];
    for (int k = 0; k < 6; ++k) w[KW_SV + 6 * j + k] = S[k] * vj;
  }
  g.sync();
  QM_PFOR(g, j, QM_NJ) {
    double V[6] = {0, 0, 0, 0, 0, 0};
    const uint32_t mask = M.pathmask[j];
    // ancestors = base joints (among 0..5) and at most six limb joints lo .. lo + 5: twelve predicated steps in the same order
    const uint32_t high = mask >> 6;
    const int lo = 6 + (high ? count_trailing_zeros(high) : 0);
    if ((mask >> lo) < 64u) {
      QM_UNROLL
      for (int d = 0; d < 6; ++d)
        if ((mask >> d) & 1u)
          for (int c = 0; c < 6; ++c) V[c] += w[KW_SV + 6 * d + c];
      QM_UNROLL
      for (int d = 0; d < 6; ++d)
        if ((mask >> (lo + d)) & 1u)
          for (int c = 0; c < 6; ++c) V[c] += w[KW_SV + 6 * (lo + d) + c];
    } else {
      for (int k = 0; k < QM_NJ; ++k)
        if ((mask >> k) & 1u)
          for (int c = 0; c < 6; ++c) V[c] += w[KW_SV + 6 * k + c];
    }
    for (int c = 0; c < 6; ++c) w[KW_V + 6 * j + c] = V[c];
    inertia_mul(w + KW_BODY + 10 * j, V, w + KW_HB + 6 * j);
  }
  g.sync();
  // foot velocities
  QM_PFOR(g, f, QM_NFEET) {
    const double* V = w + KW_V + 6 * M.foot_joint[f];
    double t[3];
    cross3(V, w + KW_FPOS + 3 * f, t);
    for (int r = 0; r < 3; ++r) w[KW_FVEL + 3 * f + r] = V[3 + r] + t[r];
  }
  if (!deriv) {
    g.sync();
    return;
  }
  // P8: subtree momenta (into KW_SV), one (joint, component) pair per work item
  g.sync(); QM_TICK(20);
  QM_PFOR(g, idx, QM_NJ * 6) {
    const int j = idx / 6, c = idx - 6 * j;
    double acc = 0.0;
    const uint32_t mask = M.submask[j];
    if (subtree_is_window(mask, j)) {
      QM_UNROLL
      for (int d = 0; d < 6; ++d)
        if ((mask >> (j + d)) & 1u) acc += w[KW_HB + 6 * (j + d) + c];
    } else {
      for (int i = 0; i < QM_NJ; ++i)
        if ((mask >> i) & 1u) acc += w[KW_HB + 6 * i + c];
    }
    w[KW_SV + idx] = acc;
  }
  g.sync(); QM_TICK(21);
  // P9: d(A v)/dq_k and d(J_i v)/dq_k at fixed generalized velocity
  QM_PFOR(g, k, QM_NJ) {
    double S[6], X[6], IX[6], dL[3], dp[3], t[3];
    joint_S(M, w, k, S);
    const double* Vk = w + KW_V + 6 * k;
    const double* H = w + KW_SV + 6 * k;       // (L0, p) of the subtree
    // X = S x V_k (motion cross product)
    cross3(S, Vk, X);
    cross3(S, Vk + 3, X + 3);
    cross3_add(S + 3, Vk, X + 3);
    inertia_mul(w + KW_COMP + 10 * k, X, IX);
    // S x* H : (w x L0 + vO x p ; w x p)
    cross3(S, H, dL);
    cross3_add(S + 3, H + 3, dL);
    cross3(S, H + 3, dp);
    for (int r = 0; r < 3; ++r) { dL[r] -= IX[r]; dp[r] -= IX[3 + r]; }
    // to the centroidal frame: L = L0 - c x p
    const double* ptot = w + KW_SV + 3;        // linear momentum of the whole tree (joint 0 subtree)
    double dc[3];
    const double im = 1.0 / M.total_mass;
    for (int r = 0; r < 3; ++r) dc[r] = w[KW_ACM + r * QM_NJ + k] * im;
    cross3(dc, ptot, t);
    double t2[3];
    cross3(w + KW_COM, dp, t2);
    for (int r = 0; r < 3; ++r) {
      w[KW_DH + r * QM_NJ + k] = dp[r];
      w[KW_DH + (3 + r) * QM_NJ + k] = dL[r] - t[r] - t2[r];
    }
    if (deriv < 2) continue;
    for (int f = 0; f < QM_NFEET; ++f) {
      const int b = M.foot_joint[f];
      double du[3] = {0, 0, 0};
      if ((M.pathmask[b] >> k) & 1u) {
        const double* Vb = w + KW_V + 6 * b;
        double D[6], Y[6];
        for (int c = 0; c < 6; ++c) D[c] = Vb[c] - Vk[c];
        cross3(S, D, Y);
        cross3(S, D + 3, Y + 3);
        cross3_add(S + 3, D, Y + 3);
        const double* pos = w + KW_FPOS + 3 * f;
        double dpos[3] = {w[KW_FJ + (3 * f) * QM_NJ + k], w[KW_FJ + (3 * f + 1) * QM_NJ + k], w[KW_FJ + (3 * f + 2) * QM_NJ + k]};
        cross3(Y, pos, du);
        cross3_add(Vb, dpos, du);
        du[0] += Y[3]; du[1] += Y[4]; du[2] += Y[5];
      }
      for (int r = 0; r < 3; ++r) w[KW_DFV + (3 * f + r) * QM_NJ + k] = du[r];
    }
  }
  g.sync(); QM_TICK(22);
}

// Forward kinematics + centroidal quantities of one configuration, optionally with the generalized
// velocity implied by (x,u) and the q-derivatives at fixed velocity needed for the linearisation.
//   q = x[6:30];   u != nullptr -> velocity level;   deriv 1 -> KW_DH, 2 -> KW_DH and KW_DFV as well
template <class G>
QM_HDN void kin_eval(G g, const qmb200_model_desc& M, const double* x, const double* u, int deriv, double* w,
                     bool jac = true) {
  kin_positions(g, M, x + 6, w, jac || deriv != 0);
  if (u == nullptr) return;
  centroidal_velocity(g, M, x, u, w);
  kin_velocities(g, M, deriv, w);
}

}  // namespace qm
