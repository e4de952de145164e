// Micro-benchmark of one-warp SPD inverse variants (development aid).
#include <cstdio>
#include <cuda_runtime.h>
#define QM_NUT 18
template <int V>
__device__ __forceinline__ void inv(double* Gm, int n, double* col) {
  const int lane = threadIdx.x & 31;
  const bool active = lane < n;
  double r[QM_NUT];
#pragma unroll
  for (int i = 0; i < QM_NUT; ++i) r[i] = (active && i < n) ? Gm[QM_NUT * i + lane] : ((i == lane) ? 1.0 : 0.0);
#pragma unroll
  for (int c = 0; c < QM_NUT; ++c) {
    if (c < n) {
      double* cb = col + 20 * (c & 1);
      const double rc = r[c];
      if (V == 0) {          // current: owner lane scales and publishes
        if (lane == c) {
          const double ip = 1.0 / rc;
#pragma unroll
          for (int i = 0; i < QM_NUT; ++i) cb[i] = (i == c) ? ip : -r[i] * ip;
        }
        __syncwarp();
      } else if (V == 1) {   // all lanes compute, owner stores (no divergent region)
        const double ip = 1.0 / ((lane == c) ? rc : 1.0);
#pragma unroll
        for (int i = 0; i < QM_NUT; ++i) { const double v = (i == c) ? ip : -r[i] * ip; if (lane == c) cb[i] = v; }
        __syncwarp();
      } else if (V == 2) {   // no division (timing only)
        const double ip = rc * 0.01;
#pragma unroll
        for (int i = 0; i < QM_NUT; ++i) { const double v = (i == c) ? ip : -r[i] * ip; if (lane == c) cb[i] = v; }
        __syncwarp();
      } else if (V == 3) {   // shuffle broadcast of the pivot, then every lane fetches its multiplier by shuffle
        const double ip = 1.0 / __shfl_sync(0xffffffffu, rc, c);
#pragma unroll
        for (int i = 0; i < QM_NUT; ++i) cb[i] = 0.0;   // unused
      }
      if (V == 3) {
        const double ip = 1.0 / __shfl_sync(0xffffffffu, rc, c);
#pragma unroll
        for (int i = 0; i < QM_NUT; ++i) {
          if (i != c && i < n) { const double m = __shfl_sync(0xffffffffu, r[i], c) * ip; r[i] = (lane == c) ? -m : r[i] - m * rc; }
        }
        r[c] = (lane == c) ? ip : rc * ip;
      } else {
        const double p = cb[c];
#pragma unroll
        for (int i = 0; i < QM_NUT; ++i) {
          if (i != c && i < n) { const double mlt = cb[i]; r[i] = (lane == c) ? mlt : r[i] + mlt * rc; }
        }
        r[c] = (lane == c) ? p : rc * p;
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < QM_NUT; ++i) if (i < n) Gm[QM_NUT * i + lane] = r[i];
  }
}
// variant 4/5: owner lane publishes the negated scaled column; all lanes fetch it with 9 unpredicated 16-byte loads and
// update all 18 rows unconditionally (rows >= n carry zero multipliers). 5: next pivot's reciprocal started early.
template <int V>
__device__ __forceinline__ void inv2(double* Gm, int n, double* col) {
  const int lane = threadIdx.x & 31;
  const bool active = lane < n;
  double r[QM_NUT];
#pragma unroll
  for (int i = 0; i < QM_NUT; ++i) r[i] = (active && i < n) ? Gm[QM_NUT * i + lane] : ((i == lane) ? 1.0 : 0.0);
  double ip = 1.0 / r[0];
#pragma unroll
  for (int c = 0; c < QM_NUT; ++c) {
    if (c < n) {
      double* cb = col + 20 * (c & 1);
      const double rc = r[c];
      if (lane == c) {
        if (V == 4) ip = 1.0 / rc;
#pragma unroll
        for (int i = 0; i < QM_NUT; i += 2) {
          double2 v;
          v.x = (i == c) ? ip : -r[i] * ip;
          v.y = (i + 1 == c) ? ip : -r[i + 1] * ip;
          reinterpret_cast<double2*>(cb)[i >> 1] = v;
        }
      }
      __syncwarp();
      double2 m2[QM_NUT / 2];
#pragma unroll
      for (int q = 0; q < QM_NUT / 2; ++q) m2[q] = reinterpret_cast<const double2*>(cb)[q];
      if (V == 5 && c + 1 < QM_NUT) {
        const double mlt = ((c + 1) & 1) ? m2[(c + 1) >> 1].y : m2[(c + 1) >> 1].x;
        r[c + 1] = (lane == c) ? mlt : r[c + 1] + mlt * rc;
        ip = 1.0 / ((lane == c + 1) ? r[c + 1] : 1.0);
      }
#pragma unroll
      for (int i = 0; i < QM_NUT; ++i) {
        if (i != c && !(V == 5 && i == c + 1)) {
          const double mlt = (i & 1) ? m2[i >> 1].y : m2[i >> 1].x;
          r[i] = (lane == c) ? mlt : r[i] + mlt * rc;
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
}
__device__ __forceinline__ double rcp_fast(double x) {     // branch-free reciprocal: MUFU.RCP64H seed + Newton steps
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double t = fma(-x, y, 1.0);
  t = fma(t, t, t);
  y = fma(y, t, y);
  t = fma(-x, y, 1.0);
  return fma(y, t, y);
}
// variant 6: raw pivot column through shared memory, reciprocal of the next pivot computed branch-free in the shadow of
// the row updates and broadcast with one shuffle
__device__ __forceinline__ void inv3(double* Gm, int n, double* col) {
  const int lane = threadIdx.x & 31;
  const bool active = lane < n;
  double r[QM_NUT];
#pragma unroll
  for (int i = 0; i < QM_NUT; ++i) r[i] = (active && i < n) ? Gm[QM_NUT * i + lane] : ((i == lane) ? 1.0 : 0.0);
  double ip = rcp_fast((lane == 0) ? r[0] : 1.0);
#pragma unroll
  for (int c = 0; c < QM_NUT; ++c) {
    if (c < n) {
      double* cb = col + 20 * (c & 1);
      const double rc = r[c];
      if (lane == c) {
#pragma unroll
        for (int i = 0; i < QM_NUT; i += 2) reinterpret_cast<double2*>(cb)[i >> 1] = make_double2(r[i], r[i + 1]);
      }
      const double ipb = __shfl_sync(0xffffffffu, ip, c);
      __syncwarp();
      double2 m2[QM_NUT / 2];
#pragma unroll
      for (int q = 0; q < QM_NUT / 2; ++q) m2[q] = reinterpret_cast<const double2*>(cb)[q];
      if (c + 1 < QM_NUT) {
        const double mlt = -(((c + 1) & 1) ? m2[(c + 1) >> 1].y : m2[(c + 1) >> 1].x) * ipb;
        r[c + 1] = (lane == c) ? mlt : r[c + 1] + mlt * rc;
        ip = rcp_fast((lane == c + 1) ? r[c + 1] : 1.0);
      }
#pragma unroll
      for (int i = 0; i < QM_NUT; ++i) {
        if (i != c && i != c + 1) {
          const double mlt = -((i & 1) ? m2[i >> 1].y : m2[i >> 1].x) * ipb;
          r[i] = (lane == c) ? mlt : r[i] + mlt * rc;
        }
      }
      r[c] = (lane == c) ? ipb : rc * ipb;
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < QM_NUT; ++i) if (i < n) Gm[QM_NUT * i + lane] = r[i];
  }
}
template <int V>
__global__ void k(double* out, long long* cyc, int n, int reps) {
  __shared__ double G[QM_NUT * QM_NUT];
  __shared__ __align__(16) double col[40];
  const int lane = threadIdx.x;
  long long tot = 0;
  for (int rep = 0; rep < reps; ++rep) {
    for (int i = lane; i < QM_NUT * QM_NUT; i += 32) { const int r = i / QM_NUT, c = i % QM_NUT; G[i] = (r == c ? 20.0 + r : 0.0) + 1.0 / (1 + r + c) + rep * 1e-3; }
    __syncwarp();
    const long long t0 = clock64();
    if (V == 6) inv3(G, n, col); else if (V >= 4) inv2<V>(G, n, col); else inv<V>(G, n, col);
    __syncwarp();
    tot += clock64() - t0;
  }
  if (lane == 0) cyc[0] = tot / reps;
  for (int i = lane; i < QM_NUT * QM_NUT; i += 32) out[i] = G[i];
}
template <int V> void run(double* out, long long* cyc) {
  for (int n : {16, 18}) {
    k<V><<<1, 32>>>(out, cyc, n, 50); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ho[324]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    printf("variant %d n=%d: %lld cycles per inverse (%.0f per sweep)  Ginv[0][0]=%.15e Ginv[5][7]=%.15e\n", V, n, h, (double)h / n, ho[0], ho[5*18+7]);
  }
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 324 * 8); cudaMalloc(&cyc, 8);
  run<4>(out, cyc); run<6>(out, cyc);
  return 0;
}
