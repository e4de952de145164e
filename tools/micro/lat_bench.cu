// Micro-benchmark: dependent-issue latencies (cycles) of the FP64 instructions the serial chains are made of (development aid).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_lat(double* out, long long* cyc, double seed) {
  __shared__ double sh[64];
  const int lane = threadIdx.x;
  sh[lane] = seed + lane; sh[lane + 32] = seed;
  __syncwarp();
  double x = seed + lane * 1e-3, y = 1.0000001, z = 1e-9;
  long long t0, t1;
  const int N = 256;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, z);
  t1 = clock64();
  if (lane == 0) cyc[0] = (t1 - t0);
  // DMUL+DADD independent x4 (throughput with ILP 4)
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z); }
  t1 = clock64();
  if (lane == 0) cyc[1] = (t1 - t0);
  x = a0 + a1 + a2 + a3;
  // reciprocal chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) x = 1.0 / (x + 1.5);
  t1 = clock64();
  if (lane == 0) cyc[2] = (t1 - t0);
  // shuffle chain (64-bit)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  t1 = clock64();
  if (lane == 0) cyc[3] = (t1 - t0);
  // LDS chain (pointer chase through doubles)
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double v = sh[idx]; idx = (int)v & 31; }
  t1 = clock64();
  if (lane == 0) cyc[4] = (t1 - t0);
  // DMMA dependent chain
  double c0 = x, c1 = y;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) dmma884(c0, c1, y, z);
  t1 = clock64();
  if (lane == 0) cyc[5] = (t1 - t0);
  // DMMA 4 independent accumulators
  double d[4][2] = {{c0, c1}, {c1, c0}, {x, y}, {y, x}};
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { dmma884(d[0][0], d[0][1], y, z); dmma884(d[1][0], d[1][1], y, z); dmma884(d[2][0], d[2][1], y, z); dmma884(d[3][0], d[3][1], y, z); }
  t1 = clock64();
  if (lane == 0) cyc[6] = (t1 - t0);
  // STS + syncwarp + LDS round trip
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { sh[lane] = x; __syncwarp(); x = sh[(lane + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64();
  if (lane == 0) cyc[7] = (t1 - t0);
  out[lane] = x + idx + c0 + c1 + d[0][0] + d[1][0] + d[2][1] + d[3][1];
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
  long long h[8];
  for (int rep = 0; rep < 2; ++rep) { k_lat<<<1, 32>>>(out, cyc, 1.0); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("one warp alone on an SM (cycles per operation)\n");
  printf("DFMA dependent chain      : %.1f\n", h[0] / 256.0);
  printf("DFMA 4 independent chains : %.1f per group of 4\n", h[1] / 256.0);
  printf("1.0/x (double) chain      : %.1f (incl. one DADD)\n", h[2] / 64.0);
  printf("SHFL.64 dependent chain   : %.1f\n", h[3] / 256.0);
  printf("LDS.64 dependent chain    : %.1f (incl. F2I + LOP)\n", h[4] / 256.0);
  printf("DMMA.884 dependent chain  : %.1f\n", h[5] / 256.0);
  printf("DMMA.884 4 independent    : %.1f per group of 4\n", h[6] / 256.0);
  printf("STS+syncwarp+LDS+DADD+syncwarp : %.1f\n", h[7] / 256.0);
  return 0;
}
