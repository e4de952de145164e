// Latency of dependent FP64 operations on one warp (development aid; nvcc -arch=sm_100a -o fp64_latency fp64_latency.cu)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  cyc[0] = t1 - t0;
  double y = a;
  t0 = clock64();
  for (int i = 0; i < n; ++i) y = fma(y, sm[(i * 7 + threadIdx.x) & 1023], a);
  t1 = clock64();
  cyc[1] = t1 - t0;
  double z = a;
  t0 = clock64();
  for (int i = 0; i < n; ++i) z = z / b + a;
  t1 = clock64();
  cyc[2] = t1 - t0;
  double w = a + 2.0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) w = sqrt(w) + a;
  t1 = clock64();
  cyc[3] = t1 - t0;
  double s = a;
  t0 = clock64();
  for (int i = 0; i < n; ++i) s += __shfl_xor_sync(0xffffffffu, s, 1 + (i & 15));
  t1 = clock64();
  cyc[4] = t1 - t0;
  // pointer chase through shared memory (LDS latency)
  int* ism = (int*)sm;
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) ism[i] = (i * 33 + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  t0 = clock64();
  for (int i = 0; i < n; ++i) p = ism[p];
  t1 = clock64();
  cyc[5] = t1 - t0;
  double h = a;
  t0 = clock64();
  for (int i = 0; i < n; ++i) h = hypot(h, b);
  t1 = clock64();
  cyc[6] = t1 - t0;
  double m = a;
  t0 = clock64();
  for (int i = 0; i < n; ++i) m = m * b;
  t1 = clock64();
  cyc[7] = t1 - t0;
  float f = (float)a;
  t0 = clock64();
  for (int i = 0; i < n; ++i) f = fmaf(f, (float)b, (float)a);
  t1 = clock64();
  cyc[8] = t1 - t0;
  out[threadIdx.x] = x + y + z + w + s + p + h + m + f;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  const int n = 4096;
  k<<<1, 32>>>(out, cyc, 0.5, 0.999, n);
  k<<<1, 32>>>(out, cyc, 0.5, 0.999, n);
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA", "LDS+DFMA", "DDIV+DADD", "DSQRT+DADD", "SHFL(double)+DADD", "LDS chase", "hypot", "DMUL", "FFMA"};
  for (int i = 0; i < 9; ++i) printf("%-20s %.1f cycles per dependent step\n", names[i], (double)h[i] / n);
  return 0;
}
