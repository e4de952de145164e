// Cost of an 18-term dot product chain out of shared memory on one warp, in the code shapes the WBC iteration uses
// (development aid; nvcc -gencode arch=compute_100a,code=sm_100a -o smem_chain smem_chain.cu)
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ double smem[];
template <int MODE>
__device__ __forceinline__ double dot(const double* J, const double* np, int n, int c) {
  double s = 0.0;
  if (MODE == 0) { for (int i = 0; i < n; ++i) s += J[i * 18 + c] * np[i]; }
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 18; ++i) if (i < n) s += J[i * 18 + c] * np[i];
  }
  if (MODE == 2) {
#pragma unroll
    for (int i = 0; i < 18; ++i) s += J[i * 18 + c] * np[i];
  }
  if (MODE == 3) {       // three partial sums
    double a = 0.0, b = 0.0, d = 0.0;
#pragma unroll
    for (int i = 0; i < 18; i += 3) { a += J[i * 18 + c] * np[i]; b += J[(i + 1) * 18 + c] * np[i + 1]; d += J[(i + 2) * 18 + c] * np[i + 2]; }
    s = (a + b) + d;
  }
  return s;
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int n, int reps) {
  double* W = smem;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) W[i] = 1.0 + 1e-3 * i;
  __syncthreads();
  double* J = W; double* np = W + 400; double* d = W + 500;
  const int c = threadIdx.x;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (c < n) d[c] = dot<MODE>(J, np, n, c);
    __syncwarp();
    if (c < n) np[c] = d[c] * 1e-3;       // dependence between repetitions
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[MODE] = t1 - t0;
  out[threadIdx.x] = d[threadIdx.x & 15];
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  const int reps = 1000;
  for (int it = 0; it < 2; ++it) {
    k<0><<<1, 32, 8192>>>(out, cyc, 18, reps);
    k<1><<<1, 32, 8192>>>(out, cyc, 18, reps);
    k<2><<<1, 32, 8192>>>(out, cyc, 18, reps);
    k<3><<<1, 32, 8192>>>(out, cyc, 18, reps);
  }
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"rolled, runtime n", "unrolled + predicate", "unrolled, fixed 18", "unrolled, 3 partial sums"};
  for (int i = 0; i < 4; ++i) printf("%-26s %.1f cycles per 18-term product + hand-over\n", names[i], (double)h[i] / reps);
  return 0;
}
