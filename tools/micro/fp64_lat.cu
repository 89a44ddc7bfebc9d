// Micro-benchmark: FP64 dependent-issue latency and per-SM throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double v[ILP];
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int threads, double* out, long long* cyc) {
  const int iters = 2000;
  k<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 8 * ILP;   // DFMA per thread
  const int warps = threads / 32;
  printf("ILP %d warps/SM %2d: %.2f cycles per DFMA per warp, %.2f DFMA warp-inst/cycle/SM\n", ILP, warps, c / n, n * warps / c);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int th : {32, 128, 256, 512, 640, 1024}) {
    run<1>(th, out, cyc); run<2>(th, out, cyc); run<4>(th, out, cyc); run<8>(th, out, cyc);
  }
  return 0;
}
