// Micro-benchmark: FP64 pipe throughput on realistic register patterns (butterflies on 16 complex values
// held in registers) against the nominal 64 lanes/clk/SM, for 2..8 warps per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pulseportraiture_b200/csrc -o fp64_mix_probe tools/micro/fp64_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fft16.cuh"
using namespace ppb;

template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  cx<double> v[16];
  for (int i = 0; i < 16; ++i) v[i] = mk<double>(threadIdx.x * 1e-3 + i, 0.5 * i - threadIdx.x * 1e-4);
  const cx<double> w = mk<double>(a, b);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { dft16(v); }                                  // 160 FP64 (144 DADD)
    if (MODE == 1) { twiddle16(v, w); dft16(v); }                 // 273 FP64
    if (MODE == 2) {                                              // 32 DADD, two distinct register operands each
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i].x = v[i].x + v[(i + 1) & 15].y; v[i].y = v[i].y - v[(i + 5) & 15].x; }
    }
    if (MODE == 3) {                                              // 32 DFMA, three distinct register operands each
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i].x = fma(v[(i + 3) & 15].y, v[(i + 7) & 15].x, v[i].x); v[i].y = fma(v[(i + 2) & 15].x, v[(i + 9) & 15].y, v[i].y); }
    }
    if (MODE == 4) {                                              // 32 DFMA, two operands shared (uniform constants)
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i].x = fma(v[i].x, a, b); v[i].y = fma(v[i].y, a, b); }
    }
    if (MODE == 5) {                                              // 32 DMUL by one register operand + constant
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i].x = v[i].x * a; v[i].y = v[i].y * a; }
    }
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int nfp64, int threads, double* out, long long* cyc) {
  const int iters = 400;
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.9999999, 1e-9);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s warps/SMSP %d: %6.2f FP64 thread-instr per cycle per SM (nominal 64)\n", name, threads / 128, (double)iters * nfp64 * threads / c);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int th : {128, 256, 384, 512, 1024}) {
    run<0>("dft16 (144 DADD + 16 mul)", 160, th, out, cyc);
    run<1>("twiddle16 + dft16", 273, th, out, cyc);
    run<2>("DADD two register operands", 32, th, out, cyc);
    run<3>("DFMA three register operands", 32, th, out, cyc);
    run<4>("DFMA one register + two uniform", 32, th, out, cyc);
    run<5>("DMUL register * uniform", 32, th, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
