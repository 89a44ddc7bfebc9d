// Micro-benchmark: per-SM throughput of the float<->double conversions (F2F on the XU pipe) against
// DFMA/DADD and against integer-ALU emulations of the same conversions, alone and mixed with DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o cvt_probe tools/micro/cvt_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double f2d_bits(float f) {
  const unsigned b = __float_as_uint(f);
  const unsigned a = b & 0x7fffffffu;
  unsigned hi = (a >> 3) + (a >= 0x00800000u ? 0x38000000u : 0u);
  hi |= b & 0x80000000u;
  return __hiloint2double((int)hi, (int)(b << 29));
}
__device__ __forceinline__ float d2f_bits(double d) {
  const unsigned hi = (unsigned)__double2hiint(d), lo = (unsigned)__double2loint(d);
  const unsigned a = hi & 0x7fffffffu;
  unsigned f = __funnelshift_l(lo, a - 0x38000000u, 3);
  const unsigned t = (lo & 0x1fffffffu) + 0x0fffffffu + (f & 1u);
  f += t >> 29;
  if (a < 0x38100000u) f = 0u;
  return __uint_as_float(f | (hi & 0x80000000u));
}

template <int MODE, int ILP>
__global__ void k(float* out, long long* cyc, int iters, float seed, double a, double b) {
  float f[ILP]; double d[ILP];
  for (int i = 0; i < ILP; ++i) { f[i] = seed + threadIdx.x * 1e-3f + i; d[i] = (double)f[i] * 1.000001; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) { d[i] = fma(d[i], a, b); }                                           // DFMA
        if (MODE == 1) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); f[i] += 1.f; }            // F2F.F64.F32 (+FADD)
        if (MODE == 2) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i])); d[i] = __hiloint2double(__double2hiint(d[i]) ^ 1, __float_as_int(f[i])); }  // F2F.F32.F64 (+2 ALU)
        if (MODE == 3) { d[i] = f2d_bits(f[i]); f[i] += 1.f; }                               // integer f->d
        if (MODE == 4) { f[i] = d2f_bits(d[i]); d[i] = __hiloint2double(__double2hiint(d[i]) ^ 1, __float_as_int(f[i])); }  // integer d->f
        if (MODE == 5) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i])); d[i] = fma(d[i], a, t); f[i] += 1.f;                  // 1 F2F.F64.F32 : 4 DFMA
                         d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); }
        if (MODE == 6) { double t = f2d_bits(f[i]); d[i] = fma(d[i], a, t); f[i] += 1.f;      // 1 integer f->d : 4 DFMA
                         d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); }
        if (MODE == 7) { d[i] = d[i] + a; }                                                  // DADD
        if (MODE == 8) { float t; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(d[i])); f[i] += t;         // 1 F2F.F32.F64 : 4 DFMA
                         d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); }
        if (MODE == 9) { float t = d2f_bits(d[i]); f[i] += t;                                 // 1 integer d->f : 4 DFMA
                         d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); d[i] = fma(d[i], a, b); }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < ILP; ++i) s += f[i] + (float)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, float* out, long long* cyc) {
  const int iters = 1000; constexpr int ILP = 8;
  k<MODE, ILP><<<148, threads>>>(out, cyc, iters, 1.5f, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 4 * ILP;   // loop bodies per thread
  printf("%-44s threads %4d: %7.2f cycles per body per warp -> %6.2f bodies (thread-level) per cycle per SM\n", name, threads, c / n / (threads / 32) * (threads / 32), n * threads / c);
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int th : {128, 512, 1024}) {
    run<0>("DFMA", th, out, cyc);
    run<7>("DADD", th, out, cyc);
    run<1>("cvt.f64.f32 (+FADD)", th, out, cyc);
    run<2>("cvt.rn.f32.f64 (+2 ALU)", th, out, cyc);
    run<3>("integer f32->f64 (+FADD)", th, out, cyc);
    run<4>("integer f64->f32 RN (+2 ALU)", th, out, cyc);
    run<5>("1 cvt.f64.f32 + 4 DFMA", th, out, cyc);
    run<6>("1 integer f32->f64 + 4 DFMA", th, out, cyc);
    run<8>("1 cvt.rn.f32.f64 + 4 DFMA", th, out, cyc);
    run<9>("1 integer f64->f32 + 4 DFMA", th, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
