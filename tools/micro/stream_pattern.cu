// Read-bandwidth probe for the k_pass2 access pattern: rows of 8 KB (1024 float2); a warp reads
// (A) 4 rows at once, 8 lanes x 16 B = 128 B per row per load (k_pass2 today), (B) 2 rows, 16 lanes x 16 B =
// 256 B per row, (C) 1 row, 32 lanes x 16 B = 512 B per load, (D) as A but each lane takes 2 adjacent
// float4 per step (256 B per row per step).  Same loads in flight per lane (8 x 16 B), same grid shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// LPR lanes per row, W float4 per lane per step
template <int LPR, int W>
__global__ void __launch_bounds__(256, 3) k(const float4* __restrict__ x, float* out, long nrows) {
  constexpr int RPW = 32 / LPR;                 // rows per warp
  constexpr int ROW4 = 512;                     // float4 per row
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long row = ((long)blockIdx.x * 8 + w) * RPW + lane / LPR;
  if (row >= nrows) return;
  const int l = lane % LPR;
  const float4* p = x + row * ROW4;
  constexpr int STEP = LPR * W;                 // float4 per row per step
  constexpr int NSTEP = ROW4 / STEP;
  constexpr int U = 8 / W;                      // steps in flight
  float acc = 0.f;
#pragma unroll 1
  for (int s0 = 0; s0 < NSTEP; s0 += U) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < W; ++q) v[u][q] = ld_stream(p + (s0 + u) * STEP + l * W + q);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < W; ++q) acc += v[u][q].x + v[u][q].y + v[u][q].z + v[u][q].w;
  }
  if (acc == 1.2345f) out[0] = acc;
}
template <int LPR, int W> void run(const char* name, const float4* x, float* out, long nrows) {
  constexpr int RPW = 32 / LPR;
  const long nblk = (nrows + 8 * RPW - 1) / (8 * RPW);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    k<LPR, W><<<(unsigned)nblk, 256>>>(x, out, nrows);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  printf("%-44s %.3f ms  %.0f GB/s  (%s)\n", name, best, nrows * 8192.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const long nrows = 2000L * 512;               // 2000 subints x 512 channels = 8.4 GB
  float4* x; float* out;
  cudaMalloc(&x, nrows * 8192); cudaMalloc(&out, 4);
  cudaMemset(x, 0, nrows * 8192);
  run<8, 1>("A: 8 lanes/row, 128 B per row-load", x, out, nrows);
  run<16, 1>("B: 16 lanes/row, 256 B", x, out, nrows);
  run<32, 1>("C: 32 lanes/row, 512 B", x, out, nrows);
  run<8, 2>("D: 8 lanes/row, 2 float4 per lane (256 B)", x, out, nrows);
  run<8, 4>("E: 8 lanes/row, 4 float4 per lane (512 B)", x, out, nrows);
  return 0;
}
