// Probe: row-FFT core of k_spectra, radix-8 plan (128 threads/row) against the
// radix-16 plan (64 threads/row), same TMA staging skeleton, no split/emit.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pulseportraiture_b200/csrc -o fft16_probe tools/micro/fft16_probe.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "fft16.cuh"
using namespace ppb;
constexpr int N = 1024;

template <int T, int MINB, bool R16>
__global__ void __launch_bounds__(T, MINB) probe(const float* data, const cx<double>* twg, int ntw, double* out, int nchan, int G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<double>* tw = reinterpret_cast<cx<double>*>(smem_raw);
  cx<double>* buf = tw + ((ntw + 1) & ~1);
  float* stage = reinterpret_cast<float*>(buf + N);
  __shared__ __align__(8) unsigned long long mbar[2];
  const int t = threadIdx.x;
  for (int i = t; i < ntw; i += T) tw[i] = twg[i];
  if (t == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
  mbar_fence_init();
  __syncthreads();
  const int s = blockIdx.y, ch0 = blockIdx.x * G;
  auto fetch = [&](int step) {
    if (t == 0 && step < G) {
      mbar_expect_tx(&mbar[step & 1], 8192);
      bulk_g2s(stage + (size_t)(step & 1) * 2 * N, data + ((size_t)s * nchan + ch0 + step) * 2 * N, 8192, &mbar[step & 1]);
    }
  };
  fetch(0); fetch(1);
  unsigned ph0 = 0, ph1 = 0;
  double accx = 0, accy = 0;
  for (int step = 0; step < G; ++step) {
    if (step & 1) { mbar_wait(&mbar[1], ph1); ph1 ^= 1u; } else { mbar_wait(&mbar[0], ph0); ph0 ^= 1u; }
    const float2* g = reinterpret_cast<const float2*>(stage + (size_t)(step & 1) * 2 * N);
    auto nop = []() {};
    if constexpr (R16) {
      fft16_rows1024<double>(buf, tw, tw, t, RowSrcF32{g}, true, []() { __syncthreads(); }, [&]() { fetch(step + 2); }, nop);
#pragma unroll
      for (int i = 0; i < 16; ++i) { const cx<double> z = buf[phys16(t + 64 * i)]; accx += z.x; accy += z.y; }
    } else {
      fft8_rows<N, double>(buf, tw, t, 0, RowSrcF32{g}, true, [&]() { fetch(step + 2); }, nop);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const cx<double> z = buf[phys(t + 128 * i)]; accx += z.x; accy += z.y; }
    }
  }
  out[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * T + t] = accx + accy;
}

int main() {
  const int nsub = 2000, nchan = 512, G = 32;
  const size_t nfl = (size_t)nsub * nchan * 2 * N;
  float* data; cudaMalloc(&data, nfl * 4);
  std::vector<float> h(1 << 24);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u >> 8) & 0xffff) / 65536.f - 0.5f;
  for (size_t o = 0; o < nfl; o += h.size()) cudaMemcpy(data + o, h.data(), std::min(h.size(), nfl - o) * 4, cudaMemcpyHostToDevice);
  // radix-8 table (TwLayout<1024>) and radix-16 table
  std::vector<cx<double>> t8(TwLayout<N>::kTotal), t16(16);
  {
    using L = TwLayout<N>; using P = Plan8<N>;
    for (int i = 1; i < P::n; ++i) { const int ns = L::ns(i), R = P::radix(i); for (int k = 0; k < ns; ++k) { double a = -2 * M_PI * k / (double)(ns * R); t8[L::off(i) + k] = {cos(a), sin(a)}; } }
    for (int p = 0; p <= N / 2; ++p) { double a = -2 * M_PI * p / (2.0 * N); t8[L::kSplitOff + p] = {cos(a), sin(a)}; }
    for (int k = 0; k < 16; ++k) { double a = -2 * M_PI * k / 256.0; t16[k] = {cos(a), sin(a)}; }
  }
  cx<double>*d8, *d16; cudaMalloc(&d8, t8.size() * 16); cudaMalloc(&d16, 16 * 16);
  cudaMemcpy(d8, t8.data(), t8.size() * 16, cudaMemcpyHostToDevice); cudaMemcpy(d16, t16.data(), 256, cudaMemcpyHostToDevice);
  double* out; cudaMalloc(&out, (size_t)nsub * 16 * 128 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern, int T, const cx<double>* tw, int ntw) {
    const size_t smem = (size_t)((ntw + 1) & ~1) * 16 + N * 16 + 2 * 2 * N * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, T, smem);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      kern<<<dim3(nchan / G, nsub), T, smem>>>(data, tw, ntw, out, nchan, G);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = std::min(best, ms);
    }
    double hsum = 0; std::vector<double> ho(1024); cudaMemcpy(ho.data(), out, 8192, cudaMemcpyDeviceToHost); for (double x : ho) hsum += x;
    printf("%-28s smem %6zu B  CTAs/SM %d  %.3f ms per %d subints (%.0f GB/s read)  chk %.6e  err %s\n", name, smem, nb, best, nsub, nfl * 4 / best / 1e6, hsum, cudaGetErrorString(cudaGetLastError()));
  };
  run("radix-8  T=128 minb5", probe<128, 5, false>, 128, d8, (int)t8.size());
  run("radix-16 T=64  minb6", probe<64, 6, true>, 64, d16, 16);
  run("radix-16 T=64  minb5", probe<64, 5, true>, 64, d16, 16);
  run("radix-16 T=64  minb4", probe<64, 4, true>, 64, d16, 16);
  return 0;
}
