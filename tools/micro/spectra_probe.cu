// Probe: variants of k_spectra<1024> (the knobs of SpecPlan16T, spectra_plan.cuh) timed side by side on
// the same synthetic rows and compared with the default plan's outputs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I pulseportraiture_b200/csrc -o spectra_probe tools/micro/spectra_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "kernels.cuh"
#include "tw_host.h"
using namespace ppb;
constexpr int N = 1024;

__global__ void k_fill(float* d, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    const int b = (int)(i & 2047);
    const float pulse = 8.f * __expf(-0.5f * (b - 700.f) * (b - 700.f) / 400.f);
    d[i] = ((float)(x >> 8) * (1.f / 16777216.f) - 0.5f) * 5.f + pulse;
  }
}

struct Outs { std::vector<float2> X, Xlo, part; std::vector<double> sigma, Ssn, Sdn; };

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int nsub = argc > 1 ? atoi(argv[1]) : 1000, nchan = 512, G = 32, gx = nchan / G, nparts = gx;
  const int ncheck = 8;   // subints compared with the default plan
  const size_t nfl = (size_t)nsub * nchan * 2 * N;
  float* data; CKC(cudaMalloc(&data, nfl * 4));
  k_fill<<<148 * 8, 256>>>(data, nfl, 12345u);
  std::vector<double2> m64((size_t)nchan * N); std::vector<float2> m32(m64.size()); std::vector<double> pn(nchan, 1.0), numean(nsub, 1500.0);
  for (size_t i = 0; i < m64.size(); ++i) {
    const int k = (int)(i % N);
    const double amp = exp(-1e-4 * (double)(k ? k : N) * (double)(k ? k : N)) + 1e-9;
    m64[i] = make_double2(amp * cos(0.37 * i), amp * sin(0.37 * i));
    m32[i] = make_float2((float)m64[i].x, (float)m64[i].y);
  }
  double2* dm64; float2* dm32; double *dpn, *dnumean, *dsigma, *dSsn, *dSdn; float2 *dX, *dXlo, *dpart;
  CKC(cudaMalloc(&dm64, m64.size() * 16)); CKC(cudaMalloc(&dm32, m32.size() * 8)); CKC(cudaMalloc(&dpn, nchan * 8)); CKC(cudaMalloc(&dnumean, nsub * 8));
  CKC(cudaMalloc(&dsigma, (size_t)nsub * nchan * 8)); CKC(cudaMalloc(&dSsn, (size_t)nsub * nchan * 8)); CKC(cudaMalloc(&dSdn, (size_t)nsub * nchan * 8));
  CKC(cudaMalloc(&dX, (size_t)nsub * nchan * N * 8)); CKC(cudaMalloc(&dXlo, (size_t)nsub * nchan * 64 * 8)); CKC(cudaMalloc(&dpart, (size_t)nsub * nparts * N * 8));
  CKC(cudaMemcpy(dm64, m64.data(), m64.size() * 16, cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dm32, m32.data(), m32.size() * 8, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dpn, pn.data(), nchan * 8, cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dnumean, numean.data(), nsub * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  Outs ref;
  auto fetch = [&](Outs& o) {
    o.X.resize((size_t)ncheck * nchan * N); o.Xlo.resize((size_t)ncheck * nchan * 64); o.part.resize((size_t)ncheck * nparts * N);
    o.sigma.resize((size_t)ncheck * nchan); o.Ssn.resize(o.sigma.size()); o.Sdn.resize(o.sigma.size());
    CKC(cudaMemcpy(o.X.data(), dX, o.X.size() * 8, cudaMemcpyDeviceToHost)); CKC(cudaMemcpy(o.Xlo.data(), dXlo, o.Xlo.size() * 8, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(o.part.data(), dpart, o.part.size() * 8, cudaMemcpyDeviceToHost)); CKC(cudaMemcpy(o.sigma.data(), dsigma, o.sigma.size() * 8, cudaMemcpyDeviceToHost));
    CKC(cudaMemcpy(o.Ssn.data(), dSsn, o.Ssn.size() * 8, cudaMemcpyDeviceToHost)); CKC(cudaMemcpy(o.Sdn.data(), dSdn, o.Sdn.size() * 8, cudaMemcpyDeviceToHost));
  };
  auto run_k = [&](const char* name, auto plan_tag, auto kern, size_t smem, int drop = 0) {
    using PL = decltype(plan_tag);
    std::vector<double2> tw; TwBuilder<PL>::build(tw);
    cx<double>* dtw; CKC(cudaMalloc(&dtw, tw.size() * 16)); CKC(cudaMemcpy(dtw, tw.data(), tw.size() * 16, cudaMemcpyHostToDevice));
    CKC(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PL::kThreads, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    SpectraArgs a; memset(&a, 0, sizeof a);
    a.data = data; a.mconj64 = reinterpret_cast<cx<double>*>(dm64); a.mconj32 = reinterpret_cast<cx<float>*>(dm32); a.pn = dpn; a.nu_mean = dnumean;
    a.X = dX; a.Xlo = dXlo; a.partial = dpart; a.sigma = dsigma; a.Ssn = dSsn; a.Sdn = dSdn; a.tw8 = dtw; a.s0 = 0; a.nchan = nchan; a.G = G; a.nparts = nparts;
    if (drop & 1) { a.X = nullptr; a.Xlo = nullptr; }   // experiment: no cross-spectrum stores / model loads
    if (drop & 2) a.partial = nullptr;
    CKC(cudaMemset(dpart, 0xff, (size_t)ncheck * nparts * N * 8));
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      kern<<<dim3(gx, nsub), PL::kThreads, smem>>>(a);
      cudaEventRecord(e1); CKC(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = std::min(best, ms);
    }
    CKC(cudaGetLastError());
    Outs o; fetch(o);
    if (ref.X.empty()) ref = o;
    size_t nx = 0, nlo = 0, ns = 0; double perr = 0, pmax = 0, xrel = 0, srel = 0;
    for (size_t i = 0; i < o.X.size(); ++i) {
      const double m = std::max(fabs((double)ref.X[i].x), fabs((double)ref.X[i].y));
      if (m > 0) xrel = std::max(xrel, std::max(fabs((double)o.X[i].x - ref.X[i].x), fabs((double)o.X[i].y - ref.X[i].y)) / m);
    }
    for (size_t i = 0; i < o.sigma.size(); ++i) {
      if (ref.sigma[i] > 0) srel = std::max(srel, fabs(o.sigma[i] / ref.sigma[i] - 1.0));
      if (ref.Sdn[i] > 0) srel = std::max(srel, fabs(o.Sdn[i] / ref.Sdn[i] - 1.0));
    }
    for (size_t i = 0; i < o.X.size(); ++i) nx += (o.X[i].x != ref.X[i].x) || (o.X[i].y != ref.X[i].y);
    for (size_t i = 0; i < o.Xlo.size(); ++i) nlo += (o.Xlo[i].x != ref.Xlo[i].x) || (o.Xlo[i].y != ref.Xlo[i].y);
    for (size_t i = 0; i < o.sigma.size(); ++i) ns += (o.sigma[i] != ref.sigma[i]) || (o.Ssn[i] != ref.Ssn[i]) || (o.Sdn[i] != ref.Sdn[i]);
    for (size_t i = 0; i < o.part.size(); ++i) {
      perr = std::max(perr, (double)std::max(fabsf(o.part[i].x - ref.part[i].x), fabsf(o.part[i].y - ref.part[i].y)));
      pmax = std::max(pmax, (double)std::max(fabsf(ref.part[i].x), fabsf(ref.part[i].y)));
    }
    printf("%-34s regs %3d smem %6zu CTAs/SM %d | %.3f ms per %d subints = %.0f GB/s (2B) | diff X %zu (rel %.1e) Xlo %zu sig %zu (rel %.1e) partial %.2e (max %.2e)\n",
           name, fa.numRegs, smem, nb, best, nsub, 2.0 * nfl * 4 / best / 1e6, nx, xrel, nlo, ns, srel, perr, pmax);
    fflush(stdout);
    cudaFree(dtw);
  };
  auto run = [&](const char* name, auto plan_tag) {
    using PL = decltype(plan_tag);
    run_k(name, plan_tag, k_spectra<N, PL, false>, spectra_smem_bytes_of<N, PL>());
  };
  //                 STAGES MINB ACC MCLATE TWTAB CVT
  run("base  st2 mb6 acc0", SpecPlan16T<2, 6, 0, false, false>{});
  run_k("k_spectra16<f32, guess>", SpecPlan16{}, k_spectra16<false, true, false>, spectra_smem_bytes_of<N, SpecPlan16>());
  const size_t sm16 = spectra_smem_bytes_of<N, SpecPlan16>();
  run_k("k16 noguess", SpecPlan16{}, k_spectra16<false, false, false>, sm16, 2);
  run_k("k16 EXP15 none of them", SpecPlan16{}, k_spectra16<false, true, false, 15>, sm16);
  return 0;
}
