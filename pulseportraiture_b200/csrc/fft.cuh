// Shared-memory real-to-complex FFT rows for sm_100a.
//
// A real row of nbin = 2N samples is viewed as N complex points
// z_j = x_{2j} + i x_{2j+1}; an N-point complex Stockham autosort FFT (radix-4
// passes, one trailing radix-2 pass when log2 N is odd) runs in shared memory
// and the half-spectrum d_k, k = 0..N, is recovered with the usual split
// d_k = E_k + W_{2N}^k O_k.  Replaces numpy.fft.rfft(axis=1) at
// pplib.py:2127-2130 / pptoaslib.py:976-979 (forward sign e^{-2 pi i jk/n}).
//
// The arithmetic type T is float or double.  numpy's float64 rfft is what the
// reference computes; a float FFT carries ~1e-7 relative error per harmonic,
// which is enough to move chi^2 by more than the 1e-8 bar, so the fit path
// (k_spectra, fft8.cuh / fft16.cuh) always runs in double; the float
// instantiation of these radix-4 rows serves rotation and noise helpers when
// the caller asks for it (pp_plan_set_fft_precision).
//
// Twiddles come from tables computed on the host in double precision:
// twN[j] = e^{-2 pi i j/N}, tw2N[k] = e^{-2 pi i k/(2N)}, k = 0..N/2.
#pragma once
#include <cuda_runtime.h>

namespace ppb {

template <typename T> struct cx { T x, y; };
template <> struct __align__(8) cx<float> { float x, y; };
template <> struct __align__(16) cx<double> { double x, y; };

template <typename T> __device__ __forceinline__ cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
  return mk<T>(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
template <typename T> __device__ __forceinline__ cx<T> cadd(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> __device__ __forceinline__ cx<T> csub(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> __device__ __forceinline__ cx<T> cconj(cx<T> a) { return mk<T>(a.x, -a.y); }

template <int N> struct Log2 { static constexpr int value = 1 + Log2<N / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// Geometry of one CTA of 256 threads working on rows of N complex points.
template <int N> struct RowGeom {
  static constexpr int kThreads = 256;
  static constexpr int kRows = (1024 / N) > 0 ? (1024 / N) : 1;  // concurrent rows per CTA
  static constexpr int kTRow = kThreads / kRows;                 // threads per row
  static constexpr int kPairs = (N / 2 + kTRow - 1) / kTRow;     // unpack pairs per thread
  static constexpr int kLoads = N / (2 * kTRow);                 // float4 loads per thread per row
  static_assert(N >= 32 && N <= 2048, "nbin must be in [64, 4096]");
  static_assert(kLoads >= 1, "row too short for float4 staging");
};

// One Stockham radix-4 pass: src -> dst, Ns = product of the radices done.
template <int N, int TROW, typename T>
__device__ __forceinline__ void pass_radix4(const cx<T>* __restrict__ src, cx<T>* __restrict__ dst,
                                            const cx<T>* __restrict__ twN, int t_row, int Ns) {
  constexpr int Q = N / 4;
#pragma unroll
  for (int j = t_row; j < Q; j += TROW) {
    const int k = j & (Ns - 1);
    const int tstep = k * (Q / Ns);  // index of e^{-2 pi i k/(4 Ns)} in twN
    cx<T> v0 = src[j];
    cx<T> v1 = src[j + Q];
    cx<T> v2 = src[j + 2 * Q];
    cx<T> v3 = src[j + 3 * Q];
    if (Ns > 1) {
      v1 = cmul(v1, twN[tstep]);
      v2 = cmul(v2, twN[2 * tstep]);
      v3 = cmul(v3, twN[3 * tstep]);
    }
    const cx<T> a0 = cadd(v0, v2), a1 = csub(v0, v2);
    const cx<T> a2 = cadd(v1, v3), a3 = csub(v1, v3);
    const cx<T> b3 = mk<T>(a3.y, -a3.x);  // -i * a3
    const int j0 = ((j - k) << 2) + k;
    dst[j0] = cadd(a0, a2);
    dst[j0 + Ns] = cadd(a1, b3);
    dst[j0 + 2 * Ns] = csub(a0, a2);
    dst[j0 + 3 * Ns] = csub(a1, b3);
  }
}

template <int N, int TROW, typename T>
__device__ __forceinline__ void pass_radix2_last(const cx<T>* __restrict__ src, cx<T>* __restrict__ dst,
                                                 const cx<T>* __restrict__ twN, int t_row) {
  constexpr int H = N / 2;  // Ns == N/2 for the last pass
#pragma unroll
  for (int j = t_row; j < H; j += TROW) {
    const cx<T> v0 = src[j];
    const cx<T> v1 = cmul(src[j + H], twN[j]);
    dst[j] = cadd(v0, v1);
    dst[j + H] = csub(v0, v1);
  }
}

// Forward complex FFT of N points held in bufA (natural order).  All threads of
// the CTA must call (uses __syncthreads()); each row-slot passes its own
// bufA/bufB.  Returns the buffer holding the natural-order result.
template <int N, int TROW, typename T>
__device__ __forceinline__ cx<T>* fft_forward(cx<T>* bufA, cx<T>* bufB, const cx<T>* __restrict__ twN, int t_row) {
  constexpr int L = Log2<N>::value;
  cx<T>* src = bufA;
  cx<T>* dst = bufB;
  int Ns = 1;
#pragma unroll
  for (int p = 0; p < L / 2; ++p) {
    pass_radix4<N, TROW, T>(src, dst, twN, t_row, Ns);
    __syncthreads();
    cx<T>* t = src; src = dst; dst = t;
    Ns <<= 2;
  }
  if (L & 1) {
    pass_radix2_last<N, TROW, T>(src, dst, twN, t_row);
    __syncthreads();
    cx<T>* t = src; src = dst; dst = t;
  }
  return src;
}

// Half-spectrum pair from the N-point complex FFT Z of the packed real row:
// returns d_p and d_{N-p} for 1 <= p <= N/2 (for p == N/2 both are the same).
template <typename T>
__device__ __forceinline__ void unpack_pair(const cx<T>* __restrict__ Z, const cx<T>* __restrict__ tw2N, int N, int p,
                                            cx<T>& dp, cx<T>& dq) {
  const cx<T> zk = Z[p];
  const cx<T> zc = cconj(Z[N - p]);
  const cx<T> E = mk<T>(T(0.5) * (zk.x + zc.x), T(0.5) * (zk.y + zc.y));
  const cx<T> D = mk<T>(T(0.5) * (zk.x - zc.x), T(0.5) * (zk.y - zc.y));
  const cx<T> O = mk<T>(D.y, -D.x);  // -i * D
  const cx<T> t = cmul(tw2N[p], O);
  dp = cadd(E, t);
  dq = cconj(csub(E, t));
}

// Inverse of unpack_pair: from half-spectrum values d_p, d_{N-p} rebuild the
// packed complex spectrum entries Z_p and Z_{N-p} (used by the C2R rotation).
template <typename T>
__device__ __forceinline__ void pack_pair(cx<T> dp, cx<T> dq, cx<T> tw /* tw2N[p] */, cx<T>& zp, cx<T>& zq) {
  // E = (d_p + conj(d_q))/2 ; t = (d_p - conj(d_q))/2 = W^p O ; O = conj(W^p) t
  const cx<T> cq = cconj(dq);
  const cx<T> E = mk<T>(T(0.5) * (dp.x + cq.x), T(0.5) * (dp.y + cq.y));
  const cx<T> t = mk<T>(T(0.5) * (dp.x - cq.x), T(0.5) * (dp.y - cq.y));
  const cx<T> O = cmul(cconj(tw), t);
  // Z_p = E + i O ; Z_{N-p} = conj(E) + i conj(O) ... (E_{N-p} = conj E, O_{N-p} = conj O)
  zp = mk<T>(E.x - O.y, E.y + O.x);
  zq = mk<T>(E.x + O.y, -E.y + O.x);
}

}  // namespace ppb
