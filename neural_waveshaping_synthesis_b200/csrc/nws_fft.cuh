// Stockham autosort FFT on complex data in shared memory (radix-4 stages, one radix-2 stage when
// log2 N is odd), shared by the noise branch (256-point frames, generators.py:25-35) and the reverb
// (four-step FFT convolution, shaping.py:161-173).  Index math validated against numpy in
// scripts/dev notes (mixed radix Stockham, Govindaraju et al. formulation).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float2 nws_cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 nws_cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 nws_csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// tw[m * tw_stride] = exp(-2*pi*i*m/N) for m < N/2; the second half of the circle is -tw[m - N/2]
template <bool INVERSE>
__device__ __forceinline__ float2 nws_twiddle(const float2* __restrict__ tw, int tw_stride, int m, int half) {
  float2 w = m < half ? tw[m * tw_stride] : tw[(m - half) * tw_stride];
  if (m >= half) { w.x = -w.x; w.y = -w.y; }
  if (INVERSE) w.y = -w.y;
  return w;
}

// `n_fft` = 1 << log_nfft independent FFTs of length N = 1 << log_n (every count is a power of two, so all index
// arithmetic is shifts and masks — runtime integer divisions made these kernels instruction-bound).  Element e of FFT f lives at buf[e * n_fft + f] when
// INTERLEAVED (consecutive threads -> consecutive FFTs: conflict-free for column transforms), else at
// buf[f * N + e].  All `n_threads` threads of the group must call (`sync` is the group's barrier: the CTA's, or a
// warpgroup's named barrier); ping-pongs between a and b and returns the buffer holding the natural-order result.
template <bool INVERSE, bool INTERLEAVED, class Sync>
__device__ __forceinline__ float2* nws_fft_smem_sync(float2* a, float2* b, const float2* __restrict__ tw, int tw_stride,
                                                     int log_n, int log_nfft, int tid, int n_threads, Sync sync) {
  const int N = 1 << log_n, half = N >> 1, n_fft = 1 << log_nfft;
  int ns = 1, s = 0;
  while (s < log_n) {
    if (log_n - s >= 2) {
      const int T = N >> 2, total = T << log_nfft, log_T = log_n - 2;
      const int tw_step = N >> (2 + s);   // N / (4 * ns), ns = 1 << s
      for (int q = tid; q < total; q += n_threads) {
        int f, j;
        if (INTERLEAVED) { f = q & (n_fft - 1); j = q >> log_nfft; } else { f = q >> log_T; j = q & (T - 1); }
        const int k = j & (ns - 1);
        float2 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = a[INTERLEAVED ? ((j + r * T) << log_nfft) + f : (f << log_n) + j + r * T];
        if (k) {
#pragma unroll
          for (int r = 1; r < 4; ++r) v[r] = nws_cmul(v[r], nws_twiddle<INVERSE>(tw, tw_stride, r * k * tw_step, half));
        }
        const float2 a0 = nws_cadd(v[0], v[2]), a1 = nws_csub(v[0], v[2]), a2 = nws_cadd(v[1], v[3]);
        const float2 d = nws_csub(v[1], v[3]);
        const float2 a3 = INVERSE ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // * (+i) or * (-i)
        const int j0 = ((j & ~(ns - 1)) << 2) + k;
        float2 o[4] = {nws_cadd(a0, a2), nws_cadd(a1, a3), nws_csub(a0, a2), nws_csub(a1, a3)};
#pragma unroll
        for (int r = 0; r < 4; ++r) b[INTERLEAVED ? ((j0 + r * ns) << log_nfft) + f : (f << log_n) + j0 + r * ns] = o[r];
      }
      ns <<= 2;
      s += 2;
    } else {
      const int total = half << log_nfft;
      for (int q = tid; q < total; q += n_threads) {
        int f, j;
        if (INTERLEAVED) { f = q & (n_fft - 1); j = q >> log_nfft; } else { f = q >> (log_n - 1); j = q & (half - 1); }
        const int k = j & (ns - 1);
        const float2 w = nws_twiddle<INVERSE>(tw, tw_stride, k << (log_n - 1 - s), half);   // k * N / (2 * ns)
        const float2 v0 = a[INTERLEAVED ? (j << log_nfft) + f : (f << log_n) + j];
        const float2 v1 = nws_cmul(a[INTERLEAVED ? ((j + half) << log_nfft) + f : (f << log_n) + j + half], w);
        const int j0 = ((j & ~(ns - 1)) << 1) + k;
        b[INTERLEAVED ? (j0 << log_nfft) + f : (f << log_n) + j0] = nws_cadd(v0, v1);
        b[INTERLEAVED ? ((j0 + ns) << log_nfft) + f : (f << log_n) + j0 + ns] = nws_csub(v0, v1);
      }
      ns <<= 1;
      s += 1;
    }
    sync();
    float2* t = a; a = b; b = t;
  }
  return a;
}

template <bool INVERSE, bool INTERLEAVED>
__device__ __forceinline__ float2* nws_fft_smem(float2* a, float2* b, const float2* __restrict__ tw, int tw_stride,
                                                int log_n, int log_nfft, int tid, int n_threads) {
  return nws_fft_smem_sync<INVERSE, INTERLEAVED>(a, b, tw, tw_stride, log_n, log_nfft, tid, n_threads, [] { __syncthreads(); });
}
