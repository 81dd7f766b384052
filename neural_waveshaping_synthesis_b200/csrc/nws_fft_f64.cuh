// Power-of-two Stockham autosort FFT (radix-4 stages, one radix-2 stage when log2 N is odd) on double-precision
// complex data — the STFT of the control-side loudness extractor (data/utils/loudness_extraction.py:11-21),
// where the reference's transform is numpy's float64 FFT (librosa 0.8.0 stft) and B200's FP64 pipe makes
// matching it cheap.  Same stage recipe as nws_fft.cuh; one transform per call.
//
// Plain C++ that compiles under nvcc (device code, all threads of the CTA call) and under g++
// (tests/cpu_harness/fft_harness.cpp runs it serially against numpy).
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define NWS_F64_HD __host__ __device__ __forceinline__
#else
#include <math.h>
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#define NWS_F64_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define NWS_F64_SYNC() __syncthreads()
#else
#define NWS_F64_SYNC() ((void)0)
#endif

NWS_F64_HD double2 nws_zmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
NWS_F64_HD double2 nws_zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
NWS_F64_HD double2 nws_zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// tw[m] = exp(-2*pi*i*m/N) for m < N/2; the second half of the circle is -tw[m - N/2]
NWS_F64_HD double2 nws_ztwiddle(const double2* tw, int m, int half) {
  double2 w = m < half ? tw[m] : tw[m - half];
  if (m >= half) { w.x = -w.x; w.y = -w.y; }
  return w;
}

// Forward transform of length N = 1 << log_n, natural order in, natural order out; ping-pongs between a and b
// and returns the buffer holding the result.  On the device every thread of the CTA must call; the input must
// be visible on entry and the result is visible on return.
NWS_F64_HD double2* nws_fft_f64(double2* a, double2* b, const double2* tw, int log_n, int tid, int n_threads) {
  const int N = 1 << log_n, half = N >> 1;
  int ns = 1, s = 0;
  while (s < log_n) {
    if (log_n - s >= 2) {
      const int T = N >> 2, tw_step = N >> (2 + s);   // N / (4 * ns)
      for (int j = tid; j < T; j += n_threads) {
        const int k = j & (ns - 1);
        double2 v0 = a[j], v1 = a[j + T], v2 = a[j + 2 * T], v3 = a[j + 3 * T];
        if (k) {
          v1 = nws_zmul(v1, nws_ztwiddle(tw, k * tw_step, half));
          v2 = nws_zmul(v2, nws_ztwiddle(tw, 2 * k * tw_step, half));
          v3 = nws_zmul(v3, nws_ztwiddle(tw, 3 * k * tw_step, half));
        }
        const double2 a0 = nws_zadd(v0, v2), a1 = nws_zsub(v0, v2), a2 = nws_zadd(v1, v3), d = nws_zsub(v1, v3);
        const double2 a3 = make_double2(d.y, -d.x);   // * (-i)
        const int j0 = ((j & ~(ns - 1)) << 2) + k;
        b[j0] = nws_zadd(a0, a2);
        b[j0 + ns] = nws_zadd(a1, a3);
        b[j0 + 2 * ns] = nws_zsub(a0, a2);
        b[j0 + 3 * ns] = nws_zsub(a1, a3);
      }
      ns <<= 2;
      s += 2;
    } else {
      for (int j = tid; j < half; j += n_threads) {
        const int k = j & (ns - 1);
        const double2 v0 = a[j], v1 = nws_zmul(a[j + half], nws_ztwiddle(tw, k << (log_n - 1 - s), half));
        const int j0 = ((j & ~(ns - 1)) << 1) + k;
        b[j0] = nws_zadd(v0, v1);
        b[j0 + ns] = nws_zsub(v0, v1);
      }
      ns <<= 1;
      s += 1;
    }
    NWS_F64_SYNC();
    double2* t = a; a = b; b = t;
  }
  return a;
}
