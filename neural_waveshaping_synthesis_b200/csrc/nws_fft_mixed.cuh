// Mixed-radix Stockham autosort FFT (radix-5 and radix-2 stages) for the two column lengths the reverb needs to
// run the reference's circular convolution at its exact length (modules/shaping.py:161-173: the transform
// length is max(N, 32000) samples — 64000 = 250 * 256 for the 4 s utterances of every benchmark config,
// 32000 = 125 * 256 for anything up to 2 s), instead of zero-padding to the next power of two (131072 / 65536).
//
// Plain C++ that compiles under nvcc (device code, all threads of the CTA call) and under g++
// (tests/cpu_harness/fft_harness.cpp runs the same stages serially against numpy, tests/test_math_cpu.py).
//
// Layout: W = 1 << log_w independent transforms, element e of transform f at buf[(e << log_w) + f]
// (consecutive threads -> consecutive transforms: conflict-free shared-memory columns).
// Stage recipe (Govindaraju et al.): T = N / R butterflies per transform; butterfly j reads a[j + r*T],
// multiplies by W_N^(r * k * N/(Ns*R)) with k = j mod Ns, and writes b[(j - k)*R + k + r*Ns]; Ns = product
// of the radices already applied.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define NWS_FM_HD __host__ __device__ __forceinline__
#else
#include <math.h>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#define NWS_FM_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define NWS_FM_SYNC() __syncthreads()
#else
#define NWS_FM_SYNC() ((void)0)
#endif

NWS_FM_HD float2 nws_fm_cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
NWS_FM_HD float2 nws_fm_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
NWS_FM_HD float2 nws_fm_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * (-i) for the forward transform, a * (+i) for the inverse
template <bool INVERSE>
NWS_FM_HD float2 nws_fm_rot(float2 a) { return INVERSE ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <int R, bool INVERSE>
struct NwsButterfly;

template <bool INVERSE>
struct NwsButterfly<2, INVERSE> {
  static NWS_FM_HD void run(float2* v) {
    const float2 a = v[0], b = v[1];
    v[0] = nws_fm_add(a, b);
    v[1] = nws_fm_sub(a, b);
  }
};

// X_m = sum_n v_n w^(mn), w = exp(-+2*pi*i/5):
//   X_1,4 = v0 + c1*t1 + c2*t2 -+ i*(s1*t3 + s2*t4),  X_2,3 = v0 + c2*t1 + c1*t2 -+ i*(s2*t3 - s1*t4)
// with t1 = v1+v4, t2 = v2+v3, t3 = v1-v4, t4 = v2-v3, c_m = cos(2*pi*m/5), s_m = sin(2*pi*m/5).
template <bool INVERSE>
struct NwsButterfly<5, INVERSE> {
  static NWS_FM_HD void run(float2* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 t1 = nws_fm_add(v[1], v[4]), t2 = nws_fm_add(v[2], v[3]);
    const float2 t3 = nws_fm_sub(v[1], v[4]), t4 = nws_fm_sub(v[2], v[3]);
    const float2 m1 = make_float2(fmaf(c2, t2.x, fmaf(c1, t1.x, v[0].x)), fmaf(c2, t2.y, fmaf(c1, t1.y, v[0].y)));
    const float2 m2 = make_float2(fmaf(c1, t2.x, fmaf(c2, t1.x, v[0].x)), fmaf(c1, t2.y, fmaf(c2, t1.y, v[0].y)));
    const float2 r1 = nws_fm_rot<INVERSE>(make_float2(fmaf(s2, t4.x, s1 * t3.x), fmaf(s2, t4.y, s1 * t3.y)));
    const float2 r2 = nws_fm_rot<INVERSE>(make_float2(fmaf(-s1, t4.x, s2 * t3.x), fmaf(-s1, t4.y, s2 * t3.y)));
    v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
    v[1] = nws_fm_add(m1, r1);
    v[4] = nws_fm_sub(m1, r1);
    v[2] = nws_fm_add(m2, r2);
    v[3] = nws_fm_sub(m2, r2);
  }
};

// One Stockham stage.  tw[m] = exp(-2*pi*i*m/N), m < N.
template <int N, int R, int NS, bool INVERSE>
NWS_FM_HD void nws_fft_mixed_stage(const float2* a, float2* b, const float2* tw, int log_w, int tid, int n_threads) {
  constexpr int T = N / R, STEP = N / (NS * R);
  const int total = T << log_w, wmask = (1 << log_w) - 1;
  for (int q = tid; q < total; q += n_threads) {
    const int f = q & wmask, j = q >> log_w;
    const int k = NS == 1 ? 0 : j % NS;   // compile-time divisor
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = a[((j + r * T) << log_w) + f];
    if (NS > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        float2 w = tw[r * k * STEP];   // r*k*STEP <= (R-1)(NS-1)N/(NS*R) < N
        if (INVERSE) w.y = -w.y;
        v[r] = nws_fm_cmul(v[r], w);
      }
    }
    NwsButterfly<R, INVERSE>::run(v);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) b[((j0 + r * NS) << log_w) + f] = v[r];
  }
}

// Column transform of length N1 in {125, 250}: returns the buffer (a or b) holding the natural-order result.
// On the device every thread of the CTA must call (the stages are separated by __syncthreads); the input in
// `a` must be visible to the CTA on entry and the result is visible on return.
template <int N1, bool INVERSE>
NWS_FM_HD float2* nws_fft_mixed(float2* a, float2* b, const float2* tw, int log_w, int tid, int n_threads) {
  static_assert(N1 == 125 || N1 == 250, "column lengths of the exact-length reverb plans");
  nws_fft_mixed_stage<N1, 5, 1, INVERSE>(a, b, tw, log_w, tid, n_threads);
  NWS_FM_SYNC();
  nws_fft_mixed_stage<N1, 5, 5, INVERSE>(b, a, tw, log_w, tid, n_threads);
  NWS_FM_SYNC();
  nws_fft_mixed_stage<N1, 5, 25, INVERSE>(a, b, tw, log_w, tid, n_threads);
  NWS_FM_SYNC();
  if constexpr (N1 == 250) {
    nws_fft_mixed_stage<N1, 2, 125, INVERSE>(b, a, tw, log_w, tid, n_threads);
    NWS_FM_SYNC();
    return a;
  }
  return b;
}
