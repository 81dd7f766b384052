// Radix-2 Stockham FFT on complex data in shared memory, shared by the noise branch (256-point
// frames, generators.py:25-35) and the reverb (four-step FFT convolution, shaping.py:161-173).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float2 nws_cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// `n_fft` independent FFTs of length N = 1 << log_n, element e of FFT f at buf[(e * n_fft + f)]
// ("interleaved": consecutive threads work on consecutive FFTs -> conflict-free) when INTERLEAVED,
// else at buf[f * N + e].  All threads of the CTA must call; ping-pongs between a and b and
// returns the buffer holding the (natural-order) result.  tw[m] = exp(-2*pi*i*m/N), m < N/2
// (tw_stride lets a longer master table be used).
template <bool INVERSE, bool INTERLEAVED>
__device__ __forceinline__ float2* nws_fft_smem(float2* a, float2* b, const float2* __restrict__ tw, int tw_stride,
                                                int log_n, int n_fft, int tid, int n_threads) {
  const int N = 1 << log_n, half = N >> 1;
  const int total = half * n_fft;
  for (int s = 0; s < log_n; ++s) {
    const int ns = 1 << s;
    for (int q = tid; q < total; q += n_threads) {
      int f, j;
      if (INTERLEAVED) { f = q % n_fft; j = q / n_fft; } else { f = q / half; j = q - f * half; }
      const int k = j & (ns - 1);
      float2 w = tw[(k << (log_n - 1 - s)) * tw_stride];
      if (INVERSE) w.y = -w.y;
      const int i0 = INTERLEAVED ? j * n_fft + f : f * N + j;
      const int i1 = INTERLEAVED ? (j + half) * n_fft + f : f * N + j + half;
      const float2 v0 = a[i0];
      const float2 v1 = nws_cmul(a[i1], w);
      const int j0 = ((j >> s) << (s + 1)) + k;
      const int o0 = INTERLEAVED ? j0 * n_fft + f : f * N + j0;
      const int o1 = INTERLEAVED ? (j0 + ns) * n_fft + f : f * N + j0 + ns;
      b[o0] = make_float2(v0.x + v1.x, v0.y + v1.y);
      b[o1] = make_float2(v0.x - v1.x, v0.y - v1.y);
    }
    __syncthreads();
    float2* t = a; a = b; b = t;
  }
  return a;
}
