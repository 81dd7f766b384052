// The noise branch's per-frame filter (FIRNoiseSynth.forward, modules/generators.py:21-35) as a CTA-level body shared
// by nws_noise_filter_kernel (nws_noise.cu) and the short-buffer MLP chain (nws_mlp_small.cu), which runs it on the band
// gains it has just computed, still in shared memory.
#pragma once
#include "nws_fft.cuh"
#include "nws_internal.cuh"

__device__ __forceinline__ void nws_load_tw256(float2* tw_s, const float2* __restrict__ tw_master, int tid, int nthreads) {
  for (int i = tid; i < 128; i += nthreads) tw_s[i] = tw_master[i * (kTwMaster / 256)];
}


// X[t][k] = rfft(xp[128 t : 128 t + 256])[k], xp = reflect_pad(noise, 128)   (generators.py:31).
// One CTA per pair of frames (all `nthreads` threads of the CTA must call).  `noise` null: the samples are this
// forward's own draw (generators.py:30), generated in place from the Philox stream (seed, offset) exactly as
// nws_rng_kernel would have written them.
__device__ __forceinline__ float nws_noise_sample(const float* __restrict__ noise, int idx, uint64_t seed, uint64_t offset) {
  if (noise) return noise[idx];
  const NwsPhilox4 r = nws_philox4x32_10(offset + (uint64_t)(idx >> 2), 1ull, seed);
  return nws_u32_to_unit(r.v[idx & 3]);
}

__device__ __forceinline__ void nws_noise_spectrum_body(int pair, int nthreads, const float* __restrict__ noise, int n_noise,
                                                        uint64_t seed, uint64_t offset, const float2* __restrict__ tw_master,
                                                        float2* __restrict__ xspec, int T) {
  __shared__ float2 buf_a[256], buf_b[256], tw_s[128];
  const int tid = threadIdx.x, ta = 2 * pair, tb = ta + 1;
  nws_load_tw256(tw_s, tw_master, tid, nthreads);
  for (int n = tid; n < 256; n += nthreads) {
    float v[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int t = q == 0 ? ta : tb;
      int idx = t * kHop + n - kIr / 2;               // index into the unpadded noise
      if (idx < 0) idx = -idx;                        // reflect (no edge repeat)
      if (idx >= n_noise) idx = 2 * (n_noise - 1) - idx;
      v[q] = t < T ? nws_noise_sample(noise, idx, seed, offset) : 0.f;
    }
    buf_a[n] = make_float2(v[0], v[1]);
  }
  __syncthreads();
  const float2* z = nws_fft_smem<false, false>(buf_a, buf_b, tw_s, 1, 8, 0, tid, nthreads);
  for (int k = tid; k <= 128; k += nthreads) {
    const float2 zk = z[k], zc = z[(256 - k) & 255];
    // Xa = (Z[k] + conj(Z[N-k])) / 2 ; Xb = (Z[k] - conj(Z[N-k])) / (2i)
    const float2 xa = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
    const float2 xb = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
    xspec[(size_t)ta * kBandsPad + k] = xa;
    if (tb < T) xspec[(size_t)tb * kBandsPad + k] = xb;
  }
}

// One CTA = one utterance x kNoiseHops output hops.  It filters frames t0-1 .. t0+kNoiseHops-1
// (kNoiseHops+1 frames = (kNoiseHops+1)/2 complex FFTs, two at a time), keeps their 256-sample
// outputs in shared memory and overlap-adds them into the kNoiseHops hops it owns.
constexpr int kNoiseHops = 15;
constexpr int kNoiseFrames = kNoiseHops + 1;

// `bands_b`: the utterance's band-gain rows [frame][kBandsPad] (global or shared memory); `out_b`: its output row.
// All 256 threads of the CTA must call (the FFT helper's barriers are CTA-wide); `blk` selects the block of hops.
__device__ __forceinline__ void nws_noise_filter_body(const float* bands_b, const float2* __restrict__ xspec,
                                                      const float2* __restrict__ tw_master, float* __restrict__ out_b,
                                                      int T, int hop_begin, int hop_end, int blk) {
  __shared__ float2 buf_a[2][256], buf_b[2][256], tw_s[128];
  __shared__ float hs[2][2][kBandsPad];           // [fft][frame of pair][band]
  __shared__ float y_s[kNoiseFrames][256];
  const int tid = threadIdx.x, t0 = hop_begin + blk * kNoiseHops;   // hops [hop_begin, hop_end)
  const int g = tid >> 7, j = tid & 127;          // g: which of the two concurrent FFTs
  const int f_lim = hop_end < T ? hop_end : T;    // frames >= hop_end are not needed (and may not be encoded yet)
  nws_load_tw256(tw_s, tw_master, tid, 256);

  for (int pair0 = 0; pair0 < kNoiseFrames / 2; pair0 += 2) {
    if (t0 - 1 + 2 * pair0 >= f_lim) break;   // short buffers: the remaining frame pairs are beyond the last hop (never read)
    const int pair = pair0 + g;
    const int fa = t0 - 1 + 2 * pair, fb = fa + 1;  // frame indices of this FFT's pair
    __syncthreads();
    for (int q = 0; q < 2; ++q) {
      const int f = q == 0 ? fa : fb;
      for (int k = j; k < kBandsPad; k += 128)
        hs[g][q][k] = (f >= 0 && f < f_lim && k < kBands) ? bands_b[(size_t)f * kBandsPad + k] : 0.f;
    }
    __syncthreads();
    // Z[k] = Ya[k] + i Yb[k], Y = X * Hw, Hermitian-extended to 256 bins
    for (int k = j; k < 256; k += 128) {
      const int kk = k <= 128 ? k : 256 - k;
      const int km = kk == 0 ? 1 : kk - 1, kp = kk == 128 ? 127 : kk + 1;
      const float sgn = (kk & 1) ? -1.f : 1.f;
      float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
      if (fa >= 0 && fa < f_lim) {
        const float hw = sgn * fmaf(0.25f, hs[g][0][km] + hs[g][0][kp], 0.5f * hs[g][0][kk]);
        const float2 x = xspec[(size_t)fa * kBandsPad + kk];
        ya = make_float2(x.x * hw, x.y * hw);
      }
      if (fb >= 0 && fb < f_lim) {
        const float hw = sgn * fmaf(0.25f, hs[g][1][km] + hs[g][1][kp], 0.5f * hs[g][1][kk]);
        const float2 x = xspec[(size_t)fb * kBandsPad + kk];
        yb = make_float2(x.x * hw, x.y * hw);
      }
      if (kk == 0 || kk == 128) { ya.y = 0.f; yb.y = 0.f; }  // irfft ignores the imaginary part of DC / Nyquist
      if (k > 128) { ya.y = -ya.y; yb.y = -yb.y; }            // conj for the mirrored half
      buf_a[g][k] = make_float2(ya.x - yb.y, ya.y + yb.x);
    }
    __syncthreads();
    // both FFTs advance in lock-step (the helper's barriers are CTA-wide)
    const float2* z = nws_fft_smem<true, false>(&buf_a[0][0], &buf_b[0][0], tw_s, 1, 8, 1, tid, 256);
    for (int n = j; n < 256; n += 128) {
      const float2 v = z[g * 256 + n];
      y_s[2 * pair][n] = v.x * (1.0f / 256.0f);
      y_s[2 * pair + 1][n] = v.y * (1.0f / 256.0f);
    }
  }
  __syncthreads();
  // overlap-add: hop t takes the first half of frame t and the second half of frame t-1, divided by
  // the number of overlapping frames (1 in the first hop, else 2)
  for (int i = tid; i < kNoiseHops * kHop; i += 256) {
    const int h = i >> 7, r = i & 127, t = t0 + h;
    if (t >= hop_end) break;
    const float cur = y_s[h + 1][r];
    const float v = t == 0 ? cur : 0.5f * (y_s[h][kHop + r] + cur);
    out_b[t * kHop + r] = v;
  }
}

