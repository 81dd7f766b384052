// Filtered-noise branch: FIRNoiseSynth.forward (modules/generators.py:21-35).
//
// The reference designs a zero-phase IR per frame (irfft of the 129 real band gains -> roll 128 ->
// periodic Hann), takes its rfft, multiplies the STFT of one shared uniform-noise vector
// (n_fft 256, hop 128, centre + reflect padding, rectangular window) and inverts with
// istft(center=False).  Two identities make this cheap without changing the result:
//   * roll by 128 multiplies bin k by (-1)^k and the Hann window is a 3-tap filter across bins, so
//     rfft(roll(irfft(H)) * hann)[k] = (-1)^k (0.5 H[k] + 0.25 (H[k-1] + H[k+1]))  (H even-extended)
//     — a REAL response: no transform is needed for the IR design;
//   * two real 256-point frames share one complex FFT (frame a in the real part, b in the imaginary).
// istft with a rectangular window is overlap-add divided by the frame-count envelope {1,2}.
#include "nws_fft.cuh"
#include "nws_hop_bodies.cuh"
#include "nws_internal.cuh"
#include "nws_noise_bodies.cuh"

__global__ void __launch_bounds__(128) nws_noise_spectrum_kernel(const float* __restrict__ noise, int n_noise,
                                                                 const float2* __restrict__ tw_master,
                                                                 float2* __restrict__ xspec, int T) {
  nws_noise_spectrum_body(blockIdx.x, 128, noise, n_noise, 0ull, 0ull, tw_master, xspec, T);
}

int nws_launch_noise_spectrum(const NwsContext* ctx, const float* noise, float2* xspec, int T, cudaStream_t s) {
  nws_noise_spectrum_kernel<<<(T + 1) / 2, 128, 0, s>>>(noise, kHop * T - 1, ctx->tw_master, xspec, T);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
// Front end of the short-buffer path in ONE launch: everything of a forward that does not depend on the encoder's
// output, next to the encoder itself.  CTA roles by block index:
//   [0, B)       the GRU recurrence of utterance b                                  (nws_gru_body)
//   [B, 2B)      the fp64 phase carries of utterance b: eight threads per hop (every partial sum of float32 values in
//                double is exact at audio-range f0, so any grouping gives the bits of nws_phase_carry_kernel)
//   [2B, ...)    one pair of noise frames each: the forward's draws (Philox, when none are injected) and the noise
//                spectrum; the first of them also writes the 101 phase-shift draws
// Replaces four launches (draws, carries, spectrum, GRU) of ~2.5-8.5 us each by one of ~8.5 us.
struct NwsFrontParams {
  const float *w_hh, *w_ih, *b_ih, *b_hh, *control;
  int ctrl_channels;
  float* hbuf;
  int B, T;
  const float* f0;
  double* carry;
  const float* noise_in;    // injected noise draw or null
  float* u_phase_out;       // where to write the phase-shift draw, or null (injected)
  uint64_t seed, offset;
  const float2* tw_master;
  float2* xspec;
};

__global__ void __launch_bounds__(kGates, 1) nws_front_kernel(const NwsFrontParams p) {
  const int role = blockIdx.x, tid = threadIdx.x, T = p.T;
  nws_pdl_launch();   // the MLP chain's weight prologue may start now (it waits for this grid before reading its results)
  if (role < p.B) {
    nws_gru_body(role, p.w_hh, p.w_ih, p.b_ih, p.b_hh, p.control, p.ctrl_channels, p.hbuf, T, 0, T, nullptr);
  } else if (role < 2 * p.B) {
    __shared__ double hop_sum[kSmallMlpMaxFrames + 8];
    const int b = role - p.B, t = tid >> 3, part = tid & 7;
    const float* f = p.f0 + (size_t)b * T;
    const float inv_hop = (float)T / (float)(T * kHop);
    double s = 0.0;
    if (t < T) {
      const float fm = f[t > 0 ? t - 1 : 0], fc = f[t], fp = f[t + 1 < T ? t + 1 : T - 1];
      for (int r = part * 16; r < part * 16 + 16; ++r) {
        const NwsLerp c = nws_lerp_coords(t * kHop + r, T, inv_hop);
        const float x0 = c.i0 == t ? fc : (c.i0 < t ? fm : fp);
        const float x1 = c.i1 == t ? fc : (c.i1 < t ? fm : fp);
        s += (double)nws_lerp_apply(c, x0, x1);
      }
    }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (t < T && part == 0) hop_sum[t] = s;
    __syncthreads();
    if (tid == 0) {
      double run = 0.0;
      for (int i = 0; i < T; ++i) {
        p.carry[(size_t)b * T + i] = run;
        run += hop_sum[i];
      }
    }
  } else {
    const int pair = role - 2 * p.B;
    if (pair == 0 && p.u_phase_out && tid < (kHarm + 3) / 4) {
      const NwsPhilox4 r = nws_philox4x32_10(p.offset + (uint64_t)tid, 0ull, p.seed);
      for (int j = 0; j < 4; ++j)
        if (4 * tid + j < kHarm) p.u_phase_out[4 * tid + j] = nws_u32_to_unit(r.v[j]);
    }
    nws_noise_spectrum_body(pair, kGates, p.noise_in, kHop * T - 1, p.seed, p.offset, p.tw_master, p.xspec, T);
  }
}

bool nws_front_ok(int B, int T) { return T >= 2 && T <= kSmallMlpMaxFrames && B >= 1 && B <= kSmallMlpMaxBatch; }

int nws_launch_front(const NwsContext* ctx, const float* control, int ctrl_channels, float* hbuf, const float* f0,
                     double* carry, const float* noise_in, float* u_phase_out, uint64_t seed, uint64_t offset,
                     float2* xspec, int B, int T, cudaStream_t s) {
  NwsFrontParams p{};
  const float* w = ctx->packed;
  p.w_hh = w + ctx->lay.gru_whh; p.w_ih = w + ctx->lay.gru_wih; p.b_ih = w + ctx->lay.gru_bih; p.b_hh = w + ctx->lay.gru_bhh;
  p.control = control; p.ctrl_channels = ctrl_channels; p.hbuf = hbuf; p.B = B; p.T = T;
  p.f0 = f0; p.carry = carry; p.noise_in = noise_in; p.u_phase_out = u_phase_out; p.seed = seed; p.offset = offset;
  p.tw_master = ctx->tw_master; p.xspec = xspec;
  nws_front_kernel<<<2 * B + (T + 1) / 2, kGates, 0, s>>>(p);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

__global__ void __launch_bounds__(256) nws_noise_filter_kernel(const float* __restrict__ bands,
                                                               const float2* __restrict__ xspec,
                                                               const float2* __restrict__ tw_master,
                                                               float* __restrict__ out, int T, int hop_begin,
                                                               int hop_end) {
  const int b = blockIdx.y;
  nws_noise_filter_body(bands + (size_t)b * T * kBandsPad, xspec, tw_master, out + (size_t)b * T * kHop, T, hop_begin, hop_end,
                        blockIdx.x);
}

int nws_launch_noise_filter(const NwsContext* ctx, const float* bands, const float2* xspec, float* out, int B, int T,
                            int hop_begin, int hop_end, cudaStream_t s) {
  dim3 grid((hop_end - hop_begin + kNoiseHops - 1) / kNoiseHops, B);
  nws_noise_filter_kernel<<<grid, 256, 0, s>>>(bands, xspec, ctx->tw_master, out, T, hop_begin, hop_end);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
