// Pieces shared by the two implementations of the fused audio-rate kernel (nws_audio.cu: fp32 SIMT
// harmonic mixer; nws_audio_tc.cu: tcgen05 harmonic mixer).
#pragma once
#include "nws_internal.cuh"

// Sine used inside the shaper MLP (1,600 evaluations per sample on the NEWT path): SFU-based with a
// full-turn reduction (nws_sinf_turn).  The LUT builder uses the accurate version so
// FastNEWT tables match the reference's to 1 ulp-level (see nws_build_lut_kernel).
// MODE 0: polynomial sine everywhere (LUT builder).  MODE 1: SFU sine with range reduction everywhere.
// MODE 2: as 1 for the first layer (its argument scales with the input), and bare sin.approx for layers 2-4,
// whose arguments are bounded by max_j(|b_j| + sum_i |W_ji|) because the previous layer's outputs are sines —
// the bound is checked when the weights are loaded (nws_load_weights) and MODE 2 is used only if it is <= 8
// (3.8 at most in the shipped checkpoints), where x/2pi keeps the error at the SFU's own 4e-7 level.
#define NWS_SHAPER_SIN(x) (MODE == 0 ? nws_sinf(x) : nws_sinf_turn(x))
#define NWS_SHAPER_SIN_INNER(x) (MODE == 0 ? nws_sinf(x) : (MODE == 1 ? nws_sinf_turn(x) : __sinf(x)))

struct NwsAudioParams {
  const float* f0;        // [B][T]
  const double* carry;    // [B][T]
  const float* film;      // [B*T][256] frame-major
  const float* u_phase;   // [101]
  const float* hmix_wt;   // [104][64]
  const float* hmix_b;    // [64]
  const float* rand_phase;// [104]
  const float* shaper;    // [64][176]
  const float* mix_w;     // [64]
  const float* mix_b;     // [1]
  const float* lut;       // [64][lut_size]
  const float2* lut2;     // [64][lut_size] (value, forward difference) pairs
  int lut_size;
  float lut_min, lut_span, lut_span_rcp;
  const float* noise_in;  // [B][N] or null: added to the mixdown (neural_waveshaping.py:85-86)
  // Noise branch inside the kernel (nws_audio_tc.cu; `noise_in` is ignored then): the band gains of the noise MLP,
  // the spectrum of the shared noise vector and the twiddle table — FIRNoiseSynth.forward, generators.py:21-35
  const float* bands;     // [B*T][kBandsPad] frame-major, or null
  const float2* xspec;    // [T][kBandsPad]
  const float2* tw_master;// [kTwMaster / 2]
  float* out;             // [B][N]
  float* exciter_out;     // [B][64][N] or null
  int B, T;
  int t_begin, t_end;     // hop range [t_begin, t_end) rendered by this launch (whole utterance: 0, T)
  int* tile_counter;      // {tiles claimed, CTAs done}: both zero at launch, reset by the last CTA (dynamic tile scheduler)
  int tile_chunk;         // tiles per scheduler claim (consecutive hops), >= 1
  uint32_t hops_magic;    // floor(2^32 / (t_end - t_begin)), saturated: tile -> (utterance, hop) without a division
};


// ------------------------------------------------------------------------------------------------
// One shaper's sine-MLP (TrainableNonlinearity.forward, shaping.py:36-37 with Sine, depth 4, width 8):
// y = sin(w4 . sin(W3 sin(W2 sin(w1*(s*x) + b1) + b2) + b3) + b4).  `wp` = packed record (kShp* offsets).
template <int MODE>
__device__ __forceinline__ float nws_shaper_mlp(const float* __restrict__ wp, float x) {
  const float4 hd = *reinterpret_cast<const float4*>(wp);
  const float u = hd.x * x;
  float h1[8], h2[8];
  {
    const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW1), wb = *reinterpret_cast<const float4*>(wp + kShpW1 + 4);
    const float4 ba = *reinterpret_cast<const float4*>(wp + kShpB1), bb = *reinterpret_cast<const float4*>(wp + kShpB1 + 4);
    h1[0] = NWS_SHAPER_SIN(fmaf(wa.x, u, ba.x)); h1[1] = NWS_SHAPER_SIN(fmaf(wa.y, u, ba.y));
    h1[2] = NWS_SHAPER_SIN(fmaf(wa.z, u, ba.z)); h1[3] = NWS_SHAPER_SIN(fmaf(wa.w, u, ba.w));
    h1[4] = NWS_SHAPER_SIN(fmaf(wb.x, u, bb.x)); h1[5] = NWS_SHAPER_SIN(fmaf(wb.y, u, bb.y));
    h1[6] = NWS_SHAPER_SIN(fmaf(wb.z, u, bb.z)); h1[7] = NWS_SHAPER_SIN(fmaf(wb.w, u, bb.w));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW2 + j * 8), wb = *reinterpret_cast<const float4*>(wp + kShpW2 + j * 8 + 4);
    float a = wp[kShpB2 + j];
    a = fmaf(wa.x, h1[0], a); a = fmaf(wa.y, h1[1], a); a = fmaf(wa.z, h1[2], a); a = fmaf(wa.w, h1[3], a);
    a = fmaf(wb.x, h1[4], a); a = fmaf(wb.y, h1[5], a); a = fmaf(wb.z, h1[6], a); a = fmaf(wb.w, h1[7], a);
    h2[j] = NWS_SHAPER_SIN_INNER(a);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW3 + j * 8), wb = *reinterpret_cast<const float4*>(wp + kShpW3 + j * 8 + 4);
    float a = wp[kShpB3 + j];
    a = fmaf(wa.x, h2[0], a); a = fmaf(wa.y, h2[1], a); a = fmaf(wa.z, h2[2], a); a = fmaf(wa.w, h2[3], a);
    a = fmaf(wb.x, h2[4], a); a = fmaf(wb.y, h2[5], a); a = fmaf(wb.z, h2[6], a); a = fmaf(wb.w, h2[7], a);
    h1[j] = NWS_SHAPER_SIN_INNER(a);
  }
  const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW4), wb = *reinterpret_cast<const float4*>(wp + kShpW4 + 4);
  float a = hd.y;
  a = fmaf(wa.x, h1[0], a); a = fmaf(wa.y, h1[1], a); a = fmaf(wa.z, h1[2], a); a = fmaf(wa.w, h1[3], a);
  a = fmaf(wb.x, h1[4], a); a = fmaf(wb.y, h1[5], a); a = fmaf(wb.z, h1[6], a); a = fmaf(wb.w, h1[7], a);
  return NWS_SHAPER_SIN_INNER(a);
}


// The same sine-MLP for TWO samples at once with one set of weight loads: the shaper weights depend on the channel
// only, so a thread that evaluates channel c for two samples halves the shared-memory traffic per shaper-sample
// (the fused kernel pairs neighbouring lanes: each lane takes one channel of a channel pair for both lanes'
// samples, see nws_audio_tc.cu).  Identical arithmetic per sample (same fma chains as nws_shaper_mlp).
template <int MODE>
__device__ __forceinline__ void nws_shaper_mlp2(const float* __restrict__ wp, float xa, float xb, float& ya, float& yb) {
  const float4 hd = *reinterpret_cast<const float4*>(wp);
  const float ua = hd.x * xa, ub = hd.x * xb;
  float h1a[8], h1b[8], h2a[8], h2b[8];
  {
    const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW1), wb = *reinterpret_cast<const float4*>(wp + kShpW1 + 4);
    const float4 ba = *reinterpret_cast<const float4*>(wp + kShpB1), bb = *reinterpret_cast<const float4*>(wp + kShpB1 + 4);
    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    const float b[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h1a[j] = NWS_SHAPER_SIN(fmaf(w[j], ua, b[j]));
      h1b[j] = NWS_SHAPER_SIN(fmaf(w[j], ub, b[j]));
    }
  }
  {
    const float4 b0 = *reinterpret_cast<const float4*>(wp + kShpB2), b1 = *reinterpret_cast<const float4*>(wp + kShpB2 + 4);
    const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW2 + j * 8), wb = *reinterpret_cast<const float4*>(wp + kShpW2 + j * 8 + 4);
      float a = bias[j], b = bias[j];
      a = fmaf(wa.x, h1a[0], a); a = fmaf(wa.y, h1a[1], a); a = fmaf(wa.z, h1a[2], a); a = fmaf(wa.w, h1a[3], a);
      a = fmaf(wb.x, h1a[4], a); a = fmaf(wb.y, h1a[5], a); a = fmaf(wb.z, h1a[6], a); a = fmaf(wb.w, h1a[7], a);
      b = fmaf(wa.x, h1b[0], b); b = fmaf(wa.y, h1b[1], b); b = fmaf(wa.z, h1b[2], b); b = fmaf(wa.w, h1b[3], b);
      b = fmaf(wb.x, h1b[4], b); b = fmaf(wb.y, h1b[5], b); b = fmaf(wb.z, h1b[6], b); b = fmaf(wb.w, h1b[7], b);
      h2a[j] = NWS_SHAPER_SIN_INNER(a);
      h2b[j] = NWS_SHAPER_SIN_INNER(b);
    }
  }
  {
    const float4 b0 = *reinterpret_cast<const float4*>(wp + kShpB3), b1 = *reinterpret_cast<const float4*>(wp + kShpB3 + 4);
    const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW3 + j * 8), wb = *reinterpret_cast<const float4*>(wp + kShpW3 + j * 8 + 4);
      float a = bias[j], b = bias[j];
      a = fmaf(wa.x, h2a[0], a); a = fmaf(wa.y, h2a[1], a); a = fmaf(wa.z, h2a[2], a); a = fmaf(wa.w, h2a[3], a);
      a = fmaf(wb.x, h2a[4], a); a = fmaf(wb.y, h2a[5], a); a = fmaf(wb.z, h2a[6], a); a = fmaf(wb.w, h2a[7], a);
      b = fmaf(wa.x, h2b[0], b); b = fmaf(wa.y, h2b[1], b); b = fmaf(wa.z, h2b[2], b); b = fmaf(wa.w, h2b[3], b);
      b = fmaf(wb.x, h2b[4], b); b = fmaf(wb.y, h2b[5], b); b = fmaf(wb.z, h2b[6], b); b = fmaf(wb.w, h2b[7], b);
      h1a[j] = NWS_SHAPER_SIN_INNER(a);
      h1b[j] = NWS_SHAPER_SIN_INNER(b);
    }
  }
  const float4 wa = *reinterpret_cast<const float4*>(wp + kShpW4), wb = *reinterpret_cast<const float4*>(wp + kShpW4 + 4);
  float a = hd.y, b = hd.y;
  a = fmaf(wa.x, h1a[0], a); a = fmaf(wa.y, h1a[1], a); a = fmaf(wa.z, h1a[2], a); a = fmaf(wa.w, h1a[3], a);
  a = fmaf(wb.x, h1a[4], a); a = fmaf(wb.y, h1a[5], a); a = fmaf(wb.z, h1a[6], a); a = fmaf(wb.w, h1a[7], a);
  b = fmaf(wa.x, h1b[0], b); b = fmaf(wa.y, h1b[1], b); b = fmaf(wa.z, h1b[2], b); b = fmaf(wa.w, h1b[3], b);
  b = fmaf(wb.x, h1b[4], b); b = fmaf(wb.y, h1b[5], b); b = fmaf(wb.z, h1b[6], b); b = fmaf(wb.w, h1b[7], b);
  ya = NWS_SHAPER_SIN_INNER(a);
  yb = NWS_SHAPER_SIN_INNER(b);
}
