// CPU harness around csrc/nws_math.h (the header the CUDA kernels use) for tests/test_math_cpu.py.
#include "../../neural_waveshaping_synthesis_b200/csrc/nws_math.h"
#include <stdint.h>
#include <string.h>

extern "C" {

void h_sinf(const float* x, float* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = nws_sinf(x[i]);
}

// exhaustive check of the Markstein division against true fp32 division for every finite float
// with lo_limit <= |a| <= limit (or a == 0); returns the number of mismatches.
long h_div_check(float d, float lo_limit, float limit, long stride) {
  const float rcp = 1.0f / d;
  long bad = 0;
  for (uint64_t bits = 0; bits < (1ull << 32); bits += (uint64_t)stride) {
    uint32_t b = (uint32_t)bits;
    float a;
    memcpy(&a, &b, 4);
    if (!(a == a) || a > limit || a < -limit) continue;
    if (a != 0.0f && fabsf(a) < lo_limit) continue;  // quotient would be subnormal: outside the kernels' domain
    volatile float t = a / d;
    float q = nws_div_markstein(a, d, rcp);
    float tt = t;
    if (memcmp(&q, &tt, 4) != 0 && !(q == 0.0f && tt == 0.0f)) ++bad;
  }
  return bad;
}

// nws_lut_idx_pow2 against the index arithmetic of nws_lut_index (multiply by the size, then divide by the span)
// for every finite float |x| <= limit (stepping `stride` bit patterns); returns the number of mismatches.
long h_idx_pow2_check(int size, float tmin, float tmax, float limit, long stride) {
  const float span = tmax - tmin, rcp = 1.0f / span;
  const float span_s = span / (float)size, rcp_s = rcp * (float)size;
  long bad = 0;
  for (uint64_t bits = 0; bits < (1ull << 32); bits += (uint64_t)stride) {
    uint32_t b = (uint32_t)bits;
    float x;
    memcpy(&x, &b, 4);
    if (!(x == x) || x > limit || x < -limit) continue;
    const float ref = nws_div_markstein(NWS_MUL((float)size, NWS_ADD(x, -tmin)), span, rcp);
    const float got = nws_lut_idx_pow2(x, tmin, span_s, rcp_s);
    if (memcmp(&ref, &got, 4) != 0 && !(ref == 0.0f && got == 0.0f)) ++bad;
  }
  return bad;
}

void h_upsample(const float* x, int T, int hop, float* y) {
  const float inv = (float)T / (float)(T * hop);
  for (int n = 0; n < T * hop; ++n) {
    NwsLerp c = nws_lerp_coords(n, T, inv);
    y[n] = nws_lerp_apply(c, x[c.i0], x[c.i1]);
  }
}

// phase pipeline of one utterance: f0 frames [T] -> csum (fp32 of a double running sum), phase, arg for harmonic k
void h_phase(const float* f0, int T, int hop, float sample_rate, float* csum, float* phase) {
  const float inv = (float)T / (float)(T * hop);
  double acc = 0.0;
  for (int n = 0; n < T * hop; ++n) {
    NwsLerp c = nws_lerp_coords(n, T, inv);
    acc += (double)nws_lerp_apply(c, f0[c.i0], f0[c.i1]);
    csum[n] = (float)acc;
    phase[n] = nws_phase_from_cumsum(csum[n], sample_rate);
  }
}

void h_harmonic_arg(const float* phase, long n, int k, float shift, float* arg) {
  for (long i = 0; i < n; ++i) arg[i] = nws_harmonic_arg(k, phase[i], shift);
}

void h_phase_shift(const float* u, const float* rp, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = nws_phase_shift(u[i], rp[i]);
}

void h_lut_index(const float* x, long n, int size, float tmin, float tmax, int* lower, int* upper, float* fract) {
  const float span = tmax - tmin;
  const float rcp = 1.0f / span;
  for (long i = 0; i < n; ++i) {
    NwsLutIdx r = nws_lut_index(x[i], size, tmin, span, rcp);
    lower[i] = r.lower; upper[i] = r.upper; fract[i] = r.fract;
  }
}

void h_lut_lerp(const float* lo, const float* up, const float* fr, long n, float* out) {
  for (long i = 0; i < n; ++i) out[i] = nws_lut_lerp(lo[i], up[i], fr[i]);
}

void h_philox(uint64_t seed, uint64_t stream, uint64_t counter0, long n4, float* out) {
  for (long i = 0; i < n4; ++i) {
    NwsPhilox4 r = nws_philox4x32_10(counter0 + (uint64_t)i, stream, seed);
    for (int j = 0; j < 4; ++j) out[4 * i + j] = nws_u32_to_unit(r.v[j]);
  }
}

void h_philox_raw(uint64_t seed, uint64_t stream, uint64_t counter, uint32_t* out) {
  NwsPhilox4 r = nws_philox4x32_10(counter, stream, seed);
  for (int j = 0; j < 4; ++j) out[j] = r.v[j];
}
}
