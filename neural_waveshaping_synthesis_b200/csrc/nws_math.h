// Scalar numerics shared by every kernel of the NWS hot path.  Plain C++ that compiles both
// under nvcc (device code) and under g++ (tests/test_math_cpu.py builds a CPU harness from this
// same header), so the numerically delicate recipes are unit-tested without a GPU.
//
// Reference semantics restated here (file:line relative to the reference repo):
//   * linear x128 upsample, align_corners=False  — models/neural_waveshaping.py:75, modules/shaping.py:69
//   * oscillator phase pipeline                  — modules/generators.py:58-66
//   * FastNEWT index arithmetic                  — modules/shaping.py:136-151
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NWS_HD __host__ __device__ __forceinline__
#else
#define NWS_HD inline
#endif

// IEEE single-rounding primitives that the compiler must never contract into FMAs.
#if defined(__CUDA_ARCH__)
#define NWS_MUL(a, b) __fmul_rn((a), (b))
#define NWS_ADD(a, b) __fadd_rn((a), (b))
#define NWS_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
static inline float nws_opaque(float x) { volatile float v = x; return v; }
#define NWS_MUL(a, b) nws_opaque((a) * (b))
#define NWS_ADD(a, b) nws_opaque((a) + (b))
#define NWS_FMA(a, b, c) fmaf((a), (b), (c))
#endif

#define NWS_TAU_F 6.28318530717958647692f   // fp32(math.tau)
#define NWS_PI_F 3.14159265358979323846f    // fp32(math.pi)

// ---------------------------------------------------------------------------------------------
// Linear upsample by `hop`, align_corners=False (torch area_pixel_compute_source_index):
// src = max(scale*(n+0.5)-0.5, 0), i0=(int)src, i1=min(i0+1,T-1), l1=src-i0, l0=1-l1,
// out = fmaf(l0, x[i0], l1*x[i1]).  `inv_hop` = fp32(T)/fp32(N) = 1/hop exactly for hop = 2^k.
struct NwsLerp {
  int i0, i1;
  float l0, l1;
};

NWS_HD NwsLerp nws_lerp_coords(int n, int T, float inv_hop) {
  float src = NWS_ADD(NWS_MUL(inv_hop, NWS_ADD((float)n, 0.5f)), -0.5f);
  src = src < 0.0f ? 0.0f : src;
  NwsLerp c;
  c.i0 = (int)src;
  c.i1 = c.i0 + 1 < T ? c.i0 + 1 : T - 1;
  c.l1 = NWS_ADD(src, -(float)c.i0);
  c.l0 = NWS_ADD(1.0f, -c.l1);
  return c;
}

NWS_HD float nws_lerp_apply(const NwsLerp& c, float x0, float x1) {
  return NWS_FMA(c.l0, x0, NWS_MUL(c.l1, x1));
}

// ---------------------------------------------------------------------------------------------
// Oscillator phase (generators.py:59): phase = (fp32(tau) * c) / 16000 — two roundings.
NWS_HD float nws_phase_from_cumsum(float c, float sample_rate) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(__fmul_rn(NWS_TAU_F, c), sample_rate);
#else
  return nws_opaque(nws_opaque(NWS_TAU_F * c) / sample_rate);
#endif
}

// arg_k = fp32(k)*phase + shift_k  (generators.py:60-61): separate multiply and add.
NWS_HD float nws_harmonic_arg(int k, float phase, float shift) { return NWS_ADD(NWS_MUL((float)k, phase), shift); }

// shift_k = u*rand_phase - pi (generators.py:55).
NWS_HD float nws_phase_shift(float u, float rand_phase) { return NWS_ADD(NWS_MUL(u, rand_phase), -NWS_PI_F); }

// ---------------------------------------------------------------------------------------------
// sin(x) for |x| up to ~2^23, <= ~1.5 ulp, branch-free (no Payne-Hanek slow path: oscillator
// arguments reach 1e5..1e6 rad, where CUDA's sinf() would diverge into its slow path).
// 3-term Cody-Waite reduction to [-pi/4, pi/4] by pi/2 (first FMA exact for |q| < 2^22), then the
// classic minimax sin/cos kernels selected by quadrant.
NWS_HD float nws_sinf(float x) {
  const float q = rintf(x * 0.63661977236758134308f);  // x * 2/pi
  float r = NWS_FMA(-q, 1.57079625129699707031e+00f, x);
  r = NWS_FMA(-q, 7.54978941586159635335e-08f, r);
  r = NWS_FMA(-q, 5.39030285815811905290e-15f, r);
  const int n = (int)q;
  const float z = r * r;
  // sin kernel: r + r*z*(S1 + z*(S2 + z*S3));  cos kernel: 1 + z*(C0 + z*(C1 + z*(C2 + z*C3)))
  const float ps = NWS_FMA(NWS_FMA(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
  const float s = NWS_FMA(ps * z, r, r);
  const float pc = NWS_FMA(NWS_FMA(NWS_FMA(2.443315711809948e-5f, z, -1.388731625493765e-3f), z,
                                   4.166664568298827e-2f), z, -0.5f);
  const float c = NWS_FMA(pc, z, 1.0f);
  float v = (n & 1) ? c : s;
  return (n & 2) ? -v : v;
}

// Same reduction, then the SFU: sin/cos.approx on the reduced argument |r| <= pi/4 (where MUFU's absolute
// error is smallest), selected by quadrant.  ~11 instructions instead of ~21; absolute error measured on
// B200 by tests/test_gpu_parity.py::test_sin_variants.  The quadrant comes from the "magic number" form of
// rint (x*2/pi + 1.5*2^23: the integer lands in the low mantissa bits), which keeps F2I/FRND off the
// quarter-rate XU pipe.  TERMS = 3 for oscillator arguments (up to ~1e6 rad), 2 for |x| < ~1e3.
template <int TERMS>
NWS_HD float nws_sinf_fast(float x) {
  const float magic = 12582912.0f;  // 1.5 * 2^23
  const float t = NWS_FMA(x, 0.63661977236758134308f, magic);
  const float q = NWS_ADD(t, -magic);
  float r = NWS_FMA(-q, 1.57079625129699707031e+00f, x);
  r = NWS_FMA(-q, 7.54978941586159635335e-08f, r);
  if (TERMS > 2) r = NWS_FMA(-q, 5.39030285815811905290e-15f, r);
#if defined(__CUDA_ARCH__)
  const unsigned n = __float_as_uint(t);
  const float sv = __sinf(r), cv = __cosf(r);
  const float v = (n & 1u) ? cv : sv;
  return __uint_as_float(__float_as_uint(v) ^ ((n & 2u) << 30));
#else
  const unsigned n = (unsigned)(int)q;
  const float v = (n & 1u) ? cosf(r) : sinf(r);
  return (n & 2u) ? -v : v;
#endif
}

// SFU sine with a full-turn reduction: q = rint(x / 2pi) by the same magic-number trick, r = x - q*2pi in two
// FMAs (2pi = C1 + C2, |r| <= pi), one MUFU.SIN.  No quadrant select, half the SFU work of nws_sinf_fast;
// the reduction rounds at |r| ~ pi instead of pi/4, so the worst case is a little looser (measured by
// nws_selftest_sin, tests/test_gpu_parity.py::test_sin_variants).  Valid for |x| < 2^22 * 2pi = 2.6e7 rad.
NWS_HD float nws_sinf_turn(float x) {
  const float magic = 12582912.0f;  // 1.5 * 2^23
  const float t = NWS_FMA(x, 0.15915494309189533577f, magic);
  const float q = NWS_ADD(t, -magic);
  float r = NWS_FMA(-q, 6.28318548202514648438f, x);       // C1 = fp32(2pi)
  r = NWS_FMA(-q, -1.7484556000744883e-07f, r);            // C2 = 2pi - C1
#if defined(__CUDA_ARCH__)
  return __sinf(r);
#else
  return sinf(r);
#endif
}

// ---------------------------------------------------------------------------------------------
// FastNEWT index arithmetic (shaping.py:137-146), bit-exact with torch CPU:
//   idx = (table_size * (x - table_min)) / (table_max - table_min)      [fp32 sub, mul, TRUE division]
//   lower = clamp(floor(idx), 0, size-1); upper = min(lower+1, size-1); fract = idx - (float)lower
// The division is done without a divide instruction: q0 = a*R, rem = fma(-q0, d, a), q = fma(rem, R, q0)
// with R = RN(1/d) is the correctly rounded a/d (Markstein) — checked exhaustively for d = 6 by
// tests/test_math_cpu.py.  `span_rcp` must be RN(1/span).
struct NwsLutIdx {
  int lower, upper;
  float fract;
};

NWS_HD float nws_div_markstein(float a, float d, float d_rcp) {
  const float q0 = NWS_MUL(a, d_rcp);
  const float rem = NWS_FMA(-q0, d, a);
  return NWS_FMA(rem, d_rcp, q0);
}

// The un-floored index of nws_lut_index for a POWER-OF-TWO table size without the multiply by the size:
// with s = x - min, idx = RN(size*s / span).  Scaling by 2^k commutes with every rounding of the Markstein
// sequence (no overflow / underflow in range), so dividing s by span/size with the reciprocal size/span * ... gives
// the bit-identical result: q0 = RN(s * (R*size)) = size*RN(s*R)... (tests/test_math_cpu.py checks every float).
// `span_over_size` = span / size and `rcp_times_size` = RN(1/span) * size, both exact scalings.
NWS_HD float nws_lut_idx_pow2(float x, float table_min, float span_over_size, float rcp_times_size) {
  return nws_div_markstein(NWS_ADD(x, -table_min), span_over_size, rcp_times_size);
}

NWS_HD NwsLutIdx nws_lut_index(float x, int table_size, float table_min, float span, float span_rcp) {
  const float a = NWS_MUL((float)table_size, NWS_ADD(x, -table_min));
  const float idx = nws_div_markstein(a, span, span_rcp);
  float fl = floorf(idx);
  const float hi = (float)(table_size - 1);
  fl = fl < 0.0f ? 0.0f : (fl > hi ? hi : fl);   // NaN propagates to lower = 0 via the int cast below
  NwsLutIdx r;
  r.lower = (int)fl;
  r.upper = r.lower + 1 < table_size ? r.lower + 1 : table_size - 1;
  r.fract = NWS_ADD(idx, -(float)r.lower);
  return r;
}

// linspace(start, end, steps)[i] in the symmetric two-sided form ATen uses (start + i*step for the
// first half, end - (steps-1-i)*step for the second).  torch's vectorised CPU kernel may differ from
// this by 1 ulp at some points, so callers that need the reference's exact grid pass it in
// (nws_build_lut's `sample_points`).
NWS_HD float nws_linspace_value(int i, int steps, float start, float end) {
  const float step = (end - start) / (float)(steps - 1);
  return i < steps / 2 ? NWS_ADD(start, NWS_MUL(step, (float)i)) : NWS_ADD(end, -NWS_MUL(step, (float)(steps - i - 1)));
}

// out = (U - L) * fract + L   (shaping.py:150): sub, mul, add — three roundings.
NWS_HD float nws_lut_lerp(float lo, float up, float fract) { return NWS_ADD(NWS_MUL(NWS_ADD(up, -lo), fract), lo); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (Salmon et al. 2011) for the forward's two uniform draws when the
// caller does not inject them (generators.py:30, :55).  24-bit mantissa uniform in [0,1), the
// resolution torch's CPU `rand` has for float32.
struct NwsPhilox4 {
  uint32_t v[4];
};

NWS_HD uint32_t nws_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

NWS_HD NwsPhilox4 nws_philox4x32_10(uint64_t counter, uint64_t stream, uint64_t seed) {
  uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
  uint32_t c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = nws_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = nws_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  NwsPhilox4 r;
  r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
  return r;
}

NWS_HD float nws_u32_to_unit(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }  // * 2^-24
