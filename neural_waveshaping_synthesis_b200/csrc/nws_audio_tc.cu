// Fused audio-rate kernel, tensor-core version: the harmonic mixer (Conv1d 101->64 at audio rate,
// neural_waveshaping.py:54,66 — 6,464 of the ~8,000 FMA per sample of the FastNEWT path) runs on the
// 5th-gen tensor cores as a 3xTF32 GEMM with the accumulator in TMEM; everything else is as in
// nws_audio.cu (same scalar recipes from nws_math.h).
//
// CTA = 4 compute warpgroups (512 threads) + 4 MMA warps, persistent, one CTA per SM.  A warpgroup owns one hop
// tile at a time: thread = sample = TMEM lane.
//   1. threads generate the oscillator bank sin(k*phase + shift_k)*mask_k for KS harmonics at a time, split each
//      value into tf32 hi/lo parts and write them straight to their own tensor-memory lane (tcgen05.st) as the A
//      operand — double-buffered stages of KS columns; nothing audio-rate touches shared memory;
//   2. the warpgroup's MMA warp (woken by the stage's named barrier) issues tcgen05.mma kind::tf32 (M=128 samples,
//      N=64 channels, K=8, A in TMEM) three times per k-step (A_hi B_hi + A_lo B_hi + A_hi B_lo) against the mixer
//      weights resident in shared memory, and commits to the stage's mbarrier — the tensor pipe works on stage s
//      while the threads already compute the sines of stage s+1, and other warpgroups fill the gaps;
//   3. when the last commit lands, each thread reads its own TMEM lane (its sample's 64 exciter channels) a few
//      columns at a time and runs FiLM -> shaper (LUT or sine MLP) -> FiLM -> mixdown.
// Nothing audio-rate is written to HBM except the final sample.
#include <stdlib.h>

#include "nws_audio_common.cuh"
#include "nws_fft.cuh"
#include "nws_tc.cuh"

// Oscillator sine: NWS_OSC_FAST=1 selects the SFU version (3.6e-7 max abs error on B200 instead of
// 1.2e-7, ~10 fewer instructions per harmonic); the choice is made by measured end-to-end parity
// (scripts/parity_report.py), see DESIGN.md.
#ifndef NWS_OSC_FAST
#define NWS_OSC_FAST 1
#endif
#if NWS_OSC_FAST
#define NWS_OSC_SIN(x) nws_sinf_turn(x)
#else
#define NWS_OSC_SIN(x) nws_sinf(x)
#endif

namespace {

constexpr int kWgs = 4;                  // compute warpgroups per CTA
constexpr int kTileChunk = 8;            // consecutive hops per scheduler claim (measured: 1 -> 8 = -1 % on random controls, -4 % on a real signal)
constexpr int kTcThreads = kWgs * 128 + kWgs * 32;   // + one MMA-issuing warp per compute warpgroup
constexpr int kWBytes = kHarmPad * kShapers * 4;       // one tf32 part of the B operand (26,624 B)
constexpr uint32_t kLboB = 8 * 128, kSbo = 128;

// Shaper weights as the tensor-core shaper path (SHP = 1) keeps them in shared memory, per channel:
//   [2 layers][32 lanes] float4: the lane's B fragment of mma.m16n8k8 for the 8x8 layers W2 / W3 (net.2 / net.4 of
//     TrainableNonlinearity, shaping.py:25-34): (hi(W[g][2t]), hi(W[g][2t+1]), lo(W[g][2t]), lo(W[g][2t+1])) with
//     g = lane / 4, t = lane % 4 and hi / lo the 3xTF32 split                                      (256 floats)
//   [4 t][24]: w1[2t], w1[2t+1], b1[2t], b1[2t+1] | (b2[2t], b2[2t+1]) x 2, twice | (b3[2t], b3[2t+1]) x 2, twice |
//              w4[2t], w4[2t+1], b4, 0                                                             (96 floats)
//     — the bias pairs are stored as ready-made accumulator quads (c0..c3 of the D fragment), one copy per 16-row
//     tile, so an accumulator is initialised by one LDS.128 instead of register moves
constexpr int kShpTcStride = 352;
constexpr int kShpTcSmall = 256;
constexpr int kShpTcSlot = 24;

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col): the warp-level tensor-core instruction.  The per-channel
// 8x8 layers of the shaper MLP are far below tcgen05's minimum tile (M = 128 needs N >= 16 and a trip through the MMA
// warp and an mbarrier per channel-layer); mma.sync keeps operands and results in the warp's own registers.
__device__ __forceinline__ void nws_mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One 8x8 layer for the warp's 32 samples as two m16n8k8 tiles, 3xTF32 (small terms first).  `h` holds this thread's
// part of the layer input in A-fragment order — h[r][u]: sample g + 8r, hidden unit 2t + u — which is also the order
// the accumulator comes back in (D fragment: row g / g + 8, columns 2t, 2t + 1), so layers chain without any data
// movement: the A operand's column k stands for unit 2k (k < 4) or 2(k - 4) + 1, and the B fragments are stored with
// the same permutation.  Returns the pre-activations (bias included) in place.
__device__ __forceinline__ void nws_shaper_layer_mma(float (&h)[4][2], const float4 bf, const float4* bias_quads) {
  const uint32_t bh0 = __float_as_uint(bf.x), bh1 = __float_as_uint(bf.y), bl0 = __float_as_uint(bf.z), bl1 = __float_as_uint(bf.w);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    uint32_t ah[4], al[4];
    const float v[4] = {h[2 * mt][0], h[2 * mt + 1][0], h[2 * mt][1], h[2 * mt + 1][1]};   // a0..a3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float hi = nws_tf32_hi(v[i]);
      ah[i] = __float_as_uint(hi);
      al[i] = __float_as_uint(v[i] - hi) & 0xffffe000u;
    }
    const float4 bq = bias_quads[mt];
    float d[4] = {bq.x, bq.y, bq.z, bq.w};
    nws_mma_tf32(d, al, bh0, bh1);
    nws_mma_tf32(d, ah, bl0, bl1);
    nws_mma_tf32(d, ah, bh0, bh1);
    h[2 * mt][0] = d[0]; h[2 * mt][1] = d[1]; h[2 * mt + 1][0] = d[2]; h[2 * mt + 1][1] = d[3];
  }
}

template <bool USE_LUT>
struct TcCfg {
  static constexpr int KS = USE_LUT ? 16 : 8;                 // harmonics per A stage
  static constexpr int NST = (kHarmPad + KS - 1) / KS;        // stages per tile (7 or 13)
  // tensor-memory columns of one warpgroup: accumulator [0,64) | A stage buffer 0: hi KS, lo KS | buffer 1: hi KS, lo KS
  static constexpr uint32_t kColA = kShapers;
  static constexpr uint32_t kTmemColsWg = 128;
  static constexpr int kChPerLd = USE_LUT ? 8 : 2;            // exciter channels per TMEM load
  // dynamic shared memory layout (bytes)
  static constexpr int oW = 0;                                // W_hi | W_lo
  static constexpr int oFilm = oW + 2 * kWBytes;              // [wg][3][256] floats
  static constexpr int oCoef = oFilm + kWgs * 3 * kFilm * 4;     // [wg][half][64][8] floats: FiLM lerp coefficients
  static constexpr int oSmall = oCoef + kWgs * 2 * kShapers * 8 * 4;  // (hmix_b, mix_w)[64] | shift[104] | input_scale[64] floats
  static constexpr int oShaper = oSmall + (kShapers + kHarmPad + kShapers + kShapers) * 4;
  static constexpr int oNoise = oShaper + (USE_LUT ? 0 : kShapers * kShpTcStride * 4);   // (kShpTcStride >= kShaperStride)
  // noise branch: twiddles [128] float2 | per warpgroup: the frame pair's spectrum / transform [256] float2 (in place) |
  // finished noise samples of this tile and the next [2][128]
  static constexpr int kNoiseWg = 256 * 8 + 2 * 128 * 4;
  static constexpr int kBytes = oNoise + 128 * 8 + kWgs * kNoiseWg;
};

// ---- Filtered-noise branch inside the fused kernel (FIRNoiseSynth.forward, generators.py:21-35; the identities of
// nws_noise.cu): the frame's response is real — (-1)^k (0.5 H[k] + 0.25 (H[k-1] + H[k+1])) — so filtering is a bin-wise
// product with the shared noise spectrum, and two real frames (fa in the real part, fa + 1 in the imaginary part) come
// back through ONE 256-point complex inverse FFT.  The work is done by the warpgroup's MMA-issuing warp while the
// compute warps are in the tile's epilogue (half of a tile's time, during which that warp has nothing to issue): it
// filters the NEXT tile's hop — a single warp, so the transform needs no CTA barrier (in-place radix-4 decimation in
// frequency, __syncwarp between passes, digit-reversed read-out) — and the compute warps never execute an instruction
// of the noise branch.  (Slicing the transform between the stages' MMA issues instead delayed those issues — a lone
// warp runs ~150 dependent instructions per slice at its own latency — and cost the kernel 14 %.)
struct NwsNoiseRegs {
  float4 ha, hb;          // band gains 4 lane .. 4 lane + 3 of frames fa, fa + 1
  float ha128, hb128;     // ... and the Nyquist band
  float4 xa[2], xb[2];    // noise spectrum bins 4 lane .. 4 lane + 3 (re, im interleaved)
  float2 xa128, xb128;
};

__device__ __forceinline__ void nws_noise_fetch(NwsNoiseRegs& r, const float* __restrict__ bands_b, const float2* __restrict__ xspec,
                                                int fa, bool va, bool vb, int lane) {
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  r.ha = r.hb = z4; r.xa[0] = r.xa[1] = r.xb[0] = r.xb[1] = z4;
  r.ha128 = r.hb128 = 0.f; r.xa128 = r.xb128 = make_float2(0.f, 0.f);
  if (va) {
    const float* h = bands_b + (size_t)fa * kBandsPad;
    const float2* x = xspec + (size_t)fa * kBandsPad;
    r.ha = *reinterpret_cast<const float4*>(h + 4 * lane); r.ha128 = h[128];
    r.xa[0] = *reinterpret_cast<const float4*>(x + 4 * lane); r.xa[1] = *reinterpret_cast<const float4*>(x + 4 * lane + 2);
    r.xa128 = x[128];
  }
  if (vb) {
    const float* h = bands_b + (size_t)(fa + 1) * kBandsPad;
    const float2* x = xspec + (size_t)(fa + 1) * kBandsPad;
    r.hb = *reinterpret_cast<const float4*>(h + 4 * lane); r.hb128 = h[128];
    r.xb[0] = *reinterpret_cast<const float4*>(x + 4 * lane); r.xb[1] = *reinterpret_cast<const float4*>(x + 4 * lane + 2);
    r.xb128 = x[128];
  }
}

// slice 0: Z[k] = Ya[k] + i Yb[k], Y = X * Hw, Hermitian-extended to 256 bins, natural order
__device__ __forceinline__ void nws_noise_spectrum_slice(float2* zb, const NwsNoiseRegs& r, int lane) {
  auto put = [&](int kk, float hwa, float hwb, float2 xa, float2 xb) {
    float2 ya = make_float2(xa.x * hwa, xa.y * hwa), yb = make_float2(xb.x * hwb, xb.y * hwb);
    if (kk == 0 || kk == 128) { ya.y = 0.f; yb.y = 0.f; }   // irfft ignores the imaginary part of DC / Nyquist
    zb[kk] = make_float2(ya.x - yb.y, ya.y + yb.x);
    if (kk > 0 && kk < 128) zb[256 - kk] = make_float2(ya.x + yb.y, yb.x - ya.y);   // conj(Ya) + i conj(Yb)
  };
  // neighbours across lanes: band 4 lane - 1 (lane 0: band 1 mirrors), band 4 lane + 4 (lane 31: band 128)
  float la = __shfl_up_sync(0xffffffffu, r.ha.w, 1), lb = __shfl_up_sync(0xffffffffu, r.hb.w, 1);
  float ra = __shfl_down_sync(0xffffffffu, r.ha.x, 1), rb = __shfl_down_sync(0xffffffffu, r.hb.x, 1);
  const float a127 = __shfl_sync(0xffffffffu, r.ha.w, 31), b127 = __shfl_sync(0xffffffffu, r.hb.w, 31);
  if (lane == 0) { la = r.ha.y; lb = r.hb.y; }
  if (lane == 31) { ra = r.ha128; rb = r.hb128; }
  const float ha[6] = {la, r.ha.x, r.ha.y, r.ha.z, r.ha.w, ra}, hb[6] = {lb, r.hb.x, r.hb.y, r.hb.z, r.hb.w, rb};
  const float2 xa[4] = {make_float2(r.xa[0].x, r.xa[0].y), make_float2(r.xa[0].z, r.xa[0].w), make_float2(r.xa[1].x, r.xa[1].y), make_float2(r.xa[1].z, r.xa[1].w)};
  const float2 xb[4] = {make_float2(r.xb[0].x, r.xb[0].y), make_float2(r.xb[0].z, r.xb[0].w), make_float2(r.xb[1].x, r.xb[1].y), make_float2(r.xb[1].z, r.xb[1].w)};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float sgn = (j & 1) ? -1.f : 1.f;   // (-1)^kk, kk = 4 lane + j
    put(4 * lane + j, sgn * fmaf(0.25f, ha[j] + ha[j + 2], 0.5f * ha[j + 1]), sgn * fmaf(0.25f, hb[j] + hb[j + 2], 0.5f * hb[j + 1]), xa[j], xb[j]);
  }
  if (lane == 0)   // Nyquist: both neighbours are band 127
    put(128, fmaf(0.25f, a127 + a127, 0.5f * r.ha128), fmaf(0.25f, b127 + b127, 0.5f * r.hb128), r.xa128, r.xb128);
}

// slices 1..4: pass `s` (0..3) of the in-place radix-4 decimation-in-frequency inverse transform; block length
// L = 256 >> 2s, two butterflies per lane.  X[4k' + m] of a block = (sum_r x[n' + r L/4] i^(r m)) W_L^(n' m), stored at
// block position m L/4 + n'; the final element order is the base-4 digit reversal.
__device__ __forceinline__ void nws_noise_fft_pass(float2* zb, const float2* tw_s, int s, int lane) {
  const int log_q = 6 - 2 * s, q = 1 << log_q;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = lane + 32 * u, pos = j & (q - 1), base = ((j >> log_q) << (log_q + 2)) + pos;
    float2 v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = zb[base + r * q];
    const float2 a0 = nws_cadd(v[0], v[2]), a1 = nws_csub(v[0], v[2]), a2 = nws_cadd(v[1], v[3]), d = nws_csub(v[1], v[3]);
    const float2 a3 = make_float2(-d.y, d.x);   // * (+i)
    float2 o[4] = {nws_cadd(a0, a2), nws_cadd(a1, a3), nws_csub(a0, a2), nws_csub(a1, a3)};
    if (s < 3 && pos) {
#pragma unroll
      for (int m = 1; m < 4; ++m) o[m] = nws_cmul(o[m], nws_twiddle<true>(tw_s, 1, (pos * m) << (2 * s), 128));   // W_L^(pos m) = W_256^(pos m 256/L)
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) zb[base + m * q] = o[m];
  }
}

// slice 5: read-out (digit reversal), scale, overlap-add.  Lane holds samples lane + 32 i (i < 4) of a hop; `keep` = second
// half of the newest frame (registers).  mode 0: frames (t-1, t) -> this hop (istft's envelope is 1 in the first hop, 2
// elsewhere); mode 1: frames (t, t+1) -> this hop from the kept tail and the next hop, finished ahead; mode 2: only the
// tail of the pair's second frame is wanted.
__device__ __forceinline__ void nws_noise_output_slice(const float2* zb, float* out_now, float (&out_next)[4], float (&keep)[4],
                                                       int mode, bool first_hop, int lane) {
  auto rev = [](int n) { return ((n & 3) << 6) | ((n & 12) << 2) | ((n & 48) >> 2) | ((n & 192) >> 6); };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = lane + 32 * i;
    const float2 lo = zb[rev(n)], hi = zb[rev(128 + n)];
    const float a_first = lo.x * (1.0f / 256.0f), a_second = hi.x * (1.0f / 256.0f);
    const float b_first = lo.y * (1.0f / 256.0f), b_second = hi.y * (1.0f / 256.0f);
    if (mode == 1) {
      out_now[n] = 0.5f * (keep[i] + a_first);
      out_next[i] = 0.5f * (a_second + b_first);
    } else if (mode == 0) {
      out_now[n] = first_hop ? b_first : 0.5f * (a_second + b_first);
    }
    keep[i] = b_second;
  }
}

// Noise state of one MMA warp.  keep[] = second half of frame t_kept of utterance u (this lane's four samples per hop);
// hold[] = the finished samples of hop `next` (or next = -1).  Frames are always paired as (2m - 1, 2m), whoever
// transforms them, so the samples do not depend on how tiles were handed out (repeat runs are bit-identical): an even
// hop t needs the pair (t-1, t); an odd hop needs the tail of (t-2, t-1) — kept from the hop before when the warpgroup
// rendered it, else transformed for its tail alone — and the pair (t, t+1), which also finishes hop t+1.  A chunk of
// consecutive hops costs one transform per two hops plus one or two at its start.
struct NwsNoiseWarp {
  const float* bands;      // [B*T][kBandsPad]
  const float2* xspec;     // [T][kBandsPad]
  const float2* tw;        // shared memory, [128]
  float2* z;               // shared memory, [256]
  int T, flim;             // frames from flim on are not needed (and may not be encoded yet)
  int u, t_kept, next;
  float keep[4], hold[4];
};

// The filtered-noise samples of hop t of utterance b -> out[128] (shared memory), then one arrival on `bar`.
__device__ __noinline__ void nws_noise_tile(NwsNoiseWarp& c, int b, int t, float* out, uint64_t* bar) {
  const int lane = threadIdx.x & 31;
  if (c.u == b && c.next == t) {   // finished by the previous transform
#pragma unroll
    for (int i = 0; i < 4; ++i) out[lane + 32 * i] = c.hold[i];
    c.next = -1;
  } else {
    const bool odd = t & 1, kept = c.u == b && c.t_kept == t - 1;
    const float* bands_b = c.bands + (size_t)b * c.T * kBandsPad;
    const int n_tr = odd && !kept ? 2 : 1;
    for (int k = 0; k < n_tr; ++k) {
      const bool last = k == n_tr - 1;
      const int fa = !last ? t - 2 : (odd ? t : t - 1);
      NwsNoiseRegs r;
      nws_noise_fetch(r, bands_b, c.xspec, fa, fa >= 0 && fa < c.flim, fa + 1 < c.flim, lane);
      nws_noise_spectrum_slice(c.z, r, lane);
      __syncwarp();
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        nws_noise_fft_pass(c.z, c.tw, s, lane);
        __syncwarp();
      }
      nws_noise_output_slice(c.z, out, c.hold, c.keep, !last ? 2 : (odd ? 1 : 0), t == 0, lane);
      __syncwarp();
    }
    c.u = b;
    c.t_kept = odd ? t + 1 : t;
    c.next = odd ? t + 1 : -1;
  }
  __syncwarp();
  if (lane == 0) nws_mbar_arrive(bar);   // (release: the warp's stores above are ordered before it)
}

__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
// Stage-operand-written barrier of (warpgroup, stage buffer): named barriers 5..12, 128 producer threads arrive
// (non-blocking; their shared-memory stores are ordered before the consumer's wake-up), the MMA warp syncs.
__device__ __forceinline__ void fill_arrive(int wg, int buf) { asm volatile("bar.arrive %0, 160;" ::"r"(5 + 2 * wg + buf) : "memory"); }
__device__ __forceinline__ void fill_wait(int wg, int buf) { asm volatile("bar.sync %0, 160;" ::"r"(5 + 2 * wg + buf) : "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T, one thread issues
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// this thread's lane, 8 consecutive columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void nws_cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(nws_smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void nws_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void nws_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// max(min(a, b), 0) in one instruction
__device__ __forceinline__ int nws_min_relu(int a, int b) {
  int d;
  asm("min.s32.relu %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld<2>(uint32_t taddr, float* v) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1);
}
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}

template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

template <bool USE_LUT, bool TAP, int MODE, int SHP>
__global__ void __launch_bounds__(kTcThreads, 1) nws_audio_tc_kernel(const NwsAudioParams p, const float* __restrict__ w_umma,
                                                                     int* __restrict__ fault) {
  using C = TcCfg<USE_LUT>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t free_bar[kWgs][2];   // the MMAs that read the stage have completed (tcgen05.commit)
  __shared__ uint64_t nz_bar[kWgs][2];     // the filtered-noise samples of a tile are in its slot of shared memory (slots and barriers alternate per tile)
  __shared__ double warp_tot[kWgs + 1][4];
  __shared__ float2 film_k[kWgs][4];       // per (warpgroup, warp): partial sums of w_c*(Ab_n, Db_n) over the warp's 32 channels
  __shared__ uint32_t tmem_base_s;
  __shared__ int tile_s[kWgs][2];           // next tile of each warpgroup (dynamic scheduler), two slots used alternately
  __shared__ volatile int done_s[kWgs];     // warpgroup has run out of tiles (tells its MMA warp to stop)

  const int tid = threadIdx.x, wg = tid >> 7, wt = tid & 127, lane = tid & 31, wwarp = wt >> 5;
  const int T = p.T, N = T * kHop;
  float* sm_small = reinterpret_cast<float*>(smem + C::oSmall);
  float2* sm_bw = reinterpret_cast<float2*>(sm_small);   // [64] (harmonic_mixer bias, mixdown weight)
  float* sm_shift = sm_small + 2 * kShapers;
  float* sm_film = reinterpret_cast<float*>(smem + C::oFilm) + (wg & 3) * 3 * kFilm;
  float* sm_coef = reinterpret_cast<float*>(smem + C::oCoef) + (wg & 3) * 2 * kShapers * 8;
  float* sm_shaper = reinterpret_cast<float*>(smem + C::oShaper);

  // ---- CTA-lifetime staging
  for (int i = tid; i < 2 * kWBytes / 16; i += kTcThreads)
    reinterpret_cast<float4*>(smem + C::oW)[i] = reinterpret_cast<const float4*>(w_umma)[i];
  if (tid < kShapers) sm_bw[tid] = make_float2(p.hmix_b[tid], p.mix_w[tid]);
  if (tid < kHarmPad) sm_shift[tid] = tid < kHarm ? nws_phase_shift(p.u_phase[tid], p.rand_phase[tid]) : 0.f;
  float* sm_scale = sm_shift + kHarmPad;                 // [64] TrainableNonlinearity.input_scale
  if (!USE_LUT && SHP == 0)
    for (int i = tid; i < kShapers * kShaperStride / 4; i += kTcThreads)
      reinterpret_cast<float4*>(sm_shaper)[i] = reinterpret_cast<const float4*>(p.shaper)[i];
  if (!USE_LUT && SHP == 1) {
    if (tid < kShapers) sm_scale[tid] = p.shaper[tid * kShaperStride + kShpScale];
    for (int i = tid; i < kShapers * 2 * 32; i += kTcThreads) {   // B fragments of the two 8x8 layers
      const int c = i >> 6, l = (i >> 5) & 1, ln = i & 31, g = ln >> 2, t = ln & 3;
      const float* W = p.shaper + c * kShaperStride + (l ? kShpW3 : kShpW2) + g * 8 + 2 * t;   // W[out g][in 2t, 2t+1]
      const float w0 = W[0], w1 = W[1], h0 = nws_tf32_hi(w0), h1 = nws_tf32_hi(w1);
      reinterpret_cast<float4*>(sm_shaper + c * kShpTcStride)[l * 32 + ln] =
          make_float4(h0, h1, nws_tf32_lo(w0, h0), nws_tf32_lo(w1, h1));
    }
    for (int i = tid; i < kShapers * 4; i += kTcThreads) {
      const int c = i >> 2, t = i & 3;
      const float* r = p.shaper + c * kShaperStride;
      float4* dst = reinterpret_cast<float4*>(sm_shaper + c * kShpTcStride + kShpTcSmall + t * kShpTcSlot);
      dst[0] = make_float4(r[kShpW1 + 2 * t], r[kShpW1 + 2 * t + 1], r[kShpB1 + 2 * t], r[kShpB1 + 2 * t + 1]);
      dst[1] = dst[2] = make_float4(r[kShpB2 + 2 * t], r[kShpB2 + 2 * t + 1], r[kShpB2 + 2 * t], r[kShpB2 + 2 * t + 1]);
      dst[3] = dst[4] = make_float4(r[kShpB3 + 2 * t], r[kShpB3 + 2 * t + 1], r[kShpB3 + 2 * t], r[kShpB3 + 2 * t + 1]);
      dst[5] = make_float4(r[kShpW4 + 2 * t], r[kShpW4 + 2 * t + 1], r[kShpB4], 0.f);
    }
  }
  // noise branch (p.bands): twiddles of the 256-point transform; per warpgroup the transform buffer and the finished
  // samples of this tile / the next one (written by the MMA warp, read by the compute threads at the end of the tile)
  float2* nz_tw = reinterpret_cast<float2*>(smem + C::oNoise);
  unsigned char* nz_wg = smem + C::oNoise + 128 * 8;   // (per-warpgroup areas, kNoiseWg bytes each)
  const bool nz_on = p.bands != nullptr;
  if (nz_on && tid < 128) nz_tw[tid] = p.tw_master[tid * (kTwMaster / 256)];
  if (tid < kWgs) done_s[tid] = 0;
  if (tid < 32) nws_tmem_alloc(&tmem_base_s, C::kTmemColsWg * kWgs);
  if (tid == 0) {
    for (int i = 0; i < kWgs * 2; ++i) {
      nws_mbar_init(&free_bar[0][0] + i, 1);
    }
    for (int i = 0; i < kWgs * 2; ++i) nws_mbar_init(&nz_bar[0][0] + i, 1);
    nws_fence_mbar_init();
  }
  nws_fence_proxy_async();   // the weight tiles were written through the generic proxy
  nws_tc_fence_before();
  __syncthreads();
  nws_tc_fence_after();
  // Short-buffer path (programmatic dependent launch): everything above overlapped the MLP chain's launch; its FiLM rows
  // and filtered noise are read from here on.  No-ops for ordinary launches.
  nws_pdl_wait();
  nws_pdl_launch();
  const uint32_t tmem_acc = tmem_base_s + (wg & 3) * C::kTmemColsWg;     // this warpgroup's columns
  const uint32_t tmem_lane = tmem_acc + ((uint32_t)(wwarp * 32) << 16);   // this warp's lane quarter
  const uint32_t idesc = nws_umma_idesc_tf32(128, 64);
  const uint32_t w_hi_addr = nws_smem_u32(smem + C::oW), w_lo_addr = w_hi_addr + kWBytes;
  const float mix_b = p.mix_b[0];
  const float inv_hop = (float)T / (float)N;
  const float lut_scale = USE_LUT ? (float)p.lut_size / p.lut_span : 1.0f;   // table positions per unit of shaper input
  const int hops = p.t_end - p.t_begin;
  const int n_tiles = p.B * hops;   // < 2^31 (checked by the launcher)
  const uint32_t hops_magic = p.hops_magic;
  uint32_t uses0 = 0, uses1 = 0;   // fills issued per stage buffer (same in every thread of the warpgroup)
  bool ok = true;

  if (wg == kWgs) {
    // ================= MMA warps: warp 16+w serves compute warpgroup w.  It sleeps on the stage's fill barrier —
    // a named hardware barrier (bar.sync; the compute warps bar.arrive), so a waiting MMA warp issues nothing:
    // polling an mbarrier here cost 15 % of the SM's issue slots — issues the 3xTF32 MMAs of the stage and
    // commits to the stage's free barrier; the last stage's commit also tells the warpgroup its accumulator is
    // complete.
    // Everything the MMAs take (tensor-memory addresses, descriptors) is made provably warp-uniform (a lane-0
    // broadcast, as CUTLASS's canonical_warp_idx_sync does) and the stage loop is fully unrolled, so the operands
    // live in uniform registers and the descriptor of every k-step is base + immediate: without this the compiler
    // wrapped each of the 39 MMAs of a tile in a 19-instruction elect / R2UR / vote loop (8 % of the SM's issue
    // slots went to the four MMA warps).
    const int w = __shfl_sync(0xffffffffu, wwarp, 0);
    const uint32_t acc = __shfl_sync(0xffffffffu, tmem_base_s, 0) + w * C::kTmemColsWg;
    const uint64_t dbh0 = nws_umma_smem_desc(w_hi_addr, kLboB, kSbo), dbl0 = nws_umma_smem_desc(w_lo_addr, kLboB, kSbo);
    const bool issuer = nws_elect_one();
    NwsNoiseWarp nzc;
    nzc.bands = p.bands; nzc.xspec = p.xspec; nzc.tw = nz_tw; nzc.z = reinterpret_cast<float2*>(nz_wg + w * C::kNoiseWg);
    nzc.T = T; nzc.flim = p.t_end < T ? p.t_end : T; nzc.u = -1; nzc.t_kept = -2; nzc.next = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) { nzc.keep[i] = 0.f; nzc.hold[i] = 0.f; }
    float* nz_out_w = reinterpret_cast<float*>(nzc.z + 256);
    int mpar = 0;
    bool nz_first = nz_on;
    auto nz_tile = [&](int slot) {   // the noise of the tile published in tile_s[w][slot], if there is one
      const int tl = tile_s[w][slot];
      if (tl >= n_tiles) return;
      int q = (int)__umulhi((uint32_t)tl, hops_magic), r = tl - q * hops;
      if (r >= hops) { ++q; r -= hops; }
      nws_noise_tile(nzc, q, p.t_begin + r, nz_out_w + slot * 128, &nz_bar[w][slot]);
    };
    for (;;) {   // one iteration per tile of warpgroup w; tiles are handed out dynamically
      fill_wait(w, 0);
      if (done_s[w]) break;   // woken by the warpgroup running out of tiles
#pragma unroll
      for (int st = 0; st < C::NST; ++st) {
        const int buf = st & 1;
        if (st > 0) fill_wait(w, buf);
        nws_tc_fence_after();
        if (issuer) {
          const int ks_here = (kHarmPad - st * C::KS) < C::KS ? (kHarmPad - st * C::KS) : C::KS;
          const uint32_t a_hi = acc + C::kColA + buf * 2 * C::KS, a_lo = a_hi + C::KS;   // A operand: tensor memory
#pragma unroll
          for (int j = 0; j < ks_here / 8; ++j) {
            // one k-step (8 harmonics) further down the weight tile = 2 * kLboB bytes = +128 in the descriptor's
            // start-address field (16-byte units; the field cannot carry: shared memory is < 2^18 bytes)
            const uint64_t adv = (uint64_t)((st * C::KS / 8 + j) * (2 * kLboB / 16));
            umma_tf32_ts(acc, a_hi + j * 8, dbh0 + adv, idesc, (st | j) ? 1u : 0u);
            umma_tf32_ts(acc, a_lo + j * 8, dbh0 + adv, idesc, 1u);
            umma_tf32_ts(acc, a_hi + j * 8, dbl0 + adv, idesc, 1u);
          }
          nws_umma_commit(&free_bar[w][buf]);
        }
        __syncwarp();
        // the warpgroup's very first tile: nobody filtered its noise ahead (the one delay of a stage issue per CTA)
        if (st == 0 && nz_first) { nz_tile(mpar); nz_first = false; }
      }
      // Every stage of this tile is issued; the compute warps now wait for the accumulator and run the epilogue: the
      // time to filter the noise of the NEXT tile (its index was published when this one started; its slot and barrier
      // are the other pair, last read at the end of the tile before this one).
      if (nz_on) nz_tile(mpar ^ 1);
      mpar ^= 1;
    }
  } else
  {
  // dynamic tile scheduler: tiles (utterance b, hop t) are claimed from a global counter, so CTAs that start
  // late (SMs still busy with the encoder of a later time block) do not hold tiles hostage
  // The three FiLM frames a tile blends (t-1, t, t+1; 3 KB) are fetched by cp.async one tile ahead: the film
  // buffer is dead once the coefficient table is built, so the next tile's rows stream into it while this
  // tile's oscillator bank and shaper loop run.
  // tile -> (utterance, hop): floor(tile / hops) by a multiply-high with floor(2^32 / hops) and one correction step
  auto split_tile = [&](int tl, int& ub, int& ut) {
    int q = (int)__umulhi((uint32_t)tl, hops_magic), r = tl - q * hops;
    if (r >= hops) { ++q; r -= hops; }
    ub = q;
    ut = p.t_begin + r;
  };
  auto film_prefetch = [&](int tl) {
    int pb, pt;
    split_tile(tl, pb, pt);
    for (int i = wt; i < 3 * kFilm / 4; i += 128) {
      const int slot = i / (kFilm / 4), fr = pt - 1 + slot;
      if (fr >= 0 && fr < T)
        nws_cp_async16(reinterpret_cast<float4*>(sm_film) + i,
                       reinterpret_cast<const float4*>(p.film + ((size_t)pb * T + fr) * kFilm) + (i - slot * (kFilm / 4)));
    }
    nws_cp_async_commit();
  };
  // Thread 0 of the warpgroup claims tiles two ahead, so the atomic's round trip is never waited for.  Tiles are
  // claimed in chunks of consecutive hops of one utterance (p.tile_chunk; the table lines a hop gathers are mostly
  // the ones the next hop needs), single tiles towards the end so the launch still drains evenly.
  int c_next = 0, c_end = 0, c_base = 0;
  const int chunk_until = n_tiles - (int)gridDim.x * kWgs * p.tile_chunk * 2;
  auto claim = [&]() -> int {
    if (c_next == c_end) {
      const int want = c_base < chunk_until ? p.tile_chunk : 1;
      c_next = c_base = atomicAdd(p.tile_counter, want);
      c_end = c_next + want;
    }
    return c_next++;
  };
  int claimed = 0;
  if (wt == 0) {
    tile_s[wg][0] = claim();
    claimed = claim();
  }
  wg_barrier(wg);
  int tile = tile_s[wg][0];
  if (tile < n_tiles) film_prefetch(tile);
  int par = 0;
  const float* nz_out = reinterpret_cast<const float*>(nz_wg + wg * C::kNoiseWg + 256 * 8);
  uint32_t nz_phase = 0;
  for (;;) {
    if (tile >= n_tiles) break;
    if (wt == 0) {
      tile_s[wg][par ^ 1] = claimed;                  // the next tile: published by the scan barrier below
      claimed = claim();                              // the one after
    }
    int b, t;
    split_tile(tile, b, t);
    // ---- f0 upsample and the cumsum of generators.py:59 (fp64 scan + per-hop carry)
    const int n = t * kHop + wt;
    const NwsLerp lc = nws_lerp_coords(n, T, inv_hop);
    const float* f0b = p.f0 + (size_t)b * T;
    // issued early so its latency hides behind the scan: the hop's fp64 phase carry
    const double carry_v = p.carry[(size_t)b * T + t];
    const float f0u = nws_lerp_apply(lc, f0b[lc.i0], f0b[lc.i1]);
    double v = (double)f0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[wg][wwarp] = v;
    nws_cp_async_wait_all();   // this thread's share of the tile's film rows has landed ...
    wg_barrier(wg);            // ... and everybody's is visible, with the warp totals and the next tile
    double pre = carry_v;
    for (int w = 0; w < wwarp; ++w) pre += warp_tot[wg][w];
    const float csum = (float)(pre + v);
    const float phase = nws_phase_from_cumsum(csum, (float)kSampleRate);
    {
      // FiLM upsample (shaping.py:69) as A + l1*(B - A): every sample of a half-hop blends the same two
      // frames, so per (half, channel) the tile tabulates everything that does not depend on the sample
      // (thread = (half, channel); the film frames were published by the barrier above).  With g = Ag + l1*Dg,
      // b = Ab + l1*Db and the exciter e = acc + bias (harmonic_mixer bias, neural_waveshaping.py:54):
      //   FiLM in   x = g_i*e + b_i      = fma(l1, fma(Dg, acc, Db'), fma(Ag, acc, Ab'))     Ab' = Ag*bias + Ab, Db' likewise
      //   FiLM out + mixdown (shaping.py:76-79)  sum_c w_c*(g_n*y_c + b_n)
      //                                  = sum_c Agw*y_c + l1 * sum_c Dgw*y_c + (KA + l1*KD)     Agw = w_c*Ag_n, KA = sum_c w_c*Ab_n
      // so the per-channel loop costs 3 + 2 FMAs instead of 4 lerps + 2 mul/add pairs + 1 FMA.
      const int h = wt >> 6, c = wt & 63;
      const int sa = h == 0 ? (t >= 1 ? 0 : 1) : 1;
      const int sb = h == 0 ? (t >= 1 ? 1 : 2) : (t + 1 < T ? 2 : 1);
      const float* fa = sm_film + sa * kFilm + c;
      const float* fb = sm_film + sb * kFilm + c;
      const float2 bw = sm_bw[c];
      float4 lo4, hi4;
      lo4.x = fa[0]; lo4.y = fb[0] - fa[0];
      lo4.z = fmaf(lo4.x, bw.x, fa[kShapers]); lo4.w = fmaf(lo4.y, bw.x, fb[kShapers] - fa[kShapers]);
      if (USE_LUT) {
        // FastNEWT: the table position idx = size * (x - min) / span (shaping.py:137) is affine in x, so it is folded
        // into the FiLM-in coefficients here, once per (half hop, channel): the per-sample loop gets idx straight out
        // of its three FMAs instead of x and a four-instruction bit-exact division after it.  idx then differs from
        // the reference's operation order by about an ulp — the size of what the rounding of x itself moves it by —
        // and where that flips floor(idx) the interpolant is continuous.  (The stage entry nws_stage_lut_lookup keeps
        // the reference's exact sequence: identical inputs -> identical indices and values.)
        lo4.x *= lut_scale; lo4.y *= lut_scale; lo4.z = (lo4.z - p.lut_min) * lut_scale; lo4.w *= lut_scale;
      }
      hi4.x = bw.y * fa[2 * kShapers]; hi4.y = bw.y * (fb[2 * kShapers] - fa[2 * kShapers]);
      hi4.z = (!USE_LUT && SHP == 1) ? sm_scale[c] : 0.f; hi4.w = 0.f;   // SHP 1: the shaper's input scale rides along
      float ka = bw.y * fa[3 * kShapers], kd = bw.y * (fb[3 * kShapers] - fa[3 * kShapers]);
      float4* dst = reinterpret_cast<float4*>(sm_coef + (h * kShapers + c) * 8);
      dst[0] = lo4;
      dst[1] = hi4;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ka += __shfl_xor_sync(0xffffffffu, ka, o);
        kd += __shfl_xor_sync(0xffffffffu, kd, o);
      }
      if (lane == 0) film_k[wg][wwarp] = make_float2(ka, kd);
    }
    wg_barrier(wg);   // coefficient table visible to the whole warpgroup before the shaper loop reads it; film rows dead
    const int tile_next = tile_s[wg][par ^ 1];
    if (tile_next < n_tiles) film_prefetch(tile_next);
    float mix_ka, mix_kd;   // this half-hop's constant part of the mixdown
    {
      const float2 ka0 = film_k[wg][(wt >> 6) * 2], ka1 = film_k[wg][(wt >> 6) * 2 + 1];
      mix_ka = ka0.x + ka1.x;
      mix_kd = ka0.y + ka1.y;
    }

    // ---- oscillator bank -> A operand stages -> tcgen05.mma
    // Fast path, decided per warp and tile: no harmonic of any lane is masked (f0 * 104 < 8000 Hz everywhere —
    // every control stream below 76.9 Hz, all of the timing scripts' inputs) — the stage loop is fully unrolled,
    // so harmonic numbers are immediates and there is no per-stage classification.  Otherwise the general loop.
    const bool tile_all_on =
        __all_sync(0xffffffffu, NWS_MUL(f0u, (float)kHarmPad) < 0.5f * kSampleRate && !(f0u != f0u));
    if (tile_all_on) {
#pragma unroll
      for (int st = 0; st < C::NST; ++st) {
        const int buf = st & 1, k0 = st * C::KS;
        const uint32_t prior = buf ? uses1 : uses0;
        if (prior > 0 && ok) ok = nws_mbar_wait(&free_bar[wg][buf], (prior - 1) & 1);
        const uint32_t col_hi = tmem_lane + C::kColA + buf * 2 * C::KS, col_lo = col_hi + C::KS;
#pragma unroll
        for (int kk = 0; kk < C::KS; kk += 8) {
          if (k0 + kk < kHarmPad) {
            float h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (k0 + kk + j < kHarm) {
                const float s = NWS_OSC_SIN(NWS_ADD(NWS_MUL((float)(k0 + kk + j + 1), phase), sm_shift[k0 + kk + j]));
                h[j] = nws_tf32_hi(s);
                l[j] = s - h[j];
              } else {   // harmonics 102..104 are padding (zero mixer weights): no sine
                h[j] = 0.f;
                l[j] = 0.f;
              }
            }
            tmem_st8(col_hi + kk, h);
            tmem_st8(col_lo + kk, l);
          }
        }
        tmem_wait_st();
        nws_tc_fence_before();
        fill_arrive(wg, buf);
        if (buf) ++uses1; else ++uses0;
      }
    } else {
#pragma unroll 1
    for (int st = 0; st < C::NST; ++st) {
      const int buf = st & 1, k0 = st * C::KS;
      const float k0f = (float)k0;
      const int ks_here = (kHarmPad - k0) < C::KS ? (kHarmPad - k0) : C::KS;   // last stage may be short
      const uint32_t prior = buf ? uses1 : uses0;
      if (prior > 0 && ok) ok = nws_mbar_wait(&free_bar[wg][buf], (prior - 1) & 1);   // MMAs that read this buffer are done
      const uint32_t col_hi = tmem_lane + C::kColA + buf * 2 * C::KS, col_lo = col_hi + C::KS;
      // Anti-alias mask (generators.py:50-52): fp32(f0*k) < 8000 is monotone in k for f0 >= 0 and always true
      // for f0 < 0, so one warp vote per stage classifies all 16 harmonics: every lane unmasked (no mask
      // arithmetic), every lane masked (the operand is zero: no sines at all), or mixed (general path).
      const float k_last = k0f + (float)ks_here, k_first = k0f + 1.0f;
      const bool all_on = __all_sync(0xffffffffu, NWS_MUL(f0u, k_last) < 0.5f * kSampleRate && !(f0u != f0u));
      const bool all_off = __all_sync(0xffffffffu, f0u >= 0.f && !(NWS_MUL(f0u, k_first) < 0.5f * kSampleRate));
      // eight harmonics (one MMA k-step) at a time: this thread's row of the A operand goes straight to its
      // tensor-memory lane (tcgen05.st), hi and lo tf32 parts
#pragma unroll
      for (int kk = 0; kk < C::KS; kk += 8) {
        if (kk < ks_here) {
          float h[8], l[8];
          if (all_off) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = 0.f; l[j] = 0.f; }
          } else if (all_on) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // harmonic number k = k0 + kk + j + 1 (k0f + const is exact: small integers).  Harmonics 102..104
              // are padding: their mixer weights are zero, so their (finite) sines are never seen.
              const float kf = k0f + (float)(kk + j + 1);
              const float s = NWS_OSC_SIN(NWS_ADD(NWS_MUL(kf, phase), sm_shift[k0 + kk + j]));   // generators.py:60-61
              h[j] = nws_tf32_hi(s);
              l[j] = s - h[j];   // exact; the tensor core ignores the 13 low bits (nws_selftest_umma checks)
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float kf = k0f + (float)(kk + j + 1);
              float s = NWS_OSC_SIN(NWS_ADD(NWS_MUL(kf, phase), sm_shift[k0 + kk + j]));
              s = NWS_MUL(f0u, kf) < 0.5f * kSampleRate ? s : 0.f;
              h[j] = nws_tf32_hi(s);
              l[j] = s - h[j];
            }
          }
          tmem_st8(col_hi + kk, h);
          tmem_st8(col_lo + kk, l);
        }
      }
      tmem_wait_st();            // this thread's operand rows are in tensor memory ...
      nws_tc_fence_before();     // ... and ordered (with the TMEM reads of the previous tile) before the MMA warp's wake-up
      fill_arrive(wg, buf);
      if (buf) ++uses1; else ++uses0;
    }
    }
    // the noise-branch sample that is added to the mixdown at the very end: its latency hides behind the shaper loop
    // the noise-branch sample that is added to the mixdown at the very end: its latency hides behind the shaper loop
    const float noise_in_v = (!nz_on && p.noise_in) ? __ldcs(p.noise_in + (size_t)b * N + n) : 0.f;   // (aliases p.out: not the read-only path; streaming: read once, keep L1 for the table)
    {  // accumulator complete when the last stage's commit lands (a commit covers all earlier MMAs)
      const int lb = (C::NST - 1) & 1;
      const uint32_t u = lb ? uses1 : uses0;
      if (ok) ok = nws_mbar_wait(&free_bar[wg][lb], (u - 1) & 1);
      nws_tc_fence_after();
    }

    // ---- FiLM -> shaper -> FiLM -> mixdown (shaping.py:67-79), exciter read from this thread's TMEM lane
    const float4* cf = reinterpret_cast<const float4*>(sm_coef + (wt >> 6) * kShapers * 8);
    const float l1 = lc.l1;
    // FastNEWT: MODE 2 = the table size is the reference's default 4096 (shaping.py:101), known at compile time
    // so a row offset is an immediate; MODE 1 = any size.  Offsets are 32-bit (a table is < 4 GB).
    const int lut_size = (USE_LUT && MODE == 2) ? 4096 : p.lut_size;
    float mix_a = 0.f, mix_d = 0.f;
    if constexpr (!USE_LUT && SHP == 1) {
      // NEWT (shaping.py:15-37) with the two 8x8 layers of every shaper on the tensor cores (mma.sync m16n8k8, 3xTF32):
      // 128 of the 144 multiply-adds per shaper-sample and all but 5 of the 46 weight loads leave the threads'
      // instruction stream.  A warp's 32 samples are two 16-row tiles; thread (g = lane / 4, t = lane % 4) works on
      // samples g + 8r (r = 0..3) and hidden units 2t, 2t + 1 — the fragment layout of the instruction — from the
      // first layer on, so nothing is transposed between layers.  Per channel: the sample's own lane applies FiLM-in and
      // the input scale, four shuffles hand each thread its four samples, and after the last layer a 4-lane
      // reduce-scatter of the 8-term dot product leaves thread (g, t) with the output of sample g + 8t, whose
      // mixdown it accumulates; one shuffle per tile returns the sums to the samples' own lanes.
      const int g = lane >> 2, t = lane & 3;
      float acc_a = 0.f, acc_d = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kShapers; c0 += 4) {
        float ev[4];
        tmem_ld<4>(tmem_lane + c0, ev);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + i;
          if (TAP) p.exciter_out[((size_t)b * kShapers + c) * N + n] = ev[i] + sm_bw[c].x;
          const float4 ci = cf[2 * c], cn = cf[2 * c + 1];
          const float u = cn.z * fmaf(l1, fmaf(ci.y, ev[i], ci.w), fmaf(ci.x, ev[i], ci.z));   // input_scale * FiLM-in
          const float* rec = sm_shaper + c * kShpTcStride;
          const float4* sl = reinterpret_cast<const float4*>(rec + kShpTcSmall + t * kShpTcSlot);
          const float4 s0 = sl[0], s2 = sl[5];
          float h[4][2];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float us = __shfl_sync(0xffffffffu, u, g + 8 * r);
            h[r][0] = NWS_SHAPER_SIN(fmaf(s0.x, us, s0.z));
            h[r][1] = NWS_SHAPER_SIN(fmaf(s0.y, us, s0.w));
          }
          nws_shaper_layer_mma(h, reinterpret_cast<const float4*>(rec)[lane], sl + 1);
#pragma unroll
          for (int r = 0; r < 4; ++r) { h[r][0] = NWS_SHAPER_SIN_INNER(h[r][0]); h[r][1] = NWS_SHAPER_SIN_INNER(h[r][1]); }
          nws_shaper_layer_mma(h, reinterpret_cast<const float4*>(rec)[32 + lane], sl + 3);
          float pr[4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
            pr[r] = fmaf(s2.y, NWS_SHAPER_SIN_INNER(h[r][1]), s2.x * NWS_SHAPER_SIN_INNER(h[r][0]));
          // 4-lane reduce-scatter: thread t ends with the full dot product of sample g + 8t
          const bool t2 = t & 2, t1 = t & 1;
          float k0 = t2 ? pr[2] : pr[0], k1 = t2 ? pr[3] : pr[1];
          k0 += __shfl_xor_sync(0xffffffffu, t2 ? pr[0] : pr[2], 2);
          k1 += __shfl_xor_sync(0xffffffffu, t2 ? pr[1] : pr[3], 2);
          const float tot = (t1 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, t1 ? k0 : k1, 1);
          const float y = NWS_SHAPER_SIN_INNER(tot + s2.z);
          acc_a = fmaf(cn.x, y, acc_a);
          acc_d = fmaf(cn.y, y, acc_d);
        }
      }
      const int src = 4 * (lane & 7) + (lane >> 3);   // thread (g, t) holds sample g + 8t
      mix_a = __shfl_sync(0xffffffffu, acc_a, src);
      mix_d = __shfl_sync(0xffffffffu, acc_d, src);
    } else if constexpr (!USE_LUT) {
      // NEWT (shaping.py:15-37).  The shaper weights depend on the channel only, so neighbouring lanes pair up: of
      // the channel pair (c0, c0+1) the even lane evaluates channel c0 and the odd lane channel c0+1, each for BOTH
      // lanes' samples — one set of weight loads per two shaper-samples (the loop was shared-memory-pipe bound at 46
      // broadcast LDS.128 per shaper-sample), at the price of one shuffle per channel pair.  FiLM-in is applied by
      // the sample's own lane before the exchange; the FiLM-out / mixdown coefficients depend on (half hop, channel)
      // and both lanes are in the same half hop, so each lane accumulates its channel's contribution to both samples
      // and the two partial mixdowns are swapped back once per tile.
      const bool odd = lane & 1;
      float own_a = 0.f, own_d = 0.f, oth_a = 0.f, oth_d = 0.f;   // partial mixdowns: own sample / the partner lane's
#pragma unroll 1
      for (int c0 = 0; c0 < kShapers; c0 += 2) {
        float ev[2];
        tmem_ld<2>(tmem_lane + c0, ev);
        float x[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = c0 + i;
          if (TAP) p.exciter_out[((size_t)b * kShapers + c) * N + n] = ev[i] + sm_bw[c].x;
          const float4 ci = cf[2 * c];
          x[i] = fmaf(l1, fmaf(ci.y, ev[i], ci.w), fmaf(ci.x, ev[i], ci.z));
        }
        const float mine = odd ? x[1] : x[0];
        const float theirs = __shfl_xor_sync(0xffffffffu, odd ? x[0] : x[1], 1);   // the partner's sample, this lane's channel
        const int c = c0 + (odd ? 1 : 0);
        float y_own, y_oth;
        nws_shaper_mlp2<MODE>(sm_shaper + c * kShaperStride, mine, theirs, y_own, y_oth);
        const float2 cn = *reinterpret_cast<const float2*>(&cf[2 * c + 1]);
        own_a = fmaf(cn.x, y_own, own_a); own_d = fmaf(cn.y, y_own, own_d);
        oth_a = fmaf(cn.x, y_oth, oth_a); oth_d = fmaf(cn.y, y_oth, oth_d);
      }
      mix_a = own_a + __shfl_xor_sync(0xffffffffu, oth_a, 1);
      mix_d = own_d + __shfl_xor_sync(0xffffffffu, oth_d, 1);
    } else {
#pragma unroll 1
    for (int c0 = 0; c0 < kShapers; c0 += C::kChPerLd) {
      float ev[C::kChPerLd];
      tmem_ld<C::kChPerLd>(tmem_lane + c0, ev);
      // row pointer of channel c0: the rows of the other channels of the group are immediates of the load
      const float2* lut_row = p.lut2 + (size_t)(c0 * lut_size);
      asm("" : "+l"(lut_row));   // keep the row pointer a 64-bit value of its own: per channel one IMAD.WIDE + load
#pragma unroll
      for (int i = 0; i < C::kChPerLd; ++i) {
        const int c = c0 + i;
        if (TAP) p.exciter_out[((size_t)b * kShapers + c) * N + n] = ev[i] + sm_bw[c].x;
        const float4 ci = cf[2 * c];
        const float2 cn = *reinterpret_cast<const float2*>(&cf[2 * c + 1]);
        // FastNEWT.shaping_fn (shaping.py:136-151): FiLM-in and the table position in one affine map (coefficient
        // table above); floor and clamp done on the integer side (F2I.FLOOR saturates, NaN -> 0 with a NaN fract, as
        // floorf/fmaxf/fminf give); the table row holds (L, U - L) pairs so one 8-byte load feeds (U - L) * fract + L
        const float idx = fmaf(l1, fmaf(ci.y, ev[i], ci.w), fmaf(ci.x, ev[i], ci.z));
        const int fi = nws_min_relu(__float2int_rd(idx), lut_size - 1);   // clamp to [0, size-1]: one VIMNMX.RELU
        const float2 t2 = __ldg(lut_row + i * lut_size + (uint32_t)fi);
        const float y = fmaf(t2.y, NWS_ADD(idx, -(float)fi), t2.x);
        mix_a = fmaf(cn.x, y, mix_a);
        mix_d = fmaf(cn.y, y, mix_d);
      }
    }
    }
    nws_tc_fence_before();   // TMEM reads ordered before the next tile's first MMA (via the warpgroup barrier)
    float o = fmaf(l1, mix_d + mix_kd, mix_a + mix_ka) + mix_b;
    float noise_v = noise_in_v;
    if (nz_on) {   // filtered by the MMA warp during the previous tile's epilogue (long done: one try_wait)
      if (ok) ok = nws_mbar_wait(&nz_bar[wg][par], (nz_phase >> par) & 1u);
      nz_phase ^= 1u << par;
      noise_v = nz_out[par * 128 + wt];
    }
    o += noise_v;
    p.out[(size_t)b * N + n] = ok ? o : __int_as_float(0x7fc00000);
    tile = tile_next;
    par ^= 1;
  }
  if (wt == 0) done_s[wg] = 1;
  fill_arrive(wg, 0);   // the MMA warp is waiting for stage 0 of a tile that will not come
  }
  if (!ok && fault) *(volatile int*)fault = 1;   // mapped host memory: a plain store, no atomic over PCIe
  nws_tc_fence_before();
  __syncthreads();
  // The scheduler's counter pair {tiles claimed, CTAs done} lives in the context and is zero between launches: the last
  // CTA to leave puts it back (no memset launch per forward; safe under CUDA-graph replay, unlike a host-side toggle).
  if (tid == 0 && atomicAdd(p.tile_counter + 1, 1) == (int)gridDim.x - 1) {
    p.tile_counter[0] = 0;
    p.tile_counter[1] = 0;
  }
  if (tid < 32) nws_tmem_dealloc(tmem_base_s, C::kTmemColsWg * kWgs);
}

// one instantiation: opt in to its dynamic shared memory once per device, launch
template <bool USE_LUT, bool TAP, int MODE, int SHP>
int launch_variant(const NwsAudioParams& p, const float* wu, int* fault, int grid, cudaStream_t s, bool pdl) {
  static bool attr_done[64] = {};
  if (nws_first_use_on_device(attr_done))
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_audio_tc_kernel<USE_LUT, TAP, MODE, SHP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TcCfg<USE_LUT>::kBytes));
  // the noise branch's area (the last of the layout) is only paid for when that branch runs in this kernel: what is not
  // shared memory is L1 for the table gathers
  const int smem_bytes = p.bands ? TcCfg<USE_LUT>::kBytes : TcCfg<USE_LUT>::oNoise;
  if (!pdl) {
    nws_audio_tc_kernel<USE_LUT, TAP, MODE, SHP><<<grid, kTcThreads, smem_bytes, s>>>(p, wu, fault);
    return NWS_OK;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int n_attr = 0;
  nws_pdl_config(&cfg, attr, &n_attr, true);
  NWS_CUDA_OK(cudaLaunchKernelEx(&cfg, nws_audio_tc_kernel<USE_LUT, TAP, MODE, SHP>, p, wu, fault));
  return NWS_OK;
}

}  // namespace

int nws_launch_audio_tc(const NwsContext* ctx, const float* f0, const double* carry, const float* film,
                        const float* u_phase, const float* noise_in, float* out, float* exciter_out, int B, int T,
                        int t_begin, int t_end, int* tile_counter, int use_lut, cudaStream_t s, int max_ctas, bool pdl,
                        const float* bands, const float2* xspec) {
  NwsAudioParams p{};
  p.bands = bands; p.xspec = xspec; p.tw_master = ctx->tw_master;
  const float* w = ctx->packed;
  p.f0 = f0; p.carry = carry; p.film = film; p.u_phase = u_phase;
  p.hmix_wt = w + ctx->lay.hmix_wt; p.hmix_b = w + ctx->lay.hmix_b; p.rand_phase = w + ctx->lay.rand_phase;
  p.shaper = w + ctx->lay.shaper; p.mix_w = w + ctx->lay.mix_w; p.mix_b = w + ctx->lay.mix_b;
  p.lut = ctx->lut; p.lut2 = ctx->lut2; p.lut_size = ctx->lut_size; p.lut_min = ctx->lut_min;
  p.lut_span = ctx->lut_max - ctx->lut_min;
  p.lut_span_rcp = 1.0f / p.lut_span;
  p.noise_in = noise_in; p.out = out; p.exciter_out = exciter_out; p.B = B; p.T = T;
  p.t_begin = t_begin; p.t_end = t_end; p.tile_counter = tile_counter;
  static const int chunk_env = getenv("NWS_TILE_CHUNK") ? atoi(getenv("NWS_TILE_CHUNK")) : 0;   // development knob
  p.tile_chunk = chunk_env >= 1 && chunk_env <= 64 ? chunk_env : kTileChunk;

  const long long tiles = (long long)B * (t_end - t_begin);
  if (tiles >= (1ll << 31) - 4096 || t_end <= t_begin) { nws_set_error("nws_launch_audio_tc: bad tile count"); return NWS_ERR_INVALID; }
  p.hops_magic = (uint32_t)((1ull << 32) / (uint64_t)(t_end - t_begin) > 0xffffffffull ? 0xffffffffull : (1ull << 32) / (uint64_t)(t_end - t_begin));
  const long long want = (tiles + kWgs - 1) / kWgs;
  const int cap = max_ctas > 0 && max_ctas < ctx->sm_count ? max_ctas : ctx->sm_count;   // SMs left to a concurrent encoder
  const int grid = (int)(want < cap ? want : cap);
  const float* wu = w + ctx->lay.hmix_umma;
  const bool direct = ctx->shaper_inner_bound <= 8.0f;   // see NWS_SHAPER_SIN_INNER
  const bool tap = exciter_out != nullptr;
  const int shp = ctx->shaper_impl;
  int rc;
  if (use_lut) {
    if (!tap && ctx->lut_size == 4096) rc = launch_variant<true, false, 2, 0>(p, wu, ctx->fault_dev, grid, s, pdl);
    else if (!tap) rc = launch_variant<true, false, 1, 0>(p, wu, ctx->fault_dev, grid, s, pdl);
    else rc = launch_variant<true, true, 1, 0>(p, wu, ctx->fault_dev, grid, s, pdl);
  } else if (shp) {
    if (!tap) rc = direct ? launch_variant<false, false, 2, 1>(p, wu, ctx->fault_dev, grid, s, pdl) : launch_variant<false, false, 1, 1>(p, wu, ctx->fault_dev, grid, s, pdl);
    else rc = direct ? launch_variant<false, true, 2, 1>(p, wu, ctx->fault_dev, grid, s, pdl) : launch_variant<false, true, 1, 1>(p, wu, ctx->fault_dev, grid, s, pdl);
  } else {
    if (!tap) rc = direct ? launch_variant<false, false, 2, 0>(p, wu, ctx->fault_dev, grid, s, pdl) : launch_variant<false, false, 1, 0>(p, wu, ctx->fault_dev, grid, s, pdl);
    else rc = direct ? launch_variant<false, true, 2, 0>(p, wu, ctx->fault_dev, grid, s, pdl) : launch_variant<false, true, 1, 0>(p, wu, ctx->fault_dev, grid, s, pdl);
  }
  if (rc) return rc;
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
