// Audio-rate part of the NWS forward, fused into ONE kernel per batch:
//   f0 upsample -> fp64-carried phase -> 101-harmonic oscillator bank -> harmonic mixer (101->64)
//   -> FiLM -> NEWT shaper MLP or FastNEWT LUT gather+lerp -> FiLM -> 64->1 mixdown (+ noise branch)
// Reference: models/neural_waveshaping.py:64-67,75-86; modules/generators.py:38-66;
// modules/shaping.py:15-37,40-79,82-151; modules/dynamic.py:6-8.
//
// Mapping (stage A, fp32 SIMT): one thread per output sample, one CTA (128 threads) per hop of one
// utterance, persistent CTAs striding over the B*T hop tiles.  Harmonic-mixer rows, phase shifts
// and shaper weights stay resident in shared memory for the CTA's lifetime; nothing audio-rate
// except the final sample is written to HBM (the reference materialises ~34 KB per sample).
#include "nws_audio_common.cuh"
#include "nws_internal.cuh"

// ------------------------------------------------------------------------------------------------
// The two uniform draws of one forward (generators.py:55 then :30) when the caller injects none.
__global__ void nws_rng_kernel(float* __restrict__ u_phase, float* __restrict__ noise, int n_noise, uint64_t seed,
                               uint64_t offset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one Philox block = 4 values
  const int n_u = (kHarm + 3) / 4, n_n = (n_noise + 3) / 4;
  if (i < n_u) {
    if (!u_phase) return;
    const NwsPhilox4 r = nws_philox4x32_10(offset + (uint64_t)i, 0ull, seed);
    for (int j = 0; j < 4; ++j)
      if (4 * i + j < kHarm) u_phase[4 * i + j] = nws_u32_to_unit(r.v[j]);
  } else if (i - n_u < n_n && noise) {
    const int q = i - n_u;
    const NwsPhilox4 r = nws_philox4x32_10(offset + (uint64_t)q, 1ull, seed);
    for (int j = 0; j < 4; ++j)
      if (4 * q + j < n_noise) noise[4 * q + j] = nws_u32_to_unit(r.v[j]);
  }
}

int nws_launch_rng(float* u_phase, float* noise, int n_noise, uint64_t seed, uint64_t offset, cudaStream_t s) {
  const int total = (kHarm + 3) / 4 + (n_noise + 3) / 4;
  nws_rng_kernel<<<(total + 255) / 256, 256, 0, s>>>(u_phase, noise, n_noise, seed, offset);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
// Weight repacking (load time).  One thread per destination float.
struct NwsPackArgs {
  const float* t[NWS_T_COUNT];
};

__global__ void nws_pack_kernel(NwsPackArgs a, NwsPackedLayout L, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.total) return;
  float v = 0.f;
  auto in = [&](int base, int n) { return i >= base && i < base + n; };
  if (in(L.gru_whh, kGates * kEmb)) v = a.t[NWS_T_GRU_W_HH][i - L.gru_whh];
  else if (in(L.gru_wih, kGates * 2)) v = a.t[NWS_T_GRU_W_IH][i - L.gru_wih];
  else if (in(L.gru_bih, kGates)) v = a.t[NWS_T_GRU_B_IH][i - L.gru_bih];
  else if (in(L.gru_bhh, kGates)) v = a.t[NWS_T_GRU_B_HH][i - L.gru_bhh];
  else if (in(L.proj_wt, kEmb * kEmb)) { const int j = i - L.proj_wt, k = j / kEmb, o = j % kEmb; v = a.t[NWS_T_PROJ_W][o * kEmb + k]; }
  else if (in(L.proj_b, kEmb)) v = a.t[NWS_T_PROJ_B][i - L.proj_b];
  else if (in(L.hmix_wt, kHarmPad * kShapers)) { const int j = i - L.hmix_wt, k = j / kShapers, c = j % kShapers; v = k < kHarm ? a.t[NWS_T_HMIX_W][c * kHarm + k] : 0.f; }
  else if (in(L.hmix_b, kShapers)) v = a.t[NWS_T_HMIX_B][i - L.hmix_b];
  else if (in(L.hmix_umma, 2 * kHarmPad * kShapers)) {
    // canonical no-swizzle K-major layout of the [64 x 104] B operand (see nws_tc.cuh): float index q ->
    // chunk (4 k's) = q / 256, 8-row group = (q % 256) / 32, row in group = (q % 32) / 4, k in chunk = q % 4
    const int j = i - L.hmix_umma, part = j / (kHarmPad * kShapers), q = j % (kHarmPad * kShapers);
    const int k = (q / 256) * 4 + (q & 3), c = ((q % 256) / 32) * 8 + ((q & 31) >> 2);
    const float w = k < kHarm ? a.t[NWS_T_HMIX_W][c * kHarm + k] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    v = part == 0 ? hi : __uint_as_float(__float_as_uint(w - hi) & 0xffffe000u);
  }
  else if (in(L.mix_w, kShapers)) v = a.t[NWS_T_MIX_W][i - L.mix_w];
  else if (in(L.mix_b, 1)) v = a.t[NWS_T_MIX_B][0];
  else if (in(L.rand_phase, kHarmPad)) { const int k = i - L.rand_phase; v = k < kHarm ? a.t[NWS_T_OSC_RAND_PHASE][k] : 0.f; }
  else if (in(L.ir, kReverbIr)) { const int k = i - L.ir; v = k == 0 ? 0.f : a.t[NWS_T_REVERB_IR][k - 1]; }
  else if (in(L.shaper, kShapers * kShaperStride)) {
    const int j = i - L.shaper, c = j / kShaperStride, q = j % kShaperStride;
    if (q == kShpScale) v = a.t[NWS_T_SHAPER_SCALE][c];
    else if (q == kShpB4) v = a.t[NWS_T_SHAPER_B4][c];
    else if (q >= kShpW1 && q < kShpW1 + 8) v = a.t[NWS_T_SHAPER_W1][c * 8 + (q - kShpW1)];
    else if (q >= kShpB1 && q < kShpB1 + 8) v = a.t[NWS_T_SHAPER_B1][c * 8 + (q - kShpB1)];
    else if (q >= kShpW2 && q < kShpW2 + 64) v = a.t[NWS_T_SHAPER_W2][c * 64 + (q - kShpW2)];  // [c*8+j][i]
    else if (q >= kShpB2 && q < kShpB2 + 8) v = a.t[NWS_T_SHAPER_B2][c * 8 + (q - kShpB2)];
    else if (q >= kShpW3 && q < kShpW3 + 64) v = a.t[NWS_T_SHAPER_W3][c * 64 + (q - kShpW3)];
    else if (q >= kShpB3 && q < kShpB3 + 8) v = a.t[NWS_T_SHAPER_B3][c * 8 + (q - kShpB3)];
    else if (q >= kShpW4 && q < kShpW4 + 8) v = a.t[NWS_T_SHAPER_W4][c * 8 + (q - kShpW4)];
  } else {
    for (int m = 0; m < 2; ++m) {
      const NwsTdMlpOffsets& o = L.mlp[m];
      const int tb = m == 0 ? NWS_T_FILM_MLP : NWS_T_NOISE_MLP;
      const int n_out = m == 0 ? kFilm : kBands;
      for (int l = 0; l < 3; ++l) {
        if (in(o.wt[l], kEmb * kEmb)) { const int j = i - o.wt[l], k = j / kEmb, oc = j % kEmb; v = a.t[tb + 4 * l][oc * kEmb + k]; }
        else if (in(o.b[l], kEmb)) v = a.t[tb + 4 * l + 1][i - o.b[l]];
        else if (in(o.g[l], kEmb)) v = a.t[tb + 4 * l + 2][i - o.g[l]];
        else if (in(o.beta[l], kEmb)) v = a.t[tb + 4 * l + 3][i - o.beta[l]];
      }
      if (in(o.wt_out, kEmb * o.ld_out)) { const int j = i - o.wt_out, k = j / o.ld_out, oc = j % o.ld_out; v = oc < n_out ? a.t[tb + 12][oc * kEmb + k] : 0.f; }
      else if (in(o.b_out, o.ld_out)) { const int oc = i - o.b_out; v = oc < n_out ? a.t[tb + 13][oc] : 0.f; }
    }
  }
  dst[i] = v;
}

int nws_launch_pack_weights(NwsContext* ctx, const float* const* tensors, cudaStream_t s) {
  NwsPackArgs a;
  for (int i = 0; i < NWS_T_COUNT; ++i) a.t[i] = tensors[i];
  nws_pack_kernel<<<(ctx->lay.total + 255) / 256, 256, 0, s>>>(a, ctx->lay, ctx->packed);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// FastNEWT._init_lookup_table (shaping.py:107-119): table[c][i] = shaper_c(linspace(min,max,size)[i]).
__global__ void nws_build_lut_kernel(const float* __restrict__ shaper, const float* __restrict__ points,
                                     float* __restrict__ lut, int table_size, float tmin, float tmax) {
  const int c = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= table_size) return;
  const float x = points ? points[i] : nws_linspace_value(i, table_size, tmin, tmax);
  lut[(size_t)c * table_size + i] = nws_shaper_mlp<0>(shaper + c * kShaperStride, x);
}

// (T[i], T[min(i+1,size-1)] - T[i]): the difference is the same fp32 subtraction FastNEWT.shaping_fn
// performs per sample (shaping.py:150), so (U - L) * fract + L stays bit-identical while the fused
// kernels fetch both operands with one 8-byte load.
__global__ void nws_pair_lut_kernel(const float* __restrict__ lut, float2* __restrict__ lut2, int size) {
  const int c = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= size) return;
  const float lo = lut[(size_t)c * size + i], up = lut[(size_t)c * size + (i + 1 < size ? i + 1 : size - 1)];
  lut2[(size_t)c * size + i] = make_float2(lo, NWS_ADD(up, -lo));
}

int nws_launch_pair_lut(NwsContext* ctx, cudaStream_t s) {
  dim3 grid((ctx->lut_size + 255) / 256, kShapers);
  nws_pair_lut_kernel<<<grid, 256, 0, s>>>(ctx->lut, ctx->lut2, ctx->lut_size);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

int nws_launch_build_lut(const NwsContext* ctx, const float* points, float* lut, int table_size, float tmin,
                         float tmax, cudaStream_t s) {
  dim3 grid((table_size + 127) / 128, kShapers);
  nws_build_lut_kernel<<<grid, 128, 0, s>>>(ctx->packed + ctx->lay.shaper, points, lut, table_size, tmin, tmax);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
constexpr int kAudioThreads = 128;
// shared memory (floats): hmix_wt | hmix_b | shift | mix_w | film[3][256] | e[64][128] | (MLP) shaper
constexpr int kSmWt = 0;
constexpr int kSmHb = kSmWt + kHarmPad * kShapers;
constexpr int kSmShift = kSmHb + kShapers;
constexpr int kSmMixW = kSmShift + kHarmPad;
constexpr int kSmFilm = kSmMixW + kShapers;
constexpr int kSmE = kSmFilm + 3 * kFilm;
constexpr int kSmShaper = kSmE + kShapers * kAudioThreads;
constexpr int kSmFloatsLut = kSmShaper;
constexpr int kSmFloatsMlp = kSmShaper + kShapers * kShaperStride;

template <bool USE_LUT>
__global__ void __launch_bounds__(kAudioThreads, USE_LUT ? 3 : 2) nws_audio_fused_kernel(const NwsAudioParams p) {
  extern __shared__ __align__(16) float sm[];
  __shared__ double warp_tot[kAudioThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T, N = T * kHop;

  // ---- CTA-lifetime staging: mixer rows (k-major), biases, phase shifts, shaper weights
  for (int i = tid; i < kHarmPad * kShapers / 4; i += kAudioThreads)
    reinterpret_cast<float4*>(sm + kSmWt)[i] = reinterpret_cast<const float4*>(p.hmix_wt)[i];
  if (tid < kShapers) {
    sm[kSmHb + tid] = p.hmix_b[tid];
    sm[kSmMixW + tid] = p.mix_w[tid];
  }
  if (tid < kHarmPad) sm[kSmShift + tid] = tid < kHarm ? nws_phase_shift(p.u_phase[tid], p.rand_phase[tid]) : 0.f;
  if (!USE_LUT)
    for (int i = tid; i < kShapers * kShaperStride / 4; i += kAudioThreads)
      reinterpret_cast<float4*>(sm + kSmShaper)[i] = reinterpret_cast<const float4*>(p.shaper)[i];
  const float mix_b = p.mix_b[0];
  const float inv_hop = (float)T / (float)N;
  const long long n_tiles = (long long)p.B * T;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = (int)(tile / T), t = (int)(tile - (long long)b * T);
    __syncthreads();  // previous tile done with film / warp_tot (also orders the staging above)
    // ---- the three film frames this hop interpolates between (frame-major rows of 256)
    for (int i = tid; i < 3 * kFilm / 4; i += kAudioThreads) {
      const int slot = i / (kFilm / 4), fr = t - 1 + slot;
      if (fr >= 0 && fr < T)
        reinterpret_cast<float4*>(sm + kSmFilm)[i] =
            reinterpret_cast<const float4*>(p.film + ((size_t)b * T + fr) * kFilm)[i - slot * (kFilm / 4)];
    }
    // ---- f0 upsample (neural_waveshaping.py:75) and the cumsum of generators.py:59
    const int n = t * kHop + tid;
    const NwsLerp lc = nws_lerp_coords(n, T, inv_hop);
    const float* f0b = p.f0 + (size_t)b * T;
    const float f0u = nws_lerp_apply(lc, f0b[lc.i0], f0b[lc.i1]);
    double v = (double)f0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    double pre = p.carry[(size_t)b * T + t];
    for (int w = 0; w < warp; ++w) pre += warp_tot[w];
    const float csum = (float)(pre + v);
    const float phase = nws_phase_from_cumsum(csum, (float)kSampleRate);

    // ---- oscillator bank + harmonic mixer: acc[c] = b[c] + sum_k W[c][k] * sin(k*phase + shift_k) * mask_k
    float acc[kShapers];
#pragma unroll
    for (int c = 0; c < kShapers; c += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(sm + kSmHb + c);
      acc[c] = bv.x; acc[c + 1] = bv.y; acc[c + 2] = bv.z; acc[c + 3] = bv.w;
    }
#pragma unroll 1
    for (int k = 1; k <= kHarm; ++k) {
      const float arg = nws_harmonic_arg(k, phase, sm[kSmShift + k - 1]);
      float s = nws_sinf(arg);
      s = NWS_MUL(f0u, (float)k) < 0.5f * kSampleRate ? s : 0.f;  // anti-alias mask, generators.py:50-52
      const float4* wr = reinterpret_cast<const float4*>(sm + kSmWt + (k - 1) * kShapers);
#pragma unroll
      for (int c4 = 0; c4 < kShapers / 4; ++c4) {
        const float4 w = wr[c4];
        acc[4 * c4] = fmaf(w.x, s, acc[4 * c4]);
        acc[4 * c4 + 1] = fmaf(w.y, s, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(w.z, s, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(w.w, s, acc[4 * c4 + 3]);
      }
    }
    // park the 64 exciter channels in this thread's private column of shared memory so the shaper
    // loop can stay rolled (dynamic channel index); no other thread touches column `tid`
    float* ecol = sm + kSmE + tid;
#pragma unroll
    for (int c = 0; c < kShapers; ++c) ecol[c * kAudioThreads] = acc[c];
    if (p.exciter_out) {
#pragma unroll 4
      for (int c = 0; c < kShapers; ++c) p.exciter_out[((size_t)b * kShapers + c) * N + n] = ecol[c * kAudioThreads];
    }

    // ---- FiLM -> shaper -> FiLM -> mixdown (shaping.py:67-79)
    const float* fa = sm + kSmFilm + (lc.i0 - (t - 1)) * kFilm;
    const float* fb = sm + kSmFilm + (lc.i1 - (t - 1)) * kFilm;
    float mix = 0.f;
#pragma unroll 2
    for (int c = 0; c < kShapers; ++c) {
      const float e = ecol[c * kAudioThreads];
      const float g_i = nws_lerp_apply(lc, fa[c], fb[c]);
      const float b_i = nws_lerp_apply(lc, fa[kShapers + c], fb[kShapers + c]);
      const float g_n = nws_lerp_apply(lc, fa[2 * kShapers + c], fb[2 * kShapers + c]);
      const float b_n = nws_lerp_apply(lc, fa[3 * kShapers + c], fb[3 * kShapers + c]);
      const float x = NWS_ADD(NWS_MUL(g_i, e), b_i);  // FiLM: gamma * x + beta (dynamic.py:8)
      float y;
      if (USE_LUT) {
        const NwsLutIdx li = nws_lut_index(x, p.lut_size, p.lut_min, p.lut_span, p.lut_span_rcp);
        const float* row = p.lut + (size_t)c * p.lut_size;
        y = nws_lut_lerp(__ldg(row + li.lower), __ldg(row + li.upper), li.fract);
      } else {
        y = nws_shaper_mlp<1>(sm + kSmShaper + c * kShaperStride, x);
      }
      const float z = NWS_ADD(NWS_MUL(g_n, y), b_n);
      mix = fmaf(sm[kSmMixW + c], z, mix);
    }
    float o = mix + mix_b;
    if (p.noise_in) o += p.noise_in[(size_t)b * N + n];
    p.out[(size_t)b * N + n] = o;
  }
}

int nws_launch_audio(const NwsContext* ctx, const float* f0, const double* carry, const float* film,
                     const float* u_phase, const float* noise_in, float* out, float* exciter_out, int B, int T,
                     int use_lut, cudaStream_t s) {
  NwsAudioParams p{};
  const float* w = ctx->packed;
  p.f0 = f0; p.carry = carry; p.film = film; p.u_phase = u_phase;
  p.hmix_wt = w + ctx->lay.hmix_wt; p.hmix_b = w + ctx->lay.hmix_b; p.rand_phase = w + ctx->lay.rand_phase;
  p.shaper = w + ctx->lay.shaper; p.mix_w = w + ctx->lay.mix_w; p.mix_b = w + ctx->lay.mix_b;
  p.lut = ctx->lut; p.lut2 = ctx->lut2; p.lut_size = ctx->lut_size; p.lut_min = ctx->lut_min;
  p.lut_span = ctx->lut_max - ctx->lut_min;
  p.lut_span_rcp = 1.0f / p.lut_span;
  p.noise_in = noise_in; p.out = out; p.exciter_out = exciter_out; p.B = B; p.T = T;

  static bool attr_done[64] = {};
  const size_t sm_lut = kSmFloatsLut * sizeof(float), sm_mlp = kSmFloatsMlp * sizeof(float);
  if (nws_first_use_on_device(attr_done)) {
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_audio_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_lut));
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_audio_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_mlp));
  }
  const long long tiles = (long long)B * T;
  const int per_sm = use_lut ? 3 : 2;
  const int grid = (int)(tiles < (long long)ctx->sm_count * per_sm ? tiles : (long long)ctx->sm_count * per_sm);
  if (use_lut)
    nws_audio_fused_kernel<true><<<grid, kAudioThreads, sm_lut, s>>>(p);
  else
    nws_audio_fused_kernel<false><<<grid, kAudioThreads, sm_mlp, s>>>(p);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
// Stand-alone TrainableNonlinearity evaluation on a shared grid: out[c][i] = shaper_c(x[i]).
// This is exactly FastNEWT._init_lookup_table (shaping.py:107-119); it needs only the nine
// shaping_fn tensors, so FastNEWT(newt) can be constructed before a full model is loaded.
__global__ void nws_pack_shaper_kernel(const float* scale, const float* w1, const float* b1, const float* w2,
                                       const float* b2, const float* w3, const float* b3, const float* w4,
                                       const float* b4, float* __restrict__ dst) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= kShapers * kShaperStride) return;
  const int c = j / kShaperStride, q = j % kShaperStride;
  float v = 0.f;
  if (q == kShpScale) v = scale[c];
  else if (q == kShpB4) v = b4[c];
  else if (q >= kShpW1 && q < kShpW1 + 8) v = w1[c * 8 + (q - kShpW1)];
  else if (q >= kShpB1 && q < kShpB1 + 8) v = b1[c * 8 + (q - kShpB1)];
  else if (q >= kShpW2 && q < kShpW2 + 64) v = w2[c * 64 + (q - kShpW2)];
  else if (q >= kShpB2 && q < kShpB2 + 8) v = b2[c * 8 + (q - kShpB2)];
  else if (q >= kShpW3 && q < kShpW3 + 64) v = w3[c * 64 + (q - kShpW3)];
  else if (q >= kShpB3 && q < kShpB3 + 8) v = b3[c * 8 + (q - kShpB3)];
  else if (q >= kShpW4 && q < kShpW4 + 8) v = w4[c * 8 + (q - kShpW4)];
  dst[j] = v;
}

extern "C" size_t nws_shaper_eval_scratch_bytes(void) { return (size_t)kShapers * kShaperStride * sizeof(float); }

extern "C" int nws_shaper_eval(const float* const* shaper_tensors, const float* x, float* out, int n_points,
                               void* scratch, void* stream) {
  if (!shaper_tensors || !x || !out || !scratch || n_points < 1) { nws_set_error("nws_shaper_eval: bad argument"); return NWS_ERR_INVALID; }
  for (int i = 0; i < 9; ++i)
    if (!shaper_tensors[i]) { nws_set_error("nws_shaper_eval: tensor %d is NULL", i); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  float* packed = (float*)scratch;
  nws_pack_shaper_kernel<<<(kShapers * kShaperStride + 255) / 256, 256, 0, s>>>(
      shaper_tensors[0], shaper_tensors[1], shaper_tensors[2], shaper_tensors[3], shaper_tensors[4], shaper_tensors[5],
      shaper_tensors[6], shaper_tensors[7], shaper_tensors[8], packed);
  NWS_LAUNCH_CHECK();
  dim3 grid((n_points + 127) / 128, kShapers);
  nws_build_lut_kernel<<<grid, 128, 0, s>>>(packed, x, out, n_points, 0.f, 1.f);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
// FastNEWT.shaping_fn on a materialised input (shaping.py:136-151): the same device functions the
// fused kernel uses (nws_lut_index / nws_lut_lerp), exposed so the parity tests can check the index
// path bit-for-bit on identical shaper inputs.  x, y: [B, 64, N].
__global__ void nws_lut_lookup_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ lut,
                                      int lut_size, float tmin, float span, float span_rcp, int N, long long total,
                                      int* __restrict__ lower_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)((i / N) % kShapers);
  const NwsLutIdx li = nws_lut_index(x[i], lut_size, tmin, span, span_rcp);
  const float* row = lut + (size_t)c * lut_size;
  y[i] = nws_lut_lerp(__ldg(row + li.lower), __ldg(row + li.upper), li.fract);
  if (lower_out) lower_out[i] = li.lower;
}

extern "C" int nws_stage_lut_lookup(NwsHandle ctx, const float* x, float* y, int* lower_out, int B, int N, void* stream) {
  if (!ctx || !x || !y || B < 1 || N < 1) { nws_set_error("nws_stage_lut_lookup: bad argument"); return NWS_ERR_INVALID; }
  if (!ctx->lut_valid) { nws_set_error("nws_stage_lut_lookup: no lookup table"); return NWS_ERR_STATE; }
  const long long total = (long long)B * kShapers * N;
  const float span = ctx->lut_max - ctx->lut_min;
  nws_lut_lookup_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, y, ctx->lut, ctx->lut_size, ctx->lut_min, span, 1.0f / span, N, total, lower_out);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
