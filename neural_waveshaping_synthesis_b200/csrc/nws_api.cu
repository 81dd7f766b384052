// C ABI of libnws_b200.so (include/nws_b200.h): handle management, weight loading, LUT management,
// workspace carving and the orchestration of the forward pass
// (NeuralWaveshaping.forward, models/neural_waveshaping.py:74-90).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "nws_internal.cuh"

thread_local uint64_t g_nws_launches = 0;
static thread_local char g_err[512] = "";

void nws_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* nws_last_error(void) { return g_err; }
extern "C" int nws_api_version(void) { return NWS_API_VERSION; }
extern "C" uint64_t nws_launch_count(int reset) {
  const uint64_t v = g_nws_launches;
  if (reset) g_nws_launches = 0;
  return v;
}

extern "C" void nws_default_config(NwsConfig* c) {
  if (!c) return;
  c->sample_rate = kSampleRate; c->control_hop = kHop; c->n_harmonics = kHarm; c->n_waveshapers = kShapers;
  c->embedding_size = kEmb; c->shaping_fn_size = 8; c->shaping_fn_depth = 4; c->noise_bands = kBands;
  c->ir_length = kIr; c->reverb_length = kReverbIr;
}

extern "C" int nws_create(const NwsConfig* cfg, NwsHandle* out) {
  if (!cfg || !out) { nws_set_error("nws_create: NULL argument"); return NWS_ERR_INVALID; }
  NwsConfig d;
  nws_default_config(&d);
  if (memcmp(cfg, &d, sizeof(d)) != 0) {
    nws_set_error("nws_create: only the gin/models/newt.gin configuration is built "
                  "(sr 16000, hop 128, 101 harmonics, 64 shapers of width 8 depth 4, 128-d embedding, "
                  "129 noise bands, 256-tap IR, 32000-tap reverb)");
    return NWS_ERR_UNSUPPORTED;
  }
  NwsContext* ctx = new NwsContext();
  ctx->cfg = *cfg;
  ctx->lay = nws_packed_layout();
  cudaError_t e = cudaGetDevice(&ctx->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->packed, (size_t)ctx->lay.total * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->mlp_tc, nws_mlp_tc_blob_floats() * sizeof(float));
  if (e != cudaSuccess) {
    nws_set_error("nws_create: %s (a CUDA device is required; there is no CPU fallback)", cudaGetErrorString(e));
    delete ctx;
    return NWS_ERR_CUDA;
  }
  e = cudaHostAlloc((void**)&ctx->fault_host, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *ctx->fault_host = 0; e = cudaHostGetDevicePointer((void**)&ctx->fault_dev, ctx->fault_host, 0); }
  if (e != cudaSuccess) { nws_set_error("nws_create: %s", cudaGetErrorString(e)); nws_destroy(ctx); return NWS_ERR_CUDA; }
  e = cudaMalloc(&ctx->dir_counters, (kDirCounters + 4 + kMaxTimeBlocks) * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(ctx->dir_counters, 0, (kDirCounters + 4 + kMaxTimeBlocks) * sizeof(int));
  ctx->tile_counters = ctx->dir_counters ? ctx->dir_counters + kDirCounters : nullptr;
  ctx->gru_done = ctx->dir_counters ? ctx->dir_counters + kDirCounters + 4 : nullptr;
  if (e != cudaSuccess) { nws_set_error("nws_create: %s", cudaGetErrorString(e)); nws_destroy(ctx); return NWS_ERR_CUDA; }
  int rc = nws_make_twiddle_master(ctx);
  if (rc) { nws_destroy(ctx); return rc; }
  // internal encoder stream + fork/join events of the pipelined forward (timing disabled: cheaper)
  {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    e = cudaStreamCreateWithPriority(&ctx->enc_stream, cudaStreamNonBlocking, prio_hi);   // GRU CTAs first when SMs free up
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  for (int i = 0; i < kMaxTimeBlocks && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_block[i], cudaEventDisableTiming);
  for (int i = 0; i < kMaxTimeBlocks && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_mlp[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_early_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_early_done, cudaEventDisableTiming);
  if (e != cudaSuccess) { nws_set_error("nws_create: %s", cudaGetErrorString(e)); nws_destroy(ctx); return NWS_ERR_CUDA; }
  *out = ctx;
  return NWS_OK;
}

extern "C" int nws_destroy(NwsHandle ctx) {
  if (!ctx) return NWS_OK;
  nws_reverb_free_plans(ctx);
  for (int i = 0; i < 2 * kStCount; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->enc_stream) { cudaStreamSynchronize(ctx->enc_stream); cudaStreamDestroy(ctx->enc_stream); }
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
  if (ctx->ev_early_ready) cudaEventDestroy(ctx->ev_early_ready);
  if (ctx->ev_early_done) cudaEventDestroy(ctx->ev_early_done);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (int i = 0; i < kMaxTimeBlocks; ++i) {
    if (ctx->ev_block[i]) cudaEventDestroy(ctx->ev_block[i]);
    if (ctx->ev_mlp[i]) cudaEventDestroy(ctx->ev_mlp[i]);
  }
  cudaFree(ctx->packed);
  cudaFree(ctx->mlp_tc);
  cudaFree(ctx->lut);
  cudaFree(ctx->lut2);
  cudaFree(ctx->tw_master);
  cudaFree(ctx->dir_counters);
  if (ctx->fault_host) cudaFreeHost(ctx->fault_host);
  delete ctx;
  return NWS_OK;
}

extern "C" size_t nws_tensor_numel(int index) {
  if (index < 0 || index >= NWS_T_COUNT) return 0;
  auto mlp = [](int i, int n_out) -> size_t {   // 14 tensors of a TimeDistributedMLP (dynamic.py:20-40), depth 4
    if (i == 12) return (size_t)n_out * kEmb;       // net.9.weight
    if (i == 13) return (size_t)n_out;              // net.9.bias
    return (i & 3) == 0 ? (size_t)kEmb * kEmb : (size_t)kEmb;   // net.{0,3,6}.weight | bias, LN weight, LN bias
  };
  if (index >= NWS_T_FILM_MLP && index < NWS_T_FILM_MLP + 14) return mlp(index - NWS_T_FILM_MLP, kFilm);
  if (index >= NWS_T_NOISE_MLP && index < NWS_T_NOISE_MLP + 14) return mlp(index - NWS_T_NOISE_MLP, kBands);
  switch (index) {
    case NWS_T_GRU_W_IH: return (size_t)kGates * 2;
    case NWS_T_GRU_W_HH: return (size_t)kGates * kEmb;
    case NWS_T_GRU_B_IH: case NWS_T_GRU_B_HH: return kGates;
    case NWS_T_PROJ_W: return (size_t)kEmb * kEmb;
    case NWS_T_PROJ_B: return kEmb;
    case NWS_T_OSC_RAND_PHASE: return kHarm;
    case NWS_T_HMIX_W: return (size_t)kShapers * kHarm;
    case NWS_T_HMIX_B: return kShapers;
    case NWS_T_SHAPER_SCALE: return kShapers;
    case NWS_T_SHAPER_W1: case NWS_T_SHAPER_B1: case NWS_T_SHAPER_B2: case NWS_T_SHAPER_B3: return (size_t)kShapers * 8;
    case NWS_T_SHAPER_W2: case NWS_T_SHAPER_W3: return (size_t)kShapers * 64;
    case NWS_T_SHAPER_W4: return (size_t)kShapers * 8;
    case NWS_T_SHAPER_B4: return kShapers;
    case NWS_T_MIX_W: return kShapers;
    case NWS_T_MIX_B: return 1;
    case NWS_T_NOISE_WINDOW: return kIr;
    case NWS_T_REVERB_IR: return kReverbIr - 1;
  }
  return 0;
}

extern "C" int nws_status(NwsHandle ctx) {
  if (!ctx) { nws_set_error("nws_status: NULL handle"); return NWS_ERR_INVALID; }
  return nws_check_fault(ctx, "nws_status");
}

extern "C" int nws_load_weights(NwsHandle ctx, const float* const* tensors, int n_tensors, void* stream) {
  if (!ctx || !tensors) { nws_set_error("nws_load_weights: NULL argument"); return NWS_ERR_INVALID; }
  if (n_tensors != NWS_T_COUNT) { nws_set_error("nws_load_weights: expected %d tensors, got %d", (int)NWS_T_COUNT, n_tensors); return NWS_ERR_INVALID; }
  for (int i = 0; i < NWS_T_COUNT; ++i)
    if (!tensors[i]) { nws_set_error("nws_load_weights: tensor %d is NULL", i); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  // the noise branch folds the window into the spectrum analytically: it must be the periodic Hann
  float win[kIr];
  NWS_CUDA_OK(cudaMemcpyAsync(win, tensors[NWS_T_NOISE_WINDOW], sizeof(win), cudaMemcpyDeviceToHost, s));
  NWS_CUDA_OK(cudaStreamSynchronize(s));
  for (int n = 0; n < kIr; ++n) {
    const double h = 0.5 - 0.5 * cos(2.0 * M_PI * n / kIr);
    if (fabs((double)win[n] - h) > 1e-6) {
      nws_set_error("nws_load_weights: noise_synth.window is not the periodic Hann window (index %d: %g vs %g)", n, win[n], h);
      return NWS_ERR_UNSUPPORTED;
    }
  }
  {
    // argument bound of the shaper MLP's inner sines (layers 2-4 see sines in [-1,1] as inputs)
    static const int kW[3] = {NWS_T_SHAPER_W2, NWS_T_SHAPER_W3, NWS_T_SHAPER_W4};
    static const int kB[3] = {NWS_T_SHAPER_B2, NWS_T_SHAPER_B3, NWS_T_SHAPER_B4};
    static const int kRows[3] = {kShapers * 8, kShapers * 8, kShapers};
    float* hw = (float*)malloc((size_t)kShapers * 8 * 9 * sizeof(float));
    if (!hw) { nws_set_error("nws_load_weights: out of host memory"); return NWS_ERR_CUDA; }
    float bound = 0.f;
    for (int l = 0; l < 3; ++l) {
      cudaError_t e = cudaMemcpyAsync(hw, tensors[kW[l]], (size_t)kRows[l] * 8 * sizeof(float), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(hw + kRows[l] * 8, tensors[kB[l]], (size_t)kRows[l] * sizeof(float), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { free(hw); nws_set_error("nws_load_weights: %s", cudaGetErrorString(e)); return NWS_ERR_CUDA; }
      for (int r = 0; r < kRows[l]; ++r) {
        float a = fabsf(hw[kRows[l] * 8 + r]);
        for (int i = 0; i < 8; ++i) a += fabsf(hw[r * 8 + i]);
        if (!(a <= bound)) bound = a;   // NaN-propagating max
      }
    }
    free(hw);
    ctx->shaper_inner_bound = bound;
  }
  int rc = nws_launch_pack_weights(ctx, tensors, s);
  if (rc) return rc;
  rc = nws_launch_mlp_tc_pack(ctx, tensors, s);
  if (rc) return rc;
  {
    // the tensor-core recurrence splits W_hh into fp16 pairs: every weight must be finite and inside the fp16 range
    // (trained values are O(1)); otherwise the fp32 kernel is used
    float* hw = (float*)malloc((size_t)kGates * kEmb * sizeof(float));
    if (!hw) { nws_set_error("nws_load_weights: out of host memory"); return NWS_ERR_CUDA; }
    cudaError_t e = cudaMemcpyAsync(hw, tensors[NWS_T_GRU_W_HH], (size_t)kGates * kEmb * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { free(hw); nws_set_error("nws_load_weights: %s", cudaGetErrorString(e)); return NWS_ERR_CUDA; }
    bool ok = true;
    for (int i = 0; i < kGates * kEmb; ++i) ok = ok && fabsf(hw[i]) <= 32768.0f;   // (false for NaN)
    free(hw);
    ctx->gru_mma_ok = ok;
    rc = nws_launch_pack_gru_mma(ctx, tensors[NWS_T_GRU_W_HH], s);
    if (rc) return rc;
  }
  ctx->weights_loaded = true;
  ctx->lut_valid = false;
  nws_reverb_invalidate(ctx);
  return NWS_OK;
}

static int ensure_lut_storage(NwsContext* ctx, int table_size) {
  if (table_size < 2 || table_size > (1 << 20)) { nws_set_error("LUT size %d out of range", table_size); return NWS_ERR_INVALID; }
  if (ctx->lut_size != table_size) {
    cudaFree(ctx->lut);
    cudaFree(ctx->lut2);
    ctx->lut = nullptr;
    ctx->lut2 = nullptr;
    ctx->lut_size = 0;
    NWS_CUDA_OK(cudaMalloc(&ctx->lut, (size_t)kShapers * table_size * sizeof(float)));
    NWS_CUDA_OK(cudaMalloc(&ctx->lut2, (size_t)kShapers * table_size * sizeof(float2)));
    ctx->lut_size = table_size;
  }
  return NWS_OK;
}

extern "C" int nws_build_lut(NwsHandle ctx, int table_size, float tmin, float tmax, const float* sample_points,
                             void* stream) {
  if (!ctx) { nws_set_error("nws_build_lut: NULL handle"); return NWS_ERR_INVALID; }
  if (!ctx->weights_loaded) { nws_set_error("nws_build_lut: weights not loaded"); return NWS_ERR_STATE; }
  if (!(tmax > tmin)) { nws_set_error("nws_build_lut: table_max must exceed table_min"); return NWS_ERR_INVALID; }
  int rc = ensure_lut_storage(ctx, table_size);
  if (rc) return rc;
  rc = nws_launch_build_lut(ctx, sample_points, ctx->lut, table_size, tmin, tmax, (cudaStream_t)stream);
  if (rc) return rc;
  rc = nws_launch_pair_lut(ctx, (cudaStream_t)stream);
  if (rc) return rc;
  ctx->lut_min = tmin; ctx->lut_max = tmax; ctx->lut_valid = true;
  return NWS_OK;
}

extern "C" int nws_set_lut(NwsHandle ctx, const float* lut, int table_size, float tmin, float tmax, void* stream) {
  if (!ctx || !lut) { nws_set_error("nws_set_lut: NULL argument"); return NWS_ERR_INVALID; }
  if (!(tmax > tmin)) { nws_set_error("nws_set_lut: table_max must exceed table_min"); return NWS_ERR_INVALID; }
  int rc = ensure_lut_storage(ctx, table_size);
  if (rc) return rc;
  NWS_CUDA_OK(cudaMemcpyAsync(ctx->lut, lut, (size_t)kShapers * table_size * sizeof(float), cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream));
  rc = nws_launch_pair_lut(ctx, (cudaStream_t)stream);
  if (rc) return rc;
  ctx->lut_min = tmin; ctx->lut_max = tmax; ctx->lut_valid = true;
  return NWS_OK;
}

extern "C" int nws_get_lut(NwsHandle ctx, float* lut_out, void* stream) {
  if (!ctx || !lut_out) { nws_set_error("nws_get_lut: NULL argument"); return NWS_ERR_INVALID; }
  if (!ctx->lut_valid) { nws_set_error("nws_get_lut: no LUT"); return NWS_ERR_STATE; }
  NWS_CUDA_OK(cudaMemcpyAsync(lut_out, ctx->lut, (size_t)kShapers * ctx->lut_size * sizeof(float),
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
NwsWorkspace nws_carve_workspace(void* base, int B, int T, int fft_len) {
  NwsWorkspace w{};
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += (bytes + 255) & ~(size_t)255; return r; };
  const size_t M = (size_t)B * T, N = (size_t)T * kHop;
  w.carry = (double*)take(M * sizeof(double));
  w.u_phase = (float*)take(kHarmPad * sizeof(float));
  w.noise = (float*)take(N * sizeof(float));
  w.hbuf = (float*)take(M * kEmb * sizeof(float));
  w.emb = (float*)take(M * kEmb * sizeof(float));
  w.act0 = (float*)take(M * kEmb * sizeof(float));
  w.act1 = (float*)take(M * kEmb * sizeof(float));
  w.film = (float*)take(M * kFilm * sizeof(float));
  w.bands = (float*)take(M * kBandsPad * sizeof(float));
  w.xspec = (float2*)take((size_t)T * kBandsPad * sizeof(float2));
  w.dry = (float*)take((size_t)B * N * sizeof(float));
  w.scratch = (float*)take(M * kFilm * sizeof(float));
  w.counters = (int*)take(kMaxTimeBlocks * sizeof(int));
  w.h_state = (float*)take((size_t)B * kEmb * sizeof(float));
  w.rev = (float2*)take((size_t)((B + 1) / 2) * fft_len * sizeof(float2));
  w.total = off;
  return w;
}

extern "C" size_t nws_workspace_bytes(NwsHandle ctx, int B, int T) {
  if (!ctx || B < 1 || T < 1) return 0;
  const int L = nws_reverb_fft_len(T * kHop);
  if (!L) return 0;
  return nws_carve_workspace(nullptr, B, T, L).total;
}

extern "C" size_t nws_reverb_workspace_bytes(NwsHandle ctx, int B, int N) {
  if (!ctx || B < 1 || N < 1) return 0;
  const int L = nws_reverb_fft_len(N);
  if (!L) return 0;
  return (size_t)((B + 1) / 2) * L * sizeof(float2) + 256;
}

// A tcgen05 kernel of an earlier call gave up waiting on an mbarrier (bounded waits, nws_tc.cuh) and wrote NaNs:
// sticky until the handle is destroyed.  Read from mapped host memory — no synchronise.
int nws_check_fault(const NwsContext* ctx, const char* who) {
  if (ctx->fault_host && *(volatile int*)ctx->fault_host) {
    nws_set_error("%s: a tensor-core kernel of this handle timed out on an mbarrier (results since then are invalid)", who);
    return NWS_ERR_CUDA;
  }
  return NWS_OK;
}

static int check_common(NwsContext* ctx, int B, int T, void* ws, size_t ws_bytes, const char* who, NwsWorkspace* out) {
  if (!ctx) { nws_set_error("%s: NULL handle", who); return NWS_ERR_INVALID; }
  if (int rc = nws_check_fault(ctx, who)) return rc;
  if (!ctx->weights_loaded) { nws_set_error("%s: weights not loaded", who); return NWS_ERR_STATE; }
  if (B < 1) { nws_set_error("%s: batch size must be >= 1 (got %d)", who, B); return NWS_ERR_INVALID; }
  if (T < 2) { nws_set_error("%s: T must be >= 2 control frames (got %d); the reference's STFT reflect padding has the same limit", who, T); return NWS_ERR_INVALID; }
  if ((long long)T * kHop >= (1ll << 23)) { nws_set_error("%s: T = %d too long (phase/upsample index arithmetic is exact below 2^23 samples)", who, T); return NWS_ERR_UNSUPPORTED; }
  const int L = nws_reverb_fft_len(T * kHop);
  if (!L) { nws_set_error("%s: T = %d too long for the reverb FFT plan", who, T); return NWS_ERR_UNSUPPORTED; }
  if (!ws) { nws_set_error("%s: NULL workspace", who); return NWS_ERR_INVALID; }
  if (((uintptr_t)ws & 255) != 0) { nws_set_error("%s: workspace must be 256-byte aligned", who); return NWS_ERR_INVALID; }
  *out = nws_carve_workspace(ws, B, T, L);
  if (out->total > ws_bytes) { nws_set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, out->total); return NWS_ERR_WORKSPACE; }
  return NWS_OK;
}

#define NWS_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

// stage timing: events bracket each stage on the launch stream when profiling is on
#define NWS_STAGE(ctx, st, s, expr)                                           \
  do {                                                                        \
    if ((ctx)->profile) cudaEventRecord((ctx)->ev[2 * (st)], (s));           \
    NWS_TRY(expr);                                                            \
    if ((ctx)->profile) { cudaEventRecord((ctx)->ev[2 * (st) + 1], (s)); (ctx)->ev_recorded[(st)] = true; } \
  } while (0)

static int launch_audio(const NwsContext* ctx, const float* f0, const double* carry, const float* film,
                        const float* u_phase, const float* noise_in, float* out, float* exciter_out, int B, int T,
                        int* tile_counter, int use_lut, cudaStream_t s) {
  return ctx->audio_impl ? nws_launch_audio_tc(ctx, f0, carry, film, u_phase, noise_in, out, exciter_out, B, T, 0, T,
                                               tile_counter, use_lut, s)
                         : nws_launch_audio(ctx, f0, carry, film, u_phase, noise_in, out, exciter_out, B, T, use_lut, s);
}

extern "C" int nws_set_mlp_impl(NwsHandle ctx, int impl) {
  if (!ctx || (impl != 0 && impl != 1)) { nws_set_error("nws_set_mlp_impl: impl must be 0 (fp32 SIMT layers) or 1 (tcgen05 chain)"); return NWS_ERR_INVALID; }
  ctx->mlp_impl = impl;
  return NWS_OK;
}

// control -> (FiLM parameters, noise band gains): get_embedding + ControlModule + both TimeDistributedMLPs
// (neural_waveshaping.py:69-72,78; shaping.py:68; neural_waveshaping.py:82), reference layouts.
extern "C" int nws_stage_control_to_params(NwsHandle ctx, const float* control, int ctrl_channels, float* film_out,
                                           float* bands_out, int B, int T, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  NwsWorkspace w;
  NWS_TRY(check_common(ctx, B, T, workspace, workspace_bytes, "nws_stage_control_to_params", &w));
  if (!control || !film_out || !bands_out || ctrl_channels < 2) { nws_set_error("nws_stage_control_to_params: bad argument"); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const int M = B * T;
  NWS_TRY(nws_launch_gru(ctx, control, ctrl_channels, w.hbuf, B, T, 0, T, nullptr, s));
  if (ctx->mlp_impl && nws_mlp_small_ok(ctx, B, T)) {
    NWS_TRY(nws_launch_mlp_small(ctx, w.hbuf, w.film, w.bands, B, T, s));
  } else if (ctx->mlp_impl) {
    NWS_TRY(nws_launch_mlp_tc(ctx, w.hbuf, w.film, w.bands, M, T, 0, T, s));
  } else {
    NWS_TRY(nws_launch_linear(w.hbuf, ctx->packed + ctx->lay.proj_wt, ctx->packed + ctx->lay.proj_b, nullptr, nullptr,
                              w.emb, M, kEmb, kEmb, kEmb, false, s));
    NWS_TRY(nws_launch_td_mlp(ctx, NWS_MLP_FILM, w.emb, w.act0, w.act1, w.film, M, s));
    NWS_TRY(nws_launch_td_mlp(ctx, NWS_MLP_NOISE, w.emb, w.act0, w.act1, w.bands, M, s));
  }
  NWS_TRY(nws_launch_rows_to_bct(w.film, film_out, B, kFilm, T, kFilm, s));
  return nws_launch_rows_to_bct(w.bands, bands_out, B, kBands, T, kBandsPad, s);
}

extern "C" int nws_set_audio_impl(NwsHandle ctx, int impl) {
  if (!ctx || (impl != 0 && impl != 1)) { nws_set_error("nws_set_audio_impl: impl must be 0 (fp32 SIMT mixer) or 1 (tcgen05 mixer)"); return NWS_ERR_INVALID; }
  ctx->audio_impl = impl;
  return NWS_OK;
}

extern "C" int nws_set_small_path(NwsHandle ctx, int enable) {
  if (!ctx) { nws_set_error("nws_set_small_path: NULL handle"); return NWS_ERR_INVALID; }
  ctx->small_path = enable != 0;
  return NWS_OK;
}

extern "C" int nws_set_reverb_direct(NwsHandle ctx, int enable) {
  if (!ctx) { nws_set_error("nws_set_reverb_direct: NULL handle"); return NWS_ERR_INVALID; }
  ctx->reverb_direct = enable != 0;
  return NWS_OK;
}

extern "C" int nws_set_gru_impl(NwsHandle ctx, int impl) {
  if (!ctx || impl < 0 || impl > 2) { nws_set_error("nws_set_gru_impl: impl must be 0 (fp32 SIMT recurrence), 1 (tensor cores from 64 utterances on) or 2 (tensor cores always)"); return NWS_ERR_INVALID; }
  ctx->gru_impl = impl;
  return NWS_OK;
}

extern "C" int nws_set_noise_fused(NwsHandle ctx, int enable) {
  if (!ctx) { nws_set_error("nws_set_noise_fused: NULL handle"); return NWS_ERR_INVALID; }
  ctx->noise_fused = enable != 0;
  return NWS_OK;
}

extern "C" int nws_set_shaper_impl(NwsHandle ctx, int impl) {
  if (!ctx || (impl != 0 && impl != 1)) { nws_set_error("nws_set_shaper_impl: impl must be 0 (fp32 FMA layers) or 1 (tensor-core 8x8 layers)"); return NWS_ERR_INVALID; }
  ctx->shaper_impl = impl;
  return NWS_OK;
}

extern "C" int nws_set_profiling(NwsHandle ctx, int enable) {
  if (!ctx) { nws_set_error("nws_set_profiling: NULL handle"); return NWS_ERR_INVALID; }
  if (enable && !ctx->ev[0]) {
    for (int i = 0; i < 2 * kStCount; ++i) NWS_CUDA_OK(cudaEventCreate(&ctx->ev[i]));
  }
  ctx->profile = enable != 0;
  for (int i = 0; i < kStCount; ++i) ctx->ev_recorded[i] = false;
  return NWS_OK;
}

extern "C" int nws_get_stage_times(NwsHandle ctx, float* ms_out, int n) {
  if (!ctx || !ms_out || n < kStCount) { nws_set_error("nws_get_stage_times: need room for %d floats", (int)kStCount); return NWS_ERR_INVALID; }
  for (int i = 0; i < kStCount; ++i) {
    ms_out[i] = 0.f;
    if (!ctx->ev_recorded[i]) continue;
    NWS_CUDA_OK(cudaEventSynchronize(ctx->ev[2 * i + 1]));
    NWS_CUDA_OK(cudaEventElapsedTime(&ms_out[i], ctx->ev[2 * i], ctx->ev[2 * i + 1]));
  }
  return NWS_OK;
}

// Development aid: NWS_TIMELINE=1 prints when each kernel of the pipelined forward finished, relative to
// the fork point (timed events on the three streams; synchronises — never enabled in measurements).
struct NwsTimeline {
  bool on = false;
  int n = 0;
  cudaEvent_t ev[128];
  const char* name[128];
  void mark(const char* nm, cudaStream_t st) {
    if (!on || n >= 128) return;
    if (!ev[n]) cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = nm;
  }
  void dump() {
    if (!on || n < 2) return;
    cudaDeviceSynchronize();
    for (int i = 1; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[0], ev[i]);
      fprintf(stderr, "[nws timeline] %-18s done at %8.3f ms\n", name[i], ms);
    }
    n = 0;
  }
};
static NwsTimeline g_tl;

extern "C" int nws_forward(NwsHandle ctx, const float* f0, const float* control, int ctrl_channels,
                           const float* u_phase, const float* noise, uint64_t seed, uint64_t offset, float* out,
                           int B, int T, int use_lut, void* workspace, size_t workspace_bytes, void* stream) {
  NwsWorkspace w;
  NWS_TRY(check_common(ctx, B, T, workspace, workspace_bytes, "nws_forward", &w));
  if (!f0 || !control || !out) { nws_set_error("nws_forward: NULL tensor"); return NWS_ERR_INVALID; }
  if (ctrl_channels < 2) { nws_set_error("nws_forward: control needs >= 2 channels (got %d)", ctrl_channels); return NWS_ERR_INVALID; }
  if (use_lut && !ctx->lut_valid) { nws_set_error("nws_forward: FastNEWT requested but no lookup table is loaded"); return NWS_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int M = B * T, N = T * kHop;

  const int n_blocks = (T + 127) / 128;
  const int gru_ctas = nws_gru_ctas(ctx, B);   // SMs the recurrence occupies
  const bool pipelined = ctx->pipeline && ctx->mlp_impl && ctx->audio_impl && !ctx->profile && ctx->enc_stream &&
                         ctx->aux_stream && n_blocks >= 3 && (long long)B * T >= 4096 && gru_ctas + 16 <= ctx->sm_count;
  // Time blocks of the pipelined order.  fp32 recurrence (B SMs for ~0.7 us per step): a 128-frame head, then the rest.
  // Tensor-core recurrence (B/8 SMs, ~1.15 us per step): equal blocks of `pipe_block` frames — the rest of the chip
  // renders block k while the encoder's SMs produce block k+1.
  int tb[kMaxTimeBlocks + 1] = {0}, nb = 0;
  const bool gru_mma = gru_ctas != B;
  bool gru_marks = false;   // one encoder launch with progress marks (below)
  if (pipelined) {
    if (!gru_mma) { tb[1] = 128; tb[2] = T; nb = 2; }
    else {
      static const int blk_env = getenv("NWS_PIPE_BLOCK") ? atoi(getenv("NWS_PIPE_BLOCK")) : 0;   // development switches
      static const int first_env = getenv("NWS_PIPE_FIRST") ? atoi(getenv("NWS_PIPE_FIRST")) : 0;
      const int blk = blk_env >= 16 ? blk_env : ctx->pipe_block, first = first_env >= 2 ? first_env : ctx->pipe_first;
      nb = 1 + (T - first + blk / 2) / blk;   // a short remainder joins the last block
      if (nb < 2) nb = 2;
      if (nb > kMaxTimeBlocks) nb = kMaxTimeBlocks;
      for (int k = 1; k < nb; ++k) tb[k] = first + (k - 1) * blk;
      tb[nb] = T;
    }
    // The GRU is T dependent steps on a few SMs; everything downstream only needs the frames already encoded.  The
    // recurrence runs on an internal high-priority stream; the blocks already encoded go through the MLP chain and the
    // noise branch and their audio hops are rendered on the SMs the GRU does not occupy (persistent CTAs, capped; tiles
    // claimed dynamically).  fp32 recurrence: one launch per block (hidden state carried in `h_state`), stream events.
    // Tensor-core recurrence: ONE launch that never gives up its SMs — a relaunch per block lost them to the waiting
    // MLP CTAs for 30 us each time — and counts its CTAs past each block boundary in `gru_done`; the consumers' streams
    // wait for the count with a one-thread kernel.
    // The encoder only reads `control`, so it is forked first: the draws, phase carries and noise spectrum
    // below run beside its first steps instead of ahead of them.
    cudaStream_t g = ctx->enc_stream;
    g_tl.on = getenv("NWS_TIMELINE") != nullptr;
    g_tl.mark("fork", s);
    // (inside a stream capture the per-block launches and events are kept: a graph does not promise that the waiting
    // kernel and the encoder run concurrently)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NWS_CUDA_OK(cudaStreamIsCapturing(s, &cap));
    gru_marks = gru_mma && cap == cudaStreamCaptureStatusNone;
    if (gru_marks) NWS_CUDA_OK(cudaMemsetAsync(ctx->gru_done, 0, kMaxTimeBlocks * sizeof(int), s));
    NWS_CUDA_OK(cudaEventRecord(ctx->ev_fork, s));
    NWS_CUDA_OK(cudaStreamWaitEvent(g, ctx->ev_fork, 0));
    if (gru_marks) {
      NwsGruMarks marks;
      marks.n = nb;
      for (int k = 0; k < nb; ++k) marks.t[k] = tb[k + 1];
      NWS_TRY(nws_launch_gru_mma(ctx, control, ctrl_channels, w.hbuf, B, T, 0, T, nullptr, g, ctx->gru_done, &marks));
      NWS_CUDA_OK(cudaEventRecord(ctx->ev_block[0], g));   // (joins the encoder stream back into the caller's at the end)
      g_tl.mark("gru", g);
    } else {
      for (int k = 0; k < nb; ++k) {
        NWS_TRY(nws_launch_gru(ctx, control, ctrl_channels, w.hbuf, B, T, tb[k], tb[k + 1], w.h_state, g));
        NWS_CUDA_OK(cudaEventRecord(ctx->ev_block[k], g));
        g_tl.mark("gru block", g);
      }
    }
  }

  // Short buffers (the sweep of scripts/time_buffer_sizes.py: 2..32 frames): latency is launches, not work.  One launch
  // does everything that precedes the MLP chain — the draws, the phase carries, the noise spectrum and the recurrence
  // run side by side as CTA roles — then the cluster MLP chain (which also filters the noise), the audio kernel and the
  // direct-form reverb: 4 launches instead of 12.
  const bool small = !pipelined && ctx->small_path && ctx->mlp_impl && ctx->audio_impl &&
                     nws_front_ok(B, T) && nws_mlp_small_ok(ctx, B, T);
  if (small) {
    static const bool no_pdl = getenv("NWS_NO_PDL") != nullptr;   // development switch: ordinary launches
    const bool pdl = !ctx->profile && !no_pdl;
    NWS_STAGE(ctx, kStGru, s, nws_launch_front(ctx, control, ctrl_channels, w.hbuf, f0, w.carry, noise, u_phase ? nullptr : w.u_phase,
                                               seed, offset, w.xspec, B, T, s));
    if (!u_phase) u_phase = w.u_phase;
    // (the noise chain's cluster also filters the noise: the band gains never leave shared memory)
    NWS_STAGE(ctx, kStMlpFilm, s, nws_launch_mlp_small(ctx, w.hbuf, w.film, nullptr, B, T, s, w.xspec, w.dry, 0, T, pdl));
    NWS_STAGE(ctx, kStAudio, s, nws_launch_audio_tc(ctx, f0, w.carry, w.film, u_phase, w.dry, w.dry, nullptr, B, T, 0, T,
                                                    ctx->tile_counters, use_lut, s, 0, pdl));
    const size_t small_rev_bytes = (size_t)((B + 1) / 2) * nws_reverb_fft_len(N) * sizeof(float2);
    if (ctx->reverb_direct && nws_reverb_direct_ok(ctx, B, N, small_rev_bytes)) {
      NWS_STAGE(ctx, kStReverb, s, nws_launch_reverb_direct(ctx, w.dry, out, (float*)w.rev, B, N, s, pdl));
    } else {
      NWS_STAGE(ctx, kStReverb, s, nws_launch_reverb(ctx, w.dry, out, w.rev, B, N, s));
    }
    return NWS_OK;
  }

  if (!u_phase || !noise) {  // the forward's own draws (generators.py:55, :30); injected ones are kept
    NWS_STAGE(ctx, kStRng, s, nws_launch_rng(u_phase ? nullptr : w.u_phase, noise ? nullptr : w.noise, N - 1, seed, offset, s));
    if (!u_phase) u_phase = w.u_phase;
    if (!noise) noise = w.noise;
  }
  // hop rate: phase carries, control encoder, FiLM parameters, noise band gains
  NWS_STAGE(ctx, kStCarry, s, nws_launch_phase_carry(f0, w.carry, B, T, s));
  NWS_STAGE(ctx, kStNoiseSpec, s, nws_launch_noise_spectrum(ctx, noise, w.xspec, T, s));
  g_tl.mark("draws+carry+spec", s);

  if (pipelined) {
    // block k's chain (MLP -> noise hops -> audio hops) alternates between the caller's stream and the auxiliary one,
    // the last block on the caller's: consecutive audio launches overlap at their tails (separate tile counters)
    cudaStream_t aux = ctx->aux_stream;
    NWS_CUDA_OK(cudaEventRecord(ctx->ev_early_ready, s));
    NWS_CUDA_OK(cudaStreamWaitEvent(aux, ctx->ev_early_ready, 0));
    bool aux_used = false;
    for (int k = 0; k < nb; ++k) {
      const bool last = k == nb - 1, on_aux = ((nb - 1 - k) & 1) != 0;
      cudaStream_t c = on_aux ? aux : s;
      if (gru_marks) NWS_TRY(nws_launch_wait_counter(ctx->gru_done + k, gru_ctas, c));
      else NWS_CUDA_OK(cudaStreamWaitEvent(c, ctx->ev_block[k], 0));
      if (k > 0) NWS_CUDA_OK(cudaStreamWaitEvent(c, ctx->ev_mlp[k - 1], 0));   // frames tb[k]-2, tb[k]-1 of the block before
      NWS_TRY(nws_launch_mlp_tc(ctx, w.hbuf, w.film, w.bands, M, T, tb[k], tb[k + 1], c));
      NWS_CUDA_OK(cudaEventRecord(ctx->ev_mlp[k], c));
      // hop h blends towards frame h + 1: the block's last frame waits for the next block
      const int hb = k == 0 ? 0 : tb[k] - 1, he = last ? T : tb[k + 1] - 1;
      // Grid of the block's audio launch.  Per-block encoder launches: capped to the SMs the encoder does not use, so
      // that its next launch finds them free.  One encoder launch (progress marks): no cap — CTAs that find no free SM
      // wait in the hardware queue and start when the encoder's CTAs leave, claiming whatever tiles remain (a cap
      // decided at launch time strangled large batches: with 128 encoder CTAs every block rendered on 20 SMs long after
      // the encoder had finished).
      const int audio_cap = (last || gru_marks) ? 0 : ctx->sm_count - gru_ctas;
      // the noise branch: its own launch writing into `dry` first (default), or inside the audio kernel
      if (!ctx->noise_fused) NWS_TRY(nws_launch_noise_filter(ctx, w.bands, w.xspec, w.dry, B, T, hb, he, c));
      NWS_TRY(nws_launch_audio_tc(ctx, f0, w.carry, w.film, u_phase, ctx->noise_fused ? nullptr : w.dry, w.dry, nullptr, B, T, hb, he,
                                  ctx->tile_counters + (on_aux ? 2 : 0), use_lut, c, audio_cap, false,
                                  ctx->noise_fused ? w.bands : nullptr, w.xspec));
      g_tl.mark(on_aux ? "block chain (aux)" : "block chain", c);
      if (on_aux) { NWS_CUDA_OK(cudaEventRecord(ctx->ev_early_done, aux)); aux_used = true; }
    }
    if (aux_used) NWS_CUDA_OK(cudaStreamWaitEvent(s, ctx->ev_early_done, 0));
    if (gru_marks) NWS_CUDA_OK(cudaStreamWaitEvent(s, ctx->ev_block[0], 0));
  } else {
    NWS_STAGE(ctx, kStGru, s, nws_launch_gru(ctx, control, ctrl_channels, w.hbuf, B, T, 0, T, nullptr, s));
    if (ctx->mlp_impl && nws_mlp_small_ok(ctx, B, T)) {
      // a handful of frames: fp32 chain, 2 CTAs per utterance, no 128-frame tile to pad
      NWS_STAGE(ctx, kStMlpFilm, s, nws_launch_mlp_small(ctx, w.hbuf, w.film, w.bands, B, T, s));
    } else if (ctx->mlp_impl) {
      // projection + both TimeDistributedMLPs in one tensor-core kernel (activations stay in TMEM)
      NWS_STAGE(ctx, kStMlpFilm, s, nws_launch_mlp_tc(ctx, w.hbuf, w.film, w.bands, M, T, 0, T, s));
    } else {
      NWS_STAGE(ctx, kStProj, s, nws_launch_linear(w.hbuf, ctx->packed + ctx->lay.proj_wt, ctx->packed + ctx->lay.proj_b,
                                                   nullptr, nullptr, w.emb, M, kEmb, kEmb, kEmb, false, s));
      NWS_STAGE(ctx, kStMlpFilm, s, nws_launch_td_mlp(ctx, NWS_MLP_FILM, w.emb, w.act0, w.act1, w.film, M, s));
      NWS_STAGE(ctx, kStMlpNoise, s, nws_launch_td_mlp(ctx, NWS_MLP_NOISE, w.emb, w.act0, w.act1, w.bands, M, s));
    }
    if (ctx->noise_fused && ctx->audio_impl) {
      // fused audio-rate kernel, noise branch included: dry = newt(exciter) + filtered noise
      NWS_STAGE(ctx, kStAudio, s, nws_launch_audio_tc(ctx, f0, w.carry, w.film, u_phase, nullptr, w.dry, nullptr, B, T, 0, T,
                                                      ctx->tile_counters, use_lut, s, 0, false, w.bands, w.xspec));
    } else {
      // noise branch -> dry, then the fused audio-rate kernel: dry = newt(exciter) + noise
      NWS_STAGE(ctx, kStNoiseFilter, s, nws_launch_noise_filter(ctx, w.bands, w.xspec, w.dry, B, T, 0, T, s));
      NWS_STAGE(ctx, kStAudio, s, launch_audio(ctx, f0, w.carry, w.film, u_phase, w.dry, w.dry, nullptr, B, T, ctx->tile_counters, use_lut, s));
    }
  }
  // reverb: direct form for short buffers (one launch instead of three 32000-point passes), FFT otherwise
  const size_t rev_bytes = (size_t)((B + 1) / 2) * nws_reverb_fft_len(N) * sizeof(float2);
  if (ctx->reverb_direct && nws_reverb_direct_ok(ctx, B, N, rev_bytes)) {
    NWS_STAGE(ctx, kStReverb, s, nws_launch_reverb_direct(ctx, w.dry, out, (float*)w.rev, B, N, s));
  } else {
    NWS_STAGE(ctx, kStReverb, s, nws_launch_reverb(ctx, w.dry, out, w.rev, B, N, s));
  }
  if (g_tl.on) { g_tl.mark("reverb", s); g_tl.dump(); }
  return NWS_OK;
}

extern "C" int nws_set_pipeline(NwsHandle ctx, int enable) {
  if (!ctx) { nws_set_error("nws_set_pipeline: NULL handle"); return NWS_ERR_INVALID; }
  ctx->pipeline = enable != 0;
  return NWS_OK;
}

extern "C" int nws_forward_host(NwsHandle ctx, const float* f0_host, const float* control_host, int ctrl_channels,
                                const float* u_phase_host, const float* noise_host, uint64_t seed, uint64_t offset,
                                float* out_host, int B, int T, int use_lut, void* workspace, size_t workspace_bytes,
                                void* stream) {
  NwsWorkspace w;
  NWS_TRY(check_common(ctx, B, T, workspace, workspace_bytes, "nws_forward_host", &w));
  if (!f0_host || !control_host || !out_host) { nws_set_error("nws_forward_host: NULL buffer"); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t M = (size_t)B * T, N = (size_t)T * kHop;
  // input staging: the layout-conversion scratch, which nws_forward itself never touches
  float* d_f0 = w.scratch;
  float* d_ctl = w.scratch + M;
  if ((size_t)(1 + ctrl_channels) * M > M * kFilm) { nws_set_error("nws_forward_host: too many control channels"); return NWS_ERR_INVALID; }
  NWS_CUDA_OK(cudaMemcpyAsync(d_f0, f0_host, M * sizeof(float), cudaMemcpyHostToDevice, s));
  NWS_CUDA_OK(cudaMemcpyAsync(d_ctl, control_host, M * ctrl_channels * sizeof(float), cudaMemcpyHostToDevice, s));
  const float* d_u = nullptr;
  const float* d_n = nullptr;
  if (u_phase_host) {
    NWS_CUDA_OK(cudaMemcpyAsync(w.u_phase, u_phase_host, kHarm * sizeof(float), cudaMemcpyHostToDevice, s));
    d_u = w.u_phase;
  }
  if (noise_host) {
    NWS_CUDA_OK(cudaMemcpyAsync(w.noise, noise_host, (N - 1) * sizeof(float), cudaMemcpyHostToDevice, s));
    d_n = w.noise;
  }
  // device-side result: the GRU state buffer is dead once the embedding is computed (long before the
  // reverb writes its output) and has exactly B*N floats
  float* d_out = w.hbuf;
  NWS_TRY(nws_forward(ctx, d_f0, d_ctl, ctrl_channels, d_u, d_n, seed, offset, d_out, B, T, use_lut, workspace,
                      workspace_bytes, stream));
  NWS_CUDA_OK(cudaMemcpyAsync(out_host, d_out, (size_t)B * N * sizeof(float), cudaMemcpyDeviceToHost, s));
  NWS_CUDA_OK(cudaStreamSynchronize(s));
  return nws_check_fault(ctx, "nws_forward_host");   // synchronised: a timeout of this very call is visible
}

// ------------------------------------------------------------------------------------------------ stages
extern "C" int nws_stage_control_embedding(NwsHandle ctx, const float* control, int ctrl_channels, float* emb, int B,
                                           int T, void* workspace, size_t workspace_bytes, void* stream) {
  NwsWorkspace w;
  NWS_TRY(check_common(ctx, B, T, workspace, workspace_bytes, "nws_stage_control_embedding", &w));
  if (!control || !emb || ctrl_channels < 2) { nws_set_error("nws_stage_control_embedding: bad argument"); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  NWS_TRY(nws_launch_gru(ctx, control, ctrl_channels, w.hbuf, B, T, 0, T, nullptr, s));
  NWS_TRY(nws_launch_linear(w.hbuf, ctx->packed + ctx->lay.proj_wt, ctx->packed + ctx->lay.proj_b, nullptr, nullptr,
                            w.emb, B * T, kEmb, kEmb, kEmb, false, s));
  return nws_launch_rows_to_bct(w.emb, emb, B, kEmb, T, kEmb, s);
}

extern "C" int nws_stage_td_mlp(NwsHandle ctx, int which, const float* emb, float* out, int B, int T, void* workspace,
                                size_t workspace_bytes, void* stream) {
  NwsWorkspace w;
  NWS_TRY(check_common(ctx, B, T, workspace, workspace_bytes, "nws_stage_td_mlp", &w));
  if (!emb || !out || (which != NWS_MLP_FILM && which != NWS_MLP_NOISE)) { nws_set_error("nws_stage_td_mlp: bad argument"); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  NWS_TRY(nws_launch_bct_to_rows(emb, w.emb, B, kEmb, T, kEmb, s));
  float* dst = which == NWS_MLP_FILM ? w.film : w.bands;
  NWS_TRY(nws_launch_td_mlp(ctx, which, w.emb, w.act0, w.act1, dst, B * T, s));
  const int C = which == NWS_MLP_FILM ? kFilm : kBands, ld = which == NWS_MLP_FILM ? kFilm : kBandsPad;
  return nws_launch_rows_to_bct(dst, out, B, C, T, ld, s);
}

extern "C" int nws_stage_audio(NwsHandle ctx, const float* f0, const float* film, const float* u_phase, float* newt_out,
                               float* exciter_out, int B, int T, int use_lut, void* workspace, size_t workspace_bytes,
                               void* stream) {
  NwsWorkspace w;
  NWS_TRY(check_common(ctx, B, T, workspace, workspace_bytes, "nws_stage_audio", &w));
  if (!f0 || !film || !u_phase || !newt_out) { nws_set_error("nws_stage_audio: NULL tensor"); return NWS_ERR_INVALID; }
  if (use_lut && !ctx->lut_valid) { nws_set_error("nws_stage_audio: no lookup table"); return NWS_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  NWS_TRY(nws_launch_bct_to_rows(film, w.film, B, kFilm, T, kFilm, s));
  NWS_TRY(nws_launch_phase_carry(f0, w.carry, B, T, s));
  return launch_audio(ctx, f0, w.carry, w.film, u_phase, nullptr, newt_out, exciter_out, B, T, ctx->tile_counters, use_lut, s);
}

extern "C" int nws_stage_noise(NwsHandle ctx, const float* H, const float* noise, float* out, int B, int T,
                               void* workspace, size_t workspace_bytes, void* stream) {
  NwsWorkspace w;
  NWS_TRY(check_common(ctx, B, T, workspace, workspace_bytes, "nws_stage_noise", &w));
  if (!H || !noise || !out) { nws_set_error("nws_stage_noise: NULL tensor"); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  NWS_TRY(nws_launch_bct_to_rows(H, w.bands, B, kBands, T, kBandsPad, s));
  NWS_TRY(nws_launch_noise_spectrum(ctx, noise, w.xspec, T, s));
  return nws_launch_noise_filter(ctx, w.bands, w.xspec, out, B, T, 0, T, s);
}

extern "C" int nws_stage_reverb(NwsHandle ctx, const float* x, float* out, int B, int N, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!ctx || !x || !out || !workspace) { nws_set_error("nws_stage_reverb: NULL argument"); return NWS_ERR_INVALID; }
  if (!ctx->weights_loaded) { nws_set_error("nws_stage_reverb: weights not loaded"); return NWS_ERR_STATE; }
  if (B < 1 || N < 1) { nws_set_error("nws_stage_reverb: bad shape"); return NWS_ERR_INVALID; }
  const size_t need = nws_reverb_workspace_bytes(ctx, B, N);
  if (!need) { nws_set_error("nws_stage_reverb: N = %d too long", N); return NWS_ERR_UNSUPPORTED; }
  if (workspace_bytes < need) { nws_set_error("nws_stage_reverb: workspace too small (%zu < %zu)", workspace_bytes, need); return NWS_ERR_WORKSPACE; }
  float2* work = (float2*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  if (ctx->reverb_direct && nws_reverb_direct_ok(ctx, B, N, need - 256))
    return nws_launch_reverb_direct(ctx, x, out, (float*)work, B, N, (cudaStream_t)stream);
  return nws_launch_reverb(ctx, x, out, work, B, N, (cudaStream_t)stream);
}
