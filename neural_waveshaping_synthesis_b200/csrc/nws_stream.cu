// Stateful streaming synthesis: the forward pass of
// NeuralWaveshaping.forward (models/neural_waveshaping.py:74-90) fed a few control frames at a time, with
// everything that crosses a buffer boundary carried in device memory between calls:
//
//   GRU hidden state                      ControlModule is recurrent (neural_waveshaping.py:21,25)
//   fp64 running sum of the upsampled f0  the oscillator phase is a cumsum over the utterance (generators.py:59)
//   the last three control frames         x128 linear upsampling blends neighbouring frames (neural_waveshaping.py:75,
//                                         shaping.py:69) and the noise branch overlap-adds neighbouring STFT frames
//                                         (generators.py:31-35)
//   the last 32000 dry samples            the reverb's 31999-tap IR (shaping.py:161-173)
//
// The reference has no streaming mode (scripts/time_buffer_sizes.py:66-72 runs independent forwards), so this
// is an extension with its own oracle: the concatenated output of the pushes equals the dry signal of ONE
// whole-utterance forward over the concatenated control frames, followed by the reverb as a causal (linear)
// convolution — the reference's circular wrap of the reverb tail (shaping.py:170-173) cannot exist in a stream.
//
// Every push re-runs the proven whole-utterance kernels on a short window [<= 3 history frames | new frames]
// and renders only the hops whose neighbours are all real: hop h needs frame h+1 (second half of the hop
// interpolates towards it), so the output lags the input by one hop until `flush`.
#include "nws_hop_bodies.cuh"
#include "nws_internal.cuh"
#include "nws_noise_bodies.cuh"

namespace {
constexpr int kHist = 3;   // history frames kept: hop r0 >= 2 never sees the window's artificial left edge
}

struct NwsStreamState {
  NwsContext* ctx;
  int B, max_frames, Tw_max;
  long long n_seen;       // control frames consumed so far
  long long n_rendered;   // hops emitted so far
  int n_hist;             // min(kHist, n_seen)
  bool flushed;
  uint64_t seed, offset;  // Philox stream of the noise draw (generators.py:30) when none is injected
  // device state
  float* f0_h;        // [B][kHist]            } the current one of two copies each (hist_buf): the fused front-end launch
  float* ctrl_h;      // [B][2][kHist]         } reads one while it writes the other
  float* hrow_h;      // [B][kHist][128] GRU outputs of the history frames
  float* hist_buf[2][3];
  int hist_cur;
  float* h_state;     // [B][128]
  double* phase_sum;  // [B]
  float* u_phase;     // [kHarmPad]
  float* rev_hist;    // [B][kReverbIr] last dry samples (the current one of the two ping-pong buffers below)
  float* rev_hist_buf[2];
  float* dir_scratch; // partial sums of the direct-form reverb (streams of short pushes), or null
  int dir_n_new_max;
  // per-push window
  float* f0_w;        // [B][Tw]
  float* ctrl_w;      // [B][2][Tw]
  float* noise_w;     // [128*Tw]
  void* ws_base;      // nws_carve_workspace(B, Tw_max)
  size_t ws_bytes;
  float* xw;          // [B][kReverbIr + 128*max_out]
  float* yw;          // same
  float2* rev_work;
  size_t rev_work_bytes;
};

// ------------------------------------------------------------------------------------------------ kernels
// window = [history | new frames]; also restores the GRU outputs of the history frames (frame-major rows)
__global__ void nws_stream_stage_in_kernel(const float* __restrict__ f0, const float* __restrict__ control, int C, int Tn,
                                           const float* __restrict__ f0_h, const float* __restrict__ ctrl_h,
                                           const float* __restrict__ hrow_h, int n_hist, float* __restrict__ f0_w,
                                           float* __restrict__ ctrl_w, float* __restrict__ hbuf, int Tw) {
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < Tw; t += blockDim.x) {
    float f, c0, c1;
    if (t < n_hist) {
      f = f0_h[b * kHist + t];
      c0 = ctrl_h[(b * 2 + 0) * kHist + t];
      c1 = ctrl_h[(b * 2 + 1) * kHist + t];
    } else {
      const int u = t - n_hist;
      f = f0[(size_t)b * Tn + u];
      c0 = control[((size_t)b * C + 0) * Tn + u];
      c1 = control[((size_t)b * C + 1) * Tn + u];
    }
    f0_w[(size_t)b * Tw + t] = f;
    ctrl_w[((size_t)b * 2 + 0) * Tw + t] = c0;
    ctrl_w[((size_t)b * 2 + 1) * Tw + t] = c1;
  }
  for (int i = threadIdx.x; i < n_hist * kEmb; i += blockDim.x)
    hbuf[((size_t)b * Tw) * kEmb + i] = hrow_h[(size_t)b * kHist * kEmb + i];
}

// after the encoder: the window's last frames become the next push's history
__global__ void nws_stream_stage_out_kernel(const float* __restrict__ f0_w, const float* __restrict__ ctrl_w,
                                            const float* __restrict__ hbuf, int Tw, int n_keep,
                                            float* __restrict__ f0_h, float* __restrict__ ctrl_h,
                                            float* __restrict__ hrow_h) {
  const int b = blockIdx.x, first = Tw - n_keep;
  for (int i = threadIdx.x; i < n_keep; i += blockDim.x) {
    f0_h[b * kHist + i] = f0_w[(size_t)b * Tw + first + i];
    ctrl_h[(b * 2 + 0) * kHist + i] = ctrl_w[((size_t)b * 2 + 0) * Tw + first + i];
    ctrl_h[(b * 2 + 1) * kHist + i] = ctrl_w[((size_t)b * 2 + 1) * Tw + first + i];
  }
  for (int i = threadIdx.x; i < n_keep * kEmb; i += blockDim.x)
    hrow_h[(size_t)b * kHist * kEmb + i] = hbuf[((size_t)b * Tw + first) * kEmb + i];
}

// carry[b][t] = phase_sum[b] + sum of the upsampled f0 over hops [r0, t), t in [r0, r1); phase_sum advances to r1.
// Same per-hop summation as nws_phase_carry_kernel (four fp64 chains over the hop's 128 samples).
__global__ void __launch_bounds__(128) nws_stream_carry_kernel(const float* __restrict__ f0_w, double* __restrict__ carry,
                                                               double* __restrict__ phase_sum, int Tw, int r0, int r1) {
  extern __shared__ double hop_sum[];
  const int b = blockIdx.x;
  const float* f = f0_w + (size_t)b * Tw;
  const float inv_hop = (float)Tw / (float)(Tw * kHop);
  for (int t = r0 + threadIdx.x; t < r1; t += blockDim.x) {
    const float fm = f[t > 0 ? t - 1 : 0], fc = f[t], fp = f[t + 1 < Tw ? t + 1 : Tw - 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int r = 0; r < kHop; r += 4) {
      double q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const NwsLerp c = nws_lerp_coords(t * kHop + r + u, Tw, inv_hop);
        const float x0 = c.i0 == t ? fc : (c.i0 < t ? fm : fp);
        const float x1 = c.i1 == t ? fc : (c.i1 < t ? fm : fp);
        q[u] = (double)nws_lerp_apply(c, x0, x1);
      }
      s0 += q[0]; s1 += q[1]; s2 += q[2]; s3 += q[3];
    }
    hop_sum[t - r0] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = phase_sum[b];
    for (int t = r0; t < r1; ++t) {
      carry[(size_t)b * Tw + t] = run;
      run += hop_sum[t - r0];
    }
    phase_sum[b] = run;
  }
}

// ------------------------------------------------------------------------------------------------
// Front end of a short push in ONE launch (the streaming counterpart of nws_front_kernel): CTA roles by block index
//   [0, B)     window assembly, the recurrence over the new frames and the next push's history, utterance b
//   [B, 2B)    the fp64 phase carries of the rendered hops (continuing phase_sum), eight threads per hop
//   [2B, ...)  one pair of noise frames each: the window's noise draw (Philox) and its spectrum
// The history arrays are double-buffered (read `*_h`, write `*_h_next`): the carry CTAs read the old f0 history while
// the window CTAs write the new one.
struct NwsStreamFrontParams {
  const float *w_hh, *w_ih, *b_ih, *b_hh;
  const float *f0, *control; int C, Tn;                       // the push's new frames
  const float *f0_h, *ctrl_h, *hrow_h; int n_hist;            // history in
  float *f0_h_next, *ctrl_h_next, *hrow_h_next; int n_keep;   // history out
  float *f0_w, *ctrl_w, *hbuf; int B, Tw;                     // window out
  float* h_state;
  double *carry, *phase_sum; int r0, r1;
  const float* noise_in; uint64_t seed, offset; const float2* tw_master; float2* xspec;
};

__global__ void __launch_bounds__(kGates, 1) nws_stream_front_kernel(const NwsStreamFrontParams p) {
  const int role = blockIdx.x, tid = threadIdx.x, Tw = p.Tw;
  nws_pdl_launch();   // the MLP chain's weight prologue may start now (it waits for this grid before reading its results)
  auto f0_at = [&](int b, int t) { return t < p.n_hist ? p.f0_h[b * kHist + t] : p.f0[(size_t)b * p.Tn + (t - p.n_hist)]; };
  if (role < p.B) {
    const int b = role;
    for (int t = tid; t < Tw; t += kGates) {
      float c0, c1;
      if (t < p.n_hist) {
        c0 = p.ctrl_h[(b * 2 + 0) * kHist + t];
        c1 = p.ctrl_h[(b * 2 + 1) * kHist + t];
      } else {
        c0 = p.control[((size_t)b * p.C + 0) * p.Tn + (t - p.n_hist)];
        c1 = p.control[((size_t)b * p.C + 1) * p.Tn + (t - p.n_hist)];
      }
      p.f0_w[(size_t)b * Tw + t] = f0_at(b, t);
      p.ctrl_w[((size_t)b * 2 + 0) * Tw + t] = c0;
      p.ctrl_w[((size_t)b * 2 + 1) * Tw + t] = c1;
    }
    for (int i = tid; i < p.n_hist * kEmb; i += kGates) p.hbuf[((size_t)b * Tw) * kEmb + i] = p.hrow_h[(size_t)b * kHist * kEmb + i];
    __syncthreads();   // the window's control rows are read back by this same CTA
    if (p.Tn > 0) nws_gru_body(b, p.w_hh, p.w_ih, p.b_ih, p.b_hh, p.ctrl_w, 2, p.hbuf, Tw, p.n_hist, Tw, p.h_state);
    __syncthreads();
    const int first = Tw - p.n_keep;
    for (int i = tid; i < p.n_keep; i += kGates) {
      p.f0_h_next[b * kHist + i] = p.f0_w[(size_t)b * Tw + first + i];
      p.ctrl_h_next[(b * 2 + 0) * kHist + i] = p.ctrl_w[((size_t)b * 2 + 0) * Tw + first + i];
      p.ctrl_h_next[(b * 2 + 1) * kHist + i] = p.ctrl_w[((size_t)b * 2 + 1) * Tw + first + i];
    }
    for (int i = tid; i < p.n_keep * kEmb; i += kGates)
      p.hrow_h_next[(size_t)b * kHist * kEmb + i] = p.hbuf[((size_t)b * Tw + first) * kEmb + i];
  } else if (role < 2 * p.B) {
    __shared__ double hop_sum[kSmallMlpMaxFrames + 8];
    const int b = role - p.B, t = p.r0 + (tid >> 3), part = tid & 7;
    const float inv_hop = (float)Tw / (float)(Tw * kHop);
    double s = 0.0;
    if (t < p.r1) {
      const float fm = f0_at(b, t > 0 ? t - 1 : 0), fc = f0_at(b, t), fp = f0_at(b, t + 1 < Tw ? t + 1 : Tw - 1);
      for (int r = part * 16; r < part * 16 + 16; ++r) {
        const NwsLerp c = nws_lerp_coords(t * kHop + r, Tw, inv_hop);
        const float x0 = c.i0 == t ? fc : (c.i0 < t ? fm : fp);
        const float x1 = c.i1 == t ? fc : (c.i1 < t ? fm : fp);
        s += (double)nws_lerp_apply(c, x0, x1);
      }
    }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (t < p.r1 && part == 0) hop_sum[t - p.r0] = s;
    __syncthreads();
    if (tid == 0) {
      double run = p.phase_sum[b];
      for (int i = p.r0; i < p.r1; ++i) {
        p.carry[(size_t)b * Tw + i] = run;
        run += hop_sum[i - p.r0];
      }
      p.phase_sum[b] = run;
    }
  } else {
    nws_noise_spectrum_body(role - 2 * p.B, kGates, p.noise_in, kHop * Tw - 1, p.seed, p.offset, p.tw_master, p.xspec, Tw);
  }
}

// reverb input window: [last 32000 dry samples | the hops rendered by this push]
__global__ void nws_stream_assemble_kernel(const float* __restrict__ rev_hist, const float* __restrict__ dry_w, int Nw_dry,
                                           int first_sample, int n_new, float* __restrict__ xw) {
  const int b = blockIdx.y, Nx = kReverbIr + n_new;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nx; i += gridDim.x * blockDim.x)
    xw[(size_t)b * Nx + i] = i < kReverbIr ? rev_hist[(size_t)b * kReverbIr + i]
                                            : dry_w[(size_t)b * Nw_dry + first_sample + (i - kReverbIr)];
}

// out <- the reverberated new samples; the history <- the last 32000 dry samples of the window.
// (xw and rev_hist are distinct buffers, so the shift has no overlap hazard.)
__global__ void nws_stream_finish_kernel(const float* __restrict__ xw, const float* __restrict__ yw, int n_new,
                                         float* __restrict__ out, float* __restrict__ rev_hist) {
  const int b = blockIdx.y, Nx = kReverbIr + n_new;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nx; i += gridDim.x * blockDim.x) {
    if (i < n_new) out[(size_t)b * n_new + i] = yw[(size_t)b * Nx + kReverbIr + i];
    if (i < kReverbIr) rev_hist[(size_t)b * kReverbIr + i] = xw[(size_t)b * Nx + n_new + i];
  }
}

// ------------------------------------------------------------------------------------------------ C ABI
#define NWS_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

extern "C" int nws_stream_destroy(NwsStreamHandle st) {
  if (!st) return NWS_OK;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) cudaFree(st->hist_buf[i][j]);
  cudaFree(st->h_state); cudaFree(st->phase_sum);
  cudaFree(st->u_phase); cudaFree(st->rev_hist_buf[0]); cudaFree(st->rev_hist_buf[1]); cudaFree(st->dir_scratch); cudaFree(st->f0_w); cudaFree(st->ctrl_w); cudaFree(st->noise_w);
  cudaFree(st->ws_base); cudaFree(st->xw); cudaFree(st->yw); cudaFree(st->rev_work);
  delete st;
  return NWS_OK;
}

extern "C" int nws_stream_create(NwsHandle ctx, int B, int max_frames, NwsStreamHandle* out) {
  if (!ctx || !out) { nws_set_error("nws_stream_create: NULL argument"); return NWS_ERR_INVALID; }
  if (B < 1 || max_frames < 2 || max_frames > 4096) { nws_set_error("nws_stream_create: need B >= 1 and 2 <= max_frames <= 4096 (got %d, %d)", B, max_frames); return NWS_ERR_INVALID; }
  if (!ctx->weights_loaded) { nws_set_error("nws_stream_create: weights not loaded"); return NWS_ERR_STATE; }
  NwsStreamState* st = new NwsStreamState();
  st->ctx = ctx; st->B = B; st->max_frames = max_frames; st->Tw_max = max_frames + kHist;
  const size_t Nx_max = (size_t)kReverbIr + (size_t)kHop * (max_frames + 1);   // a flushing push emits max_frames + 1 hops
  const int L = nws_reverb_fft_len((int)Nx_max);
  if (!L) { delete st; nws_set_error("nws_stream_create: max_frames too large for the reverb plan"); return NWS_ERR_UNSUPPORTED; }
  st->ws_bytes = nws_carve_workspace(nullptr, B, st->Tw_max, 256).total;
  st->rev_work_bytes = (size_t)((B + 1) / 2) * L * sizeof(float2);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  for (int i = 0; i < 2; ++i) {
    alloc((void**)&st->hist_buf[i][0], (size_t)B * kHist * sizeof(float));
    alloc((void**)&st->hist_buf[i][1], (size_t)B * 2 * kHist * sizeof(float));
    alloc((void**)&st->hist_buf[i][2], (size_t)B * kHist * kEmb * sizeof(float));
  }
  st->hist_cur = 0;
  st->f0_h = st->hist_buf[0][0]; st->ctrl_h = st->hist_buf[0][1]; st->hrow_h = st->hist_buf[0][2];
  alloc((void**)&st->h_state, (size_t)B * kEmb * sizeof(float));
  alloc((void**)&st->phase_sum, (size_t)B * sizeof(double));
  alloc((void**)&st->u_phase, kHarmPad * sizeof(float));
  alloc((void**)&st->rev_hist_buf[0], (size_t)B * kReverbIr * sizeof(float));
  alloc((void**)&st->rev_hist_buf[1], (size_t)B * kReverbIr * sizeof(float));
  st->rev_hist = st->rev_hist_buf[0];
  // short pushes (up to 33 hops) take the direct-form reverb: one launch instead of an overlap-save FFT over 32000 + n points
  st->dir_n_new_max = kHop * (max_frames + 1) <= kReverbDirectMaxN + kHop ? kHop * (max_frames + 1) : 0;
  if (st->dir_n_new_max) alloc((void**)&st->dir_scratch, nws_reverb_direct_causal_scratch_bytes(B, st->dir_n_new_max));
  alloc((void**)&st->f0_w, (size_t)B * st->Tw_max * sizeof(float));
  alloc((void**)&st->ctrl_w, (size_t)B * 2 * st->Tw_max * sizeof(float));
  alloc((void**)&st->noise_w, (size_t)kHop * st->Tw_max * sizeof(float));
  alloc(&st->ws_base, st->ws_bytes);
  alloc((void**)&st->xw, (size_t)B * Nx_max * sizeof(float));
  alloc((void**)&st->yw, (size_t)B * Nx_max * sizeof(float));
  alloc((void**)&st->rev_work, st->rev_work_bytes);
  if (e != cudaSuccess) { nws_set_error("nws_stream_create: %s", cudaGetErrorString(e)); nws_stream_destroy(st); return NWS_ERR_CUDA; }
  st->flushed = true;   // a stream must be reset before its first push
  *out = st;
  return NWS_OK;
}

extern "C" int nws_stream_reset(NwsStreamHandle st, const float* u_phase, uint64_t seed, uint64_t offset, void* stream) {
  if (!st) { nws_set_error("nws_stream_reset: NULL handle"); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  st->n_seen = 0; st->n_rendered = 0; st->n_hist = 0; st->flushed = false;
  st->seed = seed; st->offset = offset;
  NWS_CUDA_OK(cudaMemsetAsync(st->phase_sum, 0, (size_t)st->B * sizeof(double), s));
  NWS_CUDA_OK(cudaMemsetAsync(st->h_state, 0, (size_t)st->B * kEmb * sizeof(float), s));
  NWS_CUDA_OK(cudaMemsetAsync(st->rev_hist, 0, (size_t)st->B * kReverbIr * sizeof(float), s));
  NWS_CUDA_OK(cudaMemsetAsync(st->u_phase, 0, kHarmPad * sizeof(float), s));
  if (u_phase) {
    NWS_CUDA_OK(cudaMemcpyAsync(st->u_phase, u_phase, kHarm * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    // the phase-shift draw of generators.py:55, once per stream (a shift that changed between pushes would click)
    NWS_TRY(nws_launch_rng(st->u_phase, nullptr, 0, seed, offset, s));
  }
  return NWS_OK;
}

extern "C" int nws_stream_window(NwsStreamHandle st, int n_frames, long long* first_frame, int* window_frames) {
  if (!st || n_frames < 0) { nws_set_error("nws_stream_window: bad argument"); return NWS_ERR_INVALID; }
  if (first_frame) *first_frame = st->n_seen - st->n_hist;
  if (window_frames) *window_frames = st->n_hist + n_frames;
  return NWS_OK;
}

extern "C" int nws_stream_push(NwsStreamHandle st, const float* f0, const float* control, int ctrl_channels, int n_frames,
                               const float* noise_window, int use_lut, int flush, int apply_reverb, float* out,
                               int* n_out_frames, void* stream) {
  if (!st || !out || !n_out_frames) { nws_set_error("nws_stream_push: NULL argument"); return NWS_ERR_INVALID; }
  NwsContext* ctx = st->ctx;
  *n_out_frames = 0;
  if (st->flushed) { nws_set_error("nws_stream_push: the stream was flushed (or never reset): call nws_stream_reset first"); return NWS_ERR_STATE; }
  if (n_frames < 0 || n_frames > st->max_frames) { nws_set_error("nws_stream_push: n_frames = %d outside [0, %d]", n_frames, st->max_frames); return NWS_ERR_INVALID; }
  if (n_frames > 0 && (!f0 || !control || ctrl_channels < 2)) { nws_set_error("nws_stream_push: f0/control missing or control has < 2 channels"); return NWS_ERR_INVALID; }
  if (n_frames == 0 && !flush) return NWS_OK;
  if (use_lut && !ctx->lut_valid) { nws_set_error("nws_stream_push: FastNEWT requested but no lookup table is loaded"); return NWS_ERR_STATE; }
  if (!ctx->mlp_impl || !ctx->audio_impl) { nws_set_error("nws_stream_push: the streaming path uses the tensor-core kernels (nws_set_mlp_impl / nws_set_audio_impl must be 1)"); return NWS_ERR_UNSUPPORTED; }
  const int B = st->B, n_hist = st->n_hist, Tw = n_hist + n_frames;
  if (Tw < 2) { nws_set_error("nws_stream_push: the first push needs at least 2 frames (the reference's T >= 2 limit)"); return NWS_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const long long g_base = st->n_seen - n_hist;
  const int r0 = (int)(st->n_rendered - g_base), r1 = flush ? Tw : Tw - 1;
  const int n_out = r1 - r0;
  const NwsWorkspace w = nws_carve_workspace(st->ws_base, B, Tw, 256);
  const int M = B * Tw, Nw = Tw * kHop;

  const int n_keep = Tw < kHist ? Tw : kHist;
  const float* noise = noise_window;
  // short pushes: window assembly, recurrence, history, phase carries, noise draw and spectrum in ONE launch
  const bool fused = ctx->small_path && nws_front_ok(B, Tw) && nws_mlp_small_ok(ctx, B, Tw);
  if (fused) {
    NwsStreamFrontParams p{};
    const float* wp = ctx->packed;
    p.w_hh = wp + ctx->lay.gru_whh; p.w_ih = wp + ctx->lay.gru_wih; p.b_ih = wp + ctx->lay.gru_bih; p.b_hh = wp + ctx->lay.gru_bhh;
    p.f0 = f0; p.control = control; p.C = ctrl_channels; p.Tn = n_frames;
    p.f0_h = st->f0_h; p.ctrl_h = st->ctrl_h; p.hrow_h = st->hrow_h; p.n_hist = n_hist;
    const int nx = st->hist_cur ^ 1;
    p.f0_h_next = st->hist_buf[nx][0]; p.ctrl_h_next = st->hist_buf[nx][1]; p.hrow_h_next = st->hist_buf[nx][2]; p.n_keep = n_keep;
    p.f0_w = st->f0_w; p.ctrl_w = st->ctrl_w; p.hbuf = w.hbuf; p.B = B; p.Tw = Tw;
    p.h_state = st->h_state;
    p.carry = w.carry; p.phase_sum = st->phase_sum; p.r0 = r0; p.r1 = n_out > 0 ? r1 : r0;
    // noise[i] of the stream is Philox block (offset + i/4): the window starts at sample 128 * g_base
    p.noise_in = noise; p.seed = st->seed; p.offset = st->offset + 32ull * (uint64_t)g_base;
    p.tw_master = ctx->tw_master; p.xspec = w.xspec;
    nws_stream_front_kernel<<<2 * B + (Tw + 1) / 2, kGates, 0, s>>>(p);
    NWS_LAUNCH_CHECK();
    st->hist_cur = nx;
    st->f0_h = st->hist_buf[nx][0]; st->ctrl_h = st->hist_buf[nx][1]; st->hrow_h = st->hist_buf[nx][2];
  } else {
    // window, encoder (new frames only; the recurrence continues from h_state), history for the next push
    nws_stream_stage_in_kernel<<<B, 128, 0, s>>>(f0, control, ctrl_channels, n_frames, st->f0_h, st->ctrl_h, st->hrow_h, n_hist,
                                                 st->f0_w, st->ctrl_w, w.hbuf, Tw);
    NWS_LAUNCH_CHECK();
    if (n_frames > 0) NWS_TRY(nws_launch_gru(ctx, st->ctrl_w, 2, w.hbuf, B, Tw, n_hist, Tw, st->h_state, s));
    nws_stream_stage_out_kernel<<<B, 128, 0, s>>>(st->f0_w, st->ctrl_w, w.hbuf, Tw, n_keep, st->f0_h, st->ctrl_h, st->hrow_h);
    NWS_LAUNCH_CHECK();
  }
  st->n_seen += n_frames;
  st->n_hist = n_keep;
  if (flush) st->flushed = true;
  if (n_out <= 0) return NWS_OK;

  // hop-rate chain over the window, then the rendered hops [r0, r1)
  if (fused) {
    // cluster MLP chain with the noise filter fused (band gains stay in shared memory)
    NWS_TRY(nws_launch_mlp_small(ctx, w.hbuf, w.film, nullptr, B, Tw, s, w.xspec, w.dry, r0, r1, true));
  } else {
    if (!noise) {
      NWS_TRY(nws_launch_rng(nullptr, st->noise_w, Nw - 1, st->seed, st->offset + 32ull * (uint64_t)g_base, s));
      noise = st->noise_w;
    }
    NWS_TRY(nws_launch_noise_spectrum(ctx, noise, w.xspec, Tw, s));
    NWS_TRY(nws_launch_mlp_tc(ctx, w.hbuf, w.film, w.bands, M, Tw, 0, Tw, s));
    NWS_TRY(nws_launch_noise_filter(ctx, w.bands, w.xspec, w.dry, B, Tw, r0, r1, s));
    nws_stream_carry_kernel<<<B, 128, (size_t)n_out * sizeof(double), s>>>(st->f0_w, w.carry, st->phase_sum, Tw, r0, r1);
    NWS_LAUNCH_CHECK();
  }
  NWS_TRY(nws_launch_audio_tc(ctx, st->f0_w, w.carry, w.film, st->u_phase, w.dry, w.dry, nullptr, B, Tw, r0, r1, ctx->tile_counters,
                              use_lut, s, 0, fused));
  st->n_rendered += n_out;

  // reverb as a causal convolution over [32000 past dry samples | new], keep the new part
  const int n_new = n_out * kHop, Nx = kReverbIr + n_new;
  if (ctx->reverb_direct && st->dir_scratch && n_new <= st->dir_n_new_max) {
    // direct form: convolution, dry add and the history shift in ONE launch (nws_reverb_direct.cu)
    float* next = st->rev_hist == st->rev_hist_buf[0] ? st->rev_hist_buf[1] : st->rev_hist_buf[0];
    NWS_TRY(nws_launch_reverb_direct_causal(ctx, st->rev_hist, w.dry, (size_t)Nw, r0 * kHop, out, next, st->dir_scratch, B, n_new,
                                            apply_reverb, s, fused));
    st->rev_hist = next;
    *n_out_frames = n_out;
    return NWS_OK;
  }
  // FFT form: overlap-save through the four-step transform
  dim3 grid((Nx + 255) / 256 < 64 ? (Nx + 255) / 256 : 64, B);
  nws_stream_assemble_kernel<<<grid, 256, 0, s>>>(st->rev_hist, w.dry, Nw, r0 * kHop, n_new, st->xw);
  NWS_LAUNCH_CHECK();
  if (apply_reverb) {
    NWS_TRY(nws_launch_reverb(ctx, st->xw, st->yw, st->rev_work, B, Nx, s));
  }
  nws_stream_finish_kernel<<<grid, 256, 0, s>>>(st->xw, apply_reverb ? st->yw : st->xw, n_new, out, st->rev_hist);
  NWS_LAUNCH_CHECK();
  *n_out_frames = n_out;
  return NWS_OK;
}
