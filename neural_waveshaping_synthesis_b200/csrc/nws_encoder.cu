// Hop-rate part of the NWS forward: fp64 phase carries, the control encoder (GRU + 1x1 conv)
// and the two TimeDistributedMLPs.  Reference: models/neural_waveshaping.py:17-26,69-72,75;
// modules/dynamic.py:11-40; modules/generators.py:59 (cumsum).
//
// Layout: every hop-rate activation is frame-major, X[(b*T + t)][channel] — one frame's channels
// are contiguous so a 128-sample audio tile later reads its three frames as three 1 KB rows.
#include "nws_hop_bodies.cuh"
#include "nws_internal.cuh"

constexpr int kCarryThreads = 512;

__global__ void __launch_bounds__(kCarryThreads) nws_phase_carry_kernel(const float* __restrict__ f0,
                                                                        double* __restrict__ carry, int T) {
  nws_phase_carry_body<kCarryThreads>(blockIdx.x, f0, carry, T);
}

int nws_launch_phase_carry(const float* f0, double* carry, int B, int T, cudaStream_t s) {
  nws_phase_carry_kernel<<<B, kCarryThreads, 0, s>>>(f0, carry, T);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

__global__ void __launch_bounds__(kGates, 1)
nws_gru_kernel(const float* __restrict__ w_hh, const float* __restrict__ w_ih, const float* __restrict__ b_ih,
               const float* __restrict__ b_hh, const float* __restrict__ control, int ctrl_channels,
               float* __restrict__ hbuf, int T, int t_begin, int t_end, float* __restrict__ h_state) {
  nws_gru_body(blockIdx.x, w_hh, w_ih, b_ih, b_hh, control, ctrl_channels, hbuf, T, t_begin, t_end, h_state);
}

// Which recurrence encodes B utterances.  The fp32 kernel steps faster (0.70 us against 1.15 us: the legacy HMMA pipe
// is the tensor-core kernel's bound) but takes one SM per utterance; the tensor-core kernel takes one SM per EIGHT.
// Measured on B200 (scripts/dev_pipe.py): equal at 64 utterances x 4 s (0.99 ms per FastNEWT forward either way, NEWT
// 2.93 against 2.97 ms), 13 % faster at 256 (3.04 against 3.50 ms: the fp32 kernel needs two waves there and leaves no
// SMs to pipeline with); below 64 the chip is not full and the shorter step wins.  gru_impl 2 forces it for any B.
constexpr int kGruMmaMinBatch = 64;
static bool nws_gru_uses_mma(const NwsContext* ctx, int B) {
  return ctx->gru_mma_ok && (ctx->gru_impl == 2 || (ctx->gru_impl == 1 && B >= kGruMmaMinBatch));
}

int nws_gru_ctas(const NwsContext* ctx, int B) { return nws_gru_uses_mma(ctx, B) ? nws_gru_mma_ctas(B) : B; }

int nws_launch_gru(const NwsContext* ctx, const float* control, int ctrl_channels, float* hbuf, int B, int T,
                   int t_begin, int t_end, float* h_state, cudaStream_t s) {
  if (nws_gru_uses_mma(ctx, B)) return nws_launch_gru_mma(ctx, control, ctrl_channels, hbuf, B, T, t_begin, t_end, h_state, s);
  const float* p = ctx->packed;
  nws_gru_kernel<<<B, kGates, 0, s>>>(p + ctx->lay.gru_whh, p + ctx->lay.gru_wih, p + ctx->lay.gru_bih,
                                      p + ctx->lay.gru_bhh, control, ctrl_channels, hbuf, T, t_begin, t_end, h_state);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
// Y[M x n_out] = X[M x 128] . Wt[128 x ldw] + bias, optionally followed by LayerNorm(128, eps 1e-5)
// and LeakyReLU(0.01) — one Conv1d(k=1) [+ TimeDistributedLayerNorm + LeakyReLU] of
// modules/dynamic.py:28-37 on frame-major activations.  CTA tile 64 frames x 128 outputs, 256
// threads, each 4 frames x 8 outputs; fp32 FMA (3xTF32 tensor-core version: see DESIGN.md roadmap).
constexpr int kLinFrames = 64;
constexpr int kLinXs = kEmb + 4;  // padded row of the X tile

template <bool LN_ACT>
__global__ void __launch_bounds__(256, 2)
nws_linear128_kernel(const float* __restrict__ X, const float* __restrict__ Wt, const float* __restrict__ bias,
                     const float* __restrict__ ln_g, const float* __restrict__ ln_b, float* __restrict__ Y, int M,
                     int n_out, int ldw, int ldy) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                    // [128][128]
  float* Xs = smem + kEmb * kEmb;      // [64][132]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kLinFrames, n0 = blockIdx.y * 128;

  for (int i = tid; i < kEmb * 32; i += 256) {  // weights: 128 rows x 32 float4
    const int k = i >> 5, c4 = (i & 31) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + c4 < ldw) v = *reinterpret_cast<const float4*>(Wt + (size_t)k * ldw + n0 + c4);
    *reinterpret_cast<float4*>(Ws + k * kEmb + c4) = v;
  }
  for (int i = tid; i < kLinFrames * 32; i += 256) {
    const int f = i >> 5, c4 = (i & 31) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + f < M) v = *reinterpret_cast<const float4*>(X + (size_t)(m0 + f) * kEmb + c4);
    *reinterpret_cast<float4*>(Xs + f * kLinXs + c4) = v;
  }
  __syncthreads();

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll 2
  for (int k4 = 0; k4 < kEmb; k4 += 4) {
    float4 xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(Xs + (ty * 4 + i) * kLinXs + k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 wa = *reinterpret_cast<const float4*>(Ws + (k4 + kk) * kEmb + tx * 4);
      const float4 wb = *reinterpret_cast<const float4*>(Ws + (k4 + kk) * kEmb + 64 + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = kk == 0 ? xv[i].x : (kk == 1 ? xv[i].y : (kk == 2 ? xv[i].z : xv[i].w));
        acc[i][0] = fmaf(x, wa.x, acc[i][0]); acc[i][1] = fmaf(x, wa.y, acc[i][1]);
        acc[i][2] = fmaf(x, wa.z, acc[i][2]); acc[i][3] = fmaf(x, wa.w, acc[i][3]);
        acc[i][4] = fmaf(x, wb.x, acc[i][4]); acc[i][5] = fmaf(x, wb.y, acc[i][5]);
        acc[i][6] = fmaf(x, wb.z, acc[i][6]); acc[i][7] = fmaf(x, wb.w, acc[i][7]);
      }
    }
  }

  float bv[8], gv[8], bev[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
    bv[j] = col < n_out ? bias[col] : 0.f;
    if (LN_ACT) { gv[j] = ln_g[col]; bev[j] = ln_b[col]; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[i][j] + bv[j];
    if (LN_ACT) {
      // LayerNorm over the 128 channels of this frame: 16 threads (one half-warp) hold 8 each
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.0f / kEmb);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = 1.0f / sqrtf(q * (1.0f / kEmb) + 1e-5f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = fmaf((v[j] - mean) * rstd, gv[j], bev[j]);
        v[j] = y > 0.f ? y : 0.01f * y;
      }
    }
    const int m = m0 + ty * 4 + i;
    if (m < M) {
      float* yr = Y + (size_t)m * ldy + n0;
      if (n0 + tx * 4 + 3 < ldy) *reinterpret_cast<float4*>(yr + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
      if (n0 + 64 + tx * 4 + 3 < ldy) *reinterpret_cast<float4*>(yr + 64 + tx * 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

int nws_launch_linear(const float* X, const float* Wt, const float* bias, const float* ln_g, const float* ln_b,
                      float* Y, int M, int n_out, int ldw, int ldy, bool ln_act, cudaStream_t s) {
  const size_t smem = (size_t)(kEmb * kEmb + kLinFrames * kLinXs) * sizeof(float);
  static bool attr_done[64] = {};
  if (nws_first_use_on_device(attr_done)) {
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_linear128_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_linear128_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((M + kLinFrames - 1) / kLinFrames, (ldy + 127) / 128);
  if (ln_act) {
    if (n_out != kEmb) { nws_set_error("LayerNorm epilogue needs n_out == 128"); return NWS_ERR_INVALID; }
    nws_linear128_kernel<true><<<grid, 256, smem, s>>>(X, Wt, bias, ln_g, ln_b, Y, M, n_out, ldw, ldy);
  } else {
    nws_linear128_kernel<false><<<grid, 256, smem, s>>>(X, Wt, bias, nullptr, nullptr, Y, M, n_out, ldw, ldy);
  }
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// TimeDistributedMLP, depth 4 (shaping.py:53-55 / neural_waveshaping.py:58 with newt.gin:20-23).
int nws_launch_td_mlp(const NwsContext* ctx, int which, const float* emb, float* act0, float* act1, float* out, int M,
                      cudaStream_t s) {
  const NwsTdMlpOffsets& o = ctx->lay.mlp[which];
  const float* p = ctx->packed;
  const float* in = emb;
  float* bufs[2] = {act0, act1};
  for (int l = 0; l < 3; ++l) {
    float* dst = bufs[l & 1];
    int rc = nws_launch_linear(in, p + o.wt[l], p + o.b[l], p + o.g[l], p + o.beta[l], dst, M, kEmb, kEmb, kEmb, true, s);
    if (rc) return rc;
    in = dst;
  }
  const int n_out = which == NWS_MLP_FILM ? kFilm : kBands;
  return nws_launch_linear(in, p + o.wt_out, p + o.b_out, nullptr, nullptr, out, M, n_out, o.ld_out, o.ld_out, false, s);
}

// ------------------------------------------------------------------------------------------------
// Layout conversion for the stage entry points: reference [B,C,T] <-> frame-major rows [B*T][ld].
__global__ void nws_bct_to_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? in[((size_t)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < ld) out[((size_t)b * T + t) * ld + c] = c < C ? tile[tx][i] : 0.f;
  }
}

__global__ void nws_rows_to_bct_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    tile[i][tx] = (c < C && t < T) ? in[((size_t)b * T + t) * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (c < C && t < T) out[((size_t)b * C + c) * T + t] = tile[tx][i];
  }
}

int nws_launch_bct_to_rows(const float* in, float* out, int B, int C, int T, int ld_out, cudaStream_t s) {
  dim3 grid((T + 31) / 32, (ld_out + 31) / 32, B), block(32, 8);
  nws_bct_to_rows_kernel<<<grid, block, 0, s>>>(in, out, C, T, ld_out);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

int nws_launch_rows_to_bct(const float* in, float* out, int B, int C, int T, int ld_in, cudaStream_t s) {
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  nws_rows_to_bct_kernel<<<grid, block, 0, s>>>(in, out, C, T, ld_in);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
