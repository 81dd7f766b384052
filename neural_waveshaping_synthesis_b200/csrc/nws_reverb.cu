// Learned reverb: Reverb.forward (modules/shaping.py:161-173):
//     out = x + irfft(rfft(pad(x)) * rfft(pad([0, ir])))[:N],  circular length Lc = max(N, 32000).
// Restated as a LINEAR convolution followed by the wrap  out[n] = x[n] + y[n] + y[n + Lc]
//  so the transform length can be a power of two: L = n1 * 256 >= N + 31999.
//
// Four-step FFT over the [n1][256] view of the padded signal, two utterances per transform (one in
// the real part, one in the imaginary part — the IR is real, so they never mix):
//   R1  column FFTs (length n1, stride 256) + twiddle             -> work
//   R2  row FFT (256) * IR spectrum, inverse row FFT, in place    -> work
//   R3  conj twiddle + inverse column FFTs, 1/L                   -> work (time domain, complex pair)
//   R4  fold the circular wrap and add the dry signal             -> out
// The IR spectrum (same four-step layout) and the big twiddle table are cached per transform length.
//
// Exact-length plans.  When Lc itself is n1 * 256 with n1 = 250 (N = 64000: the 4 s utterances of every
// benchmark config), n1 = 125 (Lc = 32000: everything up to 2 s) or a power of two (N = 32768), the circular
// convolution runs at exactly Lc points — mixed-radix column transforms (nws_fft_mixed.cuh) — and there is
// no wrap to fold: half the points of the padded transform (64000 instead of 131072).
#include <math.h>
#include <stdlib.h>

#include "nws_fft.cuh"
#include "nws_fft_mixed.cuh"
#include "nws_internal.cuh"

int nws_reverb_fft_len(int N) {
  const long long need = (long long)N + kReverbIr - 1;
  for (int n1 = 128; n1 <= kTwMaster; n1 <<= 1)
    if ((long long)n1 * 256 >= need) return n1 * 256;
  return 0;
}

int nws_reverb_exact_len(int N) {
  const int Lc = N > kReverbIr ? N : kReverbIr;
  if (Lc % 256) return 0;
  const int n1 = Lc / 256;
  if (n1 == 125 || n1 == 250) return Lc;
  if (n1 >= 128 && n1 <= kTwMaster && (n1 & (n1 - 1)) == 0) return Lc;
  return 0;
}

// ---------------------------------------------------------------------------------------------- R1
// grid (256 / W, n_pairs); dynamic smem 2 * n1 * W float2 + n1/2 float2.
__global__ void __launch_bounds__(256) nws_reverb_cols_fwd_kernel(const float* __restrict__ x, int B, int N,
                                                                  float2* __restrict__ work,
                                                                  const float2* __restrict__ tw_big,
                                                                  const float2* __restrict__ tw_master, int n1,
                                                                  int log_n1, int log_w) {
  extern __shared__ __align__(16) float2 smem2[];
  const int W = 1 << log_w;   // columns per CTA (a power of two: index arithmetic is shifts and masks)
  float2* a = smem2;
  float2* bb = smem2 + n1 * W;
  float2* tw_s = smem2 + 2 * n1 * W;
  const int tid = threadIdx.x, pair = blockIdx.y, c0 = blockIdx.x * W;
  const size_t L = (size_t)n1 * 256;
  for (int i = tid; i < n1 / 2; i += 256) tw_s[i] = tw_master[i * (kTwMaster / n1)];
  const float* xa = x + (size_t)(2 * pair) * N;
  const float* xb = 2 * pair + 1 < B ? x + (size_t)(2 * pair + 1) * N : nullptr;
#pragma unroll 4
  for (int i = tid; i < n1 * W; i += 256) {
    const int r = i >> log_w, c = i & (W - 1);
    const long long n = (long long)r * 256 + c0 + c;
    a[i] = n < N ? make_float2(__ldg(xa + n), xb ? __ldg(xb + n) : 0.f) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const float2* z = nws_fft_smem<false, true>(a, bb, tw_s, 1, log_n1, log_w, tid, 256);
  float2* dst = work + (size_t)pair * L;
#pragma unroll 4
  for (int i = tid; i < n1 * W; i += 256) {
    const int k1 = i >> log_w, c = i & (W - 1);
    const size_t idx = (size_t)k1 * 256 + c0 + c;
    dst[idx] = nws_cmul(z[i], __ldg(tw_big + idx));
  }
}


// ---------------------------------------------------------------------------------------------- R1 / R3, mixed radix
// Exact-length plans (n1 = N1 in {125, 250}): grid (256 / W, n_pairs); dynamic smem (2 * N1 * W + N1) float2.
template <int N1>
__global__ void __launch_bounds__(256) nws_reverb_cols_fwd_mixed_kernel(const float* __restrict__ x, int B, int N,
                                                                        float2* __restrict__ work,
                                                                        const float2* __restrict__ tw_big,
                                                                        const float2* __restrict__ tw_cols, int log_w) {
  extern __shared__ __align__(16) float2 smem2[];
  const int W = 1 << log_w;
  float2* a = smem2;
  float2* bb = smem2 + N1 * W;
  float2* tw_s = smem2 + 2 * N1 * W;
  const int tid = threadIdx.x, pair = blockIdx.y, c0 = blockIdx.x * W;
  for (int i = tid; i < N1; i += 256) tw_s[i] = tw_cols[i];
  const float* xa = x + (size_t)(2 * pair) * N;
  const float* xb = 2 * pair + 1 < B ? x + (size_t)(2 * pair + 1) * N : nullptr;
#pragma unroll 4
  for (int i = tid; i < N1 * W; i += 256) {
    const int r = i >> log_w, c = i & (W - 1);
    const int n = r * 256 + c0 + c;
    a[i] = n < N ? make_float2(__ldg(xa + n), xb ? __ldg(xb + n) : 0.f) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const float2* z = nws_fft_mixed<N1, false>(a, bb, tw_s, log_w, tid, 256);
  float2* dst = work + (size_t)pair * (N1 * 256);
#pragma unroll 4
  for (int i = tid; i < N1 * W; i += 256) {
    const int k1 = i >> log_w, c = i & (W - 1);
    const int idx = k1 * 256 + c0 + c;
    dst[idx] = nws_cmul(z[i], __ldg(tw_big + idx));
  }
}

// conj twiddle + inverse column transforms, 1/L, dry add: out[b][n] = x[b][n] + y[n] (the transform length IS the
// circular length: nothing to fold)
template <int N1>
__global__ void __launch_bounds__(256) nws_reverb_cols_inv_mixed_kernel(const float2* __restrict__ work,
                                                                        const float2* __restrict__ tw_big,
                                                                        const float2* __restrict__ tw_cols, int log_w,
                                                                        const float* __restrict__ x,
                                                                        float* __restrict__ out, int B, int N) {
  extern __shared__ __align__(16) float2 smem2[];
  const int W = 1 << log_w;
  float2* a = smem2;
  float2* bb = smem2 + N1 * W;
  float2* tw_s = smem2 + 2 * N1 * W;
  const int tid = threadIdx.x, c0 = blockIdx.x * W, pair = blockIdx.y;
  const float2* wk = work + (size_t)pair * (N1 * 256);
  for (int i = tid; i < N1; i += 256) tw_s[i] = tw_cols[i];
#pragma unroll 4
  for (int i = tid; i < N1 * W; i += 256) {
    const int k1 = i >> log_w, c = i & (W - 1);
    const int idx = k1 * 256 + c0 + c;
    float2 t = __ldg(tw_big + idx);
    t.y = -t.y;
    a[i] = nws_cmul(wk[idx], t);
  }
  __syncthreads();
  const float2* z = nws_fft_mixed<N1, true>(a, bb, tw_s, log_w, tid, 256);
  const float scale = 1.0f / (float)(N1 * 256);
  const int rows = (N + 255) / 256;   // <= N1
  const int b0 = 2 * pair, b1 = 2 * pair + 1;
  for (int i = tid; i < rows * W; i += 256) {
    const int r = i >> log_w, c = i & (W - 1);
    const int n = r * 256 + c0 + c;
    if (n >= N) continue;
    const float2 v = z[i];
    out[(size_t)b0 * N + n] = x[(size_t)b0 * N + n] + v.x * scale;
    if (b1 < B) out[(size_t)b1 * N + n] = x[(size_t)b1 * N + n] + v.y * scale;
  }
}

// ---------------------------------------------------------------------------------------------- R2
// grid ((n1 + 1) / 2, n_pairs), 256 threads = two rows (the second one idles in the last CTA of an odd n1).
// mode 0: fwd FFT, * ir_spec, inverse FFT.  mode 1 (plan building): fwd FFT only (the result IS the IR spectrum).
__global__ void __launch_bounds__(256) nws_reverb_rows_kernel(float2* __restrict__ work, const float2* __restrict__ ir_spec,
                                                              const float2* __restrict__ tw_master, int n1, int mode) {
  __shared__ float2 buf_a[2][256], buf_b[2][256], tw_s[128];
  const int tid = threadIdx.x, g = tid >> 7, j = tid & 127;
  const size_t L = (size_t)n1 * 256;
  const bool live = 2 * (int)blockIdx.x + g < n1;
  const size_t row = live ? (size_t)(2 * blockIdx.x + g) * 256 : 0;
  float2* w = work + (size_t)blockIdx.y * L + row;
  for (int i = tid; i < 128; i += 256) tw_s[i] = tw_master[i * (kTwMaster / 256)];
  buf_a[g][j] = live ? w[j] : make_float2(0.f, 0.f);
  buf_a[g][j + 128] = live ? w[j + 128] : make_float2(0.f, 0.f);
  __syncthreads();
  float2* z = nws_fft_smem<false, false>(&buf_a[0][0], &buf_b[0][0], tw_s, 1, 8, 1, tid, 256);
  if (mode == 1) {
    if (live) {
      w[j] = z[g * 256 + j];
      w[j + 128] = z[g * 256 + j + 128];
    }
    return;
  }
  float2* other = z == &buf_a[0][0] ? &buf_b[0][0] : &buf_a[0][0];
  z[g * 256 + j] = nws_cmul(z[g * 256 + j], ir_spec[row + j]);
  z[g * 256 + j + 128] = nws_cmul(z[g * 256 + j + 128], ir_spec[row + j + 128]);
  __syncthreads();
  const float2* y = nws_fft_smem<true, false>(z, other, tw_s, 1, 8, 1, tid, 256);
  if (live) {
    w[j] = y[g * 256 + j];
    w[j + 128] = y[g * 256 + j + 128];
  }
}

// ---------------------------------------------------------------------------------------------- R3
// FUSE_FOLD (possible when the circular length Lc is a multiple of 256, i.e. the wrapped sample n + Lc sits in
// the same column, Lc/256 rows further down, inside this CTA's shared memory): the wrap, the dry add and the
// final store happen here and the time-domain work array is never written or re-read.
template <bool FUSE_FOLD>
__global__ void __launch_bounds__(256) nws_reverb_cols_inv_kernel(float2* __restrict__ work,
                                                                  const float2* __restrict__ tw_big,
                                                                  const float2* __restrict__ tw_master, int n1,
                                                                  int log_n1, int log_w, const float* __restrict__ x,
                                                                  float* __restrict__ out, int B, int N, int Lc) {
  extern __shared__ __align__(16) float2 smem2[];
  const int W = 1 << log_w;   // columns per CTA (a power of two: index arithmetic is shifts and masks)
  float2* a = smem2;
  float2* bb = smem2 + n1 * W;
  float2* tw_s = smem2 + 2 * n1 * W;
  const int tid = threadIdx.x, c0 = blockIdx.x * W, pair = blockIdx.y;
  const size_t L = (size_t)n1 * 256;
  float2* wk = work + (size_t)pair * L;
  for (int i = tid; i < n1 / 2; i += 256) tw_s[i] = tw_master[i * (kTwMaster / n1)];
#pragma unroll 4
  for (int i = tid; i < n1 * W; i += 256) {
    const int k1 = i >> log_w, c = i & (W - 1);
    const size_t idx = (size_t)k1 * 256 + c0 + c;
    float2 t = __ldg(tw_big + idx);
    t.y = -t.y;
    a[i] = nws_cmul(wk[idx], t);
  }
  __syncthreads();
  const float2* z = nws_fft_smem<true, true>(a, bb, tw_s, 1, log_n1, log_w, tid, 256);
  const float scale = 1.0f / (float)L;
  if (FUSE_FOLD) {
    // out[b][n] = x[b][n] + y[n] + y[n + Lc], y = linear convolution (zero beyond N + 31998)
    const int rows = (N + 255) / 256, shift = Lc / 256;
    const long long ylen = (long long)N + kReverbIr - 1;
    const int b0 = 2 * pair, b1 = 2 * pair + 1;
    for (int i = tid; i < rows * W; i += 256) {
      const int r = i >> log_w, c = i & (W - 1);
      const long long n = (long long)r * 256 + c0 + c;
      if (n >= N) continue;
      float2 v = z[i];
      if (n + Lc < ylen) {   // r + shift < n1 always holds here: n + Lc < N + 32000 <= L
        const float2 u = z[(r + shift) * W + c];
        v.x += u.x; v.y += u.y;
      }
      out[(size_t)b0 * N + n] = x[(size_t)b0 * N + n] + v.x * scale;
      if (b1 < B) out[(size_t)b1 * N + n] = x[(size_t)b1 * N + n] + v.y * scale;
    }
  } else {
    for (int i = tid; i < n1 * W; i += 256) {
      const int r = i >> log_w, c = i & (W - 1);
      const float2 v = z[i];
      wk[(size_t)r * 256 + c0 + c] = make_float2(v.x * scale, v.y * scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------- R4
// out[b][n] = x[b][n] + y[n] + y[n + Lc]   (y = linear convolution, zero beyond N + 31998)
__global__ void __launch_bounds__(256) nws_reverb_fold_kernel(const float* __restrict__ x, const float2* __restrict__ work,
                                                              float* __restrict__ out, int B, int N, size_t L, int Lc) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const float2* y = work + (size_t)(b >> 1) * L;
  const int ylen = N + kReverbIr - 1;
  const float2 v0 = y[n];
  float acc = (b & 1) ? v0.y : v0.x;
  const long long m = (long long)n + Lc;
  if (m < ylen) {
    const float2 v1 = y[m];
    acc += (b & 1) ? v1.y : v1.x;
  }
  out[(size_t)b * N + n] = x[(size_t)b * N + n] + acc;
}

// ---------------------------------------------------------------------------------------------- plans
static size_t cols_smem_bytes(int n1, int W) { return ((size_t)2 * n1 * W + n1 / 2) * sizeof(float2); }
static size_t cols_mixed_smem_bytes(int n1, int W) { return ((size_t)2 * n1 * W + n1) * sizeof(float2); }

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int pick_cols(int n1) {
  int W = 4096 / n1;  // 64 KB of ping-pong buffers: three CTAs per SM keep more loads in flight
  if (W > 16) W = 16;
  if (W < 1) W = 1;
  return W;
}

int nws_make_twiddle_master(NwsContext* ctx) {
  const int n = kTwMaster / 2;
  float2* h = (float2*)malloc(n * sizeof(float2));
  if (!h) return NWS_ERR_CUDA;
  for (int m = 0; m < n; ++m) {
    const double a = -2.0 * M_PI * (double)m / (double)kTwMaster;
    h[m] = make_float2((float)cos(a), (float)sin(a));
  }
  cudaError_t e = cudaMalloc(&ctx->tw_master, n * sizeof(float2));
  if (e == cudaSuccess) e = cudaMemcpy(ctx->tw_master, h, n * sizeof(float2), cudaMemcpyHostToDevice);
  free(h);
  if (e != cudaSuccess) { nws_set_error("twiddle table: %s", cudaGetErrorString(e)); return NWS_ERR_CUDA; }
  return NWS_OK;
}

void nws_reverb_invalidate(NwsContext* ctx) {
  for (int i = 0; i < ctx->n_plans; ++i) ctx->plans[i].ir_valid = false;
}

void nws_reverb_free_plans(NwsContext* ctx) {
  for (int i = 0; i < ctx->n_plans; ++i) {
    cudaFree(ctx->plans[i].tw_big);
    cudaFree(ctx->plans[i].ir_spec);
    cudaFree(ctx->plans[i].tw_cols);
  }
  ctx->n_plans = 0;
}

// Columns per CTA of the mixed-radix column kernels, chosen per launch so the grid is one balanced wave:
// 8 columns = 32 KB of ping-pong buffers -> 7 CTAs per SM = 1036 slots for the 1024 CTAs of a 64-utterance batch
// (16 columns gave 512 CTAs on 444 slots: a second wave 13 % full); 4 columns for a handful of utterances, where
// the number of CTAs in flight is what bounds the latency.
static int mixed_cols(int n_pairs) { return n_pairs <= 4 ? 4 : 8; }

static int launch_cols_fwd(NwsContext* ctx, NwsReverbPlan* pl, const float* x, int B, int N, float2* work, cudaStream_t s) {
  const int W = pl->mixed ? mixed_cols((B + 1) / 2) : pl->cols_per_cta;
  dim3 grid(256 / W, (B + 1) / 2);
  if (pl->mixed) {
    const size_t smem = cols_mixed_smem_bytes(pl->n1, W);
    if (pl->n1 == 125) nws_reverb_cols_fwd_mixed_kernel<125><<<grid, 256, smem, s>>>(x, B, N, work, pl->tw_big, pl->tw_cols, ilog2(W));
    else nws_reverb_cols_fwd_mixed_kernel<250><<<grid, 256, smem, s>>>(x, B, N, work, pl->tw_big, pl->tw_cols, ilog2(W));
    NWS_LAUNCH_CHECK();
    return NWS_OK;
  }
  nws_reverb_cols_fwd_kernel<<<grid, 256, cols_smem_bytes(pl->n1, W), s>>>(x, B, N, work, pl->tw_big, ctx->tw_master,
                                                                          pl->n1, pl->log_n1, ilog2(W));
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

int nws_reverb_get_plan(NwsContext* ctx, int fft_len, cudaStream_t s, NwsReverbPlan** out) {
  const int n1 = fft_len / 256;
  NwsReverbPlan* pl = nullptr;
  for (int i = 0; i < ctx->n_plans; ++i)
    if (ctx->plans[i].n1 == n1) pl = &ctx->plans[i];
  if (!pl) {
    if (ctx->n_plans == kMaxPlans) { nws_set_error("too many distinct reverb transform lengths"); return NWS_ERR_UNSUPPORTED; }
    pl = &ctx->plans[ctx->n_plans];
    *pl = NwsReverbPlan();
    pl->n1 = n1;
    pl->mixed = n1 == 125 || n1 == 250;
    for (pl->log_n1 = 0; (1 << pl->log_n1) < n1; ++pl->log_n1) {}
    pl->cols_per_cta = pl->mixed ? 8 : pick_cols(n1);    // (mixed plans choose per launch: mixed_cols)
    const size_t L = (size_t)fft_len;
    if (pl->mixed) {
      float2 hc[250];
      for (int m = 0; m < n1; ++m) {
        const double a = -2.0 * M_PI * (double)m / (double)n1;
        hc[m] = make_float2((float)cos(a), (float)sin(a));
      }
      NWS_CUDA_OK(cudaMalloc(&pl->tw_cols, n1 * sizeof(float2)));
      NWS_CUDA_OK(cudaMemcpy(pl->tw_cols, hc, n1 * sizeof(float2), cudaMemcpyHostToDevice));
      NWS_CUDA_OK(cudaFuncSetAttribute(nws_reverb_cols_fwd_mixed_kernel<125>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      NWS_CUDA_OK(cudaFuncSetAttribute(nws_reverb_cols_fwd_mixed_kernel<250>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      NWS_CUDA_OK(cudaFuncSetAttribute(nws_reverb_cols_inv_mixed_kernel<125>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      NWS_CUDA_OK(cudaFuncSetAttribute(nws_reverb_cols_inv_mixed_kernel<250>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    NWS_CUDA_OK(cudaMalloc(&pl->tw_big, L * sizeof(float2)));
    NWS_CUDA_OK(cudaMalloc(&pl->ir_spec, L * sizeof(float2)));
    float2* h = (float2*)malloc(L * sizeof(float2));
    if (!h) { nws_set_error("out of host memory"); return NWS_ERR_CUDA; }
    for (int k1 = 0; k1 < n1; ++k1)
      for (int n2 = 0; n2 < 256; ++n2) {
        const double a = -2.0 * M_PI * (double)((long long)k1 * n2 % (long long)L) / (double)L;
        h[(size_t)k1 * 256 + n2] = make_float2((float)cos(a), (float)sin(a));
      }
    cudaError_t e = cudaMemcpy(pl->tw_big, h, L * sizeof(float2), cudaMemcpyHostToDevice);
    free(h);
    NWS_CUDA_OK(e);
    const size_t smem = cols_smem_bytes(n1, pl->cols_per_cta);
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_reverb_cols_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_reverb_cols_inv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_reverb_cols_inv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    (void)smem;
    ++ctx->n_plans;
  }
  if (!pl->ir_valid) {
    // spectrum of [0, ir] (shaping.py:162) in the four-step layout: R1 then the forward half of R2
    int rc = launch_cols_fwd(ctx, pl, ctx->packed + ctx->lay.ir, 1, kReverbIr, pl->ir_spec, s);
    if (rc) return rc;
    nws_reverb_rows_kernel<<<dim3((n1 + 1) / 2, 1), 256, 0, s>>>(pl->ir_spec, nullptr, ctx->tw_master, n1, 1);
    NWS_LAUNCH_CHECK();
    pl->ir_valid = true;
  }
  *out = pl;
  return NWS_OK;
}

int nws_launch_reverb(NwsContext* ctx, const float* x, float* out, float2* work, int B, int N, cudaStream_t s) {
  // exact: the transform length is the circular length itself (never longer than the padded plan the workspace is
  // sized for)
  const int exact = nws_reverb_exact_len(N);
  const int L = exact ? exact : nws_reverb_fft_len(N);
  if (!L) { nws_set_error("reverb: N = %d too long for the FFT plan (max %d samples)", N, kTwMaster * 256 - kReverbIr); return NWS_ERR_UNSUPPORTED; }
  NwsReverbPlan* pl = nullptr;
  int rc = nws_reverb_get_plan(ctx, L, s, &pl);
  if (rc) return rc;
  const int n_pairs = (B + 1) / 2, W = pl->mixed ? mixed_cols(n_pairs) : pl->cols_per_cta;
  rc = launch_cols_fwd(ctx, pl, x, B, N, work, s);
  if (rc) return rc;
  nws_reverb_rows_kernel<<<dim3((pl->n1 + 1) / 2, n_pairs), 256, 0, s>>>(work, pl->ir_spec, ctx->tw_master, pl->n1, 0);
  NWS_LAUNCH_CHECK();
  if (pl->mixed) {
    const size_t smem = cols_mixed_smem_bytes(pl->n1, W);
    if (pl->n1 == 125)
      nws_reverb_cols_inv_mixed_kernel<125><<<dim3(256 / W, n_pairs), 256, smem, s>>>(work, pl->tw_big, pl->tw_cols, ilog2(W), x, out, B, N);
    else
      nws_reverb_cols_inv_mixed_kernel<250><<<dim3(256 / W, n_pairs), 256, smem, s>>>(work, pl->tw_big, pl->tw_cols, ilog2(W), x, out, B, N);
    NWS_LAUNCH_CHECK();
    return NWS_OK;
  }
  // an exact power-of-two plan has no wrap either: a circular length beyond every index disables the fold
  const int Lc = exact ? (1 << 30) : (N > kReverbIr ? N : kReverbIr);
  if (Lc % 256 == 0) {
    // (x + y*scale: the scale is applied to the sum y[n] + y[n+Lc] — same value up to one rounding)
    nws_reverb_cols_inv_kernel<true><<<dim3(256 / W, n_pairs), 256, cols_smem_bytes(pl->n1, W), s>>>(
        work, pl->tw_big, ctx->tw_master, pl->n1, pl->log_n1, ilog2(W), x, out, B, N, Lc);
    NWS_LAUNCH_CHECK();
  } else {
    nws_reverb_cols_inv_kernel<false><<<dim3(256 / W, n_pairs), 256, cols_smem_bytes(pl->n1, W), s>>>(
        work, pl->tw_big, ctx->tw_master, pl->n1, pl->log_n1, ilog2(W), x, out, B, N, Lc);
    NWS_LAUNCH_CHECK();
    nws_reverb_fold_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(x, work, out, B, N, (size_t)L, Lc);
    NWS_LAUNCH_CHECK();
  }
  return NWS_OK;
}
