// Learned reverb in DIRECT form, for short buffers (Reverb.forward, modules/shaping.py:161-173).
//
// The FFT path (nws_reverb.cu) transforms at least 32000 points whatever the buffer length, because the reference pads
// to max(N, 32000) (shaping.py:163-167): three launches and ~16 us for a 256-sample buffer, ~50 us for a streaming push
// (overlap-save over 32000 + n samples).  For short buffers the convolution itself is tiny:
//   * stateless forward, N <= 4096:   out[n] = x[n] + sum_{m < N} x[m] * ir_[(n - m) mod 32000]       (N^2 MACs; the
//     circular wrap of shaping.py:170-173 puts the IR's LAST taps in front of a short buffer — reproduced literally)
//   * streaming push of n new samples: out[n] = x[n] + sum_{k = 1}^{31999} ir_[k] * x[n - k]            (32000 n MACs)
// with ir_ = [0, ir].  One launch: a CTA multiplies a tile of 256 outputs x 1024 inputs (256 threads: 4 consecutive
// outputs x one quarter of the inputs per thread, taps and inputs staged in shared memory: 3 LDS.128 per 16 FMAs); input
// tiles of one output block are spread over CTAs, partial sums go to a scratch array and the LAST CTA of an output block
// to finish adds them in a fixed order (deterministic: repeat runs are bit-identical) together with the dry signal.  The streaming variant also
// advances the 32000-sample history in the same launch (ping-pong buffers).
#include "nws_internal.cuh"

namespace {

constexpr int kDirOut = 256;      // outputs per CTA
constexpr int kDirIn = 1024;      // inputs per CTA: four quarters of 256, one per 64-thread group
constexpr int kDirThreads = 256;

struct DirectParams {
  // input signal of one utterance = [segment A | segment B], total NX samples
  const float* a; size_t a_stride; int len_a;                    // A: a + b*a_stride, len_a samples (streaming: the history)
  const float* bsrc; size_t b_stride; int b_off;                 // B: bsrc + b*b_stride + b_off, NX - len_a samples
  int NX;
  const float* ir;     // [32000] = [0, ir]
  int off;             // tap index of (output o, input j) is o + off - j
  int n_out;
  float* out; size_t out_stride;
  float* partial;      // [B][S][n_out]
  int* counters;       // [B][n_blocks] zero on entry, zero on exit
  int S;
  // streaming: next history [B][32000] <- last 32000 samples of the input signal; null otherwise
  float* hist_next;
  int dry_only;        // skip the convolution (the streaming API's apply_reverb = 0): out = dry
};

__device__ __forceinline__ float dir_input(const DirectParams& p, int b, int j) {
  if (j < 0 || j >= p.NX) return 0.f;
  return j < p.len_a ? p.a[(size_t)b * p.a_stride + j] : p.bsrc[(size_t)b * p.b_stride + p.b_off + (j - p.len_a)];
}

template <bool WRAP>
__global__ void __launch_bounds__(kDirThreads) nws_reverb_direct_kernel(const DirectParams p) {
  __shared__ __align__(16) float taps_s[kDirOut + kDirIn];
  __shared__ __align__(16) float x_s[kDirIn];
  __shared__ __align__(16) float red_s[3][kDirOut];
  __shared__ int last_s;
  const int tid = threadIdx.x, b = blockIdx.z, tq = tid & 63, grp = tid >> 6;
  const int n_blocks = (p.n_out + kDirOut - 1) / kDirOut;
  if ((int)blockIdx.x >= n_blocks) {
    nws_pdl_wait();
    // history CTAs (streaming): hist_next[i] = input[NX - 32000 + i]
    const int n_hist_ctas = gridDim.x - n_blocks, c = blockIdx.x - n_blocks;
    if (blockIdx.y == 0 && p.hist_next)
      for (int i = c * kDirThreads + tid; i < kReverbIr; i += n_hist_ctas * kDirThreads)
        p.hist_next[(size_t)b * kReverbIr + i] = dir_input(p, b, p.NX - kReverbIr + i);
    return;
  }
  const int o0 = blockIdx.x * kDirOut, j0 = blockIdx.y * kDirIn;
  // taps_s[q] = tap(o0 + off - j0 - (kDirIn - 1) + q): everything the tile's (output, input) pairs can touch
  if (!p.dry_only) {
    for (int q = tid; q < kDirOut + kDirIn; q += kDirThreads) {
      int i = o0 + p.off - j0 - (kDirIn - 1) + q;
      float v = 0.f;
      if (WRAP) {
        if (i < 0) i += kReverbIr;            // |i| < 32000 in the stateless case (N <= 4096)
        v = p.ir[i];
      } else if (i >= 0 && i < kReverbIr) {
        v = p.ir[i];
      }
      taps_s[q] = v;
    }
  }
  nws_pdl_wait();   // the taps above are constants; the signal comes from the preceding launch
  if (!p.dry_only)
    for (int q = tid; q < kDirIn; q += kDirThreads) x_s[q] = dir_input(p, b, j0 + q);
  __syncthreads();
  // thread (tq, grp): outputs 4 tq .. 4 tq + 3, inputs 256 grp .. 256 grp + 255 of the tile
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (!p.dry_only) {
#pragma unroll 4
    for (int jj = 256 * grp; jj < 256 * grp + 256; jj += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x_s + jj);
      // tap slot of (output 4 tq + i, input jj + q) = (4 tq + i) - (jj + q) + (kDirIn - 1) = base + 3 + i - q
      const float* tb = taps_s + (4 * tq - jj + (kDirIn - 4));
      const float4 t0 = *reinterpret_cast<const float4*>(tb), t1 = *reinterpret_cast<const float4*>(tb + 4);
      const float tp[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i] = fmaf(xs[q], tp[3 + i - q], acc[i]);
    }
  }
  // the four input quarters of an output, added in a fixed order
  if (grp > 0) *reinterpret_cast<float4*>(&red_s[grp - 1][4 * tq]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  __syncthreads();
  const int o = o0 + 4 * tq;
  const int dry_at = WRAP ? 0 : p.off;   // the dry sample of output o is input o (stateless) / off + o (streaming)
  if (grp == 0) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 r = *reinterpret_cast<const float4*>(&red_s[g][4 * tq]);
      acc[0] += r.x; acc[1] += r.y; acc[2] += r.z; acc[3] += r.w;
    }
    if (p.S == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (o + i < p.n_out) p.out[(size_t)b * p.out_stride + o + i] = dir_input(p, b, dry_at + o + i) + acc[i];
    } else {
      float* part = p.partial + ((size_t)b * p.S + blockIdx.y) * p.n_out;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (o + i < p.n_out) part[o + i] = acc[i];
    }
  }
  if (p.S == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = atomicAdd(&p.counters[b * n_blocks + blockIdx.x], 1) == p.S - 1;
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  // the last CTA of the output block: partial rows s = grp, grp + 4, ... by thread group, eight loads in flight,
  // then the four groups' sums in a fixed order
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s0 = grp; s0 < p.S; s0 += 32) {
    float v[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sidx = s0 + 4 * u;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        v[u][i] = (sidx < p.S && o + i < p.n_out) ? __ldcg(p.partial + ((size_t)b * p.S + sidx) * p.n_out + o + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) sum[i] += v[u][i];
  }
  __syncthreads();
  if (grp > 0) *reinterpret_cast<float4*>(&red_s[grp - 1][4 * tq]) = make_float4(sum[0], sum[1], sum[2], sum[3]);
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 r = *reinterpret_cast<const float4*>(&red_s[g][4 * tq]);
      sum[0] += r.x; sum[1] += r.y; sum[2] += r.z; sum[3] += r.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (o + i < p.n_out) p.out[(size_t)b * p.out_stride + o + i] = dir_input(p, b, dry_at + o + i) + sum[i];
  }
  if (tid == 0) p.counters[b * n_blocks + blockIdx.x] = 0;   // ready for the next launch (stream-ordered)
}

}  // namespace

// ---- stateless circular form
size_t nws_reverb_direct_scratch_bytes(int B, int N) {
  const int S = (N + kDirIn - 1) / kDirIn;
  return S > 1 ? (size_t)B * S * N * sizeof(float) : 0;
}

bool nws_reverb_direct_ok(const NwsContext* ctx, int B, int N, size_t scratch_bytes) {
  if (!ctx->dir_counters || N < 1 || N > kReverbDirectMaxN) return false;
  if ((long long)B * ((N + kDirOut - 1) / kDirOut) > kDirCounters || B > 65535) return false;
  return scratch_bytes >= nws_reverb_direct_scratch_bytes(B, N);
}

static int launch_direct(bool wrap, dim3 grid, const DirectParams& p, cudaStream_t s, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kDirThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int n_attr = 0;
  nws_pdl_config(&cfg, attr, &n_attr, pdl);
  if (wrap) NWS_CUDA_OK(cudaLaunchKernelEx(&cfg, nws_reverb_direct_kernel<true>, p));
  else NWS_CUDA_OK(cudaLaunchKernelEx(&cfg, nws_reverb_direct_kernel<false>, p));
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

int nws_launch_reverb_direct(NwsContext* ctx, const float* x, float* out, float* scratch, int B, int N, cudaStream_t s, bool pdl) {
  DirectParams p{};
  p.a = nullptr; p.a_stride = 0; p.len_a = 0;
  p.bsrc = x; p.b_stride = (size_t)N; p.b_off = 0; p.NX = N;
  p.ir = ctx->packed + ctx->lay.ir; p.off = 0; p.n_out = N;
  p.out = out; p.out_stride = (size_t)N;
  p.partial = scratch; p.counters = ctx->dir_counters;
  p.S = (N + kDirIn - 1) / kDirIn;
  p.hist_next = nullptr;
  dim3 grid((N + kDirOut - 1) / kDirOut, p.S, B);
  return launch_direct(true, grid, p, s, pdl);
}

// ---- streaming causal form: input = [hist (32000) | dry_new (n_new)], outputs = the n_new new samples
size_t nws_reverb_direct_causal_scratch_bytes(int B, int n_new_max) {
  const int S = (kReverbIr + n_new_max + kDirIn - 1) / kDirIn;
  return (size_t)B * S * n_new_max * sizeof(float);
}

int nws_launch_reverb_direct_causal(NwsContext* ctx, const float* hist, const float* dry, size_t dry_stride, int first_sample,
                                    float* out, float* hist_next, float* scratch, int B, int n_new, int apply_reverb,
                                    cudaStream_t s, bool pdl) {
  DirectParams p{};
  p.a = hist; p.a_stride = kReverbIr; p.len_a = kReverbIr;
  p.bsrc = dry; p.b_stride = dry_stride; p.b_off = first_sample; p.NX = kReverbIr + n_new;
  p.ir = ctx->packed + ctx->lay.ir; p.off = kReverbIr; p.n_out = n_new;
  p.out = out; p.out_stride = (size_t)n_new;
  p.partial = scratch; p.counters = ctx->dir_counters;
  p.S = (p.NX + kDirIn - 1) / kDirIn;
  p.hist_next = hist_next;
  const int n_blocks = (n_new + kDirOut - 1) / kDirOut;
  if ((long long)B * n_blocks > kDirCounters || B > 65535) { nws_set_error("reverb (direct): batch too large"); return NWS_ERR_UNSUPPORTED; }
  if (!apply_reverb) { p.S = 1; p.dry_only = 1; }   // dry output (and the history update) through the same launch
  dim3 grid(n_blocks + 4, p.S, B);   // + 4 history CTAs per utterance (blockIdx.y == 0 only)
  return launch_direct(false, grid, p, s, pdl);
}
