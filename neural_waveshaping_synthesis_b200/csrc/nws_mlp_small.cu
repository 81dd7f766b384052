// Hop-rate MLP chain for a HANDFUL of frames (the buffer sweep of scripts/time_buffer_sizes.py and streaming pushes:
// 2..40 frames per utterance): embedding.proj + NEWT.mlp + h_generator (neural_waveshaping.py:24-26,58,78,82;
// dynamic.py:20-40; shaping.py:53-55,68) in one launch.
//
// The tensor-core chain (nws_mlp_tc.cu) works on 128-frame tiles and streams 1.3 MB of pre-split weights per tile
// through one SM: ~35 us for ONE tile however few of its 128 rows are real, because the five dependent layers of a
// chain each wait for their 64-128 KB of weights.  For a few frames the arithmetic is nothing; what matters is that
// every layer's weights are already on chip when its input arrives.  So each chain (proj -> FiLM MLP | proj -> noise
// MLP; the projection is recomputed: 16 k MACs per frame) runs on a CLUSTER of four CTAs:
//   * CTA r of the cluster owns output channels 32r..32r+31 of every hidden layer (64 / 33 of the output layer): its
//     slices of all five layers' weights (98 KB) are loaded into shared memory once, up front, in parallel;
//   * per layer every CTA computes its channels for all frames (thread = channel x every 8th frame, fp32 FMA, four
//     interleaved accumulators), stores them into the activation buffer of ALL four CTAs (st.shared::cluster) and
//     the cluster barrier publishes them; LayerNorm + LeakyReLU (two-pass variance, eps 1e-5, slope 0.01 — the
//     arithmetic of nws_linear128_kernel) is then done redundantly by each CTA on the gathered rows.
// Six cluster barriers per chain, no weight traffic after the prologue.
#include "nws_internal.cuh"
#include "nws_noise_bodies.cuh"
#include "nws_tc.cuh"

namespace {

constexpr int kClu = 4;              // CTAs per chain
constexpr int kCluThreads = 256;
constexpr int kSlice = kEmb / kClu;  // 32 hidden channels per CTA
constexpr int kMaxF = (kSmallMlpMaxFrames + 7) / 8;   // frames per warp

struct CluParams {
  const float* packed;
  int proj_wt, proj_b;
  NwsTdMlpOffsets mlp[2];
  const float* hbuf;   // [B*T][128] GRU outputs, frame-major
  float* film;         // [B*T][256]
  float* bands;        // [B*T][kBandsPad], or null when only the filtered noise is wanted
  int T;
  // fused noise branch (FIRNoiseSynth.forward, generators.py:21-35): the noise chain's cluster filters the hops
  // [hop_begin, hop_end) with the band gains it has just computed (gathered in shared memory), 15 hops per CTA
  const float2* xspec;      // [T][kBandsPad] noise spectrum, or null: no fused filter
  const float2* tw_master;
  float* dry;               // [B][128 T]
  int hop_begin, hop_end;
};

// dynamic shared memory (floats): weight slices of the 5 layers | X | Y0 | Y1
constexpr int kWHid = kEmb * kSlice;                 // 4096
constexpr int kWOutFilm = kEmb * (kFilm / kClu);     // 128 x 64
constexpr int kOutNoise = kBandsPad / kClu;          // 33 channels per CTA
constexpr int kOutNoisePad = 36;
constexpr int kWOutMax = kWOutFilm;
constexpr int kActs = kSmallMlpMaxFrames * kEmb;
constexpr int kCluSmemFloats = 4 * kWHid + kWOutMax + 3 * kActs;

__device__ __forceinline__ uint32_t clu_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t clu_map(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void clu_store(uint32_t caddr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory"); }
// 16-byte asynchronous global -> shared copy: the whole prologue (98 KB of weight slices per CTA) is in flight at once
__device__ __forceinline__ void clu_cp16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(nws_smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void clu_cp4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(nws_smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void clu_cp_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void clu_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// How the 8 warps of a CTA share a layer: with few frames the 128-term dot products are split over `ks` warps (k
// ranges of 128 / ks) so that a two-frame buffer still uses all eight warps; partial sums meet in shared memory and are
// added in a fixed order.  Warp -> (k part kq, frame lane fw); frames of a warp: fw, fw + fstride, ...
struct CluSplit {
  int ks, fstride, kq, fw;
};
__device__ __forceinline__ CluSplit clu_split(int T, int warp) {
  CluSplit c;
  c.ks = T <= 2 ? 4 : (T <= 4 ? 2 : 1);
  c.fstride = 8 / c.ks;
  c.kq = warp / c.fstride;
  c.fw = warp % c.fstride;
  return c;
}

// acc[f] = sum over this warp's k range of X[m_f][k] * Ws[k][col], frames m_f = fw + fstride f; NW = row length of the slice
template <int NW>
__device__ __forceinline__ void clu_dot(const float* Ws, int col, const float* X, int T, const CluSplit c, float (&acc)[kMaxF]) {
  float part[kMaxF][4];
#pragma unroll
  for (int f = 0; f < kMaxF; ++f) part[f][0] = part[f][1] = part[f][2] = part[f][3] = 0.f;
  const int k_lo = c.kq * (kEmb / c.ks), k_hi = k_lo + kEmb / c.ks;
#pragma unroll 4
  for (int k = k_lo; k < k_hi; k += 4) {
    const float w0 = Ws[(k + 0) * NW + col], w1 = Ws[(k + 1) * NW + col], w2 = Ws[(k + 2) * NW + col], w3 = Ws[(k + 3) * NW + col];
#pragma unroll
    for (int f = 0; f < kMaxF; ++f) {
      const int m = c.fw + c.fstride * f;
      if (m < T) {
        const float4 x = *reinterpret_cast<const float4*>(X + m * kEmb + k);   // warp-wide broadcast
        part[f][0] = fmaf(x.x, w0, part[f][0]); part[f][1] = fmaf(x.y, w1, part[f][1]);
        part[f][2] = fmaf(x.z, w2, part[f][2]); part[f][3] = fmaf(x.w, w3, part[f][3]);
      }
    }
  }
#pragma unroll
  for (int f = 0; f < kMaxF; ++f) acc[f] = (part[f][0] + part[f][1]) + (part[f][2] + part[f][3]);
}

// complete the dot products of a k-split layer: parts kq > 0 park their sums in `red` [ks - 1][frames <= 4][64],
// part 0 adds them in order.  Returns true for the threads that hold complete sums (all threads when ks == 1).
__device__ __forceinline__ bool clu_reduce(float (&acc)[kMaxF], float* red, int col, int T, const CluSplit c) {
  if (c.ks == 1) return true;
  const int m = c.fw;                      // ks > 1 means T <= 4: one frame per warp
  if (c.kq > 0 && m < T) red[((c.kq - 1) * 4 + m) * 64 + col] = acc[0];
  __syncthreads();
  if (c.kq == 0 && m < T)
    for (int q = 0; q < c.ks - 1; ++q) acc[0] += red[(q * 4 + m) * 64 + col];
  __syncthreads();                          // `red` is reused by the next layer
  return c.kq == 0;
}

// TimeDistributedLayerNorm + LeakyReLU (dynamic.py:11-17,36) of Y rows into X rows: one warp per frame
__device__ __forceinline__ void clu_ln_act(const float* Y, float* X, int T, const float* __restrict__ g,
                                           const float* __restrict__ beta, int warp, int lane) {
  const float4 gv = *reinterpret_cast<const float4*>(g + 4 * lane), bv = *reinterpret_cast<const float4*>(beta + 4 * lane);
  for (int m = warp; m < T; m += kCluThreads / 32) {
    const float4 v = *reinterpret_cast<const float4*>(Y + m * kEmb + 4 * lane);
    float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / kEmb);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    float q = fmaf(dw, dw, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / kEmb) + 1e-5f);
    float4 y;
    y.x = fmaf(dx * rstd, gv.x, bv.x); y.y = fmaf(dy * rstd, gv.y, bv.y);
    y.z = fmaf(dz * rstd, gv.z, bv.z); y.w = fmaf(dw * rstd, gv.w, bv.w);
    y.x = y.x > 0.f ? y.x : 0.01f * y.x; y.y = y.y > 0.f ? y.y : 0.01f * y.y;
    y.z = y.z > 0.f ? y.z : 0.01f * y.z; y.w = y.w > 0.f ? y.w : 0.01f * y.w;
    *reinterpret_cast<float4*>(X + m * kEmb + 4 * lane) = y;
  }
}

__global__ void __launch_bounds__(kCluThreads, 1) nws_mlp_small_kernel(const CluParams p) {
  extern __shared__ __align__(16) float sm[];
  float* Wh = sm;                        // [4][128][32]: proj, hidden 1..3
  float* Wo = sm + 4 * kWHid;            // output layer slice
  float* X = Wo + kWOutMax;              // [T][128] input of the current layer
  float* Y0 = X + kActs;                 // gathered outputs, alternating
  float* Y1 = Y0 + kActs;
  __shared__ float red[3 * 4 * 64];
  const int chain = blockIdx.y, b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, T = p.T;
  const int rank = (int)clu_rank();
  const CluSplit sp = clu_split(T, warp);
  const float* w = p.packed;
  const NwsTdMlpOffsets& o = p.mlp[chain];

  // ---- prologue: this CTA's slices of all five layers, and the GRU rows
  const int wt_off[4] = {p.proj_wt, o.wt[0], o.wt[1], o.wt[2]};
#pragma unroll
  for (int l = 0; l < 4; ++l)
    for (int i = tid; i < kWHid / 4; i += kCluThreads) {        // 128 rows x 8 float4
      const int k = i >> 3, c4 = (i & 7) * 4;
      clu_cp16(Wh + l * kWHid + k * kSlice + c4, w + wt_off[l] + (size_t)k * kEmb + rank * kSlice + c4);
    }
  if (chain == 0) {
    for (int i = tid; i < kWOutFilm / 4; i += kCluThreads) {     // 128 rows x 16 float4
      const int k = i >> 4, c4 = (i & 15) * 4;
      clu_cp16(Wo + k * 64 + c4, w + o.wt_out + (size_t)k * kFilm + rank * 64 + c4);
    }
  } else {
    for (int i = tid; i < kEmb * kOutNoise; i += kCluThreads) {
      const int k = i / kOutNoise, j = i % kOutNoise;
      clu_cp4(Wo + k * kOutNoisePad + j, w + o.wt_out + (size_t)k * kBandsPad + rank * kOutNoise + j);
    }
  }
  nws_pdl_wait();     // the front-end launch (GRU rows, noise spectrum) has completed ...
  nws_pdl_launch();   // ... so the audio kernel may start its own prologue
  const float* rows = p.hbuf + (size_t)b * T * kEmb;
  for (int i = tid; i < T * kEmb / 4; i += kCluThreads) clu_cp16(reinterpret_cast<float4*>(X) + i, reinterpret_cast<const float4*>(rows) + i);
  clu_cp_wait_all();
  __syncthreads();
  clu_sync();   // every CTA of the cluster is running: its shared memory may be written from now on

  const uint32_t y_addr[2] = {nws_smem_u32(Y0), nws_smem_u32(Y1)};
  const int n = rank * kSlice + lane;   // this thread's hidden channel
  // ---- proj (no activation) and the three hidden layers
  const float* in = X;
  for (int l = 0; l < 4; ++l) {
    float acc[kMaxF];
    clu_dot<kSlice>(Wh + l * kWHid, lane, in, T, sp, acc);
    const bool owner = clu_reduce(acc, red, lane, T, sp);
    const float bias = w[(l == 0 ? p.proj_b : o.b[l - 1]) + n];
    const int buf = l & 1;
#pragma unroll
    for (int f = 0; f < kMaxF; ++f) {
      const int m = sp.fw + sp.fstride * f;
      if (m < T && owner) {
        const float v = acc[f] + bias;
        const uint32_t a = y_addr[buf] + (uint32_t)(m * kEmb + n) * 4u;
#pragma unroll
        for (int r = 0; r < kClu; ++r) clu_store(clu_map(a, r), v);
      }
    }
    clu_sync();
    const float* Y = buf ? Y1 : Y0;
    if (l == 0) {
      in = Y;                 // the embedding feeds the first hidden layer as it is
    } else {
      clu_ln_act(Y, X, T, w + o.g[l - 1], w + o.beta[l - 1], warp, lane);
      __syncthreads();
      in = X;
    }
  }
  // ---- output layer: straight to global memory (256 FiLM parameters / 129 band gains per frame)
  if (chain == 0) {
    float* out = p.film + (size_t)b * T * kFilm;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float acc[kMaxF];
      clu_dot<64>(Wo, lane + 32 * h, in, T, sp, acc);
      const bool owner = clu_reduce(acc, red, lane + 32 * h, T, sp);
      const int c = rank * 64 + lane + 32 * h;
      const float bias = w[o.b_out + c];
#pragma unroll
      for (int f = 0; f < kMaxF; ++f) {
        const int m = sp.fw + sp.fstride * f;
        if (m < T && owner) out[(size_t)m * kFilm + c] = acc[f] + bias;
      }
    }
  } else {
    float* out = p.bands ? p.bands + (size_t)b * T * kBandsPad : nullptr;
    const uint32_t bands_addr = y_addr[0];        // Y0 | Y1 are free now: band rows [T][kBandsPad] for the fused filter
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = h == 0 ? lane : 32;     // channel within the slice: 0..32 (the second pass is the one extra column,
      float acc[kMaxF];                     //  computed by every lane so that the block barriers stay uniform)
      clu_dot<kOutNoisePad>(Wo, j, in, T, sp, acc);
      const bool owner = clu_reduce(acc, red, j, T, sp) && (h == 0 || lane == 0);
      const int c = rank * kOutNoise + j;
      const float bias = c < kBands ? w[o.b_out + c] : 0.f;
#pragma unroll
      for (int f = 0; f < kMaxF; ++f) {
        const int m = sp.fw + sp.fstride * f;
        if (m < T && c < kBands && owner) {
          const float v = acc[f] + bias;
          if (out) out[(size_t)m * kBandsPad + c] = v;
          if (p.xspec) {
            const uint32_t a = bands_addr + (uint32_t)(m * kBandsPad + c) * 4u;
#pragma unroll
            for (int r = 0; r < kClu; ++r) clu_store(clu_map(a, r), v);
          }
        }
      }
    }
    if (p.xspec) {
      __syncthreads();
      clu_sync();   // every CTA of the cluster holds all 129 band gains of every frame
      if (rank * kNoiseHops < p.hop_end - p.hop_begin)
        nws_noise_filter_body(Y0, p.xspec, p.tw_master, p.dry + (size_t)b * T * kHop, T, p.hop_begin, p.hop_end, rank);
    }
  }
  __syncthreads();
  clu_sync();   // no CTA leaves while a peer could still be writing into its shared memory
}

}  // namespace

bool nws_mlp_small_ok(const NwsContext* ctx, int B, int T) {
  return ctx->small_path && T >= 1 && T <= kSmallMlpMaxFrames && B >= 1 && B <= kSmallMlpMaxBatch;
}

int nws_launch_mlp_small(const NwsContext* ctx, const float* hbuf, float* film, float* bands, int B, int T, cudaStream_t s,
                         const float2* xspec, float* dry, int hop_begin, int hop_end, bool pdl) {
  CluParams p{};
  if (xspec && (hop_end - hop_begin > kClu * kNoiseHops || !dry)) { nws_set_error("nws_launch_mlp_small: bad fused noise range"); return NWS_ERR_INVALID; }
  p.xspec = xspec; p.tw_master = ctx->tw_master; p.dry = dry; p.hop_begin = hop_begin; p.hop_end = hop_end;
  p.packed = ctx->packed; p.proj_wt = ctx->lay.proj_wt; p.proj_b = ctx->lay.proj_b;
  p.mlp[0] = ctx->lay.mlp[0]; p.mlp[1] = ctx->lay.mlp[1];
  p.hbuf = hbuf; p.film = film; p.bands = bands; p.T = T;
  const size_t smem = (size_t)kCluSmemFloats * sizeof(float);
  static bool attr_done[64] = {};
  if (nws_first_use_on_device(attr_done))
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_mlp_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kClu, 2, B);
  cfg.blockDim = dim3(kCluThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClu; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  int n_attr = 1;
  nws_pdl_config(&cfg, attr, &n_attr, pdl);
  NWS_CUDA_OK(cudaLaunchKernelEx(&cfg, nws_mlp_small_kernel, p));
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
