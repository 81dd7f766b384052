// Hop-rate MLP chain on the tensor cores: embedding projection + the two TimeDistributedMLPs
// (neural_waveshaping.py:22,26; dynamic.py:20-40; shaping.py:53-55,68; neural_waveshaping.py:58,82)
// for a tile of 128 frames in ONE kernel, activations never leaving the SM:
//
//   A operand  = activations, in TENSOR MEMORY (lane = frame, column = channel), tf32 hi / lo parts
//   B operand  = weights, streamed from L2 into a shared-memory ring by cp.async.bulk (1-D TMA),
//                pre-split into tf32 hi / lo and pre-arranged in the canonical UMMA K-major layout
//   D          = accumulator in TMEM (two 128-column regions, alternating per weight block)
//   D += A_hi B_hi + A_lo B_hi + A_hi B_lo  (3xTF32: fp32-level accuracy)
//   epilogue   = two threads per frame row, 64 channels each (the two epilogue warpgroups read the same TMEM lanes):
//                TMEM -> registers, bias, LayerNorm over the row's 128 channels (in-thread partial sums, the two halves
//                exchanged through shared memory), LeakyReLU, hi/lo split, tcgen05.st back as the next A.  One thread per
//                row was a lone warp per scheduler running 1,700 dependent instructions per block at IPC 0.2 — the
//                epilogue, not the MMAs (34 % of the time), set the kernel's pace.
//
// Warp roles (320 threads): warps 0-7 two epilogue warpgroups (thread = frame x channel half), warp 8 weight
// producer, warp 9 MMA issuer.  Block sequence per tile: proj, A1, A2, A3, Aout[0:128], Aout[128:256],
// proj (again, from a reload of the GRU states), B1, B2, B3, Bout[0:128], Bout[128:144].
#include "nws_internal.cuh"
#include "nws_tc.cuh"

namespace {

constexpr int kMlpThreads = 320;
constexpr int kEpiThreads = 256;           // two epilogue warpgroups
constexpr int kHalf = kEmb / 2;              // channels per epilogue thread
constexpr int kSlots = 4;
constexpr int kChunkK = 32;                 // K elements per ring chunk (4 MMA k-steps)
constexpr int kSlotBytes = 2 * 128 * kChunkK * 4;  // hi + lo of a [128 x 32] chunk = 32 KB
constexpr int kBlocks = 12;
constexpr uint32_t kColAhi = 0, kColAlo = 128, kColD = 256;

struct MlpBlockDesc {
  int w_off;     // float offset of the block's chunks inside the TC weight blob
  int n;         // MMA N (128 or 16)
  int kind;      // 0: +bias -> next A (proj)   1: +bias, LN, LeakyReLU -> next A
                 // 2: +bias -> film[:, col0:col0+128]   3: +bias -> bands[:, col0 : col0+n_store]
  int col0;      // output column offset for kinds 2/3
  int new_a;     // 1 if this block needs an A version published after the previous block's epilogue
  int n_bias;    // valid bias entries
  int bias, g, beta;  // float offsets into the packed blob (copied to shared memory at kernel start)
};
constexpr int kVecStride = 3 * 128;   // per block: bias[128] | gamma[128] | beta[128]

struct MlpTcParams {
  MlpBlockDesc blk[kBlocks];
  const float* packed;    // ctx->packed (bias / LN vectors)
  const float* w_tc;      // TC weight blob
  const float* h;         // [M][128] GRU states
  float* film;            // [M][256]
  float* bands;           // [M][132]
  int M;
  int T, t_begin, t_end;  // frames [t_begin, t_end) of every utterance (T frames each); tiles never straddle utterances
                          // when a sub-range is processed; t_begin = 0, t_end = T, T = M processes the flat [M] array
  int split;              // 1: even CTAs run blocks 0-5 (proj + FiLM chain), odd CTAs blocks 6-11 (proj + noise chain) of a tile
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nws_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   nws_smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(nws_smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
      "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
      "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// write 16 activations (columns c0..c0+15 of this thread's row) as the next A operand: hi and lo parts
__device__ __forceinline__ void store_a16(uint32_t tmem_row, int c0, const float* y) {
  float hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    hi[i] = nws_tf32_hi(y[i]);
    lo[i] = nws_tf32_lo(y[i], hi[i]);
  }
  tmem_st16(tmem_row + kColAhi + c0, hi);
  tmem_st16(tmem_row + kColAlo + c0, lo);
}

__global__ void __launch_bounds__(kMlpThreads, 1) nws_mlp_tc_kernel(const MlpTcParams p, int* __restrict__ fault) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[kSlots], empty_bar[kSlots], a_ready, d_ready[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int fault_s;
  __shared__ __align__(16) float vec_s[kBlocks * kVecStride];
  __shared__ float ln_sum[2][128], ln_sq[2][128];   // LayerNorm partial sums of the two channel halves of a row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seg = p.t_end - p.t_begin, tiles_per_seg = (seg + 127) / 128;
  const int n_tiles = (p.M / p.T) * tiles_per_seg;
  // work items: a tile, or (split) half of a tile's block sequence — the two halves are self-contained
  const int item0 = p.split ? blockIdx.x >> 1 : blockIdx.x, item_step = p.split ? gridDim.x >> 1 : gridDim.x;
  const int b_begin = p.split ? (blockIdx.x & 1) * 6 : 0, b_end = p.split ? b_begin + 6 : kBlocks;
  for (int i = tid; i < kBlocks * kVecStride; i += kMlpThreads) {
    const int b = i / kVecStride, r = i % kVecStride, which = r >> 7, c = r & 127;
    const MlpBlockDesc& d = p.blk[b];
    float v = 0.f;
    if (which == 0) v = c < d.n_bias ? p.packed[d.bias + c] : 0.f;
    else if (d.kind == 1) v = p.packed[(which == 1 ? d.g : d.beta) + c];
    vec_s[i] = v;
  }

  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) { nws_mbar_init(&full_bar[s], 1); nws_mbar_init(&empty_bar[s], 1); }
    nws_mbar_init(&a_ready, kEpiThreads);
    nws_mbar_init(&d_ready[0], 1);
    nws_mbar_init(&d_ready[1], 1);
    nws_fence_mbar_init();
    fault_s = 0;
  }
  if (warp == 0) nws_tmem_alloc(&tmem_base_s, 512);
  nws_tc_fence_before();
  __syncthreads();
  nws_tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 8) {
    // ================= weight producer: one lane streams every block's chunks through the ring
    if (lane == 0) {
      uint32_t n_fill = 0;   // chunks issued so far (slot = n_fill % kSlots)
      bool ok = true;
      for (int tile = item0; tile < n_tiles && ok; tile += item_step) {
        for (int b = b_begin; b < b_end && ok; ++b) {
          const MlpBlockDesc d = p.blk[b];
          const uint32_t chunk_bytes = 2u * d.n * kChunkK * 4u;
          for (int ck = 0; ck < kEmb / kChunkK && ok; ++ck, ++n_fill) {
            const int s = n_fill % kSlots;
            const uint32_t use = n_fill / kSlots;
            if (use > 0) ok = nws_mbar_wait(&empty_bar[s], (use - 1) & 1);
            if (!ok) break;
            mbar_expect_tx(&full_bar[s], chunk_bytes);
            bulk_g2s(smem + s * kSlotBytes, p.w_tc + d.w_off + (size_t)ck * (chunk_bytes / 4), chunk_bytes, &full_bar[s]);
          }
        }
      }
      if (!ok) fault_s = 1;
    }
  } else if (warp == 9) {
    // ================= MMA issuer: the whole warp walks the block sequence (waits included) with warp-uniform
    // operands — tensor-memory base broadcast from lane 0, descriptors from kernel parameters and constants — and
    // one elected lane issues, so the MMAs go out back to back from uniform registers (a lone lane-0 thread made
    // the compiler wrap every MMA in an elect / R2UR / vote loop, about as long as the MMA itself).
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const bool issuer = nws_elect_one();
      uint32_t n_use = 0, a_ver = 0, n_blk = 0;
      bool ok = true;
      for (int tile = item0; tile < n_tiles && ok; tile += item_step) {
        for (int b = b_begin; b < b_end && ok; ++b, ++n_blk) {
          const MlpBlockDesc d = p.blk[b];
          if (d.new_a) {   // wait for the epilogue warpgroup to publish this block's A
            ok = __all_sync(0xffffffffu, nws_mbar_wait(&a_ready, a_ver & 1));
            ++a_ver;
            if (!ok) break;
          }
          nws_tc_fence_after();
          const uint32_t idesc = nws_umma_idesc_tf32(128, d.n);
          const uint32_t lbo = (uint32_t)(d.n / 8) * 128u, part = (uint32_t)d.n * kChunkK * 4u;
          const uint32_t dcol = tmem_u + kColD + (n_blk & 1) * 128;
          for (int ck = 0; ck < kEmb / kChunkK && ok; ++ck, ++n_use) {
            const int s = n_use % kSlots;
            ok = __all_sync(0xffffffffu, nws_mbar_wait(&full_bar[s], (n_use / kSlots) & 1));
            if (!ok) break;
            nws_tc_fence_after();
            const uint32_t sb = nws_smem_u32(smem + s * kSlotBytes);
            if (issuer) {
#pragma unroll
              for (int j = 0; j < kChunkK / 8; ++j) {
                const uint32_t acol = ck * kChunkK + j * 8;
                const uint64_t bh = nws_umma_smem_desc(sb + j * 2 * lbo, lbo, 128);
                const uint64_t bl = nws_umma_smem_desc(sb + part + j * 2 * lbo, lbo, 128);
                umma_tf32_ts(dcol, tmem_u + kColAhi + acol, bh, idesc, (ck | j) ? 1u : 0u);
                umma_tf32_ts(dcol, tmem_u + kColAlo + acol, bh, idesc, 1u);
                umma_tf32_ts(dcol, tmem_u + kColAhi + acol, bl, idesc, 1u);
              }
              nws_umma_commit(&empty_bar[s]);   // slot reusable once these MMAs have read it
            }
            __syncwarp();
          }
          if (ok && issuer) nws_umma_commit(&d_ready[n_blk & 1]);
          __syncwarp();
        }
      }
      if (!ok && lane == 0) fault_s = 1;
    }
  } else {
    // ================= epilogue warpgroups: thread = (frame row = TMEM lane, channel half)
    const int half = tid >> 7, r = tid & 127, ch0 = half * kHalf;
    const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    auto epi_barrier = [] { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); };
    uint32_t n_blk = 0;
    bool ok = true;
    for (int tile = item0; tile < n_tiles && ok; tile += item_step) {
      const int ub = tile / tiles_per_seg, tk = tile - ub * tiles_per_seg;
      const int row = ub * p.T + p.t_begin + tk * 128 + r;
      const bool valid = tk * 128 + r < seg;
      const float* hrow = p.h + (size_t)(valid ? row : 0) * kEmb;
      for (int b = b_begin; b < b_end && ok; ++b, ++n_blk) {
        const MlpBlockDesc d = p.blk[b];
        if (b == 0 || b == 6) {
          // A <- this frame's GRU state (the embedding projection's input), hi/lo split
#pragma unroll 1
          for (int c0 = ch0; c0 < ch0 + kHalf; c0 += 16) {
            float y[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 v = valid ? *reinterpret_cast<const float4*>(hrow + c0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
              y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
            }
            store_a16(tmem_row, c0, y);
          }
          tmem_wait_st();
          nws_tc_fence_before();
          mbar_arrive(&a_ready);
        }
        // wait for this block's accumulator
        ok = nws_mbar_wait(&d_ready[n_blk & 1], (n_blk >> 1) & 1);
        if (!ok) break;
        nws_tc_fence_after();
        const uint32_t drow = tmem_row + kColD + (n_blk & 1) * 128;
        const float* bias = vec_s + b * kVecStride;
        if (d.kind == 1) {
          // this thread's 64 channels of the row in registers: one burst of TMEM loads, then LayerNorm (dynamic.py:11-17)
          // with four independent accumulation chains; the row's other half adds its sums through shared memory
          const float* g = bias + 128;
          const float* be = bias + 256;
          float x[kHalf];
#pragma unroll
          for (int c0 = 0; c0 < kHalf; c0 += 16) nws_tmem_ld16_nowait(drow + ch0 + c0, x + c0);
          nws_tmem_wait_ld();
#pragma unroll
          for (int c0 = 0; c0 < kHalf; c0 += 16) nws_reg_fence16(x + c0);
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int c = 0; c < kHalf; c += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bias + ch0 + c);
            x[c] += bv.x; x[c + 1] += bv.y; x[c + 2] += bv.z; x[c + 3] += bv.w;
            s0 += x[c]; s1 += x[c + 1]; s2 += x[c + 2]; s3 += x[c + 3];
          }
          ln_sum[half][r] = (s0 + s1) + (s2 + s3);
          epi_barrier();
          const float mean = (ln_sum[0][r] + ln_sum[1][r]) * (1.0f / kEmb);
          float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
          for (int c = 0; c < kHalf; c += 4) {
            x[c] -= mean; x[c + 1] -= mean; x[c + 2] -= mean; x[c + 3] -= mean;
            q0 = fmaf(x[c], x[c], q0); q1 = fmaf(x[c + 1], x[c + 1], q1);
            q2 = fmaf(x[c + 2], x[c + 2], q2); q3 = fmaf(x[c + 3], x[c + 3], q3);
          }
          ln_sq[half][r] = (q0 + q1) + (q2 + q3);
          epi_barrier();
          const float rstd = 1.0f / sqrtf((ln_sq[0][r] + ln_sq[1][r]) * (1.0f / kEmb) + 1e-5f);
#pragma unroll
          for (int c0 = 0; c0 < kHalf; c0 += 16) {
            float y[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 gv = *reinterpret_cast<const float4*>(g + ch0 + c0 + i);
              const float4 bv = *reinterpret_cast<const float4*>(be + ch0 + c0 + i);
              const float t0 = fmaf(x[c0 + i] * rstd, gv.x, bv.x), t1 = fmaf(x[c0 + i + 1] * rstd, gv.y, bv.y);
              const float t2 = fmaf(x[c0 + i + 2] * rstd, gv.z, bv.z), t3 = fmaf(x[c0 + i + 3] * rstd, gv.w, bv.w);
              y[i] = t0 > 0.f ? t0 : 0.01f * t0; y[i + 1] = t1 > 0.f ? t1 : 0.01f * t1;
              y[i + 2] = t2 > 0.f ? t2 : 0.01f * t2; y[i + 3] = t3 > 0.f ? t3 : 0.01f * t3;
            }
            store_a16(tmem_row, ch0 + c0, y);
          }
          tmem_wait_st();
          nws_tc_fence_before();
          mbar_arrive(&a_ready);
        } else if (d.kind == 0) {
          float x[kHalf];
#pragma unroll
          for (int c0 = 0; c0 < kHalf; c0 += 16) nws_tmem_ld16_nowait(drow + ch0 + c0, x + c0);
          nws_tmem_wait_ld();
#pragma unroll
          for (int c0 = 0; c0 < kHalf; c0 += 16) nws_reg_fence16(x + c0);
#pragma unroll
          for (int c0 = 0; c0 < kHalf; c0 += 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[c0 + i] += bias[ch0 + c0 + i];
            store_a16(tmem_row, ch0 + c0, x + c0);
          }
          tmem_wait_st();
          nws_tc_fence_before();
          mbar_arrive(&a_ready);
        } else {
          const bool film = d.kind == 2;
          float* orow = film ? p.film + (size_t)(valid ? row : 0) * kFilm + d.col0
                             : p.bands + (size_t)(valid ? row : 0) * kBandsPad + d.col0;
          const int n_store = film ? 128 : (d.n == 128 ? 128 : kBandsPad - 128);
          // a 128-column block: each half its 64 columns; the 16-column tail of the band gains: the first half only
          const int c_lo = d.n == 128 ? ch0 : 0, c_hi = d.n == 128 ? ch0 + kHalf : (half == 0 ? d.n : 0);
#pragma unroll 1
          for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            float v[16];
            nws_tmem_ld16(drow + c0, v);
            if (valid) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (c0 + 4 * q < n_store) {
                  float4 o;
                  o.x = v[4 * q] + bias[c0 + 4 * q]; o.y = v[4 * q + 1] + bias[c0 + 4 * q + 1];
                  o.z = v[4 * q + 2] + bias[c0 + 4 * q + 2]; o.w = v[4 * q + 3] + bias[c0 + 4 * q + 3];
                  *reinterpret_cast<float4*>(orow + c0 + 4 * q) = o;
                }
              }
            }
          }
          nws_tc_fence_before();
        }
      }
    }
    if (!ok) fault_s = 1;
  }
  nws_tc_fence_before();
  __syncthreads();
  if (fault_s && fault && tid == 0) *(volatile int*)fault = 1;
  if (warp == 0) nws_tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// TC weight blob: for each weight block, chunks of 32 K-columns, each chunk = [hi part | lo part],
// each part in the canonical no-swizzle K-major layout of an [n x 32] operand.
struct MlpTcPackArgs {
  const float* w[11];   // source weights [n_rows][128] (PyTorch layout), 11 distinct blocks
  int row0[11];         // first source row of the block
  int rows_valid[11];   // rows beyond this are zero padding
  int n[11];            // block N (128 or 16)
  int off[11];          // float offset of the block in the blob
  int total;
};

__global__ void nws_mlp_tc_pack_kernel(MlpTcPackArgs a, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int b = 0;
  while (b + 1 < 11 && i >= a.off[b + 1]) ++b;
  const int n = a.n[b], j = i - a.off[b];
  const int chunk_floats = 2 * n * kChunkK;
  const int ck = j / chunk_floats, r = j % chunk_floats;
  const int part = r / (n * kChunkK), q = r % (n * kChunkK);
  const int per_kc = (n / 8) * 32;                  // floats per 16-byte k-chunk column of core matrices
  const int kc = q / per_kc, rem = q % per_kc;
  const int row = (rem / 32) * 8 + ((rem & 31) >> 2), k = ck * kChunkK + kc * 4 + (rem & 3);
  const float w = row < a.rows_valid[b] ? a.w[b][(size_t)(a.row0[b] + row) * kEmb + k] : 0.f;
  const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
  dst[i] = part == 0 ? hi : __uint_as_float(__float_as_uint(w - hi) & 0xffffe000u);
}

}  // namespace

size_t nws_mlp_tc_blob_floats() { return (size_t)10 * 2 * 128 * kEmb + (size_t)2 * 16 * kEmb; }

int nws_launch_mlp_tc_pack(NwsContext* ctx, const float* const* tensors, cudaStream_t s) {
  MlpTcPackArgs a{};
  // distinct blocks: proj, A1, A2, A3, Aout[0:128], Aout[128:256], B1, B2, B3, Bout[0:128], Bout[128:144]
  const float* src[11] = {tensors[NWS_T_PROJ_W],
                          tensors[NWS_T_FILM_MLP + 0], tensors[NWS_T_FILM_MLP + 4], tensors[NWS_T_FILM_MLP + 8],
                          tensors[NWS_T_FILM_MLP + 12], tensors[NWS_T_FILM_MLP + 12],
                          tensors[NWS_T_NOISE_MLP + 0], tensors[NWS_T_NOISE_MLP + 4], tensors[NWS_T_NOISE_MLP + 8],
                          tensors[NWS_T_NOISE_MLP + 12], tensors[NWS_T_NOISE_MLP + 12]};
  const int row0[11] = {0, 0, 0, 0, 0, 128, 0, 0, 0, 0, 128};
  const int rows_valid[11] = {128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 1};
  const int nn[11] = {128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 16};
  int off = 0;
  for (int b = 0; b < 11; ++b) {
    a.w[b] = src[b]; a.row0[b] = row0[b]; a.rows_valid[b] = rows_valid[b]; a.n[b] = nn[b]; a.off[b] = off;
    ctx->mlp_tc_off[b] = off;
    off += 2 * nn[b] * kEmb;
  }
  a.total = off;
  nws_mlp_tc_pack_kernel<<<(off + 255) / 256, 256, 0, s>>>(a, ctx->mlp_tc);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

int nws_launch_mlp_tc(const NwsContext* ctx, const float* hbuf, float* film, float* bands, int M, int T, int t_begin,
                      int t_end, cudaStream_t s) {
  MlpTcParams p{};
  const NwsPackedLayout& L = ctx->lay;
  // block -> (distinct weight block, N, kind, col0, new_a, bias, gamma, beta)
  const int wsel[kBlocks] = {0, 1, 2, 3, 4, 5, 0, 6, 7, 8, 9, 10};
  const int nn[kBlocks] = {128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 16};
  const int kind[kBlocks] = {0, 1, 1, 1, 2, 2, 0, 1, 1, 1, 3, 3};
  const int col0[kBlocks] = {0, 0, 0, 0, 0, 128, 0, 0, 0, 0, 0, 128};
  const int new_a[kBlocks] = {1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 0};
  for (int b = 0; b < kBlocks; ++b) {
    MlpBlockDesc& d = p.blk[b];
    d.w_off = ctx->mlp_tc_off[wsel[b]]; d.n = nn[b]; d.kind = kind[b]; d.col0 = col0[b]; d.new_a = new_a[b];
    d.g = d.beta = 0;
    d.n_bias = 128;
  }
  p.blk[11].n_bias = kBandsPad - 128;
  p.blk[0].bias = p.blk[6].bias = L.proj_b;
  for (int l = 0; l < 3; ++l) {
    p.blk[1 + l].bias = L.mlp[0].b[l]; p.blk[1 + l].g = L.mlp[0].g[l]; p.blk[1 + l].beta = L.mlp[0].beta[l];
    p.blk[7 + l].bias = L.mlp[1].b[l]; p.blk[7 + l].g = L.mlp[1].g[l]; p.blk[7 + l].beta = L.mlp[1].beta[l];
  }
  p.blk[4].bias = L.mlp[0].b_out; p.blk[5].bias = L.mlp[0].b_out + 128;
  p.blk[10].bias = L.mlp[1].b_out; p.blk[11].bias = L.mlp[1].b_out + 128;
  p.packed = ctx->packed; p.w_tc = ctx->mlp_tc; p.h = hbuf; p.film = film; p.bands = bands; p.M = M;
  if (t_begin == 0 && t_end == T) { p.T = M; p.t_begin = 0; p.t_end = M; }   // whole batch: one flat array of M frames
  else { p.T = T; p.t_begin = t_begin; p.t_end = t_end; }

  static bool attr_done[64] = {};
  const int smem = kSlots * kSlotBytes;
  if (nws_first_use_on_device(attr_done)) {
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const int tiles = (M / p.T) * ((p.t_end - p.t_begin + 127) / 128);
  // A tile's block sequence is two self-contained halves (proj + FiLM chain | proj + noise chain).  Handing out
  // half-tiles halves the dependent chain per CTA when there are few tiles (latency) and shortens the last wave
  // otherwise: 192 tiles (the pipelined forward's second block) are 2 rounds of whole tiles on 148 SMs but
  // 3 rounds of half tiles = 1.5.  Split whenever it takes strictly fewer tile-times.
  const int sms_even = ctx->sm_count & ~1;
  const int grid_whole = tiles < ctx->sm_count ? tiles : ctx->sm_count;
  const int grid_split = 2 * tiles < sms_even ? 2 * tiles : sms_even;
  const int rounds_whole2 = 2 * ((tiles + grid_whole - 1) / grid_whole);   // in half-tile times
  const int rounds_split2 = (2 * tiles + grid_split - 1) / grid_split;
  p.split = rounds_split2 < rounds_whole2 ? 1 : 0;
  const int grid = p.split ? grid_split : grid_whole;
  nws_mlp_tc_kernel<<<grid, kMlpThreads, smem, s>>>(p, ctx->fault_dev);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
