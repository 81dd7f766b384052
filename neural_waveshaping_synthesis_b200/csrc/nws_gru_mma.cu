// The GRU recurrence (neural_waveshaping.py:17-26, torch.nn.GRU(2 -> 128), gate order r,z,n) on the tensor cores.
//
// The recurrent product W_hh[384 x 128] . h is a matrix-vector product per utterance, but the same W_hh serves
// every utterance: with eight utterances per CTA it is a [384 x 128] x [128 x 8] contraction per step, and that
// is what this kernel hands to the tensor pipe (mma.sync m16n8k16, fp32 accumulation).  fp32 parity is kept by a
// two-way fp16 split of both operands,
//     w = w1 + w2 / 2048,  w1 = fp16(w), w2 = fp16((w - w1) * 2048)        (same for h, |h| < 1)
//     W.h ~= W1.h1 + (W1.h2 + W2.h1) / 2048                                 (22 significand bits per operand)
// whose products are exact in fp32; the dropped W2.h2 term is 2^-22 relative, the size of an fp32 rounding
// (emulated on the CPU against the fp32 recurrence before this was built: the two differ from a float64
// dot-product GRU by the same amount, 3e-5 in the vn checkpoint's embedding after 500 steps).
//
// Layout.  256 threads = 8 warps; warp w owns hidden units 16w..16w+15 and holds THEIR r, z and n rows of W_hh —
// three 16-row A tiles x 8 k-tiles x {w1, w2} = 192 registers per thread, resident for all T steps (the packed
// fragments come from nws_pack_gru_mma_kernel, four registers per coalesced 16-byte load).  The accumulator
// fragment of a thread is then (units g, g+8) x (utterances 2c, 2c+1) for all three gates: r, z, n and the
// thread's own four h values meet in registers, the gate arithmetic needs no exchange at all, and the only
// communication per step is the new h, written as fp16 pairs straight into the B-fragment layout in shared
// memory (double buffered: ONE __syncthreads per step).  The k axis is permuted so that the pair (g, g+8) a
// thread produces is exactly one 32-bit B register of the reader ((2c', 2c'+1) of k-tile w), and a reader's
// four registers of a k-tile ({b0, b1} x {h1, h2}) are one conflict-free LDS.128.
//
// Per step and CTA: 72 HMMA per warp (576 per SM) against 64 packed FMAs + 31 shuffles per thread for ONE
// utterance in the fp32 kernel (nws_encoder.cu / nws_hop_bodies.cuh).  Measured on B200 (profiles/r2_ncu_gru_mma.txt):
// 1.18 us per step — the legacy HMMA pipe, one m16n8k16 per 8.5 cycles and scheduler, is 53 % of it, the gate
// arithmetic (four cells per thread) the rest — against 0.70 us for the fp32 kernel, which needs one SM per utterance
// where this one needs one per eight.  So the fp32 kernel stays for batches that do not fill the chip (fewer than 64
// utterances: nws_gru_uses_mma), inside the fused short-buffer front end and the streaming path, and as the
// cross-check (nws_set_gru_impl(0)); from 64 utterances on this kernel encodes, in ONE launch that publishes its
// progress (marks) so that the rest of the forward can render the frames already encoded (nws_forward).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "nws_hop_bodies.cuh"
#include "nws_internal.cuh"

constexpr int kGruMmaUtts = 8;            // utterances per CTA = N of the MMA
constexpr int kGruMmaThreads = 256;
constexpr int kGruHsRow = 144;            // words per utterance row of the h buffer: 8 k-tiles x 16 words + 16 (bank offset of odd rows)
constexpr float kGruSplitScale = 2048.0f; // 2^11: the fp16 residual is stored scaled so that it stays normal

// ---------------------------------------------------------------------------------------------- load-time packing
// word q of the region: q = ((reg / 4) * 256 + thread) * 4 + (reg & 3), reg = ((gate * 8 + kt) * 2 + split) * 4 + a;
// A fragment a0..a3 of mma.m16n8k16 (row g / g+8, k positions 2c..2c+1 / +8), k position p of tile kt <-> hidden
// unit 16 kt + ((p & 7) >> 1) + 4 (p >> 3) + 8 (p & 1).
__global__ void nws_pack_gru_mma_kernel(const float* __restrict__ w_hh, uint32_t* __restrict__ dst) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= kGates * kEmb) return;
  const int a_lo = q & 3, thread = (q >> 2) & 255, rq = q >> 10;
  const int reg = rq * 4 + a_lo;
  const int a = reg & 3, split = (reg >> 2) & 1, kt = (reg >> 3) & 7, gate = reg >> 6;
  const int w = thread >> 5, lane = thread & 31, g = lane >> 2, c = lane & 3;
  const int row = gate * kEmb + 16 * w + g + 8 * (a & 1);
  const int col0 = 16 * kt + c + 4 * (a >> 1), col1 = col0 + 8;
  float v[2] = {w_hh[(size_t)row * kEmb + col0], w_hh[(size_t)row * kEmb + col1]};
  __half h[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const __half w1 = __float2half_rn(v[e]);
    h[e] = split == 0 ? w1 : __float2half_rn((v[e] - __half2float(w1)) * kGruSplitScale);
  }
  dst[q] = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
}

int nws_launch_pack_gru_mma(NwsContext* ctx, const float* w_hh, cudaStream_t s) {
  nws_pack_gru_mma_kernel<<<(kGates * kEmb + 255) / 256, 256, 0, s>>>(w_hh, reinterpret_cast<uint32_t*>(ctx->packed + ctx->lay.gru_mma));
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ---------------------------------------------------------------------------------------------- the recurrence
__device__ __forceinline__ void nws_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// h (two values: units g and g+8 of one utterance) -> the two fp16 words {h1 pair, h2 pair}
__device__ __forceinline__ uint2 nws_gru_split_pair(float ha, float hb) {
  const __half2 h1 = __floats2half2_rn(ha, hb);
  const float2 f = __half22float2(h1);
  const __half2 h2 = __floats2half2_rn((ha - f.x) * kGruSplitScale, (hb - f.y) * kGruSplitScale);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&h1), *reinterpret_cast<const uint32_t*>(&h2));
}

// Gate activations.  ACT 0: expf / IEEE division / tanhf as the fp32 kernel.  ACT 1: the sigmoids through ex2.approx and
// rcp.approx — the logistic's slope (<= 1/4) attenuates the 2-ulp error of the exponential, so the absolute error stays
// at 1.5e-7, the size of an fp32 rounding of the result — and tanhf.  ACT 2: tanh through the same exponential as well
// (3e-7 absolute; development only).
template <int ACT>
__device__ __forceinline__ float nws_gru_sigmoid(float x) {
  if (ACT == 0) return nws_sigmoid(x);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
template <int ACT>
__device__ __forceinline__ float nws_gru_tanh(float x) {
  if (ACT < 2) return tanhf(x);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(-2.0f, r, 1.0f);
}

template <int ACT>
__global__ void __launch_bounds__(kGruMmaThreads, 1)
nws_gru_mma_kernel(const uint4* __restrict__ wfrag, const float* __restrict__ w_ih, const float* __restrict__ b_ih,
                   const float* __restrict__ b_hh, const float* __restrict__ control, int ctrl_channels,
                   float* __restrict__ hbuf, int B, int T, int t_begin, int t_end, float* __restrict__ h_state,
                   int* __restrict__ done, NwsGruMarks marks) {
  __shared__ __align__(16) uint32_t hs[2][kGruMmaUtts * kGruHsRow];
  __shared__ float4 cst[3][kEmb];   // (w_ih[row][0], w_ih[row][1], b_ih[row], b_hh[row]) per gate row
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const int b0 = blockIdx.x * kGruMmaUtts;

  uint32_t wa[3][8][2][4];
#pragma unroll
  for (int rq = 0; rq < 48; ++rq) {
    const uint4 v = wfrag[rq * kGruMmaThreads + tid];
    uint32_t* dst = &wa[0][0][0][0] + rq * 4;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  for (int i = tid; i < kGates; i += kGruMmaThreads)
    cst[i >> 7][i & 127] = make_float4(w_ih[2 * i], w_ih[2 * i + 1], b_ih[i], b_hh[i]);

  // this thread's (unit, utterance) cells: unit u_i = 16 w + g + 8 i, utterance b_j = b0 + 2 c + j; cell index 2 i + j
  // is the accumulator fragment's register index
  const int u0 = 16 * w + g;
  float h[4];
  // utterance j = 1 is the one after j = 0: one pointer each for the inputs and the outputs, uniform strides
  const int bj0 = b0 + 2 * c;
  const bool live[2] = {bj0 < B, bj0 + 1 < B};
  const int bb = live[0] ? bj0 : B - 1;   // (keeps the addresses valid; never read or written when !live)
  const float* xp = control + (size_t)bb * ctrl_channels * T;
  float* hp = hbuf + (size_t)bb * T * kEmb + u0;
  const int xstride = ctrl_channels * T, hstride = T * kEmb;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i)
      h[2 * i + j] = (t_begin > 0 && live[j]) ? h_state[(size_t)(bb + j) * kEmb + u0 + 8 * i] : 0.0f;
  // word of (utterance n, k-tile w, reader lane c' = g & 3, b = g >> 2): {h1 pair, h2 pair} as one 8-byte store
  const int wr_word = w * 16 + (g & 3) * 4 + (g >> 2) * 2;
  const int rd_word = g * kGruHsRow + c * 4;
#pragma unroll
  for (int j = 0; j < 2; ++j)
    *reinterpret_cast<uint2*>(&hs[t_begin & 1][(2 * c + j) * kGruHsRow + wr_word]) = nws_gru_split_pair(h[j], h[2 + j]);
  float x[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    x[j][0] = live[j] && t_begin < t_end ? xp[j * xstride + t_begin] : 0.0f;
    x[j][1] = live[j] && t_begin < t_end ? xp[j * xstride + T + t_begin] : 0.0f;
  }
  __syncthreads();
  int mark = 0, mark_t = marks.n > 0 ? marks.t[0] : -1;

  for (int t = t_begin; t < t_end; ++t) {
    const uint32_t* hcur = hs[t & 1];
    float nx[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {   // next step's inputs (a step is far longer than the load)
      const bool ok = live[j] && t + 1 < t_end;
      nx[j][0] = ok ? xp[j * xstride + t + 1] : 0.0f;
      nx[j][1] = ok ? xp[j * xstride + T + t + 1] : 0.0f;
    }
    float acc0[3][4], acc1[3][4];
#pragma unroll
    for (int G = 0; G < 3; ++G)
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc0[G][e] = 0.0f; acc1[G][e] = 0.0f; }
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      const uint4 bf = *reinterpret_cast<const uint4*>(hcur + rd_word + kt * 16);   // {b0 h1, b0 h2, b1 h1, b1 h2}
#pragma unroll
      for (int G = 0; G < 3; ++G) {
        nws_mma16816(acc0[G], wa[G][kt][0], bf.x, bf.z);   // W1 . h1
        nws_mma16816(acc1[G], wa[G][kt][0], bf.y, bf.w);   // W1 . h2
        nws_mma16816(acc1[G], wa[G][kt][1], bf.x, bf.z);   // W2 . h1
      }
    }
    // gates (A.6): r, z, n of a cell are registers of this thread
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float4 kr = cst[0][u0 + 8 * i], kz = cst[1][u0 + 8 * i], kn = cst[2][u0 + 8 * i];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int e = 2 * i + j;
        const float ghr = fmaf(acc1[0][e], 1.0f / kGruSplitScale, acc0[0][e]) + kr.w;
        const float ghz = fmaf(acc1[1][e], 1.0f / kGruSplitScale, acc0[1][e]) + kz.w;
        const float ghn = fmaf(acc1[2][e], 1.0f / kGruSplitScale, acc0[2][e]) + kn.w;
        const float gir = fmaf(kr.y, x[j][1], fmaf(kr.x, x[j][0], kr.z));
        const float giz = fmaf(kz.y, x[j][1], fmaf(kz.x, x[j][0], kz.z));
        const float gin = fmaf(kn.y, x[j][1], fmaf(kn.x, x[j][0], kn.z));
        const float rg = nws_gru_sigmoid<ACT>(gir + ghr);
        const float zg = nws_gru_sigmoid<ACT>(giz + ghz);
        const float ng = nws_gru_tanh<ACT>(fmaf(rg, ghn, gin));
        h[e] = fmaf(zg, h[e] - ng, ng);   // (1-z)*n + z*h
      }
    }
    uint32_t* hnxt = hs[(t + 1) & 1];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      *reinterpret_cast<uint2*>(&hnxt[(2 * c + j) * kGruHsRow + wr_word]) = nws_gru_split_pair(h[j], h[2 + j]);
      if (live[j]) {
        hp[(size_t)j * hstride + (size_t)t * kEmb] = h[j];
        hp[(size_t)j * hstride + (size_t)t * kEmb + 8] = h[2 + j];
      }
      x[j][0] = nx[j][0]; x[j][1] = nx[j][1];
    }
    __syncthreads();
    if (t + 1 == mark_t) {
      // frames [0, mark_t) of this CTA's utterances are in hbuf: every thread's stores precede the barrier, the fence
      // makes them visible device-wide before the count (nws_wait_counter_kernel acquires it)
      if (tid == 0) { __threadfence(); atomicAdd(done + mark, 1); }
      ++mark;
      mark_t = mark < marks.n ? marks.t[mark] : -1;
    }
  }
  if (h_state) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (live[j]) {
        h_state[(size_t)(bj0 + j) * kEmb + u0] = h[j];
        h_state[(size_t)(bj0 + j) * kEmb + u0 + 8] = h[2 + j];
      }
  }
}

int nws_gru_mma_ctas(int B) { return (B + kGruMmaUtts - 1) / kGruMmaUtts; }

// Stream-ordered wait for a device-side count (the encoder's progress marks): one thread polls with acquire loads.
__global__ void nws_wait_counter_kernel(const int* counter, int target) {
  int v;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v < target) __nanosleep(200);
  } while (v < target);
}

int nws_launch_wait_counter(const int* counter, int target, cudaStream_t s) {
  nws_wait_counter_kernel<<<1, 1, 0, s>>>(counter, target);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

int nws_launch_gru_mma(const NwsContext* ctx, const float* control, int ctrl_channels, float* hbuf, int B, int T,
                       int t_begin, int t_end, float* h_state, cudaStream_t s, int* done, const NwsGruMarks* marks) {
  const float* p = ctx->packed;
  static const int act = getenv("NWS_GRU_ACT") ? atoi(getenv("NWS_GRU_ACT")) : 1;   // development switch
  auto kern = act == 0 ? nws_gru_mma_kernel<0> : (act == 2 ? nws_gru_mma_kernel<2> : nws_gru_mma_kernel<1>);
  kern<<<nws_gru_mma_ctas(B), kGruMmaThreads, 0, s>>>(
      reinterpret_cast<const uint4*>(p + ctx->lay.gru_mma), p + ctx->lay.gru_wih, p + ctx->lay.gru_bih, p + ctx->lay.gru_bhh,
      control, ctrl_channels, hbuf, B, T, t_begin, t_end, h_state, done, marks ? *marks : NwsGruMarks{});
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
