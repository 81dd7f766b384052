// Self-test of the tcgen05 plumbing in nws_tc.cuh: D[128 x 64] = A[128 x K] . B[64 x K]^T with the
// 3xTF32 split, operands written by threads in the canonical no-swizzle K-major layout, accumulator
// in TMEM.  Exposed through the C ABI so tests/test_gpu_parity.py can check it against fp64.
#include "nws_internal.cuh"
#include "nws_tc.cuh"

__global__ void __launch_bounds__(128) nws_selftest_umma_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                float* __restrict__ D, int K, int swap, int* status) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool raw_lo = (swap & 2) != 0;   // A's low part stored unmasked: the tensor core must ignore the 13 low bits
  swap &= 1;
  const uint32_t a_bytes = 128 * K * 4, b_bytes = 64 * K * 4;
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;

  for (int k = 0; k < K; ++k) {
    const float a = A[tid * K + k], h = nws_tf32_hi(a);
    *reinterpret_cast<float*>(a_hi + nws_umma_offset(tid, k, 128)) = h;
    *reinterpret_cast<float*>(a_lo + nws_umma_offset(tid, k, 128)) = raw_lo ? a - h : nws_tf32_lo(a, h);
    if (tid < 64) {
      const float b = B[tid * K + k], hb = nws_tf32_hi(b);
      *reinterpret_cast<float*>(b_hi + nws_umma_offset(tid, k, 64)) = hb;
      *reinterpret_cast<float*>(b_lo + nws_umma_offset(tid, k, 64)) = nws_tf32_lo(b, hb);
    }
  }
  if (warp == 0) nws_tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) {
    nws_mbar_init(&bar, 1);
    nws_fence_mbar_init();
  }
  nws_fence_proxy_async();
  nws_tc_fence_before();
  __syncthreads();
  nws_tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = nws_umma_idesc_tf32(128, 64);
    const uint32_t lbo_a = 16 * 128, lbo_b = 8 * 128, sbo = 128;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t oa = ks * 2 * lbo_a, ob = ks * 2 * lbo_b;
      const uint64_t dah = swap ? nws_umma_smem_desc(nws_smem_u32(a_hi) + oa, sbo, lbo_a) : nws_umma_smem_desc(nws_smem_u32(a_hi) + oa, lbo_a, sbo);
      const uint64_t dal = swap ? nws_umma_smem_desc(nws_smem_u32(a_lo) + oa, sbo, lbo_a) : nws_umma_smem_desc(nws_smem_u32(a_lo) + oa, lbo_a, sbo);
      const uint64_t dbh = swap ? nws_umma_smem_desc(nws_smem_u32(b_hi) + ob, sbo, lbo_b) : nws_umma_smem_desc(nws_smem_u32(b_hi) + ob, lbo_b, sbo);
      const uint64_t dbl = swap ? nws_umma_smem_desc(nws_smem_u32(b_lo) + ob, sbo, lbo_b) : nws_umma_smem_desc(nws_smem_u32(b_lo) + ob, lbo_b, sbo);
      nws_umma_tf32(tmem, dah, dbh, idesc, ks > 0 ? 1u : 0u);
      nws_umma_tf32(tmem, dal, dbh, idesc, 1u);
      nws_umma_tf32(tmem, dah, dbl, idesc, 1u);
    }
    nws_umma_commit(&bar);
  }
  const bool ok = nws_mbar_wait(&bar, 0, 1u << 22);
  nws_tc_fence_after();
  if (ok) {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float v[16];
      nws_tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) D[tid * 64 + c0 + i] = v[i];
    }
  }
  if (tid == 0) status[0] = ok ? 1 : -1;
  nws_tc_fence_before();
  __syncthreads();
  if (warp == 0) nws_tmem_dealloc(tmem, 64);
}

extern "C" int nws_selftest_umma(const float* A, const float* B, float* D, int K, int swap, int* status, void* stream) {
  if (!A || !B || !D || !status || K < 8 || K > 104 || (K & 7)) { nws_set_error("nws_selftest_umma: bad argument"); return NWS_ERR_INVALID; }
  const size_t smem = (size_t)(128 + 64) * K * 4 * 2;
  NWS_CUDA_OK(cudaFuncSetAttribute(nws_selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nws_selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, swap, status);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// Accuracy probe of the two device sine implementations (nws_math.h) on caller-chosen arguments.
__global__ void nws_selftest_sin_kernel(const float* __restrict__ x, float* __restrict__ y_acc, float* __restrict__ y_fast3,
                                        float* __restrict__ y_fast2, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y_acc[i] = nws_sinf(x[i]);
  y_fast3[i] = nws_sinf_fast<3>(x[i]);
  y_fast2[i] = nws_sinf_turn(x[i]);
}

extern "C" int nws_selftest_sin(const float* x, float* y_acc, float* y_fast3, float* y_fast2, long long n, void* stream) {
  if (!x || !y_acc || !y_fast3 || !y_fast2 || n < 1) { nws_set_error("nws_selftest_sin: bad argument"); return NWS_ERR_INVALID; }
  nws_selftest_sin_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y_acc, y_fast3, y_fast2, n);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 FMA issue-rate probe: the denominator of the fused audio kernel's compute roofline (SURVEY.md §8(d):
// "measure an FFMA-peak microbenchmark on the box").  Every thread runs `iters` rounds of eight independent
// fmaf chains; flops = grid * block * iters * 8 * 2.  bench.py times it with CUDA events.
__global__ void __launch_bounds__(256) nws_ffma_peak_kernel(float* __restrict__ out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" int nws_selftest_ffma_peak(float* out, int n_blocks, int iters, double* flops_out, void* stream) {
  if (!out || n_blocks < 1 || iters < 1) { nws_set_error("nws_selftest_ffma_peak: bad argument"); return NWS_ERR_INVALID; }
  nws_ffma_peak_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 1e-3f);
  NWS_LAUNCH_CHECK();
  if (flops_out) *flops_out = (double)n_blocks * 256.0 * (double)iters * 16.0 * 8.0 * 2.0;
  return NWS_OK;
}
