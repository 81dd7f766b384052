// Control-side loudness extractors (SURVEY.md §8(f) rank 4 — the step before the hot path in the timbre-transfer
// use case): extract_perceptual_loudness / compute_power_spectrogram / extract_rms of
// neural_waveshaping_synthesis/data/utils/loudness_extraction.py:11-90.
//
// extract_perceptual_loudness is  mean_k( amplitude_to_db(|stft(audio)|, ref=max, amin=eps, top_db=80) )  per frame
// (the A-weighting is computed but NOT added by the reference, :39), optionally (x + 80) / 80.  All of its
// arithmetic is librosa 0.8.0's: a centred, reflect-padded STFT with a periodic Hann window, evaluated by numpy in
// float64 and stored as complex64; float32 from there on.  The same recipe here:
//   L1  one CTA per pair of frames: window * frame in float64, both real frames through one complex float64 FFT
//       (nws_fft_f64.cuh — B200 has a real FP64 pipe), components rounded to float32, power = |.|^2 in float32
//       -> workspace [B][F][K]; running maximum per segment by atomicMax (the ref=np.max of :19-21);
//   L2  one CTA per frame: 10 log10(max(amin^2, p)) - 10 log10(max(amin^2, pmax)), floor at -80 dB below the
//       maximum (which is 0 dB by construction), mean over the K bins, normalise.
#include <math.h>

#include "nws_fft_f64.cuh"
#include "nws_internal.cuh"

namespace {

__device__ __forceinline__ int reflect_index(long long i, int N) {
  // np.pad(mode="reflect") for a pad shorter than the signal: no edge repeat
  if (i < 0) i = -i;
  if (i >= N) i = 2ll * (N - 1) - i;
  return (int)i;
}

// grid ((F + 1) / 2, B), 256 threads; dynamic smem (2 * n_fft + n_fft / 2) double2.
__global__ void __launch_bounds__(256) nws_loudness_power_kernel(const float* __restrict__ audio, int N, int log_n,
                                                                 int hop, int F, float* __restrict__ power,
                                                                 unsigned int* __restrict__ pmax_bits) {
  extern __shared__ __align__(16) double2 smem_z[];
  const int n_fft = 1 << log_n, half = n_fft >> 1, K = half + 1;
  double2* a = smem_z;
  double2* b = smem_z + n_fft;
  double2* tw = smem_z + 2 * n_fft;
  const int tid = threadIdx.x, seg = blockIdx.y, fa = 2 * blockIdx.x, fb = fa + 1;
  const float* x = audio + (size_t)seg * N;
  for (int m = tid; m < half; m += 256) {
    double s, c;
    sincospi(-2.0 * (double)m / (double)n_fft, &s, &c);
    tw[m] = make_double2(c, s);
  }
  for (int n = tid; n < n_fft; n += 256) {
    // scipy.signal.get_window("hann", n_fft, fftbins=True): 0.5 - 0.5 cos(2 pi n / n_fft), float64
    const double w = 0.5 - 0.5 * cospi(2.0 * (double)n / (double)n_fft);
    const double xa = (double)x[reflect_index((long long)fa * hop + n - half, N)];
    const double xb = fb < F ? (double)x[reflect_index((long long)fb * hop + n - half, N)] : 0.0;
    a[n] = make_double2(w * xa, w * xb);
  }
  __syncthreads();
  const double2* z = nws_fft_f64(a, b, tw, log_n, tid, 256);
  float pm = 0.f;
  for (int k = tid; k < K; k += 256) {
    const double2 zk = z[k], zc = z[(n_fft - k) & (n_fft - 1)];
    // Xa = (Z[k] + conj(Z[N-k])) / 2,  Xb = (Z[k] - conj(Z[N-k])) / (2i); stored as complex64 by librosa
    const float are = (float)(0.5 * (zk.x + zc.x)), aim = (float)(0.5 * (zk.y - zc.y));
    const float bre = (float)(0.5 * (zk.y + zc.y)), bim = (float)(-0.5 * (zk.x - zc.x));
    // np.abs(complex64) (hypotf, correctly rounded) then np.square in float32
    const float ma = (float)sqrt((double)are * are + (double)aim * aim);
    const float mb = (float)sqrt((double)bre * bre + (double)bim * bim);
    const float pa = ma * ma, pb = mb * mb;
    power[((size_t)seg * F + fa) * K + k] = pa;
    pm = fmaxf(pm, pa);
    if (fb < F) {
      power[((size_t)seg * F + fb) * K + k] = pb;
      pm = fmaxf(pm, pb);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, o));
  if ((tid & 31) == 0) atomicMax(pmax_bits + seg, __float_as_uint(pm));   // powers are >= 0: the bit pattern is monotone
}

// grid (F, B), 128 threads.  db_out (optional) [B][F][K] receives the dB spectrogram (compute_power_spectrogram).
__global__ void __launch_bounds__(128) nws_loudness_db_kernel(const float* __restrict__ power,
                                                              const unsigned int* __restrict__ pmax_bits, int K, int F,
                                                              float amin2, int normalise, float* __restrict__ out,
                                                              float* __restrict__ db_out) {
  __shared__ float part[4];
  const int f = blockIdx.x, seg = blockIdx.y, tid = threadIdx.x;
  const float ref = __uint_as_float(pmax_bits[seg]);
  const float ref_db = 10.0f * log10f(fmaxf(amin2, ref));
  const float* p = power + ((size_t)seg * F + f) * K;
  float acc = 0.f;
  for (int k = tid; k < K; k += 128) {
    float db = 10.0f * log10f(fmaxf(amin2, p[k])) - ref_db;
    db = fmaxf(db, -80.0f);   // log_spec.max() - top_db: the maximum is 0 dB (it is the reference value itself)
    if (db_out) db_out[((size_t)seg * F + f) * K + k] = db;
    acc += db;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float m = ((part[0] + part[1]) + (part[2] + part[3])) / (float)K;
    if (normalise) m = (m + 80.0f) / 80.0f;
    out[(size_t)seg * F + f] = m;
  }
}

// extract_rms (loudness_extraction.py:71-90): zero-padded centred frames, sqrt(mean(x^2)).  grid (F, B), 128 threads.
__global__ void __launch_bounds__(128) nws_rms_kernel(const float* __restrict__ audio, int N, int window, int hop, int F,
                                                      float* __restrict__ out) {
  __shared__ double part[4];
  const int f = blockIdx.x, seg = blockIdx.y, tid = threadIdx.x;
  const float* x = audio + (size_t)seg * N;
  const long long start = (long long)f * hop - window / 2;
  double acc = 0.0;
  for (int n = tid; n < window; n += 128) {
    const long long i = start + n;
    if (i >= 0 && i < N) {
      const float v = x[i], sq = v * v;   // frames ** 2 is float32 in the reference
      acc += (double)sq;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) out[(size_t)seg * F + f] = sqrtf((float)(((part[0] + part[1]) + (part[2] + part[3])) / (double)window));
}

int check_stft_args(const char* who, int B, int N, int n_fft, int hop, int* log_n) {
  int l = 0;
  while ((1 << l) < n_fft) ++l;
  if (n_fft < 64 || n_fft > 4096 || (1 << l) != n_fft) {
    nws_set_error("%s: n_fft = %d unsupported (powers of two from 64 to 4096)", who, n_fft);
    return NWS_ERR_UNSUPPORTED;
  }
  if (B < 1 || B > 65535 || hop < 1) {   // B is the grid's y dimension
    nws_set_error("%s: bad shape (B = %d, 1..65535; hop_length = %d)", who, B, hop);
    return NWS_ERR_INVALID;
  }
  if (N <= n_fft / 2) {
    nws_set_error("%s: %d samples are too few for the reflect padding of n_fft = %d (need > n_fft / 2)", who, N, n_fft);
    return NWS_ERR_INVALID;
  }
  *log_n = l;
  return NWS_OK;
}

}  // namespace

extern "C" size_t nws_loudness_workspace_bytes(int B, int N, int n_fft, int hop_length) {
  if (B < 1 || N < 1 || n_fft < 2 || hop_length < 1) return 0;
  const size_t F = 1 + (size_t)N / hop_length, K = n_fft / 2 + 1;
  return ((size_t)B * F * K + (size_t)B + 64) * sizeof(float);
}

extern "C" int nws_extract_loudness(const float* audio, int B, int N, int n_fft, int hop_length, double epsilon,
                                    int normalise, float* loudness_out, float* db_out, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!audio || !loudness_out || !workspace) { nws_set_error("nws_extract_loudness: NULL argument"); return NWS_ERR_INVALID; }
  int log_n = 0;
  const int rc = check_stft_args("nws_extract_loudness", B, N, n_fft, hop_length, &log_n);
  if (rc) return rc;
  if (!(epsilon > 0.0)) { nws_set_error("nws_extract_loudness: epsilon must be positive"); return NWS_ERR_INVALID; }
  const size_t need = nws_loudness_workspace_bytes(B, N, n_fft, hop_length);
  if (workspace_bytes < need) {
    nws_set_error("nws_extract_loudness: workspace too small (%zu < %zu)", workspace_bytes, need);
    return NWS_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int F = 1 + N / hop_length, K = n_fft / 2 + 1;
  float* power = (float*)workspace;
  unsigned int* pmax = (unsigned int*)(power + (size_t)B * F * K);
  NWS_CUDA_OK(cudaMemsetAsync(pmax, 0, (size_t)B * sizeof(unsigned int), s));
  const size_t smem = ((size_t)2 * n_fft + n_fft / 2) * sizeof(double2);
  static bool attr_done[64] = {};
  if (nws_first_use_on_device(attr_done))
    NWS_CUDA_OK(cudaFuncSetAttribute(nws_loudness_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  nws_loudness_power_kernel<<<dim3((F + 1) / 2, B), 256, smem, s>>>(audio, N, log_n, hop_length, F, power, pmax);
  NWS_LAUNCH_CHECK();
  const float amin2 = (float)(epsilon * epsilon);   // amin ** 2 (a Python float) against a float32 array
  nws_loudness_db_kernel<<<dim3(F, B), 128, 0, s>>>(power, pmax, K, F, amin2, normalise, loudness_out, db_out);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

extern "C" int nws_extract_rms(const float* audio, int B, int N, int window_size, int hop_length, float* rms_out,
                               void* stream) {
  if (!audio || !rms_out) { nws_set_error("nws_extract_rms: NULL argument"); return NWS_ERR_INVALID; }
  if (B < 1 || B > 65535 || N < 1 || window_size < 1 || hop_length < 1) { nws_set_error("nws_extract_rms: bad shape"); return NWS_ERR_INVALID; }
  const int F = 1 + (N + 2 * (window_size / 2) - window_size) / hop_length;   // librosa.util.frame over the padded signal
  if (F < 1) { nws_set_error("nws_extract_rms: signal shorter than one window"); return NWS_ERR_INVALID; }
  nws_rms_kernel<<<dim3(F, B), 128, 0, (cudaStream_t)stream>>>(audio, N, window_size, hop_length, F, rms_out);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}

// ------------------------------------------------------------------------------------------------
// Frame rate -> sample rate, linear (data/utils/upsampling.py:20-36: np.interp of the frame values over
// np.linspace(0, F-1, P), P = F*hop + window - hop, then [window/2 : window/2 + original_length]).  float64 like numpy:
// target x_i = i * step (step = (F-1)/(P-1); the last point is exactly F-1), j = floor(x), slope_j = y[j+1] - y[j]
// (the source axis is the integers, so the division by x[j+1] - x[j] is by 1), out = slope_j * (x - j) + y[j] with
// separate roundings.  One thread per output sample.
__global__ void nws_interp_frames_kernel(const float* __restrict__ frames, int F, long long P, int skip, int out_len,
                                         double* __restrict__ out) {
  const int b = blockIdx.y;
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out_len) return;
  const long long i = o + skip;
  const float* y = frames + (size_t)b * F;
  double r;
  if (F == 1) {
    r = (double)y[0];
  } else {
    const double step = (double)(F - 1) / (double)(P - 1);
    const double x = i == P - 1 ? (double)(F - 1) : __dmul_rn((double)i, step);
    int j = (int)x;
    if (j >= F - 1) {
      r = (double)y[F - 1];
    } else {
      const double y0 = (double)y[j], slope = __dsub_rn((double)y[j + 1], y0);
      r = x == (double)j ? y0 : __dadd_rn(__dmul_rn(slope, __dsub_rn(x, (double)j)), y0);
    }
  }
  out[(size_t)b * out_len + o] = r;
}

extern "C" int nws_interp_frames_len(int F, int window_length, int hop_length, int original_length) {
  if (F < 1 || window_length < 1 || hop_length < 1 || original_length < 0) return 0;
  const long long P = (long long)F * hop_length + window_length - hop_length;
  if (P < 1 || P > 0x7fffffffLL) return 0;
  if (!original_length) return (int)P;
  const long long rest = P - window_length / 2;
  return (int)(rest < 0 ? 0 : (rest < original_length ? rest : original_length));
}

extern "C" int nws_interp_frames(const float* frames, int B, int F, int window_length, int hop_length,
                                 int original_length, double* out, void* stream) {
  if (!frames || !out) { nws_set_error("nws_interp_frames: NULL argument"); return NWS_ERR_INVALID; }
  if (B < 1 || B > 65535) { nws_set_error("nws_interp_frames: bad batch size %d", B); return NWS_ERR_INVALID; }
  const int out_len = nws_interp_frames_len(F, window_length, hop_length, original_length);
  if (out_len < 1) { nws_set_error("nws_interp_frames: bad shape (F %d, window %d, hop %d, original_length %d)", F, window_length, hop_length, original_length); return NWS_ERR_INVALID; }
  const long long P = (long long)F * hop_length + window_length - hop_length;
  const int skip = original_length ? window_length / 2 : 0;
  nws_interp_frames_kernel<<<dim3((out_len + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(frames, F, P, skip, out_len, out);
  NWS_LAUNCH_CHECK();
  return NWS_OK;
}
