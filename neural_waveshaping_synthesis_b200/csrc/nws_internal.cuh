// Internal declarations shared by the .cu files of libnws_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nws_b200.h"
#include "nws_math.h"

// ------------------------------------------------------------------ fixed sizes (gin/models/newt.gin)
constexpr int kSampleRate = 16000;
constexpr int kHop = 128;
constexpr int kHarm = 101;
constexpr int kHarmPad = 104;          // harmonic axis padded to a multiple of 8 (zero weights)
constexpr int kShapers = 64;
constexpr int kEmb = 128;
constexpr int kGates = 3 * kEmb;       // GRU rows, order r,z,n (torch.nn.GRU)
constexpr int kFilm = 4 * kShapers;    // gamma_index, beta_index, gamma_norm, beta_norm
constexpr int kBands = 129;
constexpr int kBandsPad = 132;         // row stride of the noise-band / spectrum arrays (16 B multiple)
constexpr int kIr = 256;
constexpr int kReverbIr = 32000;       // [0, ir] (shaping.py:162)
constexpr int kShaperStride = 176;     // packed floats per shaper (172 used)
constexpr int kSmallMlpMaxFrames = 40;  // frames per utterance the small-batch MLP chain takes (nws_mlp_small.cu)
constexpr int kSmallMlpMaxBatch = 32;
constexpr int kReverbDirectMaxN = 4096; // buffers up to this length take the direct-form reverb (nws_reverb_direct.cu)
constexpr int kDirCounters = 4096;     // completion counters of the direct-form reverb (utterances x output blocks)

// packed per-shaper record (floats): all vector groups 16-byte aligned
constexpr int kShpScale = 0, kShpB4 = 1, kShpW1 = 4, kShpB1 = 12, kShpW2 = 20, kShpB2 = 84, kShpW3 = 92,
              kShpB3 = 156, kShpW4 = 164;

// ------------------------------------------------------------------ packed weight blob (float offsets)
struct NwsTdMlpOffsets {
  int wt[3], b[3], g[3], beta[3];  // hidden layers: wt [128][128] k-major, bias, LN gamma/beta
  int wt_out, b_out, ld_out;       // output layer: wt [128][ld_out] k-major (zero padded), bias [ld_out]
};

struct NwsPackedLayout {
  int gru_whh, gru_wih, gru_bih, gru_bhh;
  int gru_mma;                     // W_hh as fp16-split mma.m16n8k16 A fragments in per-thread register order (nws_gru_mma.cu)
  int proj_wt, proj_b;
  NwsTdMlpOffsets mlp[2];
  int hmix_wt, hmix_b;             // [kHarmPad][64] k-major, [64]
  int hmix_umma;                   // [2][kHarmPad*64]: tf32 hi / lo parts in the canonical UMMA K-major layout
  int shaper;                      // [64][kShaperStride]
  int mix_w, mix_b;
  int rand_phase;                  // [kHarmPad]
  int ir;                          // [kReverbIr]  ([0, ir])
  int total;
};

inline NwsPackedLayout nws_packed_layout() {
  NwsPackedLayout L{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  L.gru_whh = take(kGates * kEmb);
  L.gru_wih = take(kGates * 2);
  L.gru_bih = take(kGates);
  L.gru_bhh = take(kGates);
  L.gru_mma = take(kGates * kEmb);
  L.proj_wt = take(kEmb * kEmb);
  L.proj_b = take(kEmb);
  for (int m = 0; m < 2; ++m) {
    for (int l = 0; l < 3; ++l) {
      L.mlp[m].wt[l] = take(kEmb * kEmb);
      L.mlp[m].b[l] = take(kEmb);
      L.mlp[m].g[l] = take(kEmb);
      L.mlp[m].beta[l] = take(kEmb);
    }
    L.mlp[m].ld_out = m == 0 ? kFilm : kBandsPad;
    L.mlp[m].wt_out = take(kEmb * L.mlp[m].ld_out);
    L.mlp[m].b_out = take(L.mlp[m].ld_out);
  }
  L.hmix_wt = take(kHarmPad * kShapers);
  L.hmix_b = take(kShapers);
  L.hmix_umma = take(2 * kHarmPad * kShapers);
  L.shaper = take(kShapers * kShaperStride);
  L.mix_w = take(kShapers);
  L.mix_b = take(4);
  L.rand_phase = take(kHarmPad);
  L.ir = take(kReverbIr);
  L.total = o;
  return L;
}

// ------------------------------------------------------------------ reverb FFT plan
struct NwsReverbPlan {
  int n1 = 0, log_n1 = 0;      // column FFT length; L = n1 * 256 (log_n1: power-of-two lengths only)
  bool mixed = false;          // n1 = 125 or 250: mixed-radix column transforms (nws_fft_mixed.cuh)
  float2* tw_cols = nullptr;   // [n1]  exp(-2*pi*i*m/n1) (mixed plans)
  int cols_per_cta = 0;        // W
  float2* tw_big = nullptr;    // [L]   exp(-2*pi*i*n2*k1/L) at index k1*256+n2
  float2* ir_spec = nullptr;   // [L]   spectrum of [0, ir] in the four-step layout
  bool ir_valid = false;
};

constexpr int kMaxPlans = 8;
constexpr int kMaxTimeBlocks = 64;   // pipelined forward: time blocks of 128 frames (longer utterances run serially)
constexpr int kTwMaster = 4096;  // master twiddle table: W_4096^m, m < 2048

struct NwsContext {
  NwsConfig cfg;
  NwsPackedLayout lay;
  float* packed = nullptr;     // device weight blob
  bool weights_loaded = false;
  float* lut = nullptr;        // [64][table_size]
  float2* lut2 = nullptr;      // [64][table_size] pairs (T[i], T[min(i+1,size-1)] - T[i]) for the fused kernels
  int lut_size = 0;
  float lut_min = 0.f, lut_max = 0.f;
  bool lut_valid = false;
  float2* tw_master = nullptr; // [kTwMaster/2]
  NwsReverbPlan plans[kMaxPlans];
  int n_plans = 0;
  int sm_count = 148;
  int mlp_impl = 1;            // 1 = tcgen05 MLP chain (nws_mlp_tc.cu), 0 = fp32 SIMT layers (nws_encoder.cu)
  float* mlp_tc = nullptr;     // TC weight blob (hi/lo parts, canonical UMMA layout, chunked)
  int mlp_tc_off[11] = {};
  float shaper_inner_bound = 1e30f;   // max_j(|b_j| + sum_i |W_ji|) over shaper layers 2-4 (set by nws_load_weights)
  int gru_impl = 1;            // 1 = tensor-core recurrence (8 utterances per CTA, nws_gru_mma.cu) from 64 utterances on, 0 = fp32 SIMT always, 2 = tensor cores always
  bool gru_mma_ok = false;     // W_hh finite and inside the fp16 range (checked by nws_load_weights)
  int noise_fused = 0;         // whole-utterance forward: 1 = the FIR noise branch runs inside nws_audio_tc_kernel (its MMA warps), 0 = nws_noise_filter_kernel first (measured faster: DESIGN.md)
  int audio_impl = 1;          // 1 = tcgen05 harmonic mixer (nws_audio_tc.cu), 0 = fp32 SIMT (nws_audio.cu)
  int shaper_impl = 1;         // NEWT shaper layers inside nws_audio_tc_kernel: 1 = mma.sync 8x8 layers, 0 = fp32 FMA (paired lanes)
  int device = 0;
  // mbarrier-timeout flag of the tcgen05 kernels: one int in mapped pinned host memory (the kernels write it
  // only on a timeout; every API call reads the host side without a synchronise and fails with NWS_ERR_CUDA)
  int* tile_counters = nullptr; // [4] = two {tiles claimed, CTAs done} pairs of nws_audio_tc_kernel's scheduler (main / early launch)
  int* gru_done = nullptr;     // [kMaxTimeBlocks] CTAs of the tensor-core recurrence that have passed each progress mark (zeroed per forward)
  int* dir_counters = nullptr; // [kDirCounters] zero between launches (nws_reverb_direct.cu)
  int small_path = 1;          // few frames: 1 = fp32 small-batch MLP chain (nws_mlp_small.cu), 0 = always the 128-frame-tile kernel
  int reverb_direct = 1;       // short buffers: 1 = direct-form reverb, 0 = always the FFT path
  int* fault_host = nullptr;
  int* fault_dev = nullptr;
  // pipelined forward: the GRU runs in time blocks on an internal stream while the main stream renders the
  // blocks already encoded
  int pipeline = 1;
  cudaStream_t enc_stream = nullptr, aux_stream = nullptr;
  cudaEvent_t ev_early_ready = nullptr, ev_early_done = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_block[kMaxTimeBlocks] = {}, ev_mlp[kMaxTimeBlocks] = {};
  int pipe_first = 32;         // ... and of the first block (the chip idles until it is encoded)
  int pipe_block = 117;        // frames per time block when the tensor-core recurrence encodes (nws_forward)
  // optional per-stage timing of nws_forward (cudaEvents on the launch stream)
  bool profile = false;
  cudaEvent_t ev[2 * 10] = {};
  bool ev_recorded[10] = {};
};

enum { kStRng = 0, kStCarry, kStGru, kStProj, kStMlpFilm, kStMlpNoise, kStNoiseSpec, kStNoiseFilter, kStAudio, kStReverb, kStCount };

// ------------------------------------------------------------------ workspace carving
struct NwsWorkspace {
  double* carry;      // [B*T]   exclusive fp64 prefix of per-hop f0 sums
  float* u_phase;     // [kHarmPad]
  float* noise;       // [128*T]
  float* hbuf;        // [M][128] GRU states
  float* emb;         // [M][128]
  float* act0;        // [M][128]
  float* act1;        // [M][128]
  float* film;        // [M][256]
  float* bands;       // [M][kBandsPad]
  float2* xspec;      // [T][kBandsPad]
  float* dry;         // [B][N]
  float* scratch;     // [M][256]  layout conversion for the stage entry points
  int* counters;      // [kMaxTimeBlocks] tile counters of the dynamic schedulers
  float* h_state;     // [B][128] GRU state carried between time blocks
  float2* rev;        // [ceil(B/2)][L]
  size_t total;
};

NwsWorkspace nws_carve_workspace(void* base, int B, int T, int fft_len);
int nws_reverb_fft_len(int N);  // L = n1*256 >= N + kReverbIr - 1, n1 a power of two >= 128; 0 if unsupported
int nws_reverb_exact_len(int N); // max(N, kReverbIr) when the circular convolution can run at exactly that length, else 0

// ------------------------------------------------------------------ error / launch bookkeeping
void nws_set_error(const char* fmt, ...);
int nws_check_fault(const NwsContext* ctx, const char* who);
extern thread_local uint64_t g_nws_launches;

// cudaFuncSetAttribute is per device: "first use" flags are kept per device ordinal (a process may drive
// several GPUs through several handles).
inline bool nws_first_use_on_device(bool* flags) {
  int d = 0;
  cudaGetDevice(&d);
  if (flags[d & 63]) return false;
  flags[d & 63] = true;
  return true;
}

#define NWS_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      nws_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NWS_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define NWS_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    ++g_nws_launches;                                                                       \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      nws_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NWS_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// ------------------------------------------------------------------ programmatic dependent launch (short-buffer path)
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the
// stream is still running: everything before nws_pdl_wait() (weight staging, TMEM allocation, barrier set-up) overlaps
// the predecessor; nws_pdl_wait() returns when the predecessor grid has completed and its writes are visible;
// nws_pdl_launch() lets the NEXT kernel of the stream start its own prologue.  Both are no-ops for ordinary launches.
#if defined(__CUDACC__)
__device__ __forceinline__ void nws_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void nws_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
// launch configuration with (optionally) the programmatic-serialization attribute; `attr` must outlive the launch call
inline void nws_pdl_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int* n_attr, bool pdl) {
  if (pdl) {
    attr[*n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[*n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++*n_attr;
  }
  cfg->attrs = attr;
  cfg->numAttrs = *n_attr;
}

// ------------------------------------------------------------------ kernel launchers (one per .cu)
// nws_encoder.cu
int nws_launch_phase_carry(const float* f0, double* carry, int B, int T, cudaStream_t s);
int nws_launch_gru(const NwsContext* ctx, const float* control, int ctrl_channels, float* hbuf, int B, int T,
                   int t_begin, int t_end, float* h_state, cudaStream_t s);
// nws_gru_mma.cu
int nws_gru_ctas(const NwsContext* ctx, int B);   // CTAs (= SMs) the recurrence of B utterances occupies
int nws_gru_mma_ctas(int B);
int nws_launch_pack_gru_mma(NwsContext* ctx, const float* w_hh, cudaStream_t s);
// progress marks of one nws_gru_mma_kernel launch: after frame t[k] - 1 every CTA adds one to done[k]
struct NwsGruMarks { int n = 0; int t[kMaxTimeBlocks] = {}; };
int nws_launch_gru_mma(const NwsContext* ctx, const float* control, int ctrl_channels, float* hbuf, int B, int T,
                       int t_begin, int t_end, float* h_state, cudaStream_t s, int* done = nullptr,
                       const NwsGruMarks* marks = nullptr);
int nws_launch_wait_counter(const int* counter, int target, cudaStream_t s);
int nws_launch_linear(const float* X, const float* Wt, const float* bias, const float* ln_g, const float* ln_b,
                      float* Y, int M, int n_out, int ldw, int ldy, bool ln_act, cudaStream_t s);
int nws_launch_td_mlp(const NwsContext* ctx, int which, const float* emb, float* act0, float* act1, float* out,
                      int M, cudaStream_t s);
int nws_launch_bct_to_rows(const float* in, float* out, int B, int C, int T, int ld_out, cudaStream_t s);
int nws_launch_rows_to_bct(const float* in, float* out, int B, int C, int T, int ld_in, cudaStream_t s);
// nws_audio.cu
int nws_launch_rng(float* u_phase, float* noise, int n_noise, uint64_t seed, uint64_t offset, cudaStream_t s);
int nws_launch_audio(const NwsContext* ctx, const float* f0, const double* carry, const float* film,
                     const float* u_phase, const float* noise_in, float* out, float* exciter_out, int B, int T,
                     int use_lut, cudaStream_t s);
int nws_launch_audio_tc(const NwsContext* ctx, const float* f0, const double* carry, const float* film,
                        const float* u_phase, const float* noise_in, float* out, float* exciter_out, int B, int T,
                        int t_begin, int t_end, int* tile_counter, int use_lut, cudaStream_t s, int max_ctas = 0,
                        bool pdl = false, const float* bands = nullptr, const float2* xspec = nullptr);
size_t nws_mlp_tc_blob_floats();
int nws_launch_mlp_tc_pack(NwsContext* ctx, const float* const* tensors, cudaStream_t s);
int nws_launch_mlp_tc(const NwsContext* ctx, const float* hbuf, float* film, float* bands, int M, int T, int t_begin,
                      int t_end, cudaStream_t s);
int nws_launch_pair_lut(NwsContext* ctx, cudaStream_t s);
int nws_launch_build_lut(const NwsContext* ctx, const float* points, float* lut, int table_size, float tmin,
                         float tmax, cudaStream_t s);
int nws_launch_pack_weights(NwsContext* ctx, const float* const* tensors, cudaStream_t s);
// nws_noise.cu
int nws_launch_noise_spectrum(const NwsContext* ctx, const float* noise, float2* xspec, int T, cudaStream_t s);
bool nws_front_ok(int B, int T);
int nws_launch_front(const NwsContext* ctx, const float* control, int ctrl_channels, float* hbuf, const float* f0,
                     double* carry, const float* noise_in, float* u_phase_out, uint64_t seed, uint64_t offset,
                     float2* xspec, int B, int T, cudaStream_t s);
int nws_launch_noise_filter(const NwsContext* ctx, const float* bands, const float2* xspec, float* out, int B, int T,
                            int hop_begin, int hop_end, cudaStream_t s);
// nws_reverb.cu
int nws_reverb_get_plan(NwsContext* ctx, int fft_len, cudaStream_t s, NwsReverbPlan** out);
int nws_launch_reverb(NwsContext* ctx, const float* x, float* out, float2* work, int B, int N, cudaStream_t s);
void nws_reverb_free_plans(NwsContext* ctx);
void nws_reverb_invalidate(NwsContext* ctx);
int nws_make_twiddle_master(NwsContext* ctx);
// nws_mlp_small.cu
bool nws_mlp_small_ok(const NwsContext* ctx, int B, int T);
int nws_launch_mlp_small(const NwsContext* ctx, const float* hbuf, float* film, float* bands, int B, int T, cudaStream_t s,
                         const float2* xspec = nullptr, float* dry = nullptr, int hop_begin = 0, int hop_end = 0,
                         bool pdl = false);
// nws_reverb_direct.cu
size_t nws_reverb_direct_scratch_bytes(int B, int N);
bool nws_reverb_direct_ok(const NwsContext* ctx, int B, int N, size_t scratch_bytes);
int nws_launch_reverb_direct(NwsContext* ctx, const float* x, float* out, float* scratch, int B, int N, cudaStream_t s,
                             bool pdl = false);
size_t nws_reverb_direct_causal_scratch_bytes(int B, int n_new_max);
int nws_launch_reverb_direct_causal(NwsContext* ctx, const float* hist, const float* dry, size_t dry_stride, int first_sample,
                                    float* out, float* hist_next, float* scratch, int B, int n_new, int apply_reverb,
                                    cudaStream_t s, bool pdl = false);
