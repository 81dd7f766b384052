/* nws_b200.h — C ABI of the B200-native Neural Waveshaping Synthesis forward pass.
 *
 * The reference (ben-hayes/neural-waveshaping-synthesis) is pure Python/PyTorch and has no
 * FFI of its own; its boundary for this path is the Python module API
 *     NeuralWaveshaping.forward(f0, control)   neural_waveshaping_synthesis/models/neural_waveshaping.py:74-90
 * and the sub-modules it calls.  Each entry point below names the reference interface it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer to fp32 (unless stated), caller-owned, borrowed for
 *     the duration of the call; nothing is retained except by nws_load_weights/nws_set_lut,
 *     which copy;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; all work is enqueued on it and the
 *     calls do not synchronise the host unless stated;
 *   - return value 0 = NWS_OK, negative = error (nws_last_error() gives a thread-local message);
 *   - a handle may be used from one thread and on one stream at a time: its internal streams, events, scheduler
 *     counters and the packed weights are per handle, so a forward must not overlap another forward or a
 *     nws_load_weights / nws_set_lut of the same handle issued on a different stream (create one handle per stream);
 *   - the tensor-core kernels wait on their mbarriers with a bound: if one ever gives up it writes NaNs and raises a
 *     sticky flag that every later entry point of the handle reports as NWS_ERR_CUDA (see nws_status);
 *   - limits: T >= 2 frames (the reference's reflect padding has the same limit) and 128*T + 31999 <= 2^20 samples
 *     (about 63.5 s at 16 kHz: the largest reverb transform built); longer inputs return NWS_ERR_UNSUPPORTED — split
 *     them or use the streaming entry points (nws_stream_*), whose reverb is causal.
 */
#ifndef NWS_B200_H_
#define NWS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NWS_API_VERSION 1

enum {
  NWS_OK = 0,
  NWS_ERR_INVALID = -1,      /* bad argument (NULL, size, T < 2, ...)                              */
  NWS_ERR_UNSUPPORTED = -2,  /* configuration outside what the kernels are built for               */
  NWS_ERR_STATE = -3,        /* weights / LUT not loaded                                            */
  NWS_ERR_CUDA = -4,         /* CUDA runtime error                                                  */
  NWS_ERR_WORKSPACE = -5     /* workspace too small                                                 */
};

/* gin/models/newt.gin:1-33 — every hyper-parameter of the path.  The kernels are specialised
 * for the values in the comments; nws_create rejects anything else with NWS_ERR_UNSUPPORTED.   */
typedef struct NwsConfig {
  int sample_rate;        /* 16000  newt.gin:1                                     */
  int control_hop;        /* 128    newt.gin:5                                     */
  int n_harmonics;        /* 101    newt.gin:7                                     */
  int n_waveshapers;      /* 64     newt.gin:4                                     */
  int embedding_size;     /* 128    newt.gin:3,16-18 (GRU hidden == embedding)     */
  int shaping_fn_size;    /* 8      newt.gin:12                                    */
  int shaping_fn_depth;   /* 4      newt.gin:14                                    */
  int noise_bands;        /* 129    newt.gin:22                                    */
  int ir_length;          /* 256    newt.gin:24                                    */
  int reverb_length;      /* 32000  newt.gin:27-28 (length_in_seconds * sr)        */
} NwsConfig;

typedef struct NwsContext* NwsHandle;

/* Order of the fp32 tensors nws_load_weights expects == state_dict() order of the reference
 * model, native PyTorch layouts.                                           */
enum NwsTensor {
  NWS_T_GRU_W_IH = 0,     /* embedding.gru.weight_ih_l0          [384,2]    */
  NWS_T_GRU_W_HH,         /* embedding.gru.weight_hh_l0          [384,128]  */
  NWS_T_GRU_B_IH,         /* embedding.gru.bias_ih_l0            [384]      */
  NWS_T_GRU_B_HH,         /* embedding.gru.bias_hh_l0            [384]      */
  NWS_T_PROJ_W,           /* embedding.proj.weight               [128,128,1]*/
  NWS_T_PROJ_B,           /* embedding.proj.bias                 [128]      */
  NWS_T_OSC_RAND_PHASE,   /* osc.rand_phase                      [1,101,1]  */
  NWS_T_HMIX_W,           /* harmonic_mixer.weight               [64,101,1] */
  NWS_T_HMIX_B,           /* harmonic_mixer.bias                 [64]       */
  NWS_T_FILM_MLP,         /* newt.mlp.net.{0,1,3,4,6,7,9}: 14 tensors in state-dict order:
                             net.0.weight, net.0.bias, net.1.layer_norm.weight, net.1.layer_norm.bias,
                             net.3.*, net.4.layer_norm.*, net.6.*, net.7.layer_norm.*, net.9.weight [256,128,1], net.9.bias */
  NWS_T_SHAPER_SCALE = NWS_T_FILM_MLP + 14, /* newt.shaping_fn.input_scale [1,64,1]  */
  NWS_T_SHAPER_W1,        /* newt.shaping_fn.net.0.weight        [512,1,1]  */
  NWS_T_SHAPER_B1,        /* newt.shaping_fn.net.0.bias          [512]      */
  NWS_T_SHAPER_W2,        /* newt.shaping_fn.net.2.weight        [512,8,1]  */
  NWS_T_SHAPER_B2,
  NWS_T_SHAPER_W3,        /* newt.shaping_fn.net.4.weight        [512,8,1]  */
  NWS_T_SHAPER_B3,
  NWS_T_SHAPER_W4,        /* newt.shaping_fn.net.6.weight        [64,8,1]   */
  NWS_T_SHAPER_B4,        /* newt.shaping_fn.net.6.bias          [64]       */
  NWS_T_MIX_W,            /* newt.mixer.0.weight                 [1,64,1]   */
  NWS_T_MIX_B,            /* newt.mixer.0.bias                   [1]        */
  NWS_T_NOISE_MLP,        /* h_generator.net.*: 14 tensors, same order as NWS_T_FILM_MLP, net.9 [129,128,1] */
  NWS_T_NOISE_WINDOW = NWS_T_NOISE_MLP + 14, /* noise_synth.window [256] (must be the periodic Hann window) */
  NWS_T_REVERB_IR,        /* reverb.ir                           [1,31999]  */
  NWS_T_COUNT
};

/* Stage selectors for nws_stage_td_mlp. */
enum { NWS_MLP_FILM = 0, NWS_MLP_NOISE = 1 };

const char* nws_last_error(void);
int nws_api_version(void);

/* Replaces: gin-configured construction `NeuralWaveshaping()` (neural_waveshaping.py:31-62). */
int nws_create(const NwsConfig* config, NwsHandle* out_handle);
int nws_destroy(NwsHandle handle);
/* Fills *config with the newt.gin values. */
void nws_default_config(NwsConfig* config);

/* Element count nws_load_weights reads from tensor `index` (NwsTensor order) under the newt.gin
 * configuration, or 0 for an index out of range.  A binding compares its tensors against these before
 * handing over raw pointers (load_state_dict's shape check, resynthesise_dataset.py:47).            */
size_t nws_tensor_numel(int index);

/* 0, or NWS_ERR_CUDA once a tensor-core kernel launched through this handle has given up waiting on
 * an mbarrier (the wait is bounded so that a bad descriptor cannot hang the device; the kernel then writes
 * NaNs).  Does not synchronise: every other entry point makes the same check on entry, and
 * nws_forward_host after its own synchronise.                                                        */
int nws_status(NwsHandle handle);

/* Replaces: load_state_dict / `.to(device)` of the module parameters (resynthesise_dataset.py:47,53).
 * tensors[i] is the device pointer of tensor i in NwsTensor order (n_tensors == NWS_T_COUNT).
 * Repacks into the kernels' layouts; synchronises `stream` once (load-time only) to validate the
 * noise window.  Invalidates the LUT and the cached reverb plans.                                 */
int nws_load_weights(NwsHandle handle, const float* const* tensors, int n_tensors, void* stream);

/* Replaces: FastNEWT._init_lookup_table (modules/shaping.py:107-119) — evaluates the 64 shaper
 * MLPs on the table grid on the device and keeps the table.  `sample_points` (device, [table_size])
 * is the grid, normally torch.linspace(table_min, table_max, table_size) computed by the caller
 * exactly as the reference does; NULL -> the grid is computed on the device (may differ from
 * torch.linspace by 1 ulp at some points).                                                        */
int nws_build_lut(NwsHandle handle, int table_size, float table_min, float table_max,
                  const float* sample_points, void* stream);
/* Same, from a caller-provided table [64, table_size] (e.g. a loaded `newt.lookup_table`).        */
int nws_set_lut(NwsHandle handle, const float* lut, int table_size, float table_min, float table_max, void* stream);
/* Copies the current table [64, table_size] to `lut_out`. */
int nws_get_lut(NwsHandle handle, float* lut_out, void* stream);

/* Scratch the forward needs for a batch of B utterances of T control frames (bytes). */
size_t nws_workspace_bytes(NwsHandle handle, int B, int T);

/* Replaces: NeuralWaveshaping.forward (neural_waveshaping.py:74-90), eval/no-grad.
 *   f0       [B,1,T] Hz                 control [B,ctrl_channels,T] (channels 0,1 used, :69-72)
 *   u_phase  [101] the rand_like draw of generators.py:55, or NULL -> Philox(seed, offset)
 *   noise    [128*T-1] the rand draw of generators.py:30, or NULL -> Philox(seed, offset)
 *   use_lut  0 = NEWT (shaping.py:40-79), 1 = FastNEWT (shaping.py:82-151; needs a LUT)
 *   out      [B, 128*T]
 * T >= 2 (the reference raises for T == 1 in torch.stft's reflect padding).                       */
int nws_forward(NwsHandle handle, const float* f0, const float* control, int ctrl_channels,
                const float* u_phase, const float* noise, uint64_t seed, uint64_t offset,
                float* out, int B, int T, int use_lut, void* workspace, size_t workspace_bytes,
                void* stream);

/* Same path with HOST buffers (pageable or pinned): copies f0/control in, runs nws_forward,
 * copies `out` back and synchronises `stream`.  u_phase/noise are host pointers or NULL.           */
int nws_forward_host(NwsHandle handle, const float* f0_host, const float* control_host, int ctrl_channels,
                     const float* u_phase_host, const float* noise_host, uint64_t seed, uint64_t offset,
                     float* out_host, int B, int T, int use_lut, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---- stage entry points: one per row of SURVEY.md §8(a), used by the parity tests and by the
 *      host-side sub-module mirrors.  Layouts are the reference's ([B,C,T] channel-major).        */

/* get_embedding + ControlModule (neural_waveshaping.py:69-72,17-26): control -> emb [B,128,T]. */
int nws_stage_control_embedding(NwsHandle handle, const float* control, int ctrl_channels, float* emb,
                                int B, int T, void* workspace, size_t workspace_bytes, void* stream);
/* TimeDistributedMLP (modules/dynamic.py:20-40): emb [B,128,T] -> [B,256,T] (NWS_MLP_FILM,
 * shaping.py:53-55,68) or [B,129,T] (NWS_MLP_NOISE, neural_waveshaping.py:58,82).                 */
int nws_stage_td_mlp(NwsHandle handle, int which, const float* emb, float* out, int B, int T,
                     void* workspace, size_t workspace_bytes, void* stream);
/* F.upsample + render_exciter + NEWT/FastNEWT (neural_waveshaping.py:75-80, generators.py:58-66,
 * shaping.py:67-79,136-151): f0 [B,1,T], film [B,256,T], u_phase [101] -> newt_out [B,128*T].
 * `exciter_out` (optional, [B,64,128*T]) receives render_exciter's output for the parity tests.    */
int nws_stage_audio(NwsHandle handle, const float* f0, const float* film, const float* u_phase,
                    float* newt_out, float* exciter_out, int B, int T, int use_lut,
                    void* workspace, size_t workspace_bytes, void* stream);
/* FastNEWT.shaping_fn (shaping.py:136-151) on a materialised input: x [B,64,N] -> y [B,64,N];
 * `lower_out` (optional, int32 [B,64,N]) receives the clamped lower table index (bit-exact target). */
int nws_stage_lut_lookup(NwsHandle handle, const float* x, float* y, int* lower_out, int B, int N, void* stream);
/* FIRNoiseSynth.forward (generators.py:21-35): H [B,129,T], noise [128*T-1] -> [B,128*T]. */
int nws_stage_noise(NwsHandle handle, const float* H, const float* noise, float* out, int B, int T,
                    void* workspace, size_t workspace_bytes, void* stream);
/* Reverb.forward (shaping.py:161-173): x [B,N] -> out [B,N] (N any positive length). */
int nws_stage_reverb(NwsHandle handle, const float* x, float* out, int B, int N,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Scratch for nws_stage_reverb when N is not 128*T. */
size_t nws_reverb_workspace_bytes(NwsHandle handle, int B, int N);

/* TrainableNonlinearity on a shared grid (FastNEWT._init_lookup_table, modules/shaping.py:107-119),
 * without a handle: shaper_tensors = the nine newt.shaping_fn tensors in state-dict order
 * (input_scale, net.0.weight, net.0.bias, net.2.weight, net.2.bias, net.4.weight, net.4.bias,
 * net.6.weight, net.6.bias), x [n_points] -> out [64, n_points].  scratch: device,
 * nws_shaper_eval_scratch_bytes() bytes.                                                          */
size_t nws_shaper_eval_scratch_bytes(void);
int nws_shaper_eval(const float* const* shaper_tensors, const float* x, float* out, int n_points,
                    void* scratch, void* stream);

/* Per-stage device timing of nws_forward (cudaEvents on the launch stream).  Stage order:
 * rng, phase_carry, gru, proj, film_mlp, noise_mlp, noise_spectrum, noise_filter, audio_fused, reverb. */
#define NWS_N_STAGES 10
int nws_set_profiling(NwsHandle handle, int enable);
/* Waits for the recorded events and writes the last forward's stage times (ms) to ms_out[0..9]. */
int nws_get_stage_times(NwsHandle handle, float* ms_out, int n);

/* control [B,C,T] -> film [B,256,T], bands [B,129,T]: get_embedding + ControlModule + newt.mlp +
 * h_generator (neural_waveshaping.py:69-72,78,82; shaping.py:68) — the whole hop-rate chain. */
int nws_stage_control_to_params(NwsHandle handle, const float* control, int ctrl_channels, float* film_out,
                                float* bands_out, int B, int T, void* workspace, size_t workspace_bytes, void* stream);
/* Pipelined forward (default on): the GRU runs on an internal stream while the caller's stream (and an auxiliary one)
 * render, block by block in time, the frames already encoded (MLP chain, noise hops, audio hops); used when the batch is
 * large enough (B*T >= 4096, T >= 257).  fp32 recurrence: a 128-frame head and the rest, one GRU launch each, stream events.
 * Tensor-core recurrence: one GRU launch that publishes its progress in device-side counters, blocks of 32 then 117 frames,
 * the consumers wait with a one-thread kernel (inside a stream capture: per-block launches and events instead).
 * 0 = strictly serial kernels on the caller's stream. */
int nws_set_pipeline(NwsHandle handle, int enable);

/* Implementation of the GRU recurrence (ControlModule, neural_waveshaping.py:17-26): 1 (default) = tensor cores from 64
 * utterances on — eight utterances per CTA as the N of a warp-level mma m16n8k16, W_hh resident in registers as
 * fp16-split fragments with fp32-equivalent products (csrc/nws_gru_mma.cu) — and fp32 SIMT, one utterance per CTA, below
 * (csrc/nws_hop_bodies.cuh: the shorter step while the chip is not full; also inside the fused short-buffer front end
 * and the streaming path); 0 = fp32 SIMT always; 2 = tensor cores for any batch (cross-checks).                  */
int nws_set_gru_impl(NwsHandle handle, int impl);

/* Implementation of the hop-rate MLP chain: 1 (default) = one tcgen05 kernel, activations in TMEM, weights
 * streamed by cp.async.bulk (csrc/nws_mlp_tc.cu); 0 = fp32 SIMT layer kernels (csrc/nws_encoder.cu). */
int nws_set_mlp_impl(NwsHandle handle, int impl);

/* Implementation of the fused audio-rate kernel's harmonic mixer: 1 (default) = tcgen05 3xTF32 with
 * the accumulator in TMEM (csrc/nws_audio_tc.cu), 0 = fp32 SIMT (csrc/nws_audio.cu, kept as the
 * in-library cross-check). */
int nws_set_audio_impl(NwsHandle handle, int impl);
/* Where the filtered-noise branch (FIRNoiseSynth.forward, generators.py:21-35) of a whole-utterance forward runs:
 * 0 (default) = nws_noise_filter_kernel writes the filtered noise into the output buffer first and the fused audio-rate
 * kernel adds its samples to it; 1 = inside the fused audio-rate kernel: the warpgroup's MMA-issuing warp filters the
 * next tile's hop while the compute warps run the current tile's epilogue (band gains and noise spectrum straight from
 * L2, one in-place 256-point inverse FFT per frame pair in shared memory, overlap-add in registers; nothing of the branch
 * touches HBM).  Both are parity-tested against each other and the oracle; the in-kernel variant is the slower one on
 * B200 (FastNEWT audio kernel 0.514 -> 0.575 ms against 0.049 ms for the separate launch), hence not the default.   */
int nws_set_noise_fused(NwsHandle handle, int enable);
/* How the fused kernel evaluates the NEWT shapers' two 8x8 hidden layers (TrainableNonlinearity, shaping.py:25-34):
 * 1 (default) = on the tensor cores (warp-level mma m16n8k8, 3xTF32), 0 = fp32 FMA with the weights shared by lane
 * pairs.  Both are parity-tested; the switch exists for cross-checks and measurements.                        */
int nws_set_shaper_impl(NwsHandle handle, int impl);
/* Reverb.forward (shaping.py:161-173) for buffers of up to 4096 samples (and for streaming pushes of up to 33 hops):
 * 1 (default) = direct-form convolution in one launch (csrc/nws_reverb_direct.cu), 0 = always the FFT path.       */
int nws_set_reverb_direct(NwsHandle handle, int enable);
/* Hop-rate MLP chain for up to 40 frames per utterance (and up to 32 utterances): 1 (default) = fp32 small-batch chain,
 * 2 CTAs per utterance (csrc/nws_mlp_small.cu); 0 = always the 128-frame-tile tensor-core kernel.                  */
int nws_set_small_path(NwsHandle handle, int enable);

/* Self-test of the tcgen05 path (csrc/nws_tc.cuh): D[128,64] = A[128,K] . B[64,K]^T, 3xTF32 in TMEM,
 * K a multiple of 8 up to 104.  status[0] = 1 on completion, -1 if the MMA never signalled.
 * swap_lbo_sbo bit 0: exchange the descriptor strides (negative test); bit 1: store A's low part without
 * masking it to tf32 (the result must not change: kind::tf32 ignores the 13 low mantissa bits). */
int nws_selftest_umma(const float* A, const float* B, float* D, int K, int swap_lbo_sbo, int* status, void* stream);

/* Accuracy probe of the device sine implementations (csrc/nws_math.h): accurate polynomial version and
 * the SFU-based versions (quarter-turn reduction + sin/cos select; full-turn reduction, one MUFU) on x[0..n). */
int nws_selftest_sin(const float* x, float* y_accurate, float* y_quarter, float* y_turn, long long n, void* stream);

/* fp32 FMA issue-rate probe (the compute-roofline denominator bench.py reports beside the HBM one): n_blocks CTAs of
 * 256 threads, each thread iters x 16 rounds of eight independent fmaf chains; out: device, n_blocks * 256 floats;
 * *flops_out (host, optional) = the flops one launch executes.  The caller times the launch. */
int nws_selftest_ffma_peak(float* out, int n_blocks, int iters, double* flops_out, void* stream);

/* ---- stateful streaming: the same forward fed a few control frames at a time.
 * The reference only times independent, stateless forwards per buffer size (scripts/time_buffer_sizes.py:
 * 35-72); a real-time caller needs what crosses a buffer boundary carried over: the GRU state
 * (neural_waveshaping.py:21,25), the running phase sum (generators.py:59), the neighbouring frames of the x128
 * linear upsampling (neural_waveshaping.py:75, shaping.py:69) and of the noise overlap-add (generators.py:31-35),
 * and the reverb's past input (shaping.py:161-173).  Contract: the concatenation of the audio returned by the
 * pushes (+ the flush) equals the dry signal of ONE nws_forward over the concatenated frames, convolved
 * causally with [0, ir] (the reference's circular wrap of the reverb tail has no streaming equivalent).
 * Output lags input by one hop: hop h interpolates towards frame h+1. */
typedef struct NwsStreamState* NwsStreamHandle;
/* State for B parallel streams, at most max_frames (2..4096) control frames per push.  Owns its device memory. */
int nws_stream_create(NwsHandle handle, int B, int max_frames, NwsStreamHandle* out_stream);
int nws_stream_destroy(NwsStreamHandle stream_state);
/* Start of an utterance: zero state.  u_phase (device, [101]) = the phase-shift draw of generators.py:55, fixed
 * for the whole stream, or NULL -> Philox(seed, offset) — the same values nws_forward(seed, offset) would draw. */
int nws_stream_reset(NwsStreamHandle stream_state, const float* u_phase, uint64_t seed, uint64_t offset, void* stream);
/* The kernels run on a window of [<= 3 history frames | n_frames new frames]: first_frame = index of the
 * window's first frame in the utterance, window_frames = its length.  For callers that inject the noise. */
int nws_stream_window(NwsStreamHandle stream_state, int n_frames, long long* first_frame, int* window_frames);
/* f0 [B,1,n_frames], control [B,ctrl_channels,n_frames] (device).  noise_window: NULL -> Philox draws indexed
 * by absolute sample (same stream as nws_forward), or device [128*window_frames - 1] = the utterance's noise
 * vector (generators.py:30) from sample 128*first_frame on.  flush != 0: these are the utterance's last frames
 * (n_frames may be 0): the pending hop is rendered with the reference's end-of-signal clamping and the stream
 * must be reset before the next push.  apply_reverb = 0 returns the dry signal.
 * out [B, 128 * *n_out_frames] (device, room for 128*(n_frames+1) samples per row); *n_out_frames (host) =
 * n_frames - 1 on the first push, n_frames afterwards, + 1 with flush.  The first push needs n_frames >= 2. */
int nws_stream_push(NwsStreamHandle stream_state, const float* f0, const float* control, int ctrl_channels,
                    int n_frames, const float* noise_window, int use_lut, int flush, int apply_reverb, float* out,
                    int* n_out_frames, void* stream);

/* ---- control-side feature extraction: what produces control channel 1 of the forward
 * in the timbre-transfer use case (colab cell 14, scripts/create_dataset.py).  No handle: these do not depend on
 * the model.  All of the reference's arithmetic here is librosa 0.8.0's (requirements.txt:5).
 *
 * Replaces: extract_perceptual_loudness / compute_power_spectrogram / perform_perceptual_weighting
 * (neural_waveshaping_synthesis/data/utils/loudness_extraction.py:43-68, :11-23, :26-40) for a batch of B
 * segments of N samples: centred reflect-padded STFT (periodic Hann window, n_fft a power of two 64..4096,
 * float64 transform stored as complex64 like librosa.stft), amplitude_to_db(ref=max over the segment,
 * amin=epsilon, top_db=80), mean over the n_fft/2+1 bins (the A-weighting is computed but not added by the
 * reference, :39), and (x + 80) / 80 when normalise != 0.
 *   audio        [B, N]                     loudness_out [B, 1 + N / hop_length]
 *   db_out       optional [B, 1 + N / hop_length, n_fft / 2 + 1] (frame-major): the dB spectrogram that
 *                compute_power_spectrogram returns (transposed)
 * N > n_fft / 2 (the reflect padding needs it).                                                            */
size_t nws_loudness_workspace_bytes(int B, int N, int n_fft, int hop_length);
int nws_extract_loudness(const float* audio, int B, int N, int n_fft, int hop_length, double epsilon,
                         int normalise, float* loudness_out, float* db_out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Replaces: extract_rms (loudness_extraction.py:71-90): zero-padded centred frames of window_size every
 * hop_length, sqrt(mean(x^2)).  audio [B, N] -> rms_out [B, 1 + (N + 2*(window_size/2) - window_size) / hop_length]. */
int nws_extract_rms(const float* audio, int B, int N, int window_size, int hop_length, float* rms_out, void* stream);

/* Replaces: linear_interpolation (data/utils/upsampling.py:20-36) — the default `interpolate_fn` of the two
 * extractors above: frame values [B,F] (device, fp32) -> one float64 value per sample, np.interp over
 * np.linspace(0, F-1, F*hop + window - hop); original_length > 0 crops to [window/2, window/2 + original_length)
 * as the reference does, 0 keeps the padded length.  nws_interp_frames_len gives the row length of `out`.  */
int nws_interp_frames_len(int F, int window_length, int hop_length, int original_length);
int nws_interp_frames(const float* frames, int B, int F, int window_length, int hop_length, int original_length,
                      double* out, void* stream);

/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
uint64_t nws_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* NWS_B200_H_ */
