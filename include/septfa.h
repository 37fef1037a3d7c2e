/*
 * septfa.h - C ABI of libseptfa.so: the B200 (sm_100a) Sep-TFAnet-VAD inference forward pass.
 *
 * This is the drop-in boundary for ONE path of the reference (BaekMS/Sep-TFAnet-VAD):
 *   - SeparationModel.__init__ / load_state_dict / forward      model/model.py:360-461
 *   - OnlineSaving.calc_online (one hop of the sliding window)  model/online_class_unknown_targets.py:72-105
 *   - PITLossWrapper('pw_pt', L1) + reorder_source_mse          model/pit_wrapper.py:149-177,261-312; model/combined_loss.py:63-78
 *
 * The reference boundary is a Python nn.Module (there is no FFI in the reference); the entry
 * points below are what a ctypes binding of that module needs: plain pointers and sizes, no
 * torch types. Conventions: every call returns 0 on success or a negative SEPTFA_E_* code and
 * never throws; septfa_last_error() gives the message. All device pointers are caller-owned
 * and must live on the handle's device. Calls are stream-ordered on `stream` (a cudaStream_t
 * passed as void*; NULL = legacy default stream) and do not synchronise the host, except the
 * *_host variants which return after their result is in host memory. One handle per device;
 * a handle is not thread-safe.
 *
 * Device tensor layouts (float32 unless noted) mirror the torch tensors of the reference:
 *   x        [B, L]                 mixture waveforms                     forward(x)          model.py:402
 *   out_wav  [B, 2, L]              separated waveforms                   out_separation      model.py:460
 *   out_vad  [B, 2, T]              VAD probabilities, or the smoothed {0,1} decisions when
 *                                   kw->return_smoothed_vad (reference shape [B,2,1,T])      model.py:424-457
 *   est_stft [B, 2, 257, T] float2  estimated STFTs (complex64), nullable  estimated_stfts    model.py:437,453
 *   mask     [B, 2, 257, T]         post-sigmoid masks, nullable           mask_per_speaker   model.py:429
 *   spectrum [B, 257, T]            gated dB spectrum, nullable            self.spectrum      model.py:412-419
 *   logits   [B, 514, T]            pre-sigmoid masks, nullable            self.masks_b       model.py:421
 * with T = 1 + L / 256 (n_fft 512, hop 256, centre/reflect padding; L >= 257).
 */
#ifndef SEPTFA_H_
#define SEPTFA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEPTFA_OK 0
#define SEPTFA_E_INVALID (-1)     /* bad argument / unsupported configuration */
#define SEPTFA_E_CUDA (-2)        /* CUDA runtime error */
#define SEPTFA_E_STATE (-3)       /* call order (weights not committed, ...) */
#define SEPTFA_E_KEY (-4)         /* unknown / missing / mis-sized state_dict key */
#define SEPTFA_E_WORKSPACE (-5)   /* workspace too small */

typedef struct septfa_handle septfa_handle;
typedef struct septfa_online septfa_online;

/* arch.args of config_with_vad.json / config_without_vad.json (config_with_vad.json:7-28),
 * consumed by SeparationModel.__init__ (model/model.py:362-391). Booleans are 0/1. */
typedef struct septfa_config {
  int32_t n_fft_bins;                 /* "n_fftBins": only 512 */
  int32_t bn_dim;                     /* "BN_dim": only 256 (= n_fftBins/2, model.py:377-380) */
  int32_t h_dim;                      /* "H_dim": only 512 */
  int32_t layer;                      /* "layer" (blocks per stack, dilation cycle) */
  int32_t stack;                      /* "stack" */
  int32_t num_spk;                    /* "num_spk": only 2 */
  int32_t skip;                       /* must be 0 */
  int32_t dilated;                    /* must be 1 */
  int32_t causal;                     /* "casual" (sic): must be 0 */
  int32_t weight_norm;                /* must be 1 */
  int32_t final_vad;                  /* VAD head present */
  int32_t final_vad_masked_speakers;  /* must be 0 */
  int32_t noisy_phase;                /* model.py:430-439 (both branches equal S*mask numerically) */
  int32_t activity_input_bool;        /* model.py:414-419 */
  int32_t tf_attention;               /* model.py:344-345 */
  int32_t apply_recursive_ln;         /* model.py:347-348 */
  int32_t apply_residual_ln;          /* model.py:349-350 */
} septfa_config;

/* inference_kw of forward (model/model.py:444-457; only_inference.py:102-108).
 * Pass NULL for the reference's empty dict (no gating, raw probabilities). */
typedef struct septfa_infer_kw {
  int32_t length_smoothing_filter;      /* accepted and ignored, like the reference (weights are overwritten with [1,0,1]) */
  float threshold_activated_vad;
  int32_t filter_signals_by_smo_vad;
  int32_t filter_signals_by_unsmo_vad;  /* also gates with the *smoothed* decisions (model.py:454-455) */
  int32_t return_smoothed_vad;
} septfa_infer_kw;

/* Engines for the dense contractions (1x1 convs): a bit mask of the contractions that run on the
 * plain fp32 CUDA-core kernels instead of the tcgen05 tensor cores (bring-up / numerics cross-check):
 * bit 0 = conv1d (256->256), bit 1 = dconv+res_out (512->256), bit 2 = output conv (256->514). */
#define SEPTFA_ENGINE_TCGEN05_F16 0 /* all on tcgen05: fp16 operands, fp32 TMEM accumulators (default) */
#define SEPTFA_ENGINE_FP32_SIMT 7   /* all on fp32 CUDA cores */

int septfa_create(septfa_handle** out, const septfa_config* cfg, int device);
void septfa_destroy(septfa_handle* h);
const char* septfa_last_error(const septfa_handle* h); /* h may be NULL: last create() error */
const char* septfa_version(void);

/* Weights: one call per reference state_dict entry, by its reference key
 * (e.g. "TCN.TCN.3.conv1d.weight_v"), raw float32 host data in torch's contiguous layout.
 * weight_norm folding (w = g*v/||v||) and tensor-core packing happen in septfa_commit_weights,
 * which fails with SEPTFA_E_KEY if any key the configuration requires is missing - the same
 * contract as load_state_dict(strict=True) (only_inference.py:57-60). */
int septfa_set_tensor(septfa_handle* h, const char* key, const float* host_data, int64_t numel);
int septfa_commit_weights(septfa_handle* h);
/* Number of keys the configuration expects, and the i-th key (for strict checking on the host side). */
int septfa_num_keys(const septfa_handle* h);
const char* septfa_key_name(const septfa_handle* h, int i);
int64_t septfa_key_numel(const septfa_handle* h, int i);

/* name = "engine" (SEPTFA_ENGINE_*), "profile" (0/1), "host_chunks" (0 = automatic, 1..8: number of
 * batch chunks septfa_forward_host pipelines over its two stream lanes); kernel-selection switches for
 * cross-checks (default 1): "fused_resid" (cluster-resident gate + residual kernel), "conv1_persist"
 * (persistent warp-specialised conv1 kernel), "pdl" (programmatic dependent launch of the kernel chain),
 * "dconv_mma" (tensor-core depthwise + res_out kernel), "dconv_pair" (... on CTA pairs, cta_group::2, weights resident),
 * "conv1_pair" (TF32 CTA-pair conv1 fed by TMA from the fp32 stream; blocks 1 .. n-1 of the recursive-LN wiring),
 * "conv1_wres" (resident weights in the persistent conv1 kernel), "l2_discard" (the hand-off buffers inside a block are
 * dropped from the L2 once they have been read);
 * "stream_half" (default 0): opt-in mode that stores the residual stream between blocks as fp16 - 11 % faster, but the
 * stream's rounding accumulates over all blocks (VAD probabilities within 2e-3 of the reference instead of 1e-3). */
/* "precision": arithmetic of the two block contractions (conv1d 256->256, res_out 512->256) on the tensor cores.
 *   SEPTFA_PRECISION_FAST      fp16 operands, one tcgen05 pass, fp16 storage of the tensors between them
 *   SEPTFA_PRECISION_ACCURATE  2-term fp16 split of both operands (three passes, fp32-accurate), fp32 storage
 *   SEPTFA_PRECISION_AUTO      (default) ACCURATE for apply_residual_ln (config_without_vad: the residual stream is
 *                              never re-normalised and accumulates the rounding of all blocks), FAST otherwise */
#define SEPTFA_PRECISION_AUTO 0
#define SEPTFA_PRECISION_FAST 1
#define SEPTFA_PRECISION_ACCURATE 2
int septfa_set_option(septfa_handle* h, const char* name, int value);
int septfa_get_option(const septfa_handle* h, const char* name);

int64_t septfa_num_frames(int64_t L); /* T = 1 + L/256 */
size_t septfa_workspace_bytes(const septfa_handle* h, int B, int64_t L);

/* SeparationModel.forward (model/model.py:402-461). Nullable outputs are skipped. */
int septfa_forward(septfa_handle* h, const float* x, int B, int64_t L, const septfa_infer_kw* kw,
                   float* out_wav, float* out_vad, void* est_stft, float* mask, float* spectrum,
                   float* logits, void* workspace, size_t workspace_bytes, void* stream);

/* Same, from/to HOST buffers: copies x host->device, runs forward, copies out_wav/out_vad back and
 * waits for completion. Page-locked caller buffers are used directly, pageable ones go through
 * pinned staging owned by the handle. The batch is processed in chunks so that copies overlap the
 * kernels (utterances are independent). This is the call a caller without device memory of its
 * own makes (only_inference.py:90-91 semantics). */
int septfa_forward_host(septfa_handle* h, const float* x_host, int B, int64_t L, const septfa_infer_kw* kw,
                        float* out_wav_host, float* out_vad_host);

/* Batch streams (only_inference.py's loop over files, model/model.py:402 per batch): the same from/to HOST forward,
 * split into an asynchronous submit and a wait so that successive batches overlap - while batch n computes, the
 * results of batch n-1 travel device->host and the input of batch n+1 host->device (copy engines and SMs all busy).
 * SEPTFA_HOST_SLOTS independent slots (0 .. 3), each with its own copy stream, device buffers and workspace (all slots
 * share one compute stream: a batch's forward has the SMs to itself). THREE slots in rotation keep the pipeline full: with
 * two, the copy-in of batch i + 1 queues behind the copy-out of batch i - 1 on the same slot and the compute stream idles
 * ~0.4 ms per 256 x 4 s step. A slot holds one batch in
 * flight: submit(slot) -> wait(slot) -> submit(slot) ... The host buffers must be page-locked (cudaHostAlloc /
 * cudaHostRegister / torch pin_memory) and stay valid and untouched until wait(slot) returns. */
#define SEPTFA_HOST_SLOTS 4
int septfa_forward_host_submit(septfa_handle* h, int slot, const float* x_host, int B, int64_t L, const septfa_infer_kw* kw,
                               float* out_wav_host, float* out_vad_host);
int septfa_forward_host_wait(septfa_handle* h, int slot);

/* The same submit with 16-bit sample formats on the host side, which halves the PCIe bytes in either direction:
 *   x_fmt   SEPTFA_FMT_F32   float32 waveforms, already normalised (what forward() takes)
 *           SEPTFA_FMT_PCM16 int16 PCM exactly as scipy.io.wavfile.read returns it (only_inference.py:68); the device then
 *                            does only_inference.py:69,81: astype(float32) and the min-max normalisation to [-0.9, 0.9],
 *                            bit-identical to the reference's numpy expression, before the forward
 *   out_fmt SEPTFA_FMT_F32   float32 waveforms
 *           SEPTFA_FMT_F16   IEEE half, round to nearest: the `-ps 16` format of save_audio (Our_utils/utlis_inference.py:30-32)
 * x_host: B*L elements of x_fmt; out_wav_host: B*2*L elements of out_fmt; out_vad_host: float32 as above. */
#define SEPTFA_FMT_F32 0
#define SEPTFA_FMT_PCM16 1
#define SEPTFA_FMT_F16 2
int septfa_forward_host_submit_fmt(septfa_handle* h, int slot, const void* x_host, int x_fmt, int B, int64_t L,
                                   const septfa_infer_kw* kw, void* out_wav_host, int out_fmt, float* out_vad_host);

/* ---- small-request serving: the forward as a CUDA graph ---------------------------
 * A forward of one short mixture is ~80 kernels of a few microseconds each: launch-bound (0.9 ms for one 4 s mixture).
 * septfa_graph_capture records septfa_forward for one (B, L, kw) on the caller's STATIC device buffers (x, out_wav,
 * out_vad, workspace: same pointers at every replay) into an executable CUDA graph; septfa_graph_launch replays it on a
 * stream (the caller refreshes x before, reads out_wav / out_vad after, stream-ordered). One graph per shape bucket.
 * The optional exports (est_stft / mask / spectrum / logits) are not part of the graph. */
typedef struct septfa_graph septfa_graph;
int septfa_graph_capture(septfa_handle* h, const float* x, int B, int64_t L, const septfa_infer_kw* kw, float* out_wav,
                         float* out_vad, void* workspace, size_t workspace_bytes, septfa_graph** out);
int septfa_graph_launch(septfa_graph* g, void* stream);
int septfa_graph_num_nodes(const septfa_graph* g);
void septfa_graph_destroy(septfa_graph* g);

/* Number of GPU kernels launched by the last forward / online step on this handle. */
int septfa_last_launch_count(const septfa_handle* h);

/* Per-kernel-class device timing. With option "profile" = 1, forward records CUDA events on its
 * launch stream around each class of kernels; septfa_profile_read waits for them and returns the
 * accumulated milliseconds and interval counts per class (arrays of SEPTFA_PROF_NCAT). */
#define SEPTFA_PROF_FRONTEND 0 /* memset + STFT/dB + activity gate (+ spectrum export) */
#define SEPTFA_PROF_CONV1 1    /* tcgen05 GEMM 256->256 with GroupNorm prologue, PReLU/stats epilogue */
#define SEPTFA_PROF_DCONV 2    /* tcgen05 GEMM 512->256 with depthwise-conv prologue */
#define SEPTFA_PROF_GATE 3     /* TF-attention gates */
#define SEPTFA_PROF_RESID 4    /* post-block GroupNorm residual kernels */
#define SEPTFA_PROF_OUTCONV 5  /* output statistics + tcgen05 GEMM 256->514 */
#define SEPTFA_PROF_VAD 6      /* VAD head + smoothing */
#define SEPTFA_PROF_ISTFT 7    /* mask + inverse STFT overlap-add */
#define SEPTFA_PROF_EXPORT 8   /* optional torch-layout exports */
#define SEPTFA_PROF_NCAT 9
int septfa_profile_read(septfa_handle* h, double* ms, int* launches, int n, int reset);

/* ---- online mode: OnlineSaving.calc_online, one hop for S independent streams -------------
 * State kept on the device per stream: the last <= 2 s of already-emitted (permutation-fixed)
 * signal. Each step takes the current 3 s windows win[S, 48000], runs forward, picks per stream
 * the speaker permutation that minimises the L1 distance between the window's overlap with the
 * emitted signal (PITLossWrapper pw_pt + L1Loss, n_src = 2, ties -> identity), reorders, and
 * emits the last 1 s: emitted[S, 2, 16000], perm[S, 2] (int32). */
int septfa_online_create(septfa_handle* h, int S, septfa_online** out);
void septfa_online_destroy(septfa_online* st);
int septfa_online_reset(septfa_online* st, void* stream);
size_t septfa_online_workspace_bytes(const septfa_online* st);
int septfa_online_step(septfa_online* st, const float* win, const septfa_infer_kw* kw, float* emitted,
                       int32_t* perm, void* workspace, size_t workspace_bytes, void* stream);
int septfa_online_hops_done(const septfa_online* st);

/* Stand-alone PIT for 2 sources (model/pit_wrapper.py:149-177,261-312): a[S,2,n], b[S,2,n] ->
 * perm[S,2] (int32), decided per stream; pw_sums (nullable, device double [S,4]) receives the
 * pairwise L1 sums sum_n|a_i - b_j| at [s][i*2+j]. */
int septfa_pit_l1(septfa_handle* h, const float* a, const float* b, int S, int64_t n, int32_t* perm, double* pw_sums,
                  void* stream);

/* Pre-processing of only_inference.py:81, batched on the device: per utterance
 *   out = 1.8 * (x - min(x)) / (max(x) - min(x)) - 0.9        (float32, the reference's operation order: bit-identical
 * to numpy; a constant signal gives NaN like the reference). x, out: device [B, L]; lengths (nullable, device int64 [B]):
 * valid samples per utterance for ragged batches - extrema over the valid part, zeros written behind it. */
int septfa_minmax_normalize(septfa_handle* h, const float* x, int B, int64_t L, const int64_t* lengths, float* out, void* stream);

/* calc_sisdr (model/combined_loss.py:16-56): SI-SDR in dB of `rows` pairs preds[rows, n] / target[rows, n] (device,
 * float32, contiguous) in one pass; out_db device [rows]; scratch: device, rows * 5 doubles; rows <= 65535. Runs on the
 * current device (no handle). The online drivers use it for the online-vs-offline quality report of test.py:223-269. */
int septfa_sisdr(const float* preds, const float* target, int64_t rows, int64_t n, int zero_mean, float* out_db, double* scratch,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEPTFA_H_ */
