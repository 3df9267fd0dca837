/*
 * sad_b200.h -- C ABI of the B200-native inference hot path of Synthetic-Audio-Detection.
 *
 * The reference (pure Python) has no FFI of its own; its boundary for this path is the module-level
 * API of modular/source/inference_runner.py and model_merger.py.  The Python mirror of that API
 * (synthetic-audio-detection_b200/inference_runner.py) binds THIS header through ctypes; every entry
 * point below names the reference function (file:line under /root/reference/modular/source) whose
 * arithmetic it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SAD_E* code otherwise; sad_last_error(ctx) has text;
 *   - "dev" pointers are CUDA device pointers on the context's device, "host" pointers are CPU memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); calls are
 *     stream-ordered and do not synchronise unless stated;
 *   - the caller owns every buffer it passes; the library owns its context, weights and workspace;
 *   - there is no CPU fallback: without a CUDA device sad_create fails with SAD_ENODEVICE;
 *   - a context owns ONE workspace: calls on a context may come from any stream (and sad_forward_host uses its own
 *     internal streams), the library orders each call after the previous one on that context with an event, so
 *     results never race -- but two calls on one context never overlap.  Use one context per concurrent stream.
 *     Calls on one context must not be issued from two host threads at the same time;
 *   - every entry point leaves the calling thread's current CUDA device as it found it.
 */
#ifndef SAD_B200_H_
#define SAD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAD_OK 0
#define SAD_EINVAL (-1)    /* bad argument                                   */
#define SAD_ENODEVICE (-2) /* no usable sm_100 CUDA device                   */
#define SAD_ECUDA (-3)     /* a CUDA runtime / driver call failed            */
#define SAD_ESTATE (-4)    /* call order violated (e.g. weights not loaded)  */

/* Fixed geometry of the path (inference_runner.py:258-259). */
#define SAD_SAMPLE_RATE 32000
#define SAD_SEGMENT_SAMPLES 128000 /* int(4.0 * 32000)                       */
#define SAD_N_FFT 2048
#define SAD_HOP 512
#define SAD_N_FRAMES 251
#define SAD_N_FREQS 1025
#define SAD_N_MELS 128
#define SAD_IMAGE 512

typedef struct sad_ctx sad_ctx;

/* ---- lifetime ------------------------------------------------------------------------------- */
/* One context per GPU.  `n_heads` = number of sub-models of the merged ensemble
 * (ModularMultiHeadClassifier, inference_runner.py:53-73); `max_batch` = segments processed per
 * internal pass (workspace is sized for it; larger batches are chunked).                        */
int sad_create(sad_ctx** out, int device, int n_heads, int max_batch);
/* Same with the backbone named explicitly -- what `--model-name` / `backbone_name` may select (model_merger.py:101,
 * inference_runner.py:77): "resnet18" (default), "resnet34" (BasicBlock) or "resnet50" / "resnet101" / "resnet152"
 * (Bottleneck).  Any other name returns SAD_EINVAL.                                                          */
int sad_create_ex(sad_ctx** out, int device, int n_heads, int max_batch, const char* backbone);
const char* sad_backbone(const sad_ctx* ctx);
int sad_destroy(sad_ctx* ctx);
const char* sad_last_error(const sad_ctx* ctx);
const char* sad_version(void);
/* "bf16" (libsad_b200.so, the default build) or "fp16" (libsad_b200_f16.so, built with -DSAD_ACT_F16): the element type of
 * activations and conv weights -- the layout every `bf16` in this header refers to.  Same tensor-core rate; fp16 keeps 3
 * more mantissa bits and is what the 50+-conv Bottleneck backbones need to meet the 2e-2 logit bound.              */
const char* sad_act_dtype(void);

/* ---- weights: replaces load_merged_model's per-head rebuild (inference_runner.py:101-114) ----- */
/* The library describes the fp32 tensors it needs for ONE BinaryClassifier (inference_runner.py:28-51),
 * in state_dict order without the int64 num_batches_tracked entries: name i is the key suffix after
 * "sub_models.<h>." (e.g. "base.layer1.0.conv1.weight", "head.3.running_var").                    */
int sad_weight_count(void);                 /* resnet18 */
const char* sad_weight_name(int i);
long long sad_weight_numel(int i);
int sad_backbone_weight_count(const char* backbone);      /* -1 for an unsupported backbone */
const char* sad_backbone_weight_name(const char* backbone, int i);
long long sad_backbone_weight_numel(const char* backbone, int i);
/* host_tensors[i] points at sad_weight_numel(i) contiguous fp32 values.  Eval-mode BatchNorm is folded
 * into the preceding conv / Linear in fp64 and the conv weights are rounded once to bf16.          */
int sad_load_weights(sad_ctx* ctx, int head, const float* const* host_tensors, int n_tensors);

/* Optional: upload the Hann window [2048] and mel filterbank [1025*128, row = FFT bin] computed by the
 * caller (torch / torchaudio) so the constants are bit-identical to the reference's
 * (torchaudio MelSpectrogram, inference_runner.py:158-166).  Defaults are computed internally.     */
int sad_set_frontend_constants(sad_ctx* ctx, const float* host_window, const float* host_mel_fb);

/* ---- front end: replaces waveform_to_spectrogram (inference_runner.py:157-174) ---------------- */
/* pcm [B,128000] fp32 -> log-mel dB [B,128,251] fp32 (after AmplitudeToDB top_db=80, :167-170) and
 * per-segment (mean, unbiased std) [B,2] used by the standardisation at :171.  Either output may be NULL. */
int sad_frontend_logmel(sad_ctx* ctx, const float* pcm_dev, int B, float* logmel_db_dev, float* mu_sigma_dev,
                        void* stream);
/* pcm [B,128000] -> standardised, 512x512 bilinear-resized image [B,512,512] fp32; the reference's three
 * channels are identical copies of it (:172-173).                                                  */
int sad_frontend_image(sad_ctx* ctx, const float* pcm_dev, int B, float* image_dev, void* stream);

/* ---- ingest: replaces preprocess_waveform (inference_runner.py:144-155) after the container is parsed ----
 * Interleaved PCM [n_frames][n_channels] on the device -> mono (mean over channels, int16 scaled by 1/32768 as
 * torchaudio.load does) -> torchaudio.transforms.Resample(sr_in, 32000) (sinc_interp_hann, width 6, rolloff 0.99; skipped
 * when sr_in == 32000) -> zero-padded to at least one window (128000).  sad_ingest_length gives the output length
 * (torchaudio's float32 ceil rule); `out_dev` must hold that many floats.  Device -> device, ordered on `stream`.
 * No alignment is required of `pcm_dev` or `out_dev` (mono / stereo streams on a 16-byte aligned base take the faster
 * kernel; results are the same to the bit).  Only the taps where torchaudio's window is not clamped are summed (the
 * others are below 1e-32), so a non-finite float32 sample reaches the ~20 outputs whose band covers it, where the
 * reference's dense kernel spreads it over its whole 2*width+orig tap window.                                        */
#define SAD_PCM_S16 0
#define SAD_PCM_F32 1
long long sad_ingest_length(long long n_frames, int sr_in);
int sad_ingest(sad_ctx* ctx, const void* pcm_dev, int sample_format, long long n_frames, int n_channels, int sr_in,
               float* out_dev, void* stream);

/* ---- slicing gate: replaces the silence test of slice_waveform (inference_runner.py:184-188) --- */
/* For window w starting at w*hop: keep[w] = !(max|x| < silence_threshold).  n_windows as computed by
 * sad_slice_count.                                                                                 */
long long sad_slice_count(long long n_samples, long long window, long long hop);
int sad_slice_gate(sad_ctx* ctx, const float* wf_dev, long long n_samples, long long window, long long hop,
                   float silence_threshold, uint8_t* keep_dev, void* stream);
/* Gather kept windows into a dense [n_kept,window] batch: dst[i] = wf[starts[i] : starts[i]+window]. */
int sad_gather_windows(sad_ctx* ctx, const float* wf_dev, const long long* starts_dev, int n_kept,
                       long long window, float* dst_dev, void* stream);

/* ---- ensemble: replaces ModularMultiHeadClassifier.forward (inference_runner.py:62-73) over
 *      BinaryClassifier.forward (:49-51) and interpret_multihead_logits (:194-214) -------------- */
/* Fused path, pcm [B,128000] fp32 in, per segment out:
 *   logits [B,N+1] = [syn_1..syn_N, mean_i(real_i)],  probs = sigmoid(logits),
 *   labels [B] int32: N means Real, 0..N-1 the arg-max synthetic head (rule at :207-213).
 * Any output pointer may be NULL.                                                                  */
int sad_forward(sad_ctx* ctx, const float* pcm_dev, int B, float threshold, float* logits_dev, float* probs_dev,
                int32_t* labels_dev, void* stream);
/* nn.Module.forward drop-in: x [B,3,512,512] fp32 NCHW (any image, channels need not be equal). */
int sad_forward_images(sad_ctx* ctx, const float* x_nchw_dev, int B, float threshold, float* logits_dev,
                       float* probs_dev, int32_t* labels_dev, void* stream);
/* End-to-end with HOST buffers: pinned staging, H2D of pcm, compute, D2H of the results, chunked and
 * double-buffered on internal streams; synchronises before returning.                              */
int sad_forward_host(sad_ctx* ctx, const float* pcm_host, int B, float threshold, float* logits_host,
                     float* probs_host, int32_t* labels_host);

/* ---- clip aggregation: replaces inference_runner.py:328-334 ------------------------------------ */
/* clip_id [B] int32 SORTED ascending (segments of a clip are contiguous), values in [0,n_clips); clip_probs [n_clips,N+1] = mean over the clip's
 * segments of probs; clip_label = rule :207-213 applied to the clip mean (-1 for a clip with no segment). */
int sad_clip_reduce(sad_ctx* ctx, const float* probs_dev, const int32_t* clip_id_dev, int B, int n_clips,
                    float threshold, float* clip_probs_dev, int32_t* clip_label_dev, void* stream);

/* ---- synthetic corpus (bench.py, tools/run_corpus.py; no reference counterpart) ----------------- */
/* out_dev [n,128000] fp32 = segments first .. first+n-1 of the seeded noise/tone corpus the benchmarks run on
 * (a_n*N(0,1) + a_t*sin(2*pi*f*t+phi), clipped).  Counter based: a segment's bytes depend only on
 * (seed, global segment index), never on chunking or on the number of GPUs sharing the corpus.  `ctx` may be NULL:
 * the kernel then runs on the calling thread's current device.                                               */
int sad_synth_segments(sad_ctx* ctx, float* out_dev, long long first, int n, unsigned long long seed, void* stream);

/* ---- introspection (tests, bench) --------------------------------------------------------------- */
int sad_n_heads(const sad_ctx* ctx);
int sad_max_batch(const sad_ctx* ctx);
/* Number of kernels this library has launched on the context since creation. */
long long sad_launch_count(const sad_ctx* ctx);
/* Live profiling (bench.py): when enabled, every kernel class launched by sad_forward* is bracketed by CUDA events
 * on the launching stream.  Kinds 0..39 = the convolutions in state_dict order (0 = stem; 20 for resnet18), then the classes
 * below.  sad_profile_read synchronises on the recorded events and returns accumulated milliseconds and the number
 * of bracketed launch groups per kind (arrays of SAD_PROF_KINDS).  Enabling resets the counters.              */
#define SAD_PROF_CONV_SLOTS 40 /* kinds 0..39: convolutions in state_dict order      */
#define SAD_PROF_FRONTEND 40   /* fill + stft_mel + db_clamp_stats                  */
#define SAD_PROF_IMAGE 41      /* standardise/resize image (+ stem im2col, 3-ch)    */
#define SAD_PROF_POOL 42       /* 3x3/2 max pool (3-channel path only)              */
#define SAD_PROF_HEAD 43       /* avg-pool + MLP + merge + decision                 */
#define SAD_PROF_KINDS 44
int sad_profile_enable(sad_ctx* ctx, int on);
int sad_profile_read(sad_ctx* ctx, double* ms_by_kind, long long* launches_by_kind);
/* Run ONE convolution layer of one head on caller buffers (NHWC bf16): layer = index into the 20 convs in
 * state_dict order (0 = stem conv1 is not available here; 1..19).  `in` [B,Hi,Wi,Cin], `residual`
 * [B,Ho,Wo,Cout] or NULL, `out` [B,Ho,Wo,Cout].  Used by the per-layer parity tests.                */
int sad_debug_conv(sad_ctx* ctx, int head, int layer, const void* in_dev, const void* residual_dev, void* out_dev,
                   int B, int relu, void* stream);
/* Run ONE whole layer1 BasicBlock of one head (convs `layer` and `layer`+1 + identity, fused on CTA pairs through
 * peer shared memory): `in` / `out` NHWC bf16 [B,128,128,64].  Must equal two sad_debug_conv calls bit for bit.   */
int sad_debug_block(sad_ctx* ctx, int head, int layer, const void* in_dev, void* out_dev, int B, void* stream);
/* Run front end + fused stem/max-pool on pcm [B,128000] (B <= max_batch) and copy the pooled stem output, NHWC bf16
 * [H*B,128,128,64] (head-major), to out_dev.  Used by the stem parity test; sad_debug_read(which=0) then returns the
 * bf16 image the stem consumed.                                                                                  */
int sad_debug_stem(sad_ctx* ctx, const float* pcm_dev, int B, void* out_dev, void* stream);
/* Copy an internal activation of the last sad_forward* chunk to the caller (tests only).
 * which: 0 = image bf16 [B,512,512]; 2 = trunk (layer4) output bf16 NHWC [H*B,16,16,512]; 3 = per-head logits fp32
 * [H*B,2] (index 0 Real, 1 Synthetic); 4 = log-mel dB fp32 [B,128,251].  (1 is reserved: the pooled stem output is
 * overwritten by the trunk -- use sad_debug_stem.)  Returns the number of bytes written or a negative SAD_E* code. */
long long sad_debug_read(sad_ctx* ctx, int which, void* dst_dev, long long capacity_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAD_B200_H_ */
