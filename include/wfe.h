/*
 * wfe.h — C ABI of the B200-native Whisper log-mel frontend + seq2seq padding collator.
 *
 * This is the drop-in boundary for ONE hot path of asr4memory/asr-finetune.  The reference is
 * pure Python and has no FFI for this path; the "interface" each entry point replaces is the
 * Python call it stands in for (ref: = /root/reference, HF: = site-packages/transformers 5.5.0,
 * reference pins 4.46.3):
 *
 *   wfe_create            <- WhisperFeatureExtractor.__init__          HF:models/whisper/feature_extraction_whisper.py:69-103
 *                            built at ref:finetune/training/models/whisper_models.py:39,66
 *   wfe_logmel            <- WhisperFeatureExtractor.__call__ -> pad -> _torch_extract_fbank_features
 *                            HF:...feature_extraction_whisper.py:189-342, :135-164;
 *                            called at ref:finetune/training/data_and_collator/datasets_and_collators.py:191-195
 *                            and ref:finetune/prepare_dataset/materialize_dataset_ray.py:39-40
 *   wfe_clip_stats        <- zero_mean_unit_var_norm (do_normalize=True)  HF:...feature_extraction_whisper.py:168-187,306-312
 *   wfe_collate           <- DataCollatorSpeechSeq2SeqWithPadding.__call__
 *                            ref:finetune/training/data_and_collator/datasets_and_collators.py:418-461
 *                            and SimpleStreamingCollator._prepare_dataset  ...:229-256 (strip_bos = 0)
 *                            and HDF5Worker.process_sample label half
 *                            ref:finetune/prepare_dataset/materialize_dataset_ray.py:43-49 (width = 448)
 *   wfe_extract_host      <- the per-clip Python loop around the extractor with HOST buffers
 *                            ref:...datasets_and_collators.py:191-195 (+ the host<->device copies the
 *                            drop-in adds); pipelined pinned staging, H2D, kernels, D2H
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures (`stream` is a cudaStream_t
 *     passed as void*; NULL = the legacy default stream).
 *   - every function returns 0 on success, a negative wfe_status on failure; the message for the
 *     calling thread is available from wfe_last_error().  Nothing throws across the ABI.
 *   - "device" pointers must be valid on the handle's CUDA device.  The library never allocates or
 *     frees caller-visible buffers; the handle owns only its constant tables (and, for
 *     wfe_extract_host, its private pinned/device staging rings).
 *   - entry points are re-entrant: all per-call state lives in caller-provided scratch (wfe_extract_host* serialise
 *     on the handle's private staging ring).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     WFE_ERR_CUDA.
 */
#ifndef WFE_H_
#define WFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFE_ABI_VERSION 3

typedef struct wfe_handle wfe_handle;

typedef enum wfe_status {
  WFE_OK = 0,
  WFE_ERR_INVALID = -1,     /* bad argument (NULL pointer, negative size, ...) */
  WFE_ERR_UNSUPPORTED = -2, /* configuration the kernels are not built for (n_fft != 400, hop != 160, ...) */
  WFE_ERR_CUDA = -3,        /* CUDA runtime error, incl. "no device" */
  WFE_ERR_NOMEM = -4
} wfe_status;

typedef enum wfe_pcm_dtype {
  WFE_PCM_F32 = 0, /* float32 samples (what the reference decodes HDF5 audio to) */
  WFE_PCM_I16 = 1, /* int16 samples, converted on load as (float)x * pcm_scale */
  WFE_PCM_F16 = 2  /* IEEE float16 samples, widened exactly on load (pcm_scale ignored) */
} wfe_pcm_dtype;

/* Element type of `input_features`.  F32 is what the reference extractor returns; F16 / BF16 fuse the cast the
 * autocast consumer applies anyway (`fp16 = True`, ref:finetune/training/configs/largev3_debug.config:8): the value is
 * the fp32 result rounded to nearest-even once, in the kernel's epilogue. */
typedef enum wfe_out_dtype {
  WFE_OUT_F32 = 0,
  WFE_OUT_F16 = 1,
  WFE_OUT_BF16 = 2
} wfe_out_dtype;

/* Mirrors the constructor arguments of WhisperFeatureExtractor (HF:...feature_extraction_whisper.py:69-80). */
typedef struct wfe_config {
  int32_t n_mel;         /* feature_size: 80 (whisper-small) or 128 (large-v3); any 1..256 */
  int32_t n_fft;         /* must be 400 */
  int32_t hop_length;    /* must be 160 */
  int32_t n_samples;     /* chunk_length * sampling_rate = 480000; any value > 200 (n_frames = n_samples / 160, rounded down) */
  int32_t sampling_rate; /* 16000 (informational; checked by the Python shim like HF does) */
  int32_t device;        /* CUDA device ordinal (LOCAL_RANK) */
} wfe_config;

/* ---- lifecycle ------------------------------------------------------------------------------- */

/* mel_filters: HOST pointer, (n_fft/2+1, n_mel) row-major float32 — `fe.mel_filters.astype(float32)`,
 * the cast HF applies at use (HF:...feature_extraction_whisper.py:152).  The handle stores it in banded
 * form (per-mel lists of non-zeros, fp32); it must be sparse like every triangular mel bank: at most 1024 entries
 * after pairing adjacent mels (the Whisper banks need about 215). */
int wfe_create(const wfe_config* cfg, const float* mel_filters, wfe_handle** out);
void wfe_destroy(wfe_handle* h);
const char* wfe_last_error(void);
int wfe_abi_version(void);

/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
uint64_t wfe_launch_count(void);

/* ---- log-mel features, device-resident input ---------------------------------------------------- */

/* Bytes of device scratch wfe_logmel needs for `batch` clips (per-tile extrema for the per-clip clamp, tile scheduler counter, error word). */
size_t wfe_logmel_scratch_bytes(const wfe_handle* h, int32_t batch);
int32_t wfe_n_frames(const wfe_handle* h); /* n_samples / hop_length (3000) */

/*
 * pcm        device, ragged concatenation of the batch's clips (dtype per pcm_dtype)
 * offsets    device, int64: first sample of clip b in pcm.  With lengths == NULL it has batch+1 entries and clip b is
 *            pcm[offsets[b] .. offsets[b+1]); with lengths != NULL it has batch entries and clip b is
 *            pcm[offsets[b] .. offsets[b] + lengths[b]) (lets callers start every clip on a 16-byte boundary, which
 *            is what enables the kernel's 128-bit load path; any alignment is still correct).  A clip longer than
 *            n_samples is truncated, a shorter one is right-zero-padded
 *            (HF:feature_extraction_sequence_utils.py:265-278,327-332)
 * lengths    device, int64[batch] or NULL
 * norm_stats device, float2[batch] = (mean, 1/sqrt(var+1e-7)) from wfe_clip_stats, or NULL (do_normalize=False)
 * out        device, float32 (batch, n_mel, n_frames) C-contiguous  == BatchFeature["input_features"]; 8-byte aligned
 * attn_mask  device, int32 (batch, n_frames) or NULL                == BatchFeature["attention_mask"]
 * scratch    device, wfe_logmel_scratch_bytes(h, batch) bytes; contents need not be initialised
 */
int wfe_logmel(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
               const int64_t* lengths, int32_t batch, const float* norm_stats, float* out, int32_t* attn_mask,
               void* scratch, void* stream);

/* Same as wfe_logmel with `out` of element type out_dtype (wfe_out_dtype); `out` must be 16-byte aligned. */
int wfe_logmel_ex(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
                  const int64_t* lengths, int32_t batch, const float* norm_stats, void* out, int32_t out_dtype,
                  int32_t* attn_mask, void* scratch, void* stream);

/* Which kernel the handle's standard configuration runs on: 1 = tcgen05 tensor-core kernel (n_samples = 480000,
 * n_mel in {80, 128} with the slaney filter bank), 0 = CUDA-core kernel (any other geometry or filter bank).  The
 * environment variable WFE_DISABLE_TC=1 forces 0 (A/B timing). */
int32_t wfe_uses_tensor_cores(const wfe_handle* h);

/* After the stream has been synchronised: non-zero if a kernel of a previous wfe_logmel call on `scratch` gave up
 * waiting on an internal barrier (a bug, never expected).  `scratch`/`batch` as passed to that call. */
int32_t wfe_debug_scratch_error(wfe_handle* h, const void* scratch, int32_t batch);

/* Per-clip (mean, rstd) over the first min(len, n_samples) samples; stats: device float2[batch]. */
int wfe_clip_stats(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
                   const int64_t* lengths, int32_t batch, float* stats, void* stream);

/* ---- collator, device-resident ------------------------------------------------------------------ */

/*
 * ONE kernel launch that does the whole padding collator:
 *   - labels: right-pad each id list to `width` (host-computed batch max, or 448), fill padding with
 *     ignore_index (-100) by LENGTH (the true EOS == pad id survives), int64 (batch, width) row-major;
 *   - bos_flag (device int32[1]): set to 1 iff every row's first id == decoder_start_token_id; the caller
 *     returns labels[:, 1:] in that case (ref ...datasets_and_collators.py:456-457).  May be NULL.
 *   - features: if feat_srcs != NULL, gather batch separate device buffers of feat_elems floats each into
 *     the contiguous (batch, feat_elems) feat_out (the `feature_extractor.pad("longest")` stack, bit-exact).
 * ids/offsets follow the same ragged layout as pcm/offsets.  feat_srcs is a DEVICE array of device pointers.
 */
int wfe_collate(wfe_handle* h, const int64_t* ids, const int64_t* offsets, int32_t batch, int32_t width,
                int64_t decoder_start_token_id, int64_t ignore_index, int64_t* labels_out, int32_t* bos_flag,
                const float* const* feat_srcs, int64_t feat_elems, float* feat_out, void* stream);

/* ---- host-buffer entry point (what the Python drop-in calls for numpy input) ---------------------- */

/*
 * clips      HOST array of `batch` host pointers (pageable or pinned), dtype per pcm_dtype
 * lengths    HOST int64[batch], samples per clip
 * out        HOST float32 (batch, n_mel, n_frames); pinned memory avoids a staging copy.  May also be DEVICE memory
 *            of the handle's GPU (16-byte aligned): the kernels then write straight into it and nothing is
 *            downloaded -- the in-loop training consumer (host clips in, CUDA tensors out)
 * attn_mask  HOST (or DEVICE, like `out`) int32 (batch, n_frames) or NULL
 * do_normalize  0/1 (zero-mean unit-variance per clip before the STFT)
 * Work is cut into chunks of clips (16; a quarter of the batch below 64 clips; WFE_HOST_CHUNK overrides) and
 * pipelined: pinned staging -> H2D -> kernels -> D2H on the handle's private streams; pageable clips are staged by a
 * few persistent copy threads (WFE_HOST_THREADS, default min(8, cores / 2)).  Clips are uploaded straight from the
 * caller's memory only when the batch STARTS with a pinned clip (the driver is then asked about every clip).  Synchronous: returns when `out` is complete.  h2d_bytes/d2h_bytes (may be NULL) receive
 * the bytes that crossed PCIe.
 */
int wfe_extract_host(wfe_handle* h, const void* const* clips, const int64_t* lengths, int32_t batch,
                     int32_t pcm_dtype, float pcm_scale, int32_t do_normalize, float* out, int32_t* attn_mask,
                     uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* Same with `out` of element type out_dtype (halves the device->host bytes for F16 / BF16). */
int wfe_extract_host_ex(wfe_handle* h, const void* const* clips, const int64_t* lengths, int32_t batch,
                        int32_t pcm_dtype, float pcm_scale, int32_t do_normalize, void* out, int32_t out_dtype,
                        int32_t* attn_mask, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* WFE_H_ */
