// C ABI (include/wfe.h) over the sm_100a kernels.  No torch types, no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/wfe.h"
#include "wfe_collate.cuh"
#include "wfe_logmel.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define WFE_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      return fail(WFE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));               \
    }                                                                                               \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess) {
      ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

constexpr int kSlots = 3;

struct HostSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  bool busy = false;
  // staging
  void* h_in = nullptr;      // pinned, chunk * n_samples * 4 B
  float* h_out = nullptr;    // pinned, chunk * n_mel * n_frames floats
  int32_t* h_mask = nullptr;
  int64_t* h_off = nullptr;  // pinned, chunk + 1
  void* d_in = nullptr;
  float* d_out = nullptr;
  int32_t* d_mask = nullptr;
  int64_t* d_off = nullptr;
  void* d_scratch = nullptr;
  float* d_stats = nullptr;
  // pending finalisation (pageable destination)
  float* user_out = nullptr;
  int32_t* user_mask = nullptr;
  int pending_clips = 0;
};

}  // namespace

struct wfe_handle {
  wfe_config cfg;
  int n_frames = 0, ntiles = 0, nnz = 0;
  int2* d_mel_tab = nullptr;
  int32_t* d_mel_start = nullptr;
  std::mutex host_mu;
  bool ring_ready = false;
  int chunk_clips = 16;
  HostSlot slots[kSlots];
};

namespace {

using wfe::bin_to_row;

int upload_constants() {
  float win[wfe::kNFft];
  float2 tw[16 * 12];
  wfe::fill_tables(win, tw);
  WFE_CUDA(cudaMemcpyToSymbol(wfe::c_win, win, sizeof(win)));
  WFE_CUDA(cudaMemcpyToSymbol(wfe::c_tw400, tw, sizeof(tw)));
  return WFE_OK;
}

size_t scratch_bytes(const wfe_handle* h, int batch) {
  return (size_t)batch * (2 * sizeof(uint32_t) + (size_t)h->ntiles * sizeof(float));
}

template <typename T>
int launch_logmel(wfe_handle* h, const void* pcm, float scale, const int64_t* offsets, int batch, const float* norm,
                  float* out, int32_t* mask, void* scratch, cudaStream_t st) {
  static_assert(sizeof(T) == 2 || sizeof(T) == 4, "pcm dtype");
  wfe::LogmelParams p;
  p.pcm = pcm;
  p.offsets = offsets;
  p.norm = reinterpret_cast<const float2*>(norm);
  p.out = out;
  p.mask = mask;
  p.clip_key = reinterpret_cast<uint32_t*>(scratch);
  p.clip_ticket = p.clip_key + batch;
  p.tile_min = reinterpret_cast<float*>(p.clip_ticket + batch);
  p.mel_tab = h->d_mel_tab;
  p.mel_start = h->d_mel_start;
  p.pcm_scale = scale;
  p.n_mel = h->cfg.n_mel;
  p.n_samples = h->cfg.n_samples;
  p.n_frames = h->n_frames;
  p.ntiles = h->ntiles;
  WFE_CUDA(cudaMemsetAsync(scratch, 0, (size_t)batch * 2 * sizeof(uint32_t), st));
  // opt in to > 48 KB dynamic shared memory (cheap; per device context, so done on every launch)
  WFE_CUDA(cudaFuncSetAttribute(wfe::logmel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)wfe::kSmemBytes));
  const long long grid = (long long)batch * h->ntiles;
  if (grid > 2147483647LL) return fail(WFE_ERR_INVALID, "batch too large for one launch");
  wfe::logmel_kernel<T><<<(unsigned)grid, wfe::kThreads, wfe::kSmemBytes, st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  WFE_CUDA(cudaGetLastError());
  return WFE_OK;
}

int check_handle(const wfe_handle* h) {
  if (h == nullptr) return fail(WFE_ERR_INVALID, "null handle");
  return WFE_OK;
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

void free_ring(wfe_handle* h) {
  for (auto& s : h->slots) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.h_mask) cudaFreeHost(s.h_mask);
    if (s.h_off) cudaFreeHost(s.h_off);
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_mask) cudaFree(s.d_mask);
    if (s.d_off) cudaFree(s.d_off);
    if (s.d_scratch) cudaFree(s.d_scratch);
    if (s.d_stats) cudaFree(s.d_stats);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = HostSlot();
  }
  h->ring_ready = false;
}

int ensure_ring(wfe_handle* h) {
  if (h->ring_ready) return WFE_OK;
  const size_t c = (size_t)h->chunk_clips;
  const size_t in_bytes = c * h->cfg.n_samples * sizeof(float);
  const size_t out_elems = c * h->cfg.n_mel * h->n_frames;
  for (auto& s : h->slots) {
    WFE_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    WFE_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    WFE_CUDA(cudaHostAlloc(&s.h_in, in_bytes, cudaHostAllocDefault));
    WFE_CUDA(cudaHostAlloc((void**)&s.h_out, out_elems * sizeof(float), cudaHostAllocDefault));
    WFE_CUDA(cudaHostAlloc((void**)&s.h_mask, c * h->n_frames * sizeof(int32_t), cudaHostAllocDefault));
    WFE_CUDA(cudaHostAlloc((void**)&s.h_off, (c + 1) * sizeof(int64_t), cudaHostAllocDefault));
    WFE_CUDA(cudaMalloc(&s.d_in, in_bytes));
    WFE_CUDA(cudaMalloc((void**)&s.d_out, out_elems * sizeof(float)));
    WFE_CUDA(cudaMalloc((void**)&s.d_mask, c * h->n_frames * sizeof(int32_t)));
    WFE_CUDA(cudaMalloc((void**)&s.d_off, (c + 1) * sizeof(int64_t)));
    WFE_CUDA(cudaMalloc(&s.d_scratch, scratch_bytes(h, (int)c)));
    WFE_CUDA(cudaMalloc((void**)&s.d_stats, c * 2 * sizeof(float)));
  }
  h->ring_ready = true;
  return WFE_OK;
}

// wait for a slot's in-flight chunk and, if its D2H landed in staging, copy to the user's pageable buffer
int retire_slot(wfe_handle* h, HostSlot& s) {
  if (!s.busy) return WFE_OK;
  WFE_CUDA(cudaEventSynchronize(s.done));
  if (s.user_out) memcpy(s.user_out, s.h_out, (size_t)s.pending_clips * h->cfg.n_mel * h->n_frames * sizeof(float));
  if (s.user_mask) memcpy(s.user_mask, s.h_mask, (size_t)s.pending_clips * h->n_frames * sizeof(int32_t));
  s.user_out = nullptr;
  s.user_mask = nullptr;
  s.busy = false;
  return WFE_OK;
}

}  // namespace

extern "C" {

const char* wfe_last_error(void) { return g_err.c_str(); }
int wfe_abi_version(void) { return WFE_ABI_VERSION; }
uint64_t wfe_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int wfe_create(const wfe_config* cfg, const float* mel_filters, wfe_handle** out) {
  if (cfg == nullptr || mel_filters == nullptr || out == nullptr) return fail(WFE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_fft != wfe::kNFft || cfg->hop_length != wfe::kHop)
    return fail(WFE_ERR_UNSUPPORTED, "kernels are specialised for n_fft=400, hop_length=160 (every Whisper checkpoint)");
  if (cfg->n_mel < 1 || cfg->n_mel > 256) return fail(WFE_ERR_UNSUPPORTED, "n_mel must be in 1..256");
  if (cfg->n_samples < wfe::kNFft || cfg->n_samples % wfe::kHop != 0)
    return fail(WFE_ERR_UNSUPPORTED, "n_samples must be a multiple of 160 and >= 400");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(WFE_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(WFE_ERR_INVALID, "bad device ordinal");
  DeviceGuard guard(cfg->device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  WFE_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(WFE_ERR_UNSUPPORTED, "built for sm_100a (B200) only");

  wfe_handle* h = new (std::nothrow) wfe_handle();
  if (h == nullptr) return fail(WFE_ERR_NOMEM, "out of host memory");
  h->cfg = *cfg;
  h->n_frames = cfg->n_samples / wfe::kHop;
  h->ntiles = (h->n_frames + wfe::kTileF - 1) / wfe::kTileF;

  // banded (CSR-by-mel) filter bank; rows index the in-place power buffer
  std::vector<int2> tab;
  std::vector<int32_t> start(cfg->n_mel + 1, 0);
  for (int m = 0; m < cfg->n_mel; ++m) {
    start[m] = (int32_t)tab.size();
    for (int k = 0; k < wfe::kBins; ++k) {
      const float w = mel_filters[(size_t)k * cfg->n_mel + m];
      if (w != 0.0f) {
        int2 e;
        e.x = bin_to_row(k) * wfe::kTileF;
        memcpy(&e.y, &w, sizeof(float));
        tab.push_back(e);
      }
    }
  }
  start[cfg->n_mel] = (int32_t)tab.size();
  h->nnz = (int)tab.size();
  if (h->nnz > 4096) {
    delete h;
    return fail(WFE_ERR_UNSUPPORTED, "mel filter bank has more than 4096 non-zeros");
  }
  int rc = upload_constants();
  if (rc != WFE_OK) {
    delete h;
    return rc;
  }
  const size_t tab_bytes = (tab.empty() ? 1 : tab.size()) * sizeof(int2);
  if (cudaMalloc((void**)&h->d_mel_tab, tab_bytes) != cudaSuccess ||
      cudaMalloc((void**)&h->d_mel_start, start.size() * sizeof(int32_t)) != cudaSuccess) {
    wfe_destroy(h);
    return fail(WFE_ERR_NOMEM, "cudaMalloc failed for filter tables");
  }
  if (!tab.empty()) cudaMemcpy(h->d_mel_tab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_mel_start, start.data(), start.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    wfe_destroy(h);
    return fail(WFE_ERR_CUDA, std::string("table upload: ") + cudaGetErrorString(e));
  }
  *out = h;
  return WFE_OK;
}

void wfe_destroy(wfe_handle* h) {
  if (h == nullptr) return;
  DeviceGuard guard(h->cfg.device);
  free_ring(h);
  if (h->d_mel_tab) cudaFree(h->d_mel_tab);
  if (h->d_mel_start) cudaFree(h->d_mel_start);
  delete h;
}

size_t wfe_logmel_scratch_bytes(const wfe_handle* h, int32_t batch) {
  if (h == nullptr || batch < 0) return 0;
  return scratch_bytes(h, batch);
}

int32_t wfe_n_frames(const wfe_handle* h) { return h ? h->n_frames : 0; }

int wfe_logmel(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
               int32_t batch, const float* norm_stats, float* out, int32_t* attn_mask, void* scratch, void* stream) {
  if (check_handle(h)) return WFE_ERR_INVALID;
  if (batch < 0) return fail(WFE_ERR_INVALID, "negative batch");
  if (batch == 0) return WFE_OK;
  if (pcm == nullptr || offsets == nullptr || out == nullptr || scratch == nullptr)
    return fail(WFE_ERR_INVALID, "null device pointer");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (pcm_dtype) {
    case WFE_PCM_F32:
      return launch_logmel<float>(h, pcm, 1.0f, offsets, batch, norm_stats, out, attn_mask, scratch, st);
    case WFE_PCM_I16:
      return launch_logmel<int16_t>(h, pcm, pcm_scale, offsets, batch, norm_stats, out, attn_mask, scratch, st);
    default:
      return fail(WFE_ERR_INVALID, "unknown pcm_dtype");
  }
}

int wfe_clip_stats(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
                   int32_t batch, float* stats, void* stream) {
  if (check_handle(h)) return WFE_ERR_INVALID;
  if (batch < 0) return fail(WFE_ERR_INVALID, "negative batch");
  if (batch == 0) return WFE_OK;
  if (pcm == nullptr || offsets == nullptr || stats == nullptr) return fail(WFE_ERR_INVALID, "null device pointer");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pcm_dtype == WFE_PCM_F32)
    wfe::clip_stats_kernel<float><<<batch, 512, 0, st>>>(pcm, 1.0f, offsets, h->cfg.n_samples, reinterpret_cast<float2*>(stats));
  else if (pcm_dtype == WFE_PCM_I16)
    wfe::clip_stats_kernel<int16_t><<<batch, 512, 0, st>>>(pcm, pcm_scale, offsets, h->cfg.n_samples, reinterpret_cast<float2*>(stats));
  else
    return fail(WFE_ERR_INVALID, "unknown pcm_dtype");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  WFE_CUDA(cudaGetLastError());
  return WFE_OK;
}

int wfe_collate(wfe_handle* h, const int64_t* ids, const int64_t* offsets, int32_t batch, int32_t width,
                int64_t decoder_start_token_id, int64_t ignore_index, int64_t* labels_out, int32_t* bos_flag,
                const float* const* feat_srcs, int64_t feat_elems, float* feat_out, void* stream) {
  if (check_handle(h)) return WFE_ERR_INVALID;
  if (batch < 0 || width < 0 || feat_elems < 0) return fail(WFE_ERR_INVALID, "negative size");
  if (batch == 0) return WFE_OK;
  if (offsets == nullptr || (width > 0 && (ids == nullptr || labels_out == nullptr)))
    return fail(WFE_ERR_INVALID, "null label pointer");
  if (feat_srcs != nullptr && feat_out == nullptr) return fail(WFE_ERR_INVALID, "feat_out is null");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  wfe::CollateParams p;
  p.ids = ids;
  p.offsets = offsets;
  p.labels = labels_out;
  p.bos_flag = bos_flag;
  p.feat_srcs = feat_srcs;
  p.feat_out = feat_out;
  p.feat_elems = feat_elems;
  p.dec_start = decoder_start_token_id;
  p.ignore_index = ignore_index;
  p.batch = batch;
  p.width = width;
  const long long label_elems = (long long)batch * width;
  long long lb = (label_elems + wfe::kCollateThreads - 1) / wfe::kCollateThreads;
  if (lb < 1) lb = 1;
  if (lb > 148 * 4) lb = 148 * 4;
  p.label_blocks = (int)lb;
  // features: ~16 KB per block-iteration of 4 x 128-bit loads; aim for >= 2 waves of 148 SMs x 8 CTAs
  int per_clip = 0;
  if (feat_srcs != nullptr && feat_elems > 0) {
    long long want = (feat_elems / 4 + (long long)wfe::kCollateThreads * 4 - 1) / ((long long)wfe::kCollateThreads * 4);
    if (want < 1) want = 1;
    long long cap = (148LL * 16 + batch - 1) / batch;
    if (cap < 1) cap = 1;
    per_clip = (int)(want < cap ? want : cap);
  }
  p.feat_blocks_per_clip = per_clip > 0 ? per_clip : 1;
  const long long grid = lb + (long long)per_clip * batch;
  wfe::collate_kernel<<<(unsigned)grid, wfe::kCollateThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  WFE_CUDA(cudaGetLastError());
  return WFE_OK;
}

int wfe_extract_host(wfe_handle* h, const void* const* clips, const int64_t* lengths, int32_t batch,
                     int32_t pcm_dtype, float pcm_scale, int32_t do_normalize, float* out, int32_t* attn_mask,
                     uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (check_handle(h)) return WFE_ERR_INVALID;
  if (batch < 0) return fail(WFE_ERR_INVALID, "negative batch");
  if (h2d_bytes) *h2d_bytes = 0;
  if (d2h_bytes) *d2h_bytes = 0;
  if (batch == 0) return WFE_OK;
  if (clips == nullptr || lengths == nullptr || out == nullptr) return fail(WFE_ERR_INVALID, "null host pointer");
  if (pcm_dtype != WFE_PCM_F32 && pcm_dtype != WFE_PCM_I16) return fail(WFE_ERR_INVALID, "unknown pcm_dtype");
  const size_t es = pcm_dtype == WFE_PCM_F32 ? 4 : 2;
  std::lock_guard<std::mutex> lock(h->host_mu);
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  int rc = ensure_ring(h);
  if (rc != WFE_OK) return rc;

  const bool out_pinned = is_pinned_host(out);
  const bool mask_pinned = attn_mask != nullptr && is_pinned_host(attn_mask);
  const size_t clip_out = (size_t)h->cfg.n_mel * h->n_frames;
  uint64_t up = 0, down = 0;
  const int chunk = h->chunk_clips;
  int slot_i = 0;
  for (int c0 = 0; c0 < batch; c0 += chunk, slot_i = (slot_i + 1) % kSlots) {
    HostSlot& s = h->slots[slot_i];
    rc = retire_slot(h, s);
    if (rc != WFE_OK) return rc;
    const int n = (batch - c0 < chunk) ? batch - c0 : chunk;
    // ragged pack: only min(len, n_samples) samples of each clip cross PCIe
    int64_t pos = 0;
    for (int i = 0; i < n; ++i) {
      s.h_off[i] = pos;
      int64_t len = lengths[c0 + i];
      if (len < 0) return fail(WFE_ERR_INVALID, "negative clip length");
      if (len > h->cfg.n_samples) len = h->cfg.n_samples;
      if (len > 0 && clips[c0 + i] == nullptr) return fail(WFE_ERR_INVALID, "null clip pointer");
      pos += len;
    }
    s.h_off[n] = pos;
    // contiguous runs of pinned clips go straight from the caller's memory; everything else is staged
    int i = 0;
    while (i < n) {
      const int64_t len_i = s.h_off[i + 1] - s.h_off[i];
      if (len_i == 0) {
        ++i;
        continue;
      }
      const char* src = static_cast<const char*>(clips[c0 + i]);
      if (is_pinned_host(src)) {
        int j = i + 1;
        int64_t run = len_i;
        while (j < n && static_cast<const char*>(clips[c0 + j]) == src + (size_t)run * es &&
               lengths[c0 + j - 1] <= h->cfg.n_samples) {
          run += s.h_off[j + 1] - s.h_off[j];
          ++j;
        }
        WFE_CUDA(cudaMemcpyAsync(static_cast<char*>(s.d_in) + (size_t)s.h_off[i] * es, src, (size_t)run * es,
                                 cudaMemcpyHostToDevice, s.stream));
        i = j;
      } else {
        // stage a maximal run of pageable clips, then one H2D for the run
        const int i0 = i;
        while (i < n && !(s.h_off[i + 1] > s.h_off[i] && is_pinned_host(clips[c0 + i]))) {
          const int64_t l = s.h_off[i + 1] - s.h_off[i];
          if (l > 0) memcpy(static_cast<char*>(s.h_in) + (size_t)s.h_off[i] * es, clips[c0 + i], (size_t)l * es);
          ++i;
        }
        const size_t bytes = (size_t)(s.h_off[i] - s.h_off[i0]) * es;
        if (bytes)
          WFE_CUDA(cudaMemcpyAsync(static_cast<char*>(s.d_in) + (size_t)s.h_off[i0] * es,
                                   static_cast<char*>(s.h_in) + (size_t)s.h_off[i0] * es, bytes,
                                   cudaMemcpyHostToDevice, s.stream));
      }
    }
    up += (uint64_t)pos * es + (uint64_t)(n + 1) * sizeof(int64_t);
    WFE_CUDA(cudaMemcpyAsync(s.d_off, s.h_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
    const float* stats = nullptr;
    if (do_normalize) {
      rc = wfe_clip_stats(h, s.d_in, pcm_dtype, pcm_scale, s.d_off, n, s.d_stats, s.stream);
      if (rc != WFE_OK) return rc;
      stats = s.d_stats;
    }
    rc = wfe_logmel(h, s.d_in, pcm_dtype, pcm_scale, s.d_off, n, stats, s.d_out, attn_mask ? s.d_mask : nullptr,
                    s.d_scratch, s.stream);
    if (rc != WFE_OK) return rc;
    float* dst = out + (size_t)c0 * clip_out;
    if (out_pinned) {
      WFE_CUDA(cudaMemcpyAsync(dst, s.d_out, (size_t)n * clip_out * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    } else {
      WFE_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, (size_t)n * clip_out * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
      s.user_out = dst;
    }
    down += (uint64_t)n * clip_out * sizeof(float);
    if (attn_mask) {
      int32_t* mdst = attn_mask + (size_t)c0 * h->n_frames;
      if (mask_pinned) {
        WFE_CUDA(cudaMemcpyAsync(mdst, s.d_mask, (size_t)n * h->n_frames * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
      } else {
        WFE_CUDA(cudaMemcpyAsync(s.h_mask, s.d_mask, (size_t)n * h->n_frames * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        s.user_mask = mdst;
      }
      down += (uint64_t)n * h->n_frames * sizeof(int32_t);
    }
    s.pending_clips = n;
    WFE_CUDA(cudaEventRecord(s.done, s.stream));
    s.busy = true;
  }
  // drain in submission order
  for (int k = 0; k < kSlots; ++k, slot_i = (slot_i + 1) % kSlots) {
    rc = retire_slot(h, h->slots[slot_i]);
    if (rc != WFE_OK) return rc;
  }
  if (h2d_bytes) *h2d_bytes = up;
  if (d2h_bytes) *d2h_bytes = down;
  return WFE_OK;
}

}  // extern "C"
