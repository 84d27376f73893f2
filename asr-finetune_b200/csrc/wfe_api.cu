// C ABI (include/wfe.h) over the sm_100a kernels.  No torch types, no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

#include "../../include/wfe.h"
#include "wfe_collate.cuh"
#include "wfe_copy_pool.h"
#include "wfe_logmel.cuh"
#include "wfe_logmel_tc.cuh"

#define WFE_TC_GEN_HOST_TABLES 1
#define WFE_TC_GEN_NMEL 80
#include "wfe_tc_epilogue_gen.inc"
#undef WFE_TC_GEN_NMEL
#define WFE_TC_GEN_NMEL 128
#include "wfe_tc_epilogue_gen.inc"
#undef WFE_TC_GEN_NMEL
#undef WFE_TC_GEN_HOST_TABLES

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

using wfe_host::CopyPool;  // wfe_copy_pool.h: staging copies on a few persistent threads

int host_copy_threads() {
  if (const char* e = getenv("WFE_HOST_THREADS")) {
    const int v = atoi(e);
    if (v >= 1) return std::min(v, 64);
  }
  int cores = (int)std::thread::hardware_concurrency();
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
  return std::max(1, std::min(8, cores / 2));
}

struct HostSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  bool busy = false;
  // staging
  void* h_in = nullptr;      // pinned, chunk * n_samples * 4 B
  void* h_out = nullptr;     // pinned, chunk * n_mel * n_frames * 4 B
  int32_t* h_mask = nullptr;
  int64_t* h_off = nullptr;  // pinned, 2 * chunk: clip starts (16-byte aligned), then clip lengths
  void* d_in = nullptr;
  void* d_out = nullptr;
  int32_t* d_mask = nullptr;
  int64_t* d_off = nullptr;
  void* d_scratch = nullptr;
  float* d_stats = nullptr;
  int64_t* d_len = nullptr;
  // pending finalisation (pageable destination)
  void* user_out = nullptr;
  int32_t* user_mask = nullptr;
  int pending_clips = 0;
  size_t pending_out_bytes = 0;
};

}  // namespace

struct wfe_handle {
  wfe_config cfg;
  int n_frames = 0, ntiles = 0, n_groups = 0, n_rows = 0, sm_count = 0;
  int mel_wrange[wfe::kMelWarps + 1] = {0};
  int ctas_per_sm[3] = {0, 0, 0};  // resident CTAs of logmel_kernel<float>, <int16_t>, <__half> (set in wfe_create)
  bool tc_ok = false;                  // the tcgen05 kernel applies (n_samples = 480000, baked slaney filter bank)
  uint4* d_tc_b = nullptr;             // DFT-100 operand, fp16 hi / lo, canonical UMMA layout
  float4* d_tc_tw = nullptr;           // W400^(n1 k2) twiddles of the epilogue
  void* tc_encode = nullptr;           // cuTensorMapEncodeTiled, fetched through the runtime (no libcuda link)
  float4* d_s1_consts = nullptr;       // [8][25]
  float4* d_mel_tab = nullptr;           // [n_rows][2 halves]
  wfe::MelGroup* d_mel_groups = nullptr; // [n_groups]
  std::mutex host_mu;
  std::unique_ptr<CopyPool> pool;        // staging copies of wfe_extract_host (created with the ring)
  bool ring_ready = false;
  int chunk_clips = 16;
  HostSlot slots[kSlots];
};

namespace {

// words of scratch per clip: the CUDA-core kernel keeps one key per 32-frame tile (zero = not published yet); the tensor-
// core path one key per 128-frame tile + the minima of its kMinBlocks blocks.  Four more words follow the per-clip part:
// tile counter, error word (+ pad to 16 B).
size_t scratch_words_per_clip(const wfe_handle* h) {
  const size_t tc_words = (size_t)wfe::tc::kNTiles * (1 + wfe::tc::kMinBlocks);
  return (size_t)h->ntiles > tc_words ? (size_t)h->ntiles : tc_words;
}
size_t scratch_bytes(const wfe_handle* h, int batch) {
  return ((size_t)batch * scratch_words_per_clip(h) + 4) * sizeof(uint32_t);
}

template <typename T>
int prepare_cc_kernel(wfe_handle* h, int which) {
  // opt in to > 48 KB dynamic shared memory (the attribute is per function, shared by every handle: set it to the
  // most any filter bank can need) and size the persistent grid from the real occupancy
  WFE_CUDA(cudaFuncSetAttribute(wfe::logmel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)wfe::logmel_smem_bytes(wfe::kMaxMelRows)));
  WFE_CUDA(cudaFuncSetAttribute(wfe::logmel_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
  int n = 0;
  WFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wfe::logmel_kernel<T>, wfe::kThreads,
                                                         wfe::logmel_smem_bytes(h->n_rows)));
  if (n < 1) return fail(WFE_ERR_CUDA, "logmel kernel does not fit on an SM");
  h->ctas_per_sm[which] = n;
  return WFE_OK;
}

// CUDA-core kernel: any geometry, any banded filter bank, fp32 output
template <typename T>
int launch_cc(wfe_handle* h, const void* pcm, float scale, const int64_t* offsets, const int64_t* lengths, int batch,
              const float* norm, float* out, int32_t* mask, void* scratch, cudaStream_t st) {
  static_assert(sizeof(T) == 2 || sizeof(T) == 4, "pcm dtype");
  const long long total = (long long)batch * h->ntiles;
  if (total > 0x7fffffffLL) return fail(WFE_ERR_INVALID, "batch too large for one launch");
  if ((reinterpret_cast<uintptr_t>(out) & 7u) != 0) return fail(WFE_ERR_INVALID, "out must be 8-byte aligned");
  wfe::LogmelParams p;
  p.pcm = pcm;
  p.offsets = offsets;
  p.lengths = lengths;
  p.norm = reinterpret_cast<const float2*>(norm);
  p.out = out;
  p.mask = mask;
  p.tile_key = reinterpret_cast<uint32_t*>(scratch);
  p.tile_counter = p.tile_key + (size_t)batch * scratch_words_per_clip(h);
  p.s1_consts = h->d_s1_consts;
  p.mel_tab = h->d_mel_tab;
  p.mel_groups = h->d_mel_groups;
  for (int w = 0; w <= wfe::kMelWarps; ++w) p.mel_wrange[w] = h->mel_wrange[w];
  p.pcm_scale = scale;
  p.n_mel = h->cfg.n_mel;
  p.n_samples = h->cfg.n_samples;
  p.n_frames = h->n_frames;
  p.ntiles = h->ntiles;
  p.n_groups = h->n_groups;
  p.n_rows = h->n_rows;
  p.total_tiles = (uint32_t)total;
  WFE_CUDA(cudaMemsetAsync(scratch, 0, scratch_bytes(h, batch), st));
  const size_t smem = wfe::logmel_smem_bytes(h->n_rows);
  const int which = std::is_same<T, float>::value ? 0 : (std::is_same<T, int16_t>::value ? 1 : 2);
  long long grid = (long long)h->sm_count * h->ctas_per_sm[which];
  if (grid > total) grid = total;
  wfe::logmel_kernel<T><<<(unsigned)grid, wfe::kThreads, smem, st>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  WFE_CUDA(cudaGetLastError());
  return WFE_OK;
}

bool tc_disabled_by_env() {
  const char* e = getenv("WFE_DISABLE_TC");
  return e != nullptr && e[0] != 0 && e[0] != '0';
}

// tensor-core kernel: n_samples = 480000, baked filter bank; any PCM dtype, any output dtype
template <typename OutT, int kNMel>
int launch_tc(wfe_handle* h, const void* pcm, int pcm_dtype, float scale, const int64_t* offsets, const int64_t* lengths,
              int batch, const float* norm, void* out, int32_t* mask, void* scratch, cudaStream_t st) {
  const long long total = (long long)batch * wfe::tc::kNTiles;
  if (total > 0x7fffffffLL) return fail(WFE_ERR_INVALID, "batch too large for one launch");
  wfe::tc::TcParams p;
  p.pcm = pcm;
  p.offsets = offsets;
  p.lengths = lengths;
  p.norm = reinterpret_cast<const float2*>(norm);
  p.out = out;
  p.mask = mask;
  p.tile_key = reinterpret_cast<uint32_t*>(scratch);
  p.tile_min = p.tile_key + (size_t)batch * wfe::tc::kNTiles;  // [B][24][kMinBlocks]
  p.tile_counter = p.tile_key + (size_t)batch * scratch_words_per_clip(h);  // (cleared below, with the error word)
  p.b_mat = h->d_tc_b;
  p.tw = h->d_tc_tw;
  p.pcm_scale = scale;
  p.pcm_dtype = pcm_dtype;
  p.n_mel = kNMel;
  p.total_tiles = (uint32_t)total;
  uint32_t* err_flag = p.tile_key + (size_t)batch * scratch_words_per_clip(h) + 1;
  // 2-D view of the PCM for the raw-tile TMA: element (c, r) = pcm[c + 160 r]; a tile is the box (164 x 130) at
  // (first sample, 0): rows of 160 samples land on a 164-float pitch.  Only float32, 16-byte aligned PCM uses it; any
  // other input goes through the generic staging path and never touches the map (then it views a table of ours).
  const bool tma_ok = pcm_dtype == WFE_PCM_F32 && (reinterpret_cast<uintptr_t>(pcm) & 15u) == 0;
  CUtensorMap tmap;
  {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t gdim[2] = {(cuuint64_t)1 << 31, (cuuint64_t)wfe::tc::kRawRows};
    const cuuint64_t gstride[1] = {(cuuint64_t)wfe::kHop * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)wfe::tc::kRawPitch, (cuuint32_t)wfe::tc::kRawRows};
    const cuuint32_t estr[2] = {1, 1};
    void* base = tma_ok ? const_cast<void*>(pcm) : static_cast<void*>(h->d_tc_b);
    const CUresult cr = reinterpret_cast<EncodeFn>(h->tc_encode)(
        &tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(WFE_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")");
  }
  if (!tma_ok && pcm_dtype == WFE_PCM_F32) p.pcm_dtype = 3;  // float32 at an odd address: generic staging only
  // (every tile word is written by the main kernel before the clamp pass reads it: only the error word needs clearing)
  WFE_CUDA(cudaMemsetAsync(p.tile_key + (size_t)batch * scratch_words_per_clip(h), 0, 4 * sizeof(uint32_t), st));
  long long grid = h->sm_count;
  if (grid > total) grid = total;
  wfe::tc::logmel_tc_kernel<OutT, kNMel><<<(unsigned)grid, wfe::tc::kThreads, wfe::tc::kSmemBytes, st>>>(p, tmap, err_flag);
  WFE_CUDA(cudaGetLastError());
  // second pass: the per-clip clamp, over the tiles that have something below their clip's floor (often none)
#ifndef WFE_CLAMP_CTAS_PER_SM
#define WFE_CLAMP_CTAS_PER_SM 16  // (8 and 4 measured 1-2 % slower on speech-like input)
#endif
  long long cgrid = (long long)h->sm_count * WFE_CLAMP_CTAS_PER_SM;
  if (cgrid > total) cgrid = total;
  wfe::tc::clamp_kernel<OutT><<<(unsigned)cgrid, wfe::tc::kClampThreads, 0, st>>>(
      reinterpret_cast<OutT*>(out), p.tile_key, p.tile_min, kNMel, (uint32_t)total);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  WFE_CUDA(cudaGetLastError());
  return WFE_OK;
}

template <typename OutT>
__global__ void cast_kernel(const float* __restrict__ src, OutT* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = wfe::tc::to_out<OutT>(src[i]);
}

// The kernel re-partitions its registers with setmaxnreg; the budget (wfe_logmel_tc.cuh: kRegs*) assumes that it was
// compiled to launch with exactly kRegsLaunch registers per thread -- a mismatch would hang, so it is checked here.
template <typename OutT, int kNMel>
int prepare_tc_kernel() {
  WFE_CUDA(cudaFuncSetAttribute(wfe::tc::logmel_tc_kernel<OutT, kNMel>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)wfe::tc::kSmemBytes));
  cudaFuncAttributes fa;
  WFE_CUDA(cudaFuncGetAttributes(&fa, wfe::tc::logmel_tc_kernel<OutT, kNMel>));
  if (fa.numRegs != wfe::tc::kRegsLaunch)
    return fail(WFE_ERR_UNSUPPORTED, "tensor-core kernel compiled with " + std::to_string(fa.numRegs) +
                                         " registers per thread, its setmaxnreg budget needs " +
                                         std::to_string(wfe::tc::kRegsLaunch));
  return WFE_OK;
}
template <typename OutT>
int prepare_tc_kernels() {
  const int rc = prepare_tc_kernel<OutT, 80>();
  return rc != WFE_OK ? rc : prepare_tc_kernel<OutT, 128>();
}

int logmel_dispatch(wfe_handle* h, const void* pcm, int pcm_dtype, float scale, const int64_t* offsets,
                    const int64_t* lengths, int batch, const float* norm, void* out, int out_dtype, int32_t* mask,
                    void* scratch, cudaStream_t st) {
  if (pcm_dtype != WFE_PCM_F32 && pcm_dtype != WFE_PCM_I16 && pcm_dtype != WFE_PCM_F16)
    return fail(WFE_ERR_INVALID, "unknown pcm_dtype");
  if (out_dtype != WFE_OUT_F32 && out_dtype != WFE_OUT_F16 && out_dtype != WFE_OUT_BF16)
    return fail(WFE_ERR_INVALID, "unknown out_dtype");
  if ((reinterpret_cast<uintptr_t>(scratch) & 3u) != 0) return fail(WFE_ERR_INVALID, "scratch must be 4-byte aligned");
  if (pcm_dtype == WFE_PCM_F32) scale = 1.0f;
  const bool tc = h->tc_ok && !tc_disabled_by_env() && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
#ifdef WFE_EXP_MINIMAL  // experiment builds: one instantiation only (fast to compile)
  if (tc && out_dtype == WFE_OUT_F32 && h->cfg.n_mel == 128)
    return launch_tc<float, 128>(h, pcm, pcm_dtype, scale, offsets, lengths, batch, norm, out, mask, scratch, st);
  if (tc) return fail(WFE_ERR_UNSUPPORTED, "minimal experiment build");
#else
  if (tc) {
    const bool big = h->cfg.n_mel == 128;
    switch (out_dtype) {
      case WFE_OUT_F32:
        return big ? launch_tc<float, 128>(h, pcm, pcm_dtype, scale, offsets, lengths, batch, norm, out, mask, scratch, st)
                   : launch_tc<float, 80>(h, pcm, pcm_dtype, scale, offsets, lengths, batch, norm, out, mask, scratch, st);
      case WFE_OUT_F16:
        return big ? launch_tc<__half, 128>(h, pcm, pcm_dtype, scale, offsets, lengths, batch, norm, out, mask, scratch, st)
                   : launch_tc<__half, 80>(h, pcm, pcm_dtype, scale, offsets, lengths, batch, norm, out, mask, scratch, st);
      default:
        return big ? launch_tc<__nv_bfloat16, 128>(h, pcm, pcm_dtype, scale, offsets, lengths, batch, norm, out, mask, scratch, st)
                   : launch_tc<__nv_bfloat16, 80>(h, pcm, pcm_dtype, scale, offsets, lengths, batch, norm, out, mask, scratch, st);
    }
  }
#endif
  // CUDA-core kernel writes fp32; a 16-bit result takes a stream-ordered temporary and one rounding pass
  float* out32 = reinterpret_cast<float*>(out);
  const size_t n_out = (size_t)batch * h->cfg.n_mel * h->n_frames;
  if (out_dtype != WFE_OUT_F32) WFE_CUDA(cudaMallocAsync((void**)&out32, n_out * sizeof(float), st));
  int rc;
  if (pcm_dtype == WFE_PCM_F32)
    rc = launch_cc<float>(h, pcm, scale, offsets, lengths, batch, norm, out32, mask, scratch, st);
  else if (pcm_dtype == WFE_PCM_I16)
    rc = launch_cc<int16_t>(h, pcm, scale, offsets, lengths, batch, norm, out32, mask, scratch, st);
  else
    rc = launch_cc<__half>(h, pcm, scale, offsets, lengths, batch, norm, out32, mask, scratch, st);
  if (out_dtype != WFE_OUT_F32) {
    if (rc == WFE_OK) {
      const unsigned blocks = (unsigned)std::min<size_t>((n_out + 255) / 256, (size_t)h->sm_count * 16);
      if (out_dtype == WFE_OUT_F16)
        cast_kernel<__half><<<blocks, 256, 0, st>>>(out32, reinterpret_cast<__half*>(out), n_out);
      else
        cast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(out32, reinterpret_cast<__nv_bfloat16*>(out), n_out);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    cudaFreeAsync(out32, st);
    if (rc == WFE_OK) WFE_CUDA(cudaGetLastError());
  }
  return rc;
}

int check_handle(const wfe_handle* h) {
  if (h == nullptr) return fail(WFE_ERR_INVALID, "null handle");
  return WFE_OK;
}

// device memory of the handle's GPU (cudaMalloc / a torch CUDA tensor)?
bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice;
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
    if (s.d_len) cudaFree(s.d_len);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = HostSlot();
  }
  h->ring_ready = false;
}

int ensure_ring(wfe_handle* h) {
  if (h->ring_ready) return WFE_OK;
  const size_t c = (size_t)h->chunk_clips;
  const size_t in_bytes = c * ((size_t)h->cfg.n_samples + 8) * sizeof(float);
  const size_t out_elems = c * h->cfg.n_mel * h->n_frames;
  for (auto& s : h->slots) {
    WFE_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    WFE_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    WFE_CUDA(cudaHostAlloc(&s.h_in, in_bytes, cudaHostAllocDefault));
    WFE_CUDA(cudaHostAlloc(&s.h_out, out_elems * sizeof(float), cudaHostAllocDefault));
    WFE_CUDA(cudaHostAlloc((void**)&s.h_mask, c * h->n_frames * sizeof(int32_t), cudaHostAllocDefault));
    WFE_CUDA(cudaHostAlloc((void**)&s.h_off, 2 * c * sizeof(int64_t), cudaHostAllocDefault));
    WFE_CUDA(cudaMalloc(&s.d_in, in_bytes));
    WFE_CUDA(cudaMalloc(&s.d_out, out_elems * sizeof(float)));
    WFE_CUDA(cudaMalloc((void**)&s.d_mask, c * h->n_frames * sizeof(int32_t)));
    WFE_CUDA(cudaMalloc((void**)&s.d_off, 2 * c * sizeof(int64_t)));
    WFE_CUDA(cudaMalloc(&s.d_scratch, scratch_bytes(h, (int)c)));
    WFE_CUDA(cudaMalloc((void**)&s.d_stats, c * 2 * sizeof(float)));
  }
  if (!h->pool) h->pool.reset(new (std::nothrow) CopyPool(host_copy_threads() - 1));
  if (!h->pool) return fail(WFE_ERR_NOMEM, "out of host memory");
  h->ring_ready = true;
  return WFE_OK;
}

// wait for a slot's in-flight chunk and, if its D2H landed in staging, copy to the user's pageable buffer
int retire_slot(wfe_handle* h, HostSlot& s) {
  if (!s.busy) return WFE_OK;
  WFE_CUDA(cudaEventSynchronize(s.done));
  if (s.user_out) h->pool->run({{static_cast<char*>(s.user_out), static_cast<const char*>(s.h_out), s.pending_out_bytes}});
  if (s.user_mask) memcpy(s.user_mask, s.h_mask, (size_t)s.pending_clips * h->n_frames * sizeof(int32_t));
  s.user_out = nullptr;
  s.user_mask = nullptr;
  s.busy = false;
  return WFE_OK;
}

// Tables of the tcgen05 kernel (wfe_logmel_tc.cuh).  The kernel's mel epilogue is generated code with the slaney
// weights baked in: it only applies when the caller's filter bank has exactly those non-zeros.
int setup_tensor_core_path(wfe_handle* h, const float* mel_filters) {
  h->tc_ok = false;
  const int n_mel = h->cfg.n_mel;
  if (h->cfg.n_samples != wfe::tc::kNSamples || (n_mel != 80 && n_mel != 128)) return WFE_OK;
  const uint32_t(*nnz)[3] = n_mel == 80 ? kTcNnz80 : kTcNnz128;
  const size_t n_nnz = n_mel == 80 ? sizeof(kTcNnz80) / sizeof(kTcNnz80[0]) : sizeof(kTcNnz128) / sizeof(kTcNnz128[0]);
  size_t found = 0;
  for (int k = 0; k < wfe::kBins; ++k)
    for (int m = 0; m < n_mel; ++m)
      if (mel_filters[(size_t)k * n_mel + m] != 0.0f) ++found;
  if (found != n_nnz) return WFE_OK;
  for (size_t i = 0; i < n_nnz; ++i) {
    uint32_t bits;
    memcpy(&bits, &mel_filters[(size_t)nnz[i][0] * n_mel + nnz[i][1]], 4);
    if (bits != nnz[i][2]) return WFE_OK;
  }
  using namespace wfe::tc;
  const double kPi = 3.14159265358979323846;
  // B[n][k]: n = 4 * pair + {re k2 = 2 pair, re k2 + 1, im k2, im k2 + 1}, k = n2; canonical layout [chunk][n][8 x fp16]
  std::vector<__half> bmat((size_t)kBBytes / 2, __float2half_rn(0.0f));
  for (int n = 0; n < kN; ++n) {
    const int pair = n / 4, c = n % 4;
    const int k2 = 2 * pair + (c & 1);
    if (pair >= 26 || k2 > 50) continue;
    for (int n2 = 0; n2 < 100; ++n2) {
      const double ang = 2.0 * kPi * (double)((n2 * k2) % 100) / 100.0;
      const float v = (float)(c < 2 ? cos(ang) : -sin(ang));
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const size_t idx = ((size_t)(n2 / 8) * kN + n) * 8 + (n2 % 8);
      bmat[idx] = hi;
      bmat[(size_t)kBBytes / 4 + idx] = lo;
    }
  }
  std::vector<float> tw((size_t)kTwBytes / 4);
  for (int pair = 0; pair < 26; ++pair)
    for (int n1 = 1; n1 <= 3; ++n1) {
      float* t = &tw[((size_t)pair * 3 + (n1 - 1)) * 4];
      for (int e2 = 0; e2 < 2; ++e2) {
        const double ang = 2.0 * kPi * (double)(n1 * (2 * pair + e2)) / 400.0;
        t[e2] = (float)cos(ang);
        t[2 + e2] = (float)sin(ang);
      }
    }
  {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return WFE_OK;  // driver without tensor maps: stay on the CUDA-core kernel
    }
    h->tc_encode = fn;
  }
  if (cudaMalloc((void**)&h->d_tc_b, kBBytes) != cudaSuccess || cudaMalloc((void**)&h->d_tc_tw, kTwBytes) != cudaSuccess)
    return fail(WFE_ERR_NOMEM, "cudaMalloc failed for the tensor-core tables");
  WFE_CUDA(cudaMemcpy(h->d_tc_b, bmat.data(), kBBytes, cudaMemcpyHostToDevice));
  WFE_CUDA(cudaMemcpy(h->d_tc_tw, tw.data(), kTwBytes, cudaMemcpyHostToDevice));
#ifdef WFE_EXP_MINIMAL
  int rc = prepare_tc_kernel<float, 128>();
#else
  int rc = prepare_tc_kernels<float>();
  if (rc == WFE_OK) rc = prepare_tc_kernels<__half>();
  if (rc == WFE_OK) rc = prepare_tc_kernels<__nv_bfloat16>();
#endif
  if (rc != WFE_OK) return rc;
  h->tc_ok = true;
  return WFE_OK;
}

}  // namespace

extern "C" {

#if WFE_EXP & 256
// timing-trace build only: copy the device trace buffer out (not part of the shipped ABI)
int wfe_debug_read_trace(unsigned long long* dst, int n) {
  return (int)cudaMemcpyFromSymbol(dst, wfe::g_trace, sizeof(unsigned long long) * n);
}
#endif

#ifdef WFE_TC_TRACE
// timing-trace build only (not part of the shipped ABI)
int wfe_debug_read_tc_trace(unsigned long long* dst, int n) {
  return (int)cudaMemcpyFromSymbol(dst, wfe::tc::g_tc_trace, sizeof(unsigned long long) * n);
}
int wfe_debug_read_tc_warps(unsigned long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, wfe::tc::g_tc_warps, sizeof(unsigned long long) * 128);
}
int wfe_debug_read_tc_cta(unsigned long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, wfe::tc::g_tc_cta, sizeof(unsigned long long) * 160 * 4);
}
int wfe_debug_read_tc_tiles(unsigned long long* dst) {
  return (int)cudaMemcpyFromSymbol(dst, wfe::tc::g_tc_tiles, sizeof(unsigned long long) * 3 * 64);
}
#endif

const char* wfe_last_error(void) { return g_err.c_str(); }
int wfe_abi_version(void) { return WFE_ABI_VERSION; }
uint64_t wfe_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int wfe_create(const wfe_config* cfg, const float* mel_filters, wfe_handle** out) {
  if (cfg == nullptr || mel_filters == nullptr || out == nullptr) return fail(WFE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_fft != wfe::kNFft || cfg->hop_length != wfe::kHop)
    return fail(WFE_ERR_UNSUPPORTED, "kernels are specialised for n_fft=400, hop_length=160 (every Whisper checkpoint)");
  if (cfg->n_mel < 1 || cfg->n_mel > 256) return fail(WFE_ERR_UNSUPPORTED, "n_mel must be in 1..256");
  if (cfg->n_samples <= wfe::kNFft / 2)
    return fail(WFE_ERR_UNSUPPORTED, "n_samples must exceed n_fft / 2 = 200 (centred reflect pad)");
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

  if (h->ntiles > wfe::kRing) {
    delete h;
    return fail(WFE_ERR_UNSUPPORTED, "n_samples too long: at most 128 tiles of 32 frames (40.96 s) per clip");
  }
  h->sm_count = prop.multiProcessorCount;

  // ---- mel projection tables (banded filter bank).  Mels are taken two at a time (m, m+1: one per half-warp); pairs
  // are sorted by band length and cut into groups of four "slots" that run in lock step over a common number of table
  // rows, so each thread carries four independent accumulators.  Every (slot, half) filter is treated as a contiguous
  // band of `trips` power rows [lo, lo + trips) (weights taken straight from the dense filter matrix, so zeros inside or
  // beyond the filter cost nothing extra); a table row holds, per half, the four weights of the slots.
  const int n_mel = cfg->n_mel;
  struct Band {
    int lo, len;
  };
  std::vector<Band> band(n_mel + 1, Band{0, 0});
  for (int m = 0; m < n_mel; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < wfe::kBins; ++k) {
      const float w = mel_filters[(size_t)k * n_mel + m];
      if (!(w >= 0.0f)) {  // the tile extrema are tracked on the bit patterns of the mel powers, which must be >= 0
        delete h;
        return fail(WFE_ERR_UNSUPPORTED, "mel_filters must be non-negative and finite");
      }
      if (w != 0.0f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    if (first >= 0) band[m] = Band{first, last - first + 1};
  }
  struct PairN {
    int m, n;
  };
  std::vector<PairN> prs;
  for (int m = 0; m < n_mel; m += 2) prs.push_back({m, std::max(band[m].len, band[m + 1].len)});
  std::sort(prs.begin(), prs.end(), [](const PairN& a, const PairN& b) { return a.n != b.n ? a.n > b.n : a.m < b.m; });
  struct GroupCost {
    int first, trips, cost;  // first pair (index into prs), table rows, issue-slot estimate
  };
  std::vector<GroupCost> gpool;
  for (size_t i = 0; i < prs.size(); i += 4) {
    const int trips = std::max(2, (prs[i].n + 1) & ~1);  // even: the loop takes two rows per step
    gpool.push_back({(int)i, trips, 10 * trips + 60});
  }
  std::vector<std::vector<GroupCost>> per_warp(wfe::kMelWarps);
  int load[wfe::kMelWarps] = {0};
  for (const GroupCost& gc : gpool) {  // longest-processing-time first (gpool is already sorted by cost, descending)
    int w = 0;
    for (int i = 1; i < wfe::kMelWarps; ++i)
      if (load[i] < load[w]) w = i;
    per_warp[w].push_back(gc);
    load[w] += gc.cost;
  }
  std::vector<float4> mtab;
  std::vector<wfe::MelGroup> groups;
  for (int w = 0; w < wfe::kMelWarps; ++w) {
    h->mel_wrange[w] = (int)groups.size();
    for (const GroupCost& gc : per_warp[w]) {
      wfe::MelGroup g;
      memset(&g, 0, sizeof(g));
      g.trips = gc.trips;
      g.tab_idx = (int32_t)mtab.size();
      int mel_of[4], lo_of[2][4];
      for (int sl = 0; sl < 4; ++sl) {
        const bool real = gc.first + sl < (int)prs.size();
        mel_of[sl] = real ? prs[gc.first + sl].m : -1;
        g.out_off[sl] = real ? mel_of[sl] * h->n_frames : 0;
        if (real) g.valid |= 1 << (2 * sl);
        if (real && mel_of[sl] + 1 < n_mel) g.valid |= 1 << (2 * sl + 1);
        for (int half = 0; half < 2; ++half) {
          // the band window must stay inside the 201 power rows of THIS tile (rows beyond hold stale data)
          const int lo = real ? band[mel_of[sl] + half].lo : 0;
          lo_of[half][sl] = std::max(0, std::min(lo, wfe::kBins - gc.trips));
          g.lo_off[half][sl] = lo_of[half][sl] * wfe::kPStride;
        }
      }
      for (int i = 0; i < gc.trips; ++i)
        for (int half = 0; half < 2; ++half) {
          float wgt[4];
          for (int sl = 0; sl < 4; ++sl) {
            const int m = mel_of[sl] >= 0 ? mel_of[sl] + half : -1;
            const int k = lo_of[half][sl] + i;
            wgt[sl] = (m >= 0 && m < n_mel && k < wfe::kBins) ? mel_filters[(size_t)k * n_mel + m] : 0.0f;
          }
          mtab.push_back(make_float4(wgt[0], wgt[1], wgt[2], wgt[3]));
        }
      groups.push_back(g);
    }
  }
  h->mel_wrange[wfe::kMelWarps] = (int)groups.size();
  h->n_groups = (int)groups.size();
  h->n_rows = (int)(mtab.size() / 2);
  if (h->n_groups > wfe::kMaxMelGroups || h->n_rows > wfe::kMaxMelRows) {
    delete h;
    return fail(WFE_ERR_UNSUPPORTED, "mel filter bank has too many non-zeros (not banded)");
  }
  std::vector<float> s1c(8 * wfe::kS1ConstVec * 4);
  wfe::fill_stage1_consts(s1c.data());
  if (cudaMalloc((void**)&h->d_s1_consts, s1c.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void**)&h->d_mel_tab, mtab.size() * sizeof(float4)) != cudaSuccess ||
      cudaMalloc((void**)&h->d_mel_groups, groups.size() * sizeof(wfe::MelGroup)) != cudaSuccess) {
    wfe_destroy(h);
    return fail(WFE_ERR_NOMEM, "cudaMalloc failed for constant tables");
  }
  cudaError_t e = cudaMemcpy(h->d_s1_consts, s1c.data(), s1c.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_mel_tab, mtab.data(), mtab.size() * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(h->d_mel_groups, groups.data(), groups.size() * sizeof(wfe::MelGroup), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    wfe_destroy(h);
    return fail(WFE_ERR_CUDA, std::string("table upload: ") + cudaGetErrorString(e));
  }
  // function attributes and occupancy once, here, so that wfe_logmel is re-entrant
  int rc = prepare_cc_kernel<float>(h, 0);
  if (rc == WFE_OK) rc = prepare_cc_kernel<int16_t>(h, 1);
  if (rc == WFE_OK) rc = prepare_cc_kernel<__half>(h, 2);
  if (rc == WFE_OK) rc = setup_tensor_core_path(h, mel_filters);
  if (rc != WFE_OK) {
    const std::string msg = g_err;
    wfe_destroy(h);
    return fail(rc, msg);
  }
  *out = h;
  return WFE_OK;
}

void wfe_destroy(wfe_handle* h) {
  if (h == nullptr) return;
  DeviceGuard guard(h->cfg.device);
  free_ring(h);
  if (h->d_s1_consts) cudaFree(h->d_s1_consts);
  if (h->d_mel_tab) cudaFree(h->d_mel_tab);
  if (h->d_mel_groups) cudaFree(h->d_mel_groups);
  if (h->d_tc_b) cudaFree(h->d_tc_b);
  if (h->d_tc_tw) cudaFree(h->d_tc_tw);
  delete h;
}

size_t wfe_logmel_scratch_bytes(const wfe_handle* h, int32_t batch) {
  if (h == nullptr || batch < 0) return 0;
  return scratch_bytes(h, batch);
}

int32_t wfe_n_frames(const wfe_handle* h) { return h ? h->n_frames : 0; }

int wfe_logmel_ex(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
                  const int64_t* lengths, int32_t batch, const float* norm_stats, void* out, int32_t out_dtype,
                  int32_t* attn_mask, void* scratch, void* stream) {
  if (check_handle(h)) return WFE_ERR_INVALID;
  if (batch < 0) return fail(WFE_ERR_INVALID, "negative batch");
  if (batch == 0) return WFE_OK;
  if (pcm == nullptr || offsets == nullptr || out == nullptr || scratch == nullptr)
    return fail(WFE_ERR_INVALID, "null device pointer");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  return logmel_dispatch(h, pcm, pcm_dtype, pcm_scale, offsets, lengths, batch, norm_stats, out, out_dtype, attn_mask,
                         scratch, reinterpret_cast<cudaStream_t>(stream));
}

int wfe_logmel(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
               const int64_t* lengths, int32_t batch, const float* norm_stats, float* out, int32_t* attn_mask,
               void* scratch, void* stream) {
  return wfe_logmel_ex(h, pcm, pcm_dtype, pcm_scale, offsets, lengths, batch, norm_stats, out, WFE_OUT_F32, attn_mask,
                       scratch, stream);
}

int32_t wfe_uses_tensor_cores(const wfe_handle* h) { return (h != nullptr && h->tc_ok && !tc_disabled_by_env()) ? 1 : 0; }

int32_t wfe_debug_scratch_error(wfe_handle* h, const void* scratch, int32_t batch) {
  if (h == nullptr || scratch == nullptr || batch <= 0) return 0;
  DeviceGuard guard(h->cfg.device);
  uint32_t v = 0;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(scratch) + (size_t)batch * scratch_words_per_clip(h) + 1;
  if (cudaMemcpy(&v, w, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int32_t)v;
}

int wfe_clip_stats(wfe_handle* h, const void* pcm, int32_t pcm_dtype, float pcm_scale, const int64_t* offsets,
                   const int64_t* lengths, int32_t batch, float* stats, void* stream) {
  if (check_handle(h)) return WFE_ERR_INVALID;
  if (batch < 0) return fail(WFE_ERR_INVALID, "negative batch");
  if (batch == 0) return WFE_OK;
  if (pcm == nullptr || offsets == nullptr || stats == nullptr) return fail(WFE_ERR_INVALID, "null device pointer");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pcm_dtype == WFE_PCM_F32)
    wfe::clip_stats_kernel<float><<<batch, 512, 0, st>>>(pcm, 1.0f, offsets, lengths, h->cfg.n_samples, reinterpret_cast<float2*>(stats));
  else if (pcm_dtype == WFE_PCM_I16)
    wfe::clip_stats_kernel<int16_t><<<batch, 512, 0, st>>>(pcm, pcm_scale, offsets, lengths, h->cfg.n_samples, reinterpret_cast<float2*>(stats));
  else if (pcm_dtype == WFE_PCM_F16)
    wfe::clip_stats_kernel<__half><<<batch, 512, 0, st>>>(pcm, 1.0f, offsets, lengths, h->cfg.n_samples, reinterpret_cast<float2*>(stats));
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
  if (lb > (long long)h->sm_count * 4) lb = (long long)h->sm_count * 4;
  p.label_blocks = (int)lb;
  // features: ~16 KB per block-iteration of 4 x 128-bit loads; aim for >= 2 waves of 148 SMs x 8 CTAs
  int per_clip = 0;
  if (feat_srcs != nullptr && feat_elems > 0) {
    long long want = (feat_elems / 4 + (long long)wfe::kCollateThreads * 4 - 1) / ((long long)wfe::kCollateThreads * 4);
    if (want < 1) want = 1;
    long long cap = ((long long)h->sm_count * 16 + batch - 1) / batch;
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

int wfe_extract_host_ex(wfe_handle* h, const void* const* clips, const int64_t* lengths, int32_t batch,
                        int32_t pcm_dtype, float pcm_scale, int32_t do_normalize, void* out, int32_t out_dtype,
                        int32_t* attn_mask, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (check_handle(h)) return WFE_ERR_INVALID;
  if (batch < 0) return fail(WFE_ERR_INVALID, "negative batch");
  if (h2d_bytes) *h2d_bytes = 0;
  if (d2h_bytes) *d2h_bytes = 0;
  if (batch == 0) return WFE_OK;
  if (clips == nullptr || lengths == nullptr || out == nullptr) return fail(WFE_ERR_INVALID, "null host pointer");
  if (pcm_dtype != WFE_PCM_F32 && pcm_dtype != WFE_PCM_I16 && pcm_dtype != WFE_PCM_F16)
    return fail(WFE_ERR_INVALID, "unknown pcm_dtype");
  if (out_dtype != WFE_OUT_F32 && out_dtype != WFE_OUT_F16 && out_dtype != WFE_OUT_BF16)
    return fail(WFE_ERR_INVALID, "unknown out_dtype");
  // validate every clip BEFORE anything is enqueued: an early return must not leave copies in flight
  for (int i = 0; i < batch; ++i) {
    if (lengths[i] < 0) return fail(WFE_ERR_INVALID, "negative clip length");
    if (lengths[i] > 0 && clips[i] == nullptr) return fail(WFE_ERR_INVALID, "null clip pointer");
  }
  const size_t es = pcm_dtype == WFE_PCM_F32 ? 4 : 2;
  const size_t os = out_dtype == WFE_OUT_F32 ? 4 : 2;
  std::lock_guard<std::mutex> lock(h->host_mu);
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok) return fail(WFE_ERR_CUDA, "cudaSetDevice failed");
  int rc = ensure_ring(h);
  if (rc != WFE_OK) {
    const std::string msg = g_err;
    free_ring(h);  // partial allocations
    return fail(rc, msg);
  }
  // on any failure below: wait for everything already enqueued and forget the caller's buffers, so that neither this
  // call's copies outlive it nor the next call writes into them
  struct RingGuard {
    wfe_handle* h;
    bool armed = true;
    ~RingGuard() {
      if (!armed) return;
      const std::string msg = g_err;
      for (auto& s : h->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        s.busy = false;
        s.user_out = nullptr;
        s.user_mask = nullptr;
      }
      cudaGetLastError();
      g_err = msg;
    }
  } ring_guard{h};

  // `out` / `attn_mask` may also be DEVICE memory (the in-loop training consumer): the kernels then write
  // straight into them and nothing comes back over PCIe
  const bool out_device = is_device_ptr(out);
  const bool mask_device = attn_mask != nullptr && is_device_ptr(attn_mask);
  const bool out_pinned = !out_device && is_pinned_host(out);
  const bool mask_pinned = attn_mask != nullptr && !mask_device && is_pinned_host(attn_mask);
  if (out_device && ((reinterpret_cast<uintptr_t>(out) & 15u) != 0))
    return fail(WFE_ERR_INVALID, "device `out` must be 16-byte aligned");
  const size_t clip_out = (size_t)h->cfg.n_mel * h->n_frames;
  // Pinned source clips are uploaded straight from the caller's memory.  Asking the driver about every clip costs ~1 us
  // each (0.5 ms for a batch of 256): the first non-empty clip decides whether the batch is looked at clip by clip at all
  // (a pinned clip in an otherwise pageable batch is merely staged like the others -- correct, one copy slower).
  bool look_for_pinned = false;
  for (int i = 0; i < batch; ++i)
    if (lengths[i] > 0) {
      look_for_pinned = is_pinned_host(clips[i]);
      break;
    }
  uint64_t up = 0, down = 0;
  // Chunks of up to `cap` clips go through the ring; a small batch is cut finer so that the staging copy of one chunk
  // overlaps the upload of the previous one (a batch of 8 as ONE chunk is staged, uploaded and computed back to back)
  const int cap = h->chunk_clips;
  int chunk = cap;
  if (const char* e = getenv("WFE_HOST_CHUNK")) {
    const int v = atoi(e);
    if (v >= 1) chunk = v < cap ? v : cap;
  } else if (batch < 4 * cap) {
    chunk = (batch + 3) / 4;
    if (chunk < 1) chunk = 1;
  }
  // (Staging chunk k + 1 on the copy threads while this thread issues chunk k's CUDA calls was built and measured: 3 %
  //  on 256 float32 clips, nothing on int16, and a noisier, slower batch of 8 -- the label collation's CUDA calls on the
  //  helper thread then contend with both.  One chunk at a time it stays.)
  int slot_i = 0;
  for (int c0 = 0; c0 < batch; c0 += chunk, slot_i = (slot_i + 1) % kSlots) {
    HostSlot& s = h->slots[slot_i];
    rc = retire_slot(h, s);
    if (rc != WFE_OK) return rc;
    const int n = (batch - c0 < chunk) ? batch - c0 : chunk;
    // ragged pack: only min(len, n_samples) samples of each clip cross PCIe; every clip starts on a 16-byte boundary
    // of the device buffer so that the kernels' vector / bulk-copy paths apply (starts in h_off[0..n), lengths after)
    int64_t* starts = s.h_off;
    int64_t* lens = s.h_off + cap;
    int64_t pos = 0, payload = 0;
    for (int i = 0; i < n; ++i) {
      int64_t len = lengths[c0 + i];
      if (len > h->cfg.n_samples) len = h->cfg.n_samples;
      starts[i] = pos;
      lens[i] = len;
      payload += len;
      pos = (pos + len + 7) & ~(int64_t)7;
    }
    // contiguous runs of pinned clips go straight from the caller's memory; everything else is staged
    int i = 0;
    while (i < n) {
      if (lens[i] == 0) {
        ++i;
        continue;
      }
      const char* src = static_cast<const char*>(clips[c0 + i]);
      if (look_for_pinned && is_pinned_host(src)) {
        // extend the run while the next clip continues the source exactly where the device layout expects it
        int j = i + 1;
        while (j < n && lens[j] > 0 &&
               static_cast<const char*>(clips[c0 + j]) == src + (size_t)(starts[j] - starts[i]) * es &&
               starts[j] == starts[j - 1] + lens[j - 1])
          ++j;
        const int64_t run = starts[j - 1] + lens[j - 1] - starts[i];
        WFE_CUDA(cudaMemcpyAsync(static_cast<char*>(s.d_in) + (size_t)starts[i] * es, src, (size_t)run * es,
                                 cudaMemcpyHostToDevice, s.stream));
        i = j;
      } else {
        // stage a maximal run of pageable clips (all copy threads), then one H2D for the run (alignment gaps travel too:
        // < 32 B each)
        const int i0 = i;
        int last = i;
        std::vector<CopyPool::Job> jobs;
        while (i < n && !(lens[i] > 0 && look_for_pinned && is_pinned_host(clips[c0 + i]))) {
          if (lens[i] > 0) {
            jobs.push_back({static_cast<char*>(s.h_in) + (size_t)starts[i] * es, static_cast<const char*>(clips[c0 + i]),
                            (size_t)lens[i] * es});
            last = i;
          }
          ++i;
        }
        h->pool->run(jobs);
        const size_t bytes = (size_t)(starts[last] + lens[last] - starts[i0]) * es;
        if (bytes)
          WFE_CUDA(cudaMemcpyAsync(static_cast<char*>(s.d_in) + (size_t)starts[i0] * es,
                                   static_cast<char*>(s.h_in) + (size_t)starts[i0] * es, bytes, cudaMemcpyHostToDevice,
                                   s.stream));
      }
    }
    up += (uint64_t)payload * es + (uint64_t)(2 * n) * sizeof(int64_t);
    WFE_CUDA(cudaMemcpyAsync(s.d_off, s.h_off, (size_t)(2 * cap) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
    const float* stats = nullptr;
    if (do_normalize) {
      rc = wfe_clip_stats(h, s.d_in, pcm_dtype, pcm_scale, s.d_off, s.d_off + cap, n, s.d_stats, s.stream);
      if (rc != WFE_OK) return rc;
      stats = s.d_stats;
    }
    char* dst = static_cast<char*>(out) + (size_t)c0 * clip_out * os;
    int32_t* mdst = attn_mask ? attn_mask + (size_t)c0 * h->n_frames : nullptr;
    rc = wfe_logmel_ex(h, s.d_in, pcm_dtype, pcm_scale, s.d_off, s.d_off + cap, n, stats, out_device ? dst : s.d_out,
                       out_dtype, attn_mask ? (mask_device ? mdst : s.d_mask) : nullptr, s.d_scratch, s.stream);
    if (rc != WFE_OK) return rc;
    s.pending_out_bytes = (size_t)n * clip_out * os;
    if (out_device) {
      // nothing to copy
    } else if (out_pinned) {
      WFE_CUDA(cudaMemcpyAsync(dst, s.d_out, s.pending_out_bytes, cudaMemcpyDeviceToHost, s.stream));
      down += (uint64_t)s.pending_out_bytes;
    } else {
      WFE_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, s.pending_out_bytes, cudaMemcpyDeviceToHost, s.stream));
      s.user_out = dst;
      down += (uint64_t)s.pending_out_bytes;
    }
    if (attn_mask && !mask_device) {
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
  ring_guard.armed = false;
  if (h2d_bytes) *h2d_bytes = up;
  if (d2h_bytes) *d2h_bytes = down;
  return WFE_OK;
}

int wfe_extract_host(wfe_handle* h, const void* const* clips, const int64_t* lengths, int32_t batch,
                     int32_t pcm_dtype, float pcm_scale, int32_t do_normalize, float* out, int32_t* attn_mask,
                     uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  return wfe_extract_host_ex(h, clips, lengths, batch, pcm_dtype, pcm_scale, do_normalize, out, WFE_OUT_F32, attn_mask,
                             h2d_bytes, d2h_bytes);
}

}  // extern "C"
