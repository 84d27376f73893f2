// Whisper log-mel frontend kernel for sm_100a (B200): ONE persistent kernel per batch.
//
// Unit of work = a tile of 32 consecutive STFT frames of one clip; LANE == FRAME in the FFT stages, so every
// shared-memory access is [row][lane] (conflict-free) and every constant is warp-uniform.  CTAs are persistent
// (grid = SMs x resident CTAs) and pull tile ids from a global counter, clip-major.
//
//   stage 0  coalesced 128-bit loads of the tile's 5360 PCM samples -> smem (truncate / right-zero-pad to n_samples,
//            centred reflect pad, optional int16 -> float and zero-mean/unit-variance)          (Appendix A steps 2-3)
//   stage 1  8 warps x (n1, n1+1): Hann window, real 25-point DFT (5x5), W400^(n1 k2) twiddle -- two n1 per thread,
//            packed f32x2 (FADD2/FMUL2/FFMA2)                                                    (steps 5-6)
//   stage 2  6 warps x (k2, k2+1) + 1 warp for k2 = 0: complex 16-point DFT (4x4) over n1, power |X|^2, stored
//            bin-major                                                                           (steps 6-7)
//   stage 3  banded slaney mel projection in exact fp32: each half-warp owns one mel and 16 frame PAIRS, so one
//            LDS.64 (power of two adjacent frames) + one FFMA2 with the weight broadcast covers two frames of a
//            non-zero; epilogue log10, (x+4)/4, full-line 64-bit stores, tile min/max            (steps 8, 9, 11)
//            (legacy mma.sync TF32 was tried and measured at CUDA-core rate on B200 - see DESIGN.md)
//   clamp    per-clip max-8 clamp (step 10) without a second pass over HBM and without fences or atomics: every tile
//            stores its own maximum into a zero-initialised word tile_key[clip][tile]; a clip is complete exactly when
//            none of its words is zero, and its maximum is the maximum of the words (each is written once, so no
//            ordering between locations is needed).  Every CTA remembers its own tiles and, once their clip is
//            complete, re-reads from L2 only those whose minimum is below the floor and fixes them; tiles that lie
//            entirely in the zero padding are written once, late, as a constant.
//
// Arithmetic restated from HF:models/whisper/feature_extraction_whisper.py:135-164;
// 400 = 16 x 25 Cooley-Tukey: n = n1 + 16*n2, k = k2 + 25*k1,
//   X[k2+25k1] = sum_n1 W16^(n1 k1) * W400^(n1 k2) * sum_n2 x[n1+16 n2] W25^(n2 k2).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "wfe_codelets.cuh"

#ifndef WFE_EXP
#define WFE_EXP 0  // bit 0: skip clamp fix-ups, bit 1: skip mel stage, bit 2: skip stage 2, bit 3: skip stage 1,
                   // bit 4: skip scheduler bookkeeping (what-if timing builds only: results are wrong)
#endif

namespace wfe {

#if WFE_EXP & 256
// timing-trace build: lane 0 of every warp of CTAs 0..3 stamps clock64() at each stage boundary of iterations 8..15;
// the scheduler lane additionally stamps the steps of its block (read back with wfe_debug_read_trace)
__device__ unsigned long long g_trace[4 * 8 * 8 * 12 + 4 * 8 * 8];
#define WFE_TRACE(pt)                                                                                        \
  do {                                                                                                       \
    if (lane == 0 && blockIdx.x < 4 && it >= 8 && it < 16)                                                   \
      g_trace[((blockIdx.x * 8 + (it - 8)) * 8 + warp) * 12 + (pt)] = clock64();                             \
  } while (0)
#define WFE_TRACE_S(pt)                                                                                      \
  do {                                                                                                       \
    if (blockIdx.x < 4 && it >= 8 && it < 16)                                                                \
      g_trace[4 * 8 * 8 * 12 + (blockIdx.x * 8 + (it - 8)) * 8 + (pt)] = clock64();                          \
  } while (0)
#else
#define WFE_TRACE(pt) \
  do {                \
  } while (0)
#define WFE_TRACE_S(pt) \
  do {                  \
  } while (0)
#endif

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMelWarps = 7;  // warps 0..6 run stages 2-3; warp 7 is the scheduler / bookkeeping warp after stage 1
constexpr int kSigLen = (kTileF - 1) * kHop + kNFft;  // 5360 padded-signal samples per tile
constexpr int kSigStride = kHop + 2;                  // +2 pad words per 160 samples: conflict-free LDS.64 across frames
constexpr int kSigSm = kSigLen + 2 * (kSigLen / kHop) + 4;   // 5430 floats per staging buffer (two of them)
constexpr int kZSm = kZPlanes * 16 * kTileF;          // 12800 floats; the power buffer (201 x 40) aliases it
constexpr int kRing = 128;                            // pending-tile ring (>= tiles per clip, see wfe_api.cu)
constexpr int kMaxMelGroups = 32;                     // groups of 4 mel pairs (n_mel <= 256)
constexpr int kMaxMelRows = 256;                      // table rows over all groups (55 for large-v3, 54 for whisper-small)

static_assert(kBins * kPStride <= kZSm, "power buffer must fit in the z buffer it aliases");
static_assert((kSigSm * 4) % 16 == 0 || true, "");

// order-preserving float <-> uint32 key (for atomic max on floats of either sign); key 0 < every float
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <typename T>
__device__ __forceinline__ float pcm_to_float(T v, float scale);
template <>
__device__ __forceinline__ float pcm_to_float<float>(float v, float) { return v; }
template <>
__device__ __forceinline__ float pcm_to_float<int16_t>(int16_t v, float scale) { return (float)v * scale; }
template <>
__device__ __forceinline__ float pcm_to_float<__half>(__half v, float) { return __half2float(v); }  // exact widening

// one mel-stage work item: FOUR mel pairs (slot s: mels m_s, m_s+1, one per half-warp) that run in lock step over `trips`
// table rows, so every thread has four independent accumulator chains.  Every (slot, half) filter is a contiguous BAND of
// `trips` power rows starting at row lo (zero weights where the band is longer than the filter), so the power loads are
// pointer + immediate and a table row is just the four weights of a half.  Pairs are grouped by similar band length.
struct alignas(16) MelGroup {
  int32_t trips;       // table rows (even)
  int32_t tab_idx;     // float4 index of the group's first row in mel_tab; row i = [half 0: w of slots 0..3][half 1: ...]
  int32_t valid;       // bit 2*s + h: mel of slot s, half h exists
  int32_t pad_;
  int32_t out_off[4];  // m_s * n_frames: element offset of slot s's first mel row inside one clip's output
  int32_t lo_off[2][4];  // [half][slot]: lo * kPStride, float offset of the band's first power row
};
constexpr int kMelUnroll = 16;  // table rows per unrolled pass of the mel loop (large-v3's longest band: 16)

struct LogmelParams {
  const void* pcm;
  const int64_t* offsets;   // [B+1] (or [B] when lengths != nullptr)
  const int64_t* lengths;   // [B] or nullptr
  const float2* norm;       // (mean, rstd) per clip or nullptr
  float* out;               // (B, n_mel, n_frames)
  int32_t* mask;            // (B, n_frames) or nullptr
  uint32_t* tile_key;       // [B][ntiles] max of y = (log10(mel)+4)/4 over the tile as an ordered key; 0 = not yet
                            //     published (zero-initialised; keys of finite floats are never 0)
  uint32_t* tile_counter;   // [1] dynamic tile scheduler (zero-initialised)
  const float4* s1_consts;  // [8][25] per-warp window/twiddle block
  const float4* mel_tab;    // [n_rows][2 halves]: the weights of slots 0..3
  const MelGroup* mel_groups;  // [n_groups], grouped by warp
  int mel_wrange[kMelWarps + 1];  // warp w < kMelWarps owns groups [mel_wrange[w], mel_wrange[w+1])
  float pcm_scale;
  int n_mel, n_samples, n_frames, ntiles, n_groups, n_rows;
  uint32_t total_tiles;
};

constexpr int kSigBuf = (kSigSm + 3) & ~3;  // 16-byte multiple
__host__ __device__ inline size_t logmel_smem_bytes(int n_rows) {
  return (size_t)(2 * kSigBuf + kZSm) * 4 + 8 * kS1ConstVec * 16 + (size_t)n_rows * 2 * 16;
}

// work item handed from the scheduler lane to the CTA through shared memory
struct alignas(16) TileDesc {
  int32_t b;      // clip; < 0: no more work
  int32_t tile;   // tile within the clip
  int32_t len;    // min(clip length, n_samples)
  int32_t mode;   // 0 = silent (all zero padding), 1 = interior + aligned (cp.async prefetch), 2 = synchronous staging
  int64_t off;    // first sample of the clip in pcm
  int64_t pad_;
};
constexpr int kModeSilent = 0, kModeAsync = 1, kModeSync = 2;
constexpr int kSilentBit = 0x40000000;  // in a pending-ring tile index: the tile lies in the zero padding

__device__ __forceinline__ void cp_async8(float* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_global_f2(float* p, float x, float y) {
  asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}
// named barrier over the first `nthreads` threads' warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// y = (log10(max(v, 1e-10)) + 4) / 4 in three instructions: MUFU.LG2, FFMA, FMNMX.  v <= 1e-10 (incl. lg2(0) = -inf)
// lands on exactly -1.5 = (log10(1e-10) + 4) / 4, so silence is bit-identical to the reference.
__device__ __forceinline__ float logmel_feature(float v) {
  return fmaxf(fmaf(lg2_approx(v), 0.25f * kLog10_2, 1.0f), -1.5f);
}

// the same without the floor: v = 0 gives -inf, which the per-clip clamp fix-up raises to max(floor, -1.5) later
__device__ __forceinline__ float logmel_feature_raw(float v) { return fmaf(lg2_approx(v), 0.25f * kLog10_2, 1.0f); }

struct FixEntry {
  int b, tile;       // tile < 0: nothing to do
  float floor_y;     // max(y_max - 2, -1.5): ((g - 8) + 4) / 4 and the absolute floor (log10(1e-10) + 4) / 4
  int silent;        // tile lies in the zero padding: store the constant instead of clamping
};

// apply the per-clip clamp to one of this CTA's own tiles (values come back from L2).  Executed by ONE warp for the mel
// rows 4*(wi + nw*j) + (lane >> 3): eight lanes cover the 128 bytes a tile occupies in a mel row with 128-bit accesses,
// eight rows in flight per thread.  In the main loop warp 7 does this alone (wi = 0, nw = 1) behind stages 2-3 of the
// other warps, so the clamp never sits on the CTA's critical path; the kernel tail splits the rows over all warps.
__device__ __forceinline__ void fix_tile(float* __restrict__ out, int n_mel, int n_frames, const FixEntry fx, int wi,
                                         int nw, int lane) {
  const int t0 = fx.tile * kTileF;
  const int nvalid = min(kTileF, n_frames - t0);
  const float fl = fx.floor_y;
  if ((n_frames & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    const int q = lane & 7, r = lane >> 3;
    if (4 * q >= nvalid) return;  // nvalid is a multiple of 4 here
    const size_t stride = (size_t)(n_frames >> 2);  // float4 per mel row
    float4* const base = reinterpret_cast<float4*>(out + (size_t)fx.b * n_mel * n_frames + t0) + q;
    const float4 c = make_float4(fl, fl, fl, fl);
    constexpr int kDeep = 8;  // rows in flight per thread (16 spills the kernel out of its 128 registers: measured slower)
    for (int m0 = 4 * wi + r; m0 < n_mel; m0 += 4 * kDeep * nw) {
      if (fx.silent) {  // (max(-10, g-8) + 4) / 4 everywhere
#pragma unroll
        for (int j = 0; j < kDeep; ++j) {
          const int m = m0 + 4 * nw * j;
          if (m < n_mel) base[(size_t)m * stride] = c;
        }
      } else {
        float4 v[kDeep];
#pragma unroll
        for (int j = 0; j < kDeep; ++j) {
          const int m = m0 + 4 * nw * j;
          v[j] = m < n_mel ? __ldcg(base + (size_t)m * stride) : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f);
        }
#pragma unroll
        for (int j = 0; j < kDeep; ++j) {
          const int m = m0 + 4 * nw * j;
          // (-inf, the log of a zero mel power, is below every floor)
          if (fminf(fminf(v[j].x, v[j].y), fminf(v[j].z, v[j].w)) < fl)
            base[(size_t)m * stride] = make_float4(fmaxf(v[j].x, fl), fmaxf(v[j].y, fl), fmaxf(v[j].z, fl), fmaxf(v[j].w, fl));
        }
      }
    }
  } else {  // generic geometry: scalar, lane = frame
    if (lane >= nvalid) return;
    float* q = out + (size_t)fx.b * n_mel * n_frames + t0 + lane;
    for (int m = wi; m < n_mel; m += nw) {
      float* e = q + (size_t)m * n_frames;
      if (fx.silent) {
        *e = fl;
      } else if (__ldcg(e) < fl) {
        *e = fl;
      }
    }
  }
}

// stage 0, synchronous: the tile's 5360 samples -> skewed smem.  kNorm selects the zero-mean/unit-variance variant.
template <typename T, bool kNorm>
__device__ __forceinline__ void stage_signal(float* __restrict__ sig, const T* __restrict__ pcm, int s_begin, int len,
                                             int n_samples, float scale, float mean, float rstd, int tid) {
  constexpr int kVec = 16 / (int)sizeof(T);  // samples per 128-bit load
  const bool fast =
      (s_begin >= 0) && (s_begin + kSigLen <= len) && ((reinterpret_cast<uintptr_t>(pcm + s_begin) & 15u) == 0);
  if (fast) {
    const uint4* src4 = reinterpret_cast<const uint4*>(pcm + s_begin);
#pragma unroll 3
    for (int v = tid; v < kSigLen / kVec; v += kThreads) {
      const uint4 raw = __ldg(src4 + v);
      const T* e = reinterpret_cast<const T*>(&raw);
      const int i = v * kVec;
      float* dst = sig + i + 2 * (i / kHop);  // 160 is a multiple of kVec: a vector never straddles a hop row
#pragma unroll
      for (int j = 0; j < kVec; j += 2) {
        float2 o;
        o.x = pcm_to_float<T>(e[j], scale);
        o.y = pcm_to_float<T>(e[j + 1], scale);
        if (kNorm) {
          o.x = (o.x - mean) * rstd;
          o.y = (o.y - mean) * rstd;
        }
        *reinterpret_cast<float2*>(dst + j) = o;
      }
    }
  } else {
    for (int i = tid; i < kSigLen; i += kThreads) {
      int s = s_begin + i;
      if (s < 0) s = -s;
      if (s >= n_samples) s = 2 * (n_samples - 1) - s;
      float v = 0.f;
      if (s >= 0 && s < len) {
        v = pcm_to_float<T>(pcm[s], scale);
        if (kNorm) v = (v - mean) * rstd;
      }
      sig[i + 2 * (i / kHop)] = v;
    }
  }
}

// work-item descriptor for tile id `id` of a clip whose (offset, available samples) are already known
template <typename T>
__device__ __forceinline__ TileDesc make_desc(const LogmelParams& p, uint32_t id, int64_t off, int64_t avail) {
  TileDesc d;
  d.pad_ = 0;
  d.off = off;
  if (id >= p.total_tiles) {
    d.b = -1;
    d.tile = 0;
    d.len = 0;
    d.mode = kModeSilent;
    return d;
  }
  d.b = (int)(id / (uint32_t)p.ntiles);
  d.tile = (int)(id - (uint32_t)d.b * (uint32_t)p.ntiles);
  d.len = (int)(avail < (int64_t)p.n_samples ? avail : (int64_t)p.n_samples);  // truncate to 30 s
  const int s_begin = d.tile * kTileF * kHop - kNFft / 2;
  const int s_hi = s_begin + kSigLen - 1;
  // lowest source sample this tile touches (right reflect maps s >= n_samples to 2(n-1)-s)
  int lowest = s_begin < 0 ? 0 : s_begin;
  if (s_hi >= p.n_samples) lowest = min(lowest, 2 * (p.n_samples - 1) - s_hi);
  if (lowest >= d.len) {
    d.mode = kModeSilent;  // every sample of every frame in the tile is zero padding
  } else {
    const T* src = reinterpret_cast<const T*>(p.pcm) + d.off + s_begin;
    const bool async_ok = sizeof(T) == 4 && p.norm == nullptr && s_begin >= 0 && s_begin + kSigLen <= d.len &&
                          (reinterpret_cast<uintptr_t>(src) & 7u) == 0;
    d.mode = async_ok ? kModeAsync : kModeSync;
  }
  return d;
}
// request (plain loads, not waited on here) the clip geometry a descriptor needs
__device__ __forceinline__ void request_clip(const LogmelParams& p, uint32_t id, int64_t& off, int64_t& avail) {
  off = 0;
  avail = 0;
  if (id < p.total_tiles) {
    const int b = (int)(id / (uint32_t)p.ntiles);
    off = __ldg(p.offsets + b);
    avail = p.lengths != nullptr ? __ldg(p.lengths + b) : __ldg(p.offsets + b + 1) - off;
  }
}

// asynchronous staging of an interior float32 tile: 2680 8-byte cp.async, no registers held.  Threads 0..239 each own
// one float2 column of three hop rows per pass, so every address is (per-thread base) + (compile-time constant).
__device__ __forceinline__ void prefetch_signal(float* __restrict__ sig, const float* __restrict__ src, int tid) {
  if (tid >= 240) return;
  const int r0 = tid / 80, c = tid - 80 * r0;
  float* d = sig + r0 * kSigStride + 2 * c;
  const float* g = src + r0 * kHop + 2 * c;
  constexpr int kRows = kSigLen / kHop;  // 33 full rows + half a row
#pragma unroll
  for (int k = 0; k < kRows / 3; ++k) cp_async8(d + 3 * k * kSigStride, g + 3 * k * kHop);
  if (r0 == 0 && c < (kSigLen - kRows * kHop) / 2) cp_async8(d + kRows * kSigStride, g + kRows * kHop);
}

// block (whole warp) until every tile of clip `b` has published its maximum; returns the clamp floor max(max - 2, -1.5)
__device__ __forceinline__ float wait_clip_floor(const LogmelParams& p, int b, int lane) {
  const uint32_t* row = p.tile_key + (size_t)b * p.ntiles;
  for (;;) {
    uint32_t m = 1u;
    bool zero = false;
    for (int w = lane; w < p.ntiles; w += 32) {
      const uint32_t k = ld_relaxed_u32(row + w);
      zero |= k == 0;
      m = max(m, k);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (!__any_sync(0xffffffffu, zero)) return fmaxf(key2f(m) - 2.0f, -1.5f);
    __nanosleep(200);
  }
}

// max / min of y over a tile from the per-warp extrema of the raw mel powers (warp-wide; result uniform)
__device__ __forceinline__ void tile_extrema(const uint32_t (&red)[2][kWarps], int lane, int silent, float& mx, float& mn) {
  uint32_t hi = lane < kMelWarps ? red[0][lane] : 0u;
  uint32_t lo = lane < kMelWarps ? red[1][lane] : 0x7f800000u;
  hi = __reduce_max_sync(0xffffffffu, hi);
  lo = __reduce_min_sync(0xffffffffu, lo);
  mx = logmel_feature_raw(__uint_as_float(hi));
  mn = logmel_feature_raw(__uint_as_float(lo));
  if (silent) {  // a tile in the zero padding: max = (log10(1e-10) + 4) / 4; min = -inf so that it is always written
    mx = -1.5f;
    mn = -__int_as_float(0x7f800000);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) logmel_kernel(const LogmelParams p) {
  extern __shared__ __align__(16) float smem[];
  float* const sigbuf = smem;               // two signal staging buffers (tile i -> buffer i & 1)
  float* const zbuf = smem + 2 * kSigBuf;   // stage 1 -> stage 2 exchange; the power buffer aliases it after stage 2
  float4* const s_cst = reinterpret_cast<float4*>(zbuf + kZSm);
  float4* const s_mtab = reinterpret_cast<float4*>(s_cst + 8 * kS1ConstVec);
  __shared__ MelGroup s_groups[kMaxMelGroups];
  __shared__ uint32_t s_red[2][2][kWarps];  // [tile parity][max, min][warp]: bit patterns of the largest / smallest mel
                                            // power of the tile (mel powers are >= 0, so uint order == float order)
  __shared__ TileDesc s_desc[2];         // descriptor of tile k lives in slot k & 1
  __shared__ FixEntry s_fix[1];          // kernel tail only: the entry all warps work on
  __shared__ int2 s_pend_bt[kRing];      // (clip, tile | kSilentBit: tile lies in the zero padding, not yet written)
  __shared__ float2 s_pend_mm[kRing];    // tile (minimum, maximum) of y (minimum -inf when a mel power is 0)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool sched = tid == 7 * 32;  // lane 0 of warp 7: tile scheduler; the whole of warp 7 keeps the clamp's books

  // warp-7 state (uniform across its lanes): ring of this CTA's pending tiles and the previous tile, whose maximum is
  // published one tile late.  Everything the block needs from global memory is requested at its start and consumed
  // at its end; it executes no fence and no atomic besides the tile-counter increment.
  int ring_head = 0, ring_count = 0;
  int prev_b = -1, prev_tile = 0, prev_silent = 0;

  // ---- one-time CTA set-up ----
  for (int i = tid; i < 8 * kS1ConstVec; i += kThreads) s_cst[i] = p.s1_consts[i];
  for (int i = tid; i < p.n_groups; i += kThreads) s_groups[i] = p.mel_groups[i];
  for (int i = tid; i < p.n_rows * 2; i += kThreads) s_mtab[i] = p.mel_tab[i];
  if (sched) {
    const uint32_t id0 = atomicAdd(p.tile_counter, 1u);
    const uint32_t id1 = atomicAdd(p.tile_counter, 1u);
    int64_t o, a;
    request_clip(p, id0, o, a);
    s_desc[0] = make_desc<T>(p, id0, o, a);
    request_clip(p, id1, o, a);
    s_desc[1] = make_desc<T>(p, id1, o, a);
  }
  __syncthreads();

  TileDesc cur = s_desc[0], nxt = s_desc[1];
  if (cur.b >= 0 && cur.mode == kModeAsync)
    prefetch_signal(sigbuf, reinterpret_cast<const float*>(p.pcm) + cur.off + cur.tile * kTileF * kHop - kNFft / 2, tid);
  cp_async_commit();
  int it = 0;

  while (cur.b >= 0) {
    const int b = cur.b, tile = cur.tile, len = cur.len;
    const int t0 = tile * kTileF;
    const int s_begin = t0 * kHop - kNFft / 2;  // unpadded sample index of sig[0]
    const int nvalid = min(kTileF, p.n_frames - t0);
    const bool silent = cur.mode == kModeSilent;
    float* const sig = sigbuf + (it & 1) * kSigBuf;

    WFE_TRACE(0);
    // ---- top: make sure this tile's signal has landed, then start the NEXT tile's loads.  The prefetch is issued AFTER
    //      the barrier: BAR.SYNC drains a warp's pending shared-memory writes, in-flight LDGSTS included, so a prefetch
    //      issued just before S1 was waited for in full there (~2500 cycles per tile).  Measured alternatives: issued
    //      by warp 7 alone behind stages 2-3 (no barrier ever waits: +2 % on noise, -6 % when that warp also has a clamp
    //      fix-up per tile) or by warps 0..6 after S3 (-2 %) ----
    uint32_t idB = 0;
    if (sched) idB = atomicAdd(p.tile_counter, 1u);  // id of tile it+2, first used in this tile's stage 2
    if (cur.mode == kModeSync) {
      const T* pcm = reinterpret_cast<const T*>(p.pcm) + cur.off;
      if (p.norm != nullptr) {
        const float2 st = __ldg(p.norm + b);
        stage_signal<T, true>(sig, pcm, s_begin, len, p.n_samples, p.pcm_scale, st.x, st.y, tid);
      } else {
        stage_signal<T, false>(sig, pcm, s_begin, len, p.n_samples, p.pcm_scale, 0.f, 1.f, tid);
      }
    }
    if (p.mask != nullptr && tid < nvalid) p.mask[(size_t)b * p.n_frames + t0 + tid] = ((t0 + tid) * kHop < len) ? 1 : 0;
    cp_async_wait<0>();  // this tile's signal (the group committed during the previous tile) is complete
    WFE_TRACE(1);
    __syncthreads();     // S1: signal visible to all warps
    WFE_TRACE(2);
    if (nxt.b >= 0 && nxt.mode == kModeAsync)  // the other buffer was last read in stage 1 of the previous tile
      prefetch_signal(sigbuf + ((it + 1) & 1) * kSigBuf,
                      reinterpret_cast<const float*>(p.pcm) + nxt.off + nxt.tile * kTileF * kHop - kNFft / 2, tid);
    cp_async_commit();

    uint32_t rmax = 0u, rmin = 0x7f800000u;  // bit patterns of the largest / smallest mel power (identity: 0, +inf)
    if (!silent) {
      // ---- stage 1: warp w owns n1 = 2w, 2w+1 ----
      WFE_TRACE(3);
      if (!(WFE_EXP & 8)) stage1_pair(sig + kSigStride * lane, s_cst + warp * kS1ConstVec, 2 * warp, zbuf + lane);
      WFE_TRACE(4);
      __syncthreads();  // S2
      WFE_TRACE(5);
    }

    // ---- stage 2 (compute half): warps 0..5 own (k2, k2+1) = (1,2)..(11,12); warp 6 owns k2 = 0;
    //      warp 7 lane 0 runs the scheduler ----
    f2 pw[16];
    if (!silent && !(WFE_EXP & 4)) {
      if (warp < 6)
        stage2_pair_compute(zbuf + lane, 2 * warp + 1, pw);
      else if (warp == 6)
        stage2_k0_compute(zbuf + lane, pw);
    }
    if (warp == 7) {
      // ---- warp 7: tile scheduler + clamp bookkeeping, overlapping stages 2 and 3 of warps 0..6 ----
      WFE_TRACE_S(0);
      // (1) requests first: geometry of tile it+2's clip (lane 0), and the tile_key words of the clips of the two oldest
      //     pending tiles (lane l reads words l, l+32, l+64, l+96); all consumed at the end of the block
      int64_t g_off = 0, g_avail = 0;
      if (lane == 0) request_clip(p, idB, g_off, g_avail);
      int chk0 = ring_count > 0 ? s_pend_bt[ring_head].x : -1;
      int chk1 = ring_count > 1 ? s_pend_bt[(ring_head + 1) & (kRing - 1)].x : -1;
      uint32_t k0[4], k1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = lane + 32 * j;
        k0[j] = (chk0 >= 0 && w < p.ntiles) ? ld_relaxed_u32(p.tile_key + (size_t)chk0 * p.ntiles + w) : 1u;
        k1[j] = (chk1 >= 0 && w < p.ntiles) ? ld_relaxed_u32(p.tile_key + (size_t)chk1 * p.ntiles + w) : 1u;
      }
      WFE_TRACE_S(1);
      // (2) previous tile: max / min over the 7 mel warps, publish the max, remember the tile
      if (prev_b >= 0) {
        float mx, mn;
        tile_extrema(s_red[(it + 1) & 1], lane, prev_silent, mx, mn);
        // ring full: cannot happen while ntiles <= kRing unless other CTAs lag a whole clip behind; the oldest entry's
        // clip then has every tile assigned to a RUNNING CTA (ids are handed out in order) whose warp 7 publishes
        // without ever waiting, so this wait terminates
        if (ring_count == kRing) {
          const int2 bt = s_pend_bt[ring_head];
          const float2 pm = s_pend_mm[ring_head];
          ring_head = (ring_head + 1) & (kRing - 1);
          --ring_count;
          const float fl = wait_clip_floor(p, bt.x, lane);
          if (pm.x < fl) {
            const FixEntry fx{bt.x, bt.y & ~kSilentBit, fl, (bt.y & kSilentBit) != 0 || pm.y <= fl};
            fix_tile(p.out, p.n_mel, p.n_frames, fx, 0, 1, lane);
          }
          // the words read in (1) belong to the entry just popped and to its successor: the head has moved, so step (3)
          // must not pair them with the NEW head this time round
          chk0 = chk1 = -1;
        }
        if (lane == 0) {
          st_relaxed_u32(p.tile_key + (size_t)prev_b * p.ntiles + prev_tile, f2key(mx));
          const int slot = (ring_head + ring_count) & (kRing - 1);
          s_pend_bt[slot] = make_int2(prev_b, prev_tile | (prev_silent ? kSilentBit : 0));
          s_pend_mm[slot] = make_float2(mn, mx);
        }
        ++ring_count;
      }
      prev_b = b;
      prev_tile = tile;
      prev_silent = silent;
      WFE_TRACE_S(2);
      // (3) consume the words: a clip is complete when none of them is zero; its max is the max of the words
      FixEntry fxa{0, -1, 0.f, 0}, fxb{0, -1, 0.f, 0};
      if (!(WFE_EXP & 16)) {
        const uint32_t m0 = __reduce_max_sync(0xffffffffu, max(max(k0[0], k0[1]), max(k0[2], k0[3])));
        const bool z0 = __any_sync(0xffffffffu, (k0[0] == 0) | (k0[1] == 0) | (k0[2] == 0) | (k0[3] == 0));
        const uint32_t m1 = __reduce_max_sync(0xffffffffu, max(max(k1[0], k1[1]), max(k1[2], k1[3])));
        const bool z1 = __any_sync(0xffffffffu, (k1[0] == 0) | (k1[1] == 0) | (k1[2] == 0) | (k1[3] == 0));
        if (chk0 >= 0 && !z0) {
          const float floor_y = fmaxf(key2f(m0) - 2.0f, -1.5f);
          const int2 bt = s_pend_bt[ring_head];
          const float2 pm = s_pend_mm[ring_head];
          ring_head = (ring_head + 1) & (kRing - 1);
          --ring_count;
          // a tile wholly at or below the floor becomes the constant without being read back
          if (pm.x < floor_y) fxa = FixEntry{bt.x, bt.y & ~kSilentBit, floor_y, (bt.y & kSilentBit) != 0 || pm.y <= floor_y};
          if (chk1 >= 0 && !z1) {
            const float floor1 = fmaxf(key2f(m1) - 2.0f, -1.5f);
            const int2 bt1 = s_pend_bt[ring_head];
            const float2 pm1 = s_pend_mm[ring_head];
            ring_head = (ring_head + 1) & (kRing - 1);
            --ring_count;
            if (pm1.x < floor1) fxb = FixEntry{bt1.x, bt1.y & ~kSilentBit, floor1, (bt1.y & kSilentBit) != 0 || pm1.y <= floor1};
          }
        }
      }
      // (4) hand tile it+2 to the CTA (slot of tile it, whose descriptor already sits in registers)
      if (lane == 0) s_desc[it & 1] = make_desc<T>(p, idB, g_off, g_avail);
      // (5) the clamp fix-ups just decided (own tiles, written at least two tiles ago, L2-resident): warp 7 alone,
      //     behind stages 2-3 of the other warps
      if (!(WFE_EXP & 1)) {
        if (fxa.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fxa, 0, 1, lane);
        if (fxb.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fxb, 0, 1, lane);
      }
      WFE_TRACE_S(3);
      WFE_TRACE_S(4);
      WFE_TRACE_S(5);
      WFE_TRACE_S(6);
    }

    WFE_TRACE(6);
    if (!silent && warp < kMelWarps) {
      if (!(WFE_EXP & 4)) {
        bar_sync_named(1, kMelWarps * 32);  // S2b (warps 0..6): all z planes have been read; the power may overwrite them
        if (warp < 6)
          stage2_pair_store(pw, 2 * warp + 1, zbuf + lane);
        else
          stage2_k0_store(pw, zbuf + lane);
      }
      WFE_TRACE(7);
      bar_sync_named(2, kMelWarps * 32);  // S3 (warps 0..6): power buffer complete.  Warp 7 only rejoins at S4, so its
                                          // scheduler block overlaps stages 2 and 3
      WFE_TRACE(8);
      // ---- stage 3: banded mel projection, exact fp32.  Half-warp h owns mel m_s + h of each of a group's four slots,
      //      lane pr owns frames 2pr, 2pr+1.  Per table row: 1 broadcast LDS.128 (the four weights of this half), 4 LDS.64
      //      (power pairs, band pointer + immediate), 4 FFMA2 with the weight broadcast: four independent chains per
      //      thread.  Epilogue: lg2, one FFMA, 64-bit full-line stores; the tile's extrema are tracked on the RAW mel
      //      powers as integers (>= 0, so uint order == float order): ALU pipe, not FMA, and one REDUX per warp ----
      if (!(WFE_EXP & 2)) {
        const int h = lane >> 4, pr = lane & 15;
        const bool full = nvalid == kTileF && (p.n_frames & 1) == 0;
        const float* const pwl = zbuf + 2 * pr;  // power pair (frames 2pr, 2pr+1) of bin row 0
        float* obase = p.out + ((size_t)b * p.n_mel + h) * p.n_frames + t0 + 2 * pr;
        asm volatile("" : "+l"(obase));  // keep it one 64-bit base: each store address is then a single IMAD.WIDE
        const int g_end = p.mel_wrange[warp + 1];
        for (int gi = p.mel_wrange[warp]; gi < g_end; ++gi) {
          const int4* const gp = reinterpret_cast<const int4*>(&s_groups[gi]);
          const int4 gd = gp[0];      // trips, tab_idx, valid
          const int4 go = gp[1];      // out_off[0..3]
          const int4 lo = gp[2 + h];  // lo_off[h][0..3]
          const float *p0 = pwl + lo.x, *p1 = pwl + lo.y, *p2 = pwl + lo.z, *p3 = pwl + lo.w;
          const float4* wr = s_mtab + gd.y + h;
          f2 a0 = mk2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
          for (int rem = gd.x;;) {
#pragma unroll
            for (int i = 0; i < kMelUnroll; i += 2) {
              if (i >= rem) break;
              const float4 w0 = wr[2 * i], w1 = wr[2 * i + 2];
              a0 = vfma(f2{*reinterpret_cast<const float2*>(p0 + i * kPStride)}, w0.x, a0);
              a1 = vfma(f2{*reinterpret_cast<const float2*>(p1 + i * kPStride)}, w0.y, a1);
              a2 = vfma(f2{*reinterpret_cast<const float2*>(p2 + i * kPStride)}, w0.z, a2);
              a3 = vfma(f2{*reinterpret_cast<const float2*>(p3 + i * kPStride)}, w0.w, a3);
              a0 = vfma(f2{*reinterpret_cast<const float2*>(p0 + (i + 1) * kPStride)}, w1.x, a0);
              a1 = vfma(f2{*reinterpret_cast<const float2*>(p1 + (i + 1) * kPStride)}, w1.y, a1);
              a2 = vfma(f2{*reinterpret_cast<const float2*>(p2 + (i + 1) * kPStride)}, w1.z, a2);
              a3 = vfma(f2{*reinterpret_cast<const float2*>(p3 + (i + 1) * kPStride)}, w1.w, a3);
            }
            rem -= kMelUnroll;
            if (rem <= 0) break;
            p0 += kMelUnroll * kPStride;
            p1 += kMelUnroll * kPStride;
            p2 += kMelUnroll * kPStride;
            p3 += kMelUnroll * kPStride;
            wr += 2 * kMelUnroll;
          }
          const f2 acc[4] = {a0, a1, a2, a3};
          const uint32_t off[4] = {(uint32_t)go.x, (uint32_t)go.y, (uint32_t)go.z, (uint32_t)go.w};
          if (full && gd.z == 0xff) {
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              const uint32_t u0 = __float_as_uint(acc[sl].v.x), u1 = __float_as_uint(acc[sl].v.y);
              rmax = max(rmax, max(u0, u1));
              rmin = min(rmin, min(u0, u1));
              st_global_f2(obase + off[sl], logmel_feature_raw(acc[sl].v.x), logmel_feature_raw(acc[sl].v.y));
            }
          } else {
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              if (!((gd.z >> (2 * sl + h)) & 1)) continue;
              float* q0 = obase + off[sl];
              const float v[2] = {acc[sl].v.x, acc[sl].v.y};
#pragma unroll
              for (int e = 0; e < 2; ++e)
                if (2 * pr + e < nvalid) {
                  q0[e] = logmel_feature_raw(v[e]);
                  rmax = max(rmax, __float_as_uint(v[e]));
                  rmin = min(rmin, __float_as_uint(v[e]));
                }
            }
          }
        }
      }
      rmax = __reduce_max_sync(0xffffffffu, rmax);
      rmin = __reduce_min_sync(0xffffffffu, rmin);
      if (lane == 0) {
        s_red[it & 1][0][warp] = rmax;
        s_red[it & 1][1][warp] = rmin;
      }
    }
    WFE_TRACE(9);
    __syncthreads();  // S4: tile written (visible to this CTA); s_desc / s_red published; smem free
    WFE_TRACE(10);
    cur = nxt;
    nxt = s_desc[it & 1];
    ++it;
  }
  cp_async_wait<0>();

  // ---- epilogue: publish the last tile, then drain the tiles this CTA still has pending (every remaining tile of their
  //      clips is owned by a running CTA whose warp 7 publishes without ever waiting: the waits terminate) ----
  if (warp == 7 && prev_b >= 0) {
    float mx, mn;
    tile_extrema(s_red[(it + 1) & 1], lane, prev_silent, mx, mn);
    if (lane == 0) st_relaxed_u32(p.tile_key + (size_t)prev_b * p.ntiles + prev_tile, f2key(mx));
    if (ring_count < kRing) {
      if (lane == 0) {
        const int slot = (ring_head + ring_count) & (kRing - 1);
        s_pend_bt[slot] = make_int2(prev_b, prev_tile | (prev_silent ? kSilentBit : 0));
        s_pend_mm[slot] = make_float2(mn, mx);
      }
      ++ring_count;
    } else {  // ring full (see above): fix this one with warp 7 alone once its clip completes
      const float fl = wait_clip_floor(p, prev_b, lane);
      if (mn < fl) {
        const FixEntry fx{prev_b, prev_tile, fl, prev_silent};
        fix_tile(p.out, p.n_mel, p.n_frames, fx, 0, 1, lane);
      }
    }
  }
  __syncthreads();
  for (;;) {
    if (warp == 7) {
      __syncwarp();  // the ring entries written by lane 0 are visible to the warp
      if (ring_count > 0) {
        const int2 bt = s_pend_bt[ring_head];
        const float2 pm = s_pend_mm[ring_head];
        ring_head = (ring_head + 1) & (kRing - 1);
        --ring_count;
        const float fl = wait_clip_floor(p, bt.x, lane);
        if (lane == 0)
          s_fix[0] = FixEntry{bt.x, pm.x < fl ? (bt.y & ~kSilentBit) : -1, fl, (bt.y & kSilentBit) != 0 || pm.y <= fl};
      } else if (lane == 0) {
        s_fix[0].tile = -2;  // -2: ring empty
      }
    }
    __syncthreads();
    const FixEntry fx = s_fix[0];
    if (fx.tile == -2) break;
    if (fx.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fx, warp, kWarps, lane);
    __syncthreads();
  }
}

// ---- per-clip mean / rstd for do_normalize (HF:...feature_extraction_whisper.py:168-187) ------------
template <typename T>
__global__ void __launch_bounds__(512) clip_stats_kernel(const void* pcm_, float scale, const int64_t* offsets,
                                                         const int64_t* lengths, int n_samples, float2* stats) {
  const int b = blockIdx.x;
  const int64_t off = offsets[b];
  const int64_t avail = lengths != nullptr ? lengths[b] : offsets[b + 1] - off;
  const int len = (int)(avail < (int64_t)n_samples ? avail : (int64_t)n_samples);
  const T* pcm = reinterpret_cast<const T*>(pcm_) + off;
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const double v = (double)pcm_to_float<T>(pcm[i], scale);
    s += v;
    ss += v * v;
  }
  __shared__ double sh[2][16];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s;
    sh[1][threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, SS = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      S += sh[0][w];
      SS += sh[1][w];
    }
    const double n = len > 0 ? (double)len : 1.0;
    const double mean = S / n;
    double var = SS / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[b] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-7)));
  }
}

}  // namespace wfe
