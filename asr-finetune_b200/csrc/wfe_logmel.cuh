// Whisper log-mel frontend kernel for sm_100a (B200): ONE persistent, warp-specialised kernel per batch.
//
// Unit of work = a tile of 32 consecutive STFT frames of one clip; LANE == FRAME in the FFT stages, so every
// shared-memory access is [row][lane] (conflict-free) and every constant is (half-)warp-uniform.  One CTA per SM
// (13 warps, ~200 KB of shared memory) pulls tile ids from a global counter, clip-major, and runs them through a
// software pipeline of three roles that only meet at mbarriers, so the FMA-bound FFT and the LDS/MUFU/STG-bound mel
// stage of different tiles overlap on the same SM:
//
//   S warp (1)     draws tile ids, builds descriptors, fills the signal ring with 1-D bulk copies (TMA, 34 rows of 640 B
//                  per tile, completion on an mbarrier), writes the attention mask, and keeps the books of the
//                  per-clip clamp: publishes tile maxima, decides and applies the fix-ups (step 10)
//   FFT team (8)   stage 1: Hann window, real 25-point DFT (5x5), W400^(n1 k2) twiddle -- two n1 per thread, packed
//                  f32x2 (FADD2/FMUL2/FFMA2) -> z ring (double-buffered: no barrier between a tile's stage 2 and the
//                  next tile's stage 1);  stage 2: complex 16-point DFT (4x4) over n1, power |X|^2 -> P ring
//                  (Appendix A steps 5-7).  Edge tiles (reflect pad, zero pad, truncation, int16, do_normalize,
//                  unaligned clips) are staged by the team itself with plain loads
//   mel team (4)   stage 3: banded slaney mel projection in exact fp32 (each half-warp owns one mel and 16 frame PAIRS:
//                  one LDS.64 + one FFMA2 with the weight broadcast covers two frames of a non-zero), log10, (x+4)/4,
//                  full-line 64-bit stores, tile extrema (steps 8, 9, 11)
//
//   clamp          per-clip max-8 clamp (step 10) without a second pass over HBM and without fences or atomics: every tile
//                  stores its own maximum into a zero-initialised word tile_key[clip][tile]; a clip is complete exactly
//                  when none of its words is zero, and its maximum is the maximum of the words (each is written once, so
//                  no ordering between locations is needed).  Every CTA remembers its own tiles and, once their clip is
//                  complete, the S warp re-reads from L2 only those whose minimum is below the floor and fixes them;
//                  tiles that lie entirely in the zero padding never enter the pipeline and are written once, late, as
//                  a constant.
//
// Arithmetic restated from HF:models/whisper/feature_extraction_whisper.py:135-164;
// 400 = 16 x 25 Cooley-Tukey: n = n1 + 16*n2, k = k2 + 25*k1,
//   X[k2+25k1] = sum_n1 W16^(n1 k1) * W400^(n1 k2) * sum_n2 x[n1+16 n2] W25^(n2 k2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wfe_codelets.cuh"

#ifndef WFE_EXP
#define WFE_EXP 0  // bit 0: skip clamp fix-ups, bit 1: skip mel stage, bit 2: skip stage 2, bit 3: skip stage 1,
                   // bit 5: skip the signal loads (what-if timing builds only: results are wrong)
#endif

namespace wfe {

constexpr int kFftTeams = 2;                   // team t takes the pipeline tiles n = t (mod 2)
constexpr int kTeamWarps = 8;
constexpr int kFftWarps = kFftTeams * kTeamWarps;
constexpr int kMelWarps = 6;                   // warps 16..19 and 22, 23 (one or two per SM sub-partition)
constexpr int kLWarp = kFftWarps + 4;          // loader warp
constexpr int kKWarp = kLWarp + 1;             // bookkeeper warp
constexpr int kIdleWarps = 0;                  // only there to bring their registers into the CTA's pool (see below)
constexpr int kWarps = kKWarp + 3;
constexpr int kThreads = kWarps * 32;          // 768 (24 warps: registers are allocated four warps at a time)
constexpr int kFftThreads = kTeamWarps * 32;   // threads of one FFT team
// register budget (setmaxnreg): 768 threads launch with 80 registers each (the host checks the compiled count).  The
// pool a warp can grow from is its OWN SM sub-partition's (warp index mod 4): six warps x 32 x 80 = 15360 registers,
// shared by four FFT warps (88 each), one mel warp (64) and the loader (40), the bookkeeper (64) or a second mel warp:
// 32 * (4*88 + 64 + 64) = 15360.
constexpr int kRegsLaunch = 80, kRegsFft = 88, kRegsMel = 64, kRegsL = 40, kRegsK = 64;
static_assert(kFftWarps == 16 && kWarps == 24, "the per-sub-partition register budget assumes 4 FFT + 2 other warps each");
static_assert(4 * kRegsFft + kRegsMel + kRegsK <= 6 * kRegsLaunch && kRegsL <= kRegsK && kRegsMel <= kRegsK,
              "register budget of an SM sub-partition");
// named barriers: 1..6 the FFT teams' own (3 each), 7 the tail, 8/9 "power buffer t full" (FFT team t arrives, the mel
// team waits), 10/11 "power buffer t free" (the mel team arrives, FFT team t waits).  Hardware barriers park the waiting
// warps; an mbarrier try_wait loop was measured to burn half of all issued instructions here.
constexpr int kBarPFull = 8, kBarPFree = 10, kPBarThreads = (kTeamWarps + kMelWarps) * 32;
constexpr int kRegsBudget = 32 * (kFftWarps * kRegsFft + kMelWarps * kRegsMel + kRegsL + kRegsK);
constexpr int kSigLen = (kTileF - 1) * kHop + kNFft;  // 5360 padded-signal samples per tile
constexpr int kSigRows = kSigLen / kHop;                               // 33 full hop rows (+ 80 samples)
constexpr int kSigBuf = (kSigLen + kSigSkew * kSigRows + 3) & ~3;      // 5428 floats per signal buffer
constexpr int kZSm = kZPlanes * 16 * kTileF;          // 12800 floats per z buffer
constexpr int kPSm = kBins * kPStride;                // 6432 floats per power buffer
constexpr int kRing = 128;                            // pending-tile ring
constexpr int kSigStages = 3;                         // signal ring depth (bulk copies in flight ahead of the FFT team)
constexpr int kDrawBatch = 4;                         // consecutive tile ids per draw from the global counter
constexpr int kDescRing = 8;                          // descriptors / extrema of the tiles in flight (<= 6, see S warp)
constexpr int kMaxMelGroups = 32;                     // groups of 4 mel pairs (n_mel <= 256)
constexpr int kMaxMelRows = 128;                      // table rows over all groups (56 for large-v3); bounded by smem

static_assert(kSigBuf % 4 == 0 && kZSm % 4 == 0 && kPSm % 4 == 0, "16-byte aligned buffers");

// order-preserving float <-> uint32 key; key 0 < every float (0 = "not published")
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

struct LogmelParams {
  const void* pcm;
  const int64_t* offsets;   // [B+1] (or [B] when lengths != nullptr)
  const int64_t* lengths;   // [B] or nullptr
  const float2* norm;       // (mean, rstd) per clip or nullptr
  float* out;               // (B, n_mel, n_frames)
  int32_t* mask;            // (B, n_frames) or nullptr
  uint32_t* tile_key;       // [B][ntiles] max of y = (log10(mel)+4)/4 over the tile as an ordered key; 0 = not yet
                            //     published (zero-initialised; keys of floats are never 0)
  uint32_t* tile_counter;   // [1] dynamic tile scheduler (zero-initialised)
  const float4* s1_consts;  // [8][25] window/twiddle block of n1 pair j = (2j, 2j+1)
  const float4* mel_tab;    // [n_rows][2 halves]: the weights of slots 0..3
  const MelGroup* mel_groups;  // [n_groups], grouped by warp
  int mel_wrange[kMelWarps + 1];  // mel warp w owns groups [mel_wrange[w], mel_wrange[w+1])
  float pcm_scale;
  int n_mel, n_samples, n_frames, ntiles, n_groups, n_rows;
  uint32_t total_tiles;
};

__host__ __device__ inline size_t logmel_smem_bytes(int n_rows) {
  return (size_t)(kSigStages * kSigBuf + kFftTeams * kZSm + 2 * kPSm) * 4 + 8 * kS1ConstVec * 16 + (size_t)n_rows * 2 * 16;
}

// work item handed from the S warp to the teams through shared memory
struct alignas(16) TileDesc {
  int32_t b;      // clip; < 0: no more work
  int32_t tile;   // tile within the clip
  int32_t len;    // min(clip length, n_samples)
  int32_t mode;   // 0 = silent (all zero padding; never enters the pipeline), 1 = bulk-copied, 2 = staged by the FFT team
  int64_t off;    // first sample of the clip in pcm
  int32_t slot;   // pending-ring slot of the tile
  int32_t pad_;
};
constexpr int kModeSilent = 0, kModeAsync = 1, kModeSync = 2;
constexpr int kSilentBit = 0x40000000;  // in a pending-ring tile index: the tile lies in the zero padding

// ---- mbarrier / bulk-copy primitives (PTX; CTA-local shared addresses) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int N>
__device__ __forceinline__ void reg_grow() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_shrink() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one probe: true when the phase with the given parity has completed (suspends for a hardware time slice when not)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// blocking wait: the suspend-time hint keeps the warp parked in hardware instead of spinning through the issue port
// (a bare try_wait loop comes back every few hundred cycles: measured 130 loop trips per warp per tile)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WFE_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WFE_DONE;\n\t"
      "bra WFE_WAIT;\n\t"
      "WFE_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// 8-byte asynchronous copy global -> shared (LDGSTS), no registers held
__device__ __forceinline__ void cp_async8(float* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// arrive on `bar` (without raising its pending count) once all of this thread's earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned both sides, size a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void bulk_g2s_u32(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
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
// named barrier over `nthreads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// non-blocking arrival at a named barrier (the waiting side uses bar_sync_named with the same id and count)
__device__ __forceinline__ void bar_arrive_named(int id, int nthreads) {
  __threadfence_block();  // bar.arrive itself promises no memory ordering: make this thread's writes visible first
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// y = (log10(v) + 4) / 4 in two instructions: MUFU.LG2, FFMA.  No floor here: v < 1e-10 (incl. lg2(0) = -inf) gives
// y < -1.5, which the per-clip clamp fix-up raises to max(floor, -1.5) later -- (log10(1e-10) + 4) / 4 = -1.5 exactly,
// so silence is bit-identical to the reference.
__device__ __forceinline__ float logmel_feature_raw(float v) { return fmaf(lg2_approx(v), 0.25f * kLog10_2, 1.0f); }

struct FixEntry {
  int b, tile;       // tile < 0: nothing to do
  float floor_y;     // max(y_max - 2, -1.5): ((g - 8) + 4) / 4 and the absolute floor (log10(1e-10) + 4) / 4
  int silent;        // tile lies in the zero padding: store the constant instead of clamping
};

// apply the per-clip clamp to one of this CTA's own tiles (values come back from L2).  Executed by ONE warp for the mel
// rows 4*(wi + nw*j) + (lane >> 3): eight lanes cover the 128 bytes a tile occupies in a mel row with 128-bit accesses,
// eight rows in flight per thread.  In the main loop the S warp does this alone (wi = 0, nw = 1) beside the teams, so
// the clamp never sits on the pipeline's critical path; the kernel tail splits the rows over all warps.
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
    constexpr int kDeep = 8;
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

// stage 0 for edge tiles, by the FFT team (256 threads): the tile's 5360 samples -> skewed smem with plain loads
// (truncate / right-zero-pad to n_samples, centred reflect pad, int16 -> float, zero-mean/unit-variance).
template <typename T, bool kNorm>
__device__ __forceinline__ void stage_signal(float* __restrict__ sig, const T* __restrict__ pcm, int s_begin, int len,
                                             int n_samples, float scale, float mean, float rstd, int tid) {
  constexpr int kVec = 16 / (int)sizeof(T);  // samples per 128-bit load
  const bool fast =
      (s_begin >= 0) && (s_begin + kSigLen <= len) && ((reinterpret_cast<uintptr_t>(pcm + s_begin) & 15u) == 0);
  if (fast) {
    const uint4* src4 = reinterpret_cast<const uint4*>(pcm + s_begin);
#pragma unroll 3
    for (int v = tid; v < kSigLen / kVec; v += kFftThreads) {
      const uint4 raw = __ldg(src4 + v);
      const T* e = reinterpret_cast<const T*>(&raw);
      const int i = v * kVec;
      float* dst = sig + sig_pos(i);  // 160 is a multiple of kVec: a vector never straddles a hop row
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
    for (int i = tid; i < kSigLen; i += kFftThreads) {
      int s = s_begin + i;
      if (s < 0) s = -s;
      if (s >= n_samples) s = 2 * (n_samples - 1) - s;
      float v = 0.f;
      if (s >= 0 && s < len) {
        v = pcm_to_float<T>(pcm[s], scale);
        if (kNorm) v = (v - mean) * rstd;
      }
      sig[sig_pos(i)] = v;
    }
  }
}

// work-item descriptor for tile id `id` of a clip whose (offset, available samples) are already known
template <typename T>
__device__ __forceinline__ TileDesc make_desc(const LogmelParams& p, uint32_t id, int64_t off, int64_t avail) {
  TileDesc d;
  d.pad_ = 0;
  d.slot = 0;
  d.off = off;
  if (id >= p.total_tiles) {
    d.b = -1;
    d.tile = 0;
    d.len = 0;
    d.mode = kModeSync;
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

// one probe of clip `b`'s tile words (whole warp): true when every tile has published its maximum; floor_y = the clamp
// floor max(max - 2, -1.5)
__device__ __forceinline__ bool clip_floor_ready(const LogmelParams& p, int b, int lane, float& floor_y) {
  const uint32_t* row = p.tile_key + (size_t)b * p.ntiles;
  uint32_t m = 1u;
  bool zero = false;
  for (int w = lane; w < p.ntiles; w += 32) {
    const uint32_t k = ld_relaxed_u32(row + w);
    zero |= k == 0;
    m = max(m, k);
  }
  m = __reduce_max_sync(0xffffffffu, m);
  floor_y = fmaxf(key2f(m) - 2.0f, -1.5f);
  return !__any_sync(0xffffffffu, zero);
}


// TRIPS table rows of one mel group: per row one broadcast LDS.128 (the four weights of this half) and, per slot, one
// LDS.64 (power pair of this lane's two frames, band pointer + immediate) and one FFMA2 with the weight broadcast
template <int TRIPS>
__device__ __forceinline__ void mel_rows(const float* __restrict__ p0, const float* __restrict__ p1,
                                         const float* __restrict__ p2, const float* __restrict__ p3,
                                         const float4* __restrict__ wr, f2& a0, f2& a1, f2& a2, f2& a3) {
#pragma unroll
  for (int i = 0; i < TRIPS; ++i) {
    const float4 w = wr[2 * i];
    a0 = vfma(f2{*reinterpret_cast<const float2*>(p0 + i * kPStride)}, w.x, a0);
    a1 = vfma(f2{*reinterpret_cast<const float2*>(p1 + i * kPStride)}, w.y, a1);
    a2 = vfma(f2{*reinterpret_cast<const float2*>(p2 + i * kPStride)}, w.z, a2);
    a3 = vfma(f2{*reinterpret_cast<const float2*>(p3 + i * kPStride)}, w.w, a3);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) logmel_kernel(const LogmelParams p) {
  static_assert(kThreads * kRegsLaunch >= kRegsBudget, "register budget");
  extern __shared__ __align__(16) float smem[];
  float* const sigbuf = smem;                         // signal ring: tile n -> buffer n % kSigStages (bulk copies)
  float* const zbuf = sigbuf + kSigStages * kSigBuf;  // stage 1 -> stage 2 exchange, one buffer per FFT team
  float* const pbuf = zbuf + kFftTeams * kZSm;        // power ring: stage 2 -> mel team (tile n -> buffer n & 1)
  float4* const s_cst = reinterpret_cast<float4*>(pbuf + 2 * kPSm);
  float4* const s_mtab = s_cst + 8 * kS1ConstVec;
  __shared__ MelGroup s_groups[kMaxMelGroups];
  __shared__ TileDesc s_desc[kDescRing];               // descriptor of pipeline tile n lives in slot n & 7
  __shared__ uint32_t s_ext[kDescRing][2][8];  // [tile slot][max, min][mel warp]: bit patterns of the largest /
                                                       // smallest mel power of the tile (>= 0: uint order == float order)
  __shared__ float s_min8[kDescRing];    // K warp: minimum of y of a published, not yet booked tile
  __shared__ FixEntry s_fix[1];          // kernel tail only: the entry all warps work on
  __shared__ int2 s_pend_bt[kRing];      // (clip, tile | kSilentBit: tile lies in the zero padding, not yet written)
  __shared__ float s_pend_min[kRing];    // tile minimum of y (-inf when a mel power is 0)
  __shared__ __align__(8) uint64_t s_sig_full[kSigStages], s_sig_empty[kSigStages], s_ext_full[kDescRing],
      s_booked[kDescRing];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- one-time CTA set-up ----
  for (int i = tid; i < 8 * kS1ConstVec; i += kThreads) s_cst[i] = p.s1_consts[i];
  for (int i = tid; i < p.n_groups; i += kThreads) s_groups[i] = p.mel_groups[i];
  for (int i = tid; i < p.n_rows * 2; i += kThreads) s_mtab[i] = p.mel_tab[i];
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kSigStages; ++i) {
      mbar_init(&s_sig_full[i], 32);           // every L-warp lane: its copies have landed (or there are none)
      mbar_init(&s_sig_empty[i], kTeamWarps);  // every warp of the tile's FFT team has its samples in registers
    }
#pragma unroll
    for (int i = 0; i < kDescRing; ++i) {
      mbar_init(&s_ext_full[i], kMelWarps);    // every mel warp has stored its rows and its extrema
      mbar_init(&s_booked[i], 1);              // K warp has the tile in its ring: descriptor slot free
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int ring_head = 0, ring_count = 0;  // K warp: ring of this CTA's pending tiles

  if (warp < kFftWarps) {
    // =============================== FFT teams ===============================
    reg_grow<kRegsFft>();
    const int team = warp / kTeamWarps, tw = warp % kTeamWarps, ttid = tid % kFftThreads;
    const int bar0 = 1 + 3 * team;  // the team's named barriers: bar0 (staging), bar0 + 1 (z full), bar0 + 2 (z free)
    // team warp w takes the n1 pair (2w, 2w + 1) of every frame of the tile
    const float4* const cst = s_cst + tw * kS1ConstVec;
    const int n1 = 2 * tw;
    // stage-2 task: team warp 0 -> k2 = 0 (real input, light), 4 -> none, the other six -> k2 pairs (a, a+1)
    const int s2a = (tw & 3) == 0 ? 0 : 2 * (tw < 4 ? tw - 1 : tw - 2) + 1;
    const int roff0 = (kHop + kSigSkew) * lane;  // frame `lane` starts here; + kSigSkew per hop row crossed
    float* const z = zbuf + team * kZSm + lane;
    for (uint32_t n = team;; n += kFftTeams) {
      const int sb = n & 1;  // == team
      const int ss = n % kSigStages;
      mbar_wait(&s_sig_full[ss], (n / kSigStages) & 1);
      const TileDesc d = s_desc[n & (kDescRing - 1)];
      if (d.b < 0) {  // no more work: pass the stop on to the mel team
        if (n >= 2) bar_sync_named(kBarPFree + sb, kPBarThreads);
        bar_arrive_named(kBarPFull + sb, kPBarThreads);
        break;
      }
      const bool silent = d.mode == kModeSilent;
      float* const sig = sigbuf + ss * kSigBuf;
      if (d.mode == kModeSync) {  // edge tile: the team stages it itself
        const T* pcm = reinterpret_cast<const T*>(p.pcm) + d.off;
        const int s_begin = d.tile * kTileF * kHop - kNFft / 2;
        if (p.norm != nullptr) {
          const float2 st = __ldg(p.norm + d.b);
          stage_signal<T, true>(sig, pcm, s_begin, d.len, p.n_samples, p.pcm_scale, st.x, st.y, ttid);
        } else {
          stage_signal<T, false>(sig, pcm, s_begin, d.len, p.n_samples, p.pcm_scale, 0.f, 1.f, ttid);
        }
        bar_sync_named(bar0, kFftThreads);
      }
      // ---- stage 1 ----
      if (!silent) {
        f2 x[25];
        const float* const rowp[3] = {sig + roff0, sig + roff0 + kSigSkew, sig + roff0 + 2 * kSigSkew};
        stage1_load(rowp, cst, n1, x);
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_sig_empty[ss]);  // the buffer may be refilled (tile n + kSigStages)
        if (!(WFE_EXP & 8)) stage1_compute_store(x, cst, n1, z);
      } else if (lane == 0) {
        mbar_arrive(&s_sig_empty[ss]);
      }
      bar_sync_named(bar0 + 1, kFftThreads);  // the team's z buffer is complete
      // ---- stage 2 ----
      float* const P = pbuf + sb * kPSm + lane;
      f2 pw[16];
      if (!silent && !(WFE_EXP & 4)) {
        if (s2a > 0)
          stage2_pair_compute(z, s2a, pw);
        else if (tw == 0)
          stage2_k0_compute(z, pw);
      }
      bar_sync_named(bar0 + 2, kFftThreads);  // every warp has read z: the next tile's stage 1 may overwrite it
      if (n >= 2) bar_sync_named(kBarPFree + sb, kPBarThreads);  // the mel team is done with tile n - 2
      if (!silent && !(WFE_EXP & 4)) {
        if (s2a > 0)
          stage2_pair_store(pw, s2a, P);
        else if (tw == 0)
          stage2_k0_store(pw, P);
      }
      bar_arrive_named(kBarPFull + sb, kPBarThreads);
    }
    __syncthreads();  // (end of kernel: no warp exits early)
    return;  // the tail belongs to the K and mel warps
  } else if (warp != kLWarp && warp != kKWarp) {
    // =============================== mel team ===============================
    reg_shrink<kRegsMel>();
    const int mw = warp < kLWarp ? warp - kFftWarps : warp - kFftWarps - 2;
    const int h = lane >> 4, pr = lane & 15;
    const int g_begin = p.mel_wrange[mw], g_end = p.mel_wrange[mw + 1];
    for (uint32_t n = 0;; ++n) {
      const int sb = n & 1;
      bar_sync_named(kBarPFull + sb, kPBarThreads);
      const TileDesc d = s_desc[n & (kDescRing - 1)];
      if (d.b < 0) {  // pass the stop on to the K warp
        if (lane == 0) mbar_arrive(&s_ext_full[n & (kDescRing - 1)]);
        break;
      }
      const int t0 = d.tile * kTileF;
      const int nvalid = min(kTileF, p.n_frames - t0);
      uint32_t rmax = 0u, rmin = 0x7f800000u;  // bit patterns of the largest / smallest mel power (identity: 0, +inf)
      // ---- stage 3: banded mel projection, exact fp32.  Half-warp h owns mel m_s + h of each of a group's four slots,
      //      lane pr owns frames 2pr, 2pr+1.  Per table row: 1 broadcast LDS.128 (the four weights of this half), 4 LDS.64
      //      (power pairs, band pointer + immediate), 4 FFMA2 with the weight broadcast: four independent chains per
      //      thread.  Epilogue: lg2, one FFMA, 64-bit full-line stores; the extrema are tracked on the RAW mel powers as
      //      integers: ALU pipe, not FMA, and one REDUX per warp ----
      if (d.mode != kModeSilent && !(WFE_EXP & 2)) {
        const bool full = nvalid == kTileF && (p.n_frames & 1) == 0;
        const float* const pwl = pbuf + sb * kPSm + 2 * pr;  // power pair (frames 2pr, 2pr+1) of bin row 0
        float* obase = p.out + ((size_t)d.b * p.n_mel + h) * p.n_frames + t0 + 2 * pr;
        asm volatile("" : "+l"(obase));  // keep it one 64-bit base: each store address is then a single IMAD.WIDE
        for (int gi = g_begin; gi < g_end; ++gi) {
          const int4* const gp = reinterpret_cast<const int4*>(&s_groups[gi]);
          const int4 gd = gp[0];      // trips, tab_idx, valid
          const int4 go = gp[1];      // out_off[0..3]
          const int4 lo = gp[2 + h];  // lo_off[h][0..3]
          const float *p0 = pwl + lo.x, *p1 = pwl + lo.y, *p2 = pwl + lo.z, *p3 = pwl + lo.w;
          const float4* wr = s_mtab + gd.y + h;
          f2 a0 = mk2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
          switch (gd.x) {  // straight-line code per band length: all of a group's loads can be in flight at once
            case 2: mel_rows<2>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            case 4: mel_rows<4>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            case 6: mel_rows<6>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            case 8: mel_rows<8>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            case 10: mel_rows<10>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            case 12: mel_rows<12>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            case 14: mel_rows<14>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            case 16: mel_rows<16>(p0, p1, p2, p3, wr, a0, a1, a2, a3); break;
            default:
              for (int rem = gd.x; rem > 0; rem -= 2) {
                mel_rows<2>(p0, p1, p2, p3, wr, a0, a1, a2, a3);
                p0 += 2 * kPStride;
                p1 += 2 * kPStride;
                p2 += 2 * kPStride;
                p3 += 2 * kPStride;
                wr += 4;
              }
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
        s_ext[n & (kDescRing - 1)][0][mw] = rmax;
        s_ext[n & (kDescRing - 1)][1][mw] = rmin;
      }
      __syncwarp();  // the lanes' stores of the tile are ordered before lane 0's arrives
      if (lane == 0) mbar_arrive(&s_ext_full[n & (kDescRing - 1)]);
      bar_arrive_named(kBarPFree + sb, kPBarThreads);
    }
  } else if (warp == kLWarp) {
    // =============================== L warp: tile ids, descriptors, signal loads ===============================
    reg_shrink<kRegsL>();
    auto draw = [&]() -> uint32_t {  // kDrawBatch consecutive tile ids; lane 0 holds the first, the others a dummy
      return lane == 0 ? atomicAdd(p.tile_counter, (uint32_t)kDrawBatch) : 0u;
    };
    auto clip_geometry = [&](uint32_t id, int64_t& off, int64_t& avail) {
      off = 0;
      avail = 0;
      if (id < p.total_tiles) {
        const int cb = (int)(id / (uint32_t)p.ntiles);
        off = __ldg(p.offsets + cb);
        avail = p.lengths != nullptr ? __ldg(p.lengths + cb) : __ldg(p.offsets + cb + 1) - off;
      }
    };
    // hand one descriptor to the pipeline (stop: d.b < 0)
    uint32_t n = 0;
    auto issue = [&](const TileDesc& d) {
      const int ds = n & (kDescRing - 1), ss = n % kSigStages;
      mbar_wait(&s_booked[ds], ((n >> 3) & 1) ^ 1);                 // K warp is done with tile n - 8
      mbar_wait(&s_sig_empty[ss], ((n / kSigStages) & 1) ^ 1);      // FFT team has tile n - kSigStages in registers
      if (lane == 0) s_desc[ds] = d;
      __syncwarp();
      if (d.b >= 0 && d.mode == kModeAsync && !(WFE_EXP & 32)) {
        // 2680 8-byte cp.async per tile, 84 per lane.  Two hop rows (160 float2) take five warp-wide copies; only in the
        // third one do the lanes split between the rows, so every address is (one of two per-lane bases) + immediate.
        const float* src = reinterpret_cast<const float*>(p.pcm) + d.off + d.tile * kTileF * kHop - kNFft / 2 + 2 * lane;
        float* dstA = sigbuf + ss * kSigBuf + 2 * lane;
        float* dstB = dstA + (lane >= 16 ? kSigSkew : 0);
#pragma unroll
        for (int g = 0; g < (kSigRows + 2) / 2; ++g) {
          constexpr int kPairDst = 2 * (kHop + kSigSkew), kPairSrc = 2 * kHop;
          // float2 index within the row pair: lane + 32 m; row 1 starts at index 80 (dst + kSigSkew)
          const int valid = g * kPairSrc < kSigLen ? (kSigLen - g * kPairSrc) / 2 : 0;  // float2 left from this pair on
          if (0 < valid) cp_async8(dstA + g * kPairDst, src + g * kPairSrc);
          if (32 < valid) cp_async8(dstA + g * kPairDst + 64, src + g * kPairSrc + 64);
          if (64 < valid && 64 + lane < valid) cp_async8(dstB + g * kPairDst + 128, src + g * kPairSrc + 128);
          if (96 < valid && 96 + lane < valid) cp_async8(dstA + g * kPairDst + 192 + kSigSkew, src + g * kPairSrc + 192);
          if (128 < valid && 128 + lane < valid) cp_async8(dstA + g * kPairDst + 256 + kSigSkew, src + g * kPairSrc + 256);
        }
        cp_async_arrive(&s_sig_full[ss]);
      } else {
        mbar_arrive(&s_sig_full[ss]);  // silent, team-staged or stop: nothing to copy
      }
      __syncwarp();
      ++n;
    };
    // Tile ids are drawn a batch ahead and the geometry of their clips is loaded a batch ahead (lane j: tile j of the
    // batch), so an iteration never waits on global memory.
    uint32_t bidA = __shfl_sync(0xffffffffu, draw(), 0);
    int64_t offA, availA;
    clip_geometry(lane < kDrawBatch ? bidA + lane : 0xffffffffu, offA, availA);
    uint32_t bidB_raw = draw();
    while (bidA < p.total_tiles) {
      const uint32_t bidB = __shfl_sync(0xffffffffu, bidB_raw, 0);  // drawn a whole batch ago
      int64_t offB, availB;
      clip_geometry(lane < kDrawBatch ? bidB + lane : 0xffffffffu, offB, availB);
      const uint32_t bidC_raw = draw();
      for (int j = 0; j < kDrawBatch; ++j) {
        const uint32_t id = bidA + j;
        if (id >= p.total_tiles) break;
        const TileDesc d = make_desc<T>(p, id, __shfl_sync(0xffffffffu, offA, j), __shfl_sync(0xffffffffu, availA, j));
        if (p.mask != nullptr) {  // feature attention mask of the tile's frames (Appendix A step 12)
          const int t = d.tile * kTileF + lane;
          if (t < p.n_frames) p.mask[(size_t)d.b * p.n_frames + t] = (t * kHop < d.len) ? 1 : 0;
        }
        issue(d);
      }
      bidA = bidB;
      offA = offB;
      availA = availB;
      bidB_raw = bidC_raw;
    }
    for (int t = 0; t < kFftTeams; ++t) issue(make_desc<T>(p, p.total_tiles, 0, 0));  // one stop per FFT team
    __syncthreads();  // (end of kernel)
    return;
  } else {
    // =============================== K warp: the books of the per-clip clamp ===============================
    reg_shrink<kRegsK>();
    uint32_t n = 0;      // next pipeline tile to take into the ring
    uint32_t n_pub = 0;  // next pipeline tile whose maximum is to be published (>= n)
    // publish the maxima of finished tiles, in order, possibly ahead of the ring (never blocks)
    auto publish_ready = [&]() {
      while (n_pub - n < (uint32_t)kDescRing &&
             mbar_test_wait(&s_ext_full[n_pub & (kDescRing - 1)], (n_pub >> 3) & 1)) {
        const int es = n_pub & (kDescRing - 1);
        const int db = s_desc[es].b, dt = s_desc[es].tile, dm = s_desc[es].mode;
        if (db < 0) break;  // the stop marker is not a tile
        uint32_t hi = lane < kMelWarps ? s_ext[es][0][lane] : 0u;
        uint32_t lo = lane < kMelWarps ? s_ext[es][1][lane] : 0x7f800000u;
        hi = __reduce_max_sync(0xffffffffu, hi);
        lo = __reduce_min_sync(0xffffffffu, lo);
        float mx = logmel_feature_raw(__uint_as_float(hi)), mn = logmel_feature_raw(__uint_as_float(lo));
        if (dm == kModeSilent) {  // a tile in the zero padding: max = (log10(1e-10) + 4) / 4, written late as a constant
          mx = -1.5f;
          mn = -__int_as_float(0x7f800000);  // below every floor: always written
        }
        if (lane == 0) {
          st_relaxed_u32(p.tile_key + (size_t)db * p.ntiles + dt, f2key(mx));
          s_min8[es] = mn;
        }
        ++n_pub;
      }
      __syncwarp();
    };
    // retire the oldest pending tile if its clip is complete (blocking probe; ring-full path only)
    auto retire_oldest = [&]() {
      if (ring_count == 0) return;
      const int2 bt = s_pend_bt[ring_head];
      float fl;
      if (!clip_floor_ready(p, bt.x, lane, fl)) return;
      const float pm = s_pend_min[ring_head];
      ring_head = (ring_head + 1) & (kRing - 1);
      --ring_count;
      if (pm < fl && !(WFE_EXP & 1))
        fix_tile(p.out, p.n_mel, p.n_frames, FixEntry{bt.x, bt.y & ~kSilentBit, fl, (bt.y & kSilentBit) != 0}, 0, 1, lane);
    };
    int chk0 = -1, chk1 = -1;
    uint32_t k0[4] = {1u, 1u, 1u, 1u}, k1[4] = {1u, 1u, 1u, 1u};
    for (;;) {
      mbar_wait(&s_ext_full[n & (kDescRing - 1)], (n >> 3) & 1);
      const int es = n & (kDescRing - 1);
      const int db = s_desc[es].b, dt = s_desc[es].tile, dm = s_desc[es].mode;
      if (db < 0) break;
      publish_ready();  // at least tile n
      // a ring slot.  Ring full: its oldest entries' clips complete without this warp's help, except for this CTA's own
      // tiles in flight -- which publish_ready publishes ahead of the ring
      while (ring_count == kRing) {
        publish_ready();
        retire_oldest();
        chk0 = chk1 = -1;  // the words requested earlier may belong to retired entries
      }
      if (lane == 0) {
        const int slot = (ring_head + ring_count) & (kRing - 1);
        s_pend_bt[slot] = make_int2(db, dt | (dm == kModeSilent ? kSilentBit : 0));
        s_pend_min[slot] = s_min8[es];
        mbar_arrive(&s_booked[es]);  // descriptor / extrema slot free for tile n + 8
      }
      ++ring_count;
      ++n;
      __syncwarp();
      // retire the oldest pending tiles whose clip is complete (tile words requested one iteration ago: no wait on a
      // load just issued), then request the words of the new oldest
      {
        const uint32_t m0 = __reduce_max_sync(0xffffffffu, max(max(k0[0], k0[1]), max(k0[2], k0[3])));
        const bool z0 = __any_sync(0xffffffffu, (k0[0] == 0) | (k0[1] == 0) | (k0[2] == 0) | (k0[3] == 0));
        const uint32_t m1 = __reduce_max_sync(0xffffffffu, max(max(k1[0], k1[1]), max(k1[2], k1[3])));
        const bool z1 = __any_sync(0xffffffffu, (k1[0] == 0) | (k1[1] == 0) | (k1[2] == 0) | (k1[3] == 0));
        FixEntry fxa{0, -1, 0.f, 0}, fxb{0, -1, 0.f, 0};
        if (chk0 >= 0 && !z0) {
          const float floor0 = fmaxf(key2f(m0) - 2.0f, -1.5f);
          const int2 bt = s_pend_bt[ring_head];
          const float pm = s_pend_min[ring_head];
          ring_head = (ring_head + 1) & (kRing - 1);
          --ring_count;
          if (pm < floor0) fxa = FixEntry{bt.x, bt.y & ~kSilentBit, floor0, (bt.y & kSilentBit) != 0};
          if (chk1 >= 0 && !z1) {
            const float floor1 = fmaxf(key2f(m1) - 2.0f, -1.5f);
            const int2 bt1 = s_pend_bt[ring_head];
            const float pm1 = s_pend_min[ring_head];
            ring_head = (ring_head + 1) & (kRing - 1);
            --ring_count;
            if (pm1 < floor1) fxb = FixEntry{bt1.x, bt1.y & ~kSilentBit, floor1, (bt1.y & kSilentBit) != 0};
          }
        }
        chk0 = ring_count > 0 ? s_pend_bt[ring_head].x : -1;
        chk1 = ring_count > 1 ? s_pend_bt[(ring_head + 1) & (kRing - 1)].x : -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int w = lane + 32 * j;
          k0[j] = (chk0 >= 0 && w < p.ntiles) ? ld_relaxed_u32(p.tile_key + (size_t)chk0 * p.ntiles + w) : 1u;
          k1[j] = (chk1 >= 0 && w < p.ntiles) ? ld_relaxed_u32(p.tile_key + (size_t)chk1 * p.ntiles + w) : 1u;
        }
        if (!(WFE_EXP & 1)) {
          if (fxa.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fxa, 0, 1, lane);
          if (fxb.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fxb, 0, 1, lane);
        }
      }
    }
  }

  // ---- tail (K and mel warps): drain the tiles this CTA still has pending, seven warps on each.  Every tile of this CTA
  //      is published by now, and the other CTAs' K warps publish theirs without ever waiting on another CTA: the
  //      waits terminate ----
  constexpr int kTailThreads = (kMelWarps + 1) * 32;
  const int tw = warp == kKWarp ? kMelWarps : (warp < kLWarp ? warp - kFftWarps : warp - kFftWarps - 2);
  bar_sync_named(7, kTailThreads);
  for (;;) {
    if (warp == kKWarp) {
      if (ring_count > 0) {
        const int2 bt = s_pend_bt[ring_head];
        const float pm = s_pend_min[ring_head];
        ring_head = (ring_head + 1) & (kRing - 1);
        --ring_count;
        float fl;
        while (!clip_floor_ready(p, bt.x, lane, fl)) __nanosleep(200);
        if (lane == 0)
          s_fix[0] = FixEntry{bt.x, (pm < fl && !(WFE_EXP & 1)) ? (bt.y & ~kSilentBit) : -1, fl, (bt.y & kSilentBit) != 0};
      } else if (lane == 0) {
        s_fix[0].tile = -2;  // -2: ring empty
      }
    }
    bar_sync_named(7, kTailThreads);
    const FixEntry fx = s_fix[0];
    if (fx.tile == -2) break;
    if (fx.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fx, tw, kMelWarps + 1, lane);
    bar_sync_named(7, kTailThreads);
  }
  __syncthreads();  // (end of kernel)
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
