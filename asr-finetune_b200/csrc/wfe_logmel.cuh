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
//   stage 3  mel projection on the tensor pipe: the slaney filter bank is banded (33 non-zero 8-mel x 8-bin blocks of
//            26 x 16), so each warp issues a handful of mma.sync m16n8k8 TF32 (frames x bins x mels) with the filter
//            fragments read from smem; epilogue log10, (x+4)/4, store, tile min/max              (steps 8, 9, 11)
//   clamp    per-clip max-8 clamp (step 10) without a second pass over HBM: every CTA remembers its own tiles and,
//            once the clip's ticket shows all of its tiles are done, re-reads only the tiles whose minimum is below
//            the floor from L2 and fixes them; tiles that lie entirely in the zero padding are written once, late.
//
// Arithmetic restated from HF:models/whisper/feature_extraction_whisper.py:135-164;
// 400 = 16 x 25 Cooley-Tukey: n = n1 + 16*n2, k = k2 + 25*k1,
//   X[k2+25k1] = sum_n1 W16^(n1 k1) * W400^(n1 k2) * sum_n2 x[n1+16 n2] W25^(n2 k2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wfe_codelets.cuh"

namespace wfe {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSigLen = (kTileF - 1) * kHop + kNFft;  // 5360 padded-signal samples per tile
constexpr int kSigStride = kHop + 2;                  // +2 pad words per 160 samples: conflict-free LDS.64 across frames
constexpr int kSigSm = kSigLen + 2 * (kSigLen / kHop) + 4;   // 5430 floats per staging buffer (two of them)
constexpr int kZSm = kZPlanes * 16 * kTileF;          // 12800 floats; the power buffer (201 x 40) aliases it
constexpr int kRing = 128;                            // pending-tile ring (>= tiles per clip, see wfe_api.cu)
constexpr int kMaxUnits = 64;                         // (8-mel tile, 16-frame tile) work units of the mel stage
constexpr int kMaxKsteps = 64;                        // non-zero 8x8 blocks of the filter bank (33 for Whisper)

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

// one mel-stage work unit: 8 mels x 16 frames, `ks` k-steps of 8 bins starting at bin `kb`
struct alignas(16) MelUnit {
  int32_t kstep0;  // first k-step (index into the B-fragment table)
  int32_t ks;      // number of k-steps
  int32_t kb;      // first bin
  int32_t nb;      // first mel
};

struct LogmelParams {
  const void* pcm;
  const int64_t* offsets;   // [B+1] (or [B] when lengths != nullptr)
  const int64_t* lengths;   // [B] or nullptr
  const float2* norm;       // (mean, rstd) per clip or nullptr
  float* out;               // (B, n_mel, n_frames)
  int32_t* mask;            // (B, n_frames) or nullptr
  uint32_t* clip_key;       // [B] running max of the scaled feature y = (log10(mel)+4)/4 as ordered key (zero-init)
  uint32_t* clip_ticket;    // [B] finished-tile counter (zero-initialised)
  uint32_t* tile_counter;   // [1] dynamic tile scheduler (zero-initialised)
  const float4* s1_consts;  // [8][25] per-warp window/twiddle block
  const float2* mel_btab;   // [n_ksteps][32] per-lane B fragments (TF32-rounded filter weights)
  const MelUnit* mel_units; // [n_units]
  float pcm_scale;
  int n_mel, n_samples, n_frames, ntiles, n_units, n_ksteps;
  uint32_t total_tiles;
};

constexpr int kSigBuf = (kSigSm + 3) & ~3;  // 16-byte multiple
__host__ __device__ inline size_t logmel_smem_bytes(int n_ksteps) {
  return (size_t)(2 * kSigBuf + kZSm) * 4 + 8 * kS1ConstVec * 16 + (size_t)n_ksteps * 32 * 8;
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

__device__ __forceinline__ void cp_async8(float* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_max_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// named barrier over the first `nthreads` threads' warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
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

struct FixEntry {
  int b, tile;       // tile < 0: nothing to do
  float floor_y;     // ((g - 8) + 4) / 4 = y_max - 2
  int silent;        // tile lies in the zero padding: store the constant instead of clamping
};

// apply the per-clip clamp to one of this CTA's own tiles (values come back from L2)
__device__ __forceinline__ void fix_tile(float* __restrict__ out, int n_mel, int n_frames, const FixEntry fx, int warp,
                                         int lane) {
  const int t0 = fx.tile * kTileF;
  if (t0 + lane >= n_frames) return;
  float* q = out + ((size_t)fx.b * n_mel) * n_frames + t0 + lane;
  if (fx.silent) {
    const float y = fmaxf(-1.5f, fx.floor_y);  // (max(-10, g-8) + 4) / 4
    for (int m = warp; m < n_mel; m += kWarps) q[(size_t)m * n_frames] = y;
  } else {
    for (int m0 = warp; m0 < n_mel; m0 += 4 * kWarps) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + j * kWarps;
        v[j] = m < n_mel ? __ldcg(q + (size_t)m * n_frames) : 3.0e38f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + j * kWarps;
        if (v[j] < fx.floor_y) q[(size_t)m * n_frames] = fx.floor_y;
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

// fill a work-item descriptor for tile id `id` (scheduler lane only): clip geometry + how its signal gets staged
template <typename T>
__device__ __forceinline__ TileDesc make_desc(const LogmelParams& p, uint32_t id) {
  TileDesc d;
  d.pad_ = 0;
  if (id >= p.total_tiles) {
    d.b = -1;
    d.tile = 0;
    d.len = 0;
    d.mode = kModeSilent;
    d.off = 0;
    return d;
  }
  d.b = (int)(id / (uint32_t)p.ntiles);
  d.tile = (int)(id - (uint32_t)d.b * (uint32_t)p.ntiles);
  d.off = __ldg(p.offsets + d.b);
  const int64_t avail = (p.lengths != nullptr ? __ldg(p.lengths + d.b) : __ldg(p.offsets + d.b + 1) - d.off);
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

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) logmel_kernel(const LogmelParams p) {
  extern __shared__ __align__(16) float smem[];
  float* const sigbuf = smem;               // two signal staging buffers (tile i -> buffer i & 1)
  float* const zbuf = smem + 2 * kSigBuf;   // stage 1 -> stage 2 exchange; the power buffer aliases it after stage 2
  float4* const s_cst = reinterpret_cast<float4*>(zbuf + kZSm);
  float2* const s_btab = reinterpret_cast<float2*>(s_cst + 8 * kS1ConstVec);
  __shared__ MelUnit s_units[kMaxUnits];
  __shared__ float s_red[2][2][kWarps];  // [tile parity][max, min][warp]
  __shared__ TileDesc s_desc[2];         // descriptor of tile k lives in slot k & 1
  __shared__ FixEntry s_fix[2];
  __shared__ int s_pend_bt[kRing];       // clip * ntiles + tile
  __shared__ float s_pend_min[kRing];    // tile minimum of y; -inf marks a silent (not yet written) tile

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float kNegInf = -__int_as_float(0x7f800000);
  const bool sched = tid == 7 * 32;  // lane 0 of warp 7 (idle in stage 2): tile scheduler + clip bookkeeping

  // ---- one-time CTA set-up ----
  for (int i = tid; i < 8 * kS1ConstVec; i += kThreads) s_cst[i] = p.s1_consts[i];
  for (int i = tid; i < p.n_ksteps * 32; i += kThreads) s_btab[i] = p.mel_btab[i];
  for (int i = tid; i < p.n_units; i += kThreads) s_units[i] = p.mel_units[i];
  if (sched) {
    const uint32_t id0 = atomicAdd(p.tile_counter, 1u);
    const uint32_t id1 = atomicAdd(p.tile_counter, 1u);
    s_desc[0] = make_desc<T>(p, id0);
    s_desc[1] = make_desc<T>(p, id1);
    s_fix[0].tile = -1;
    s_fix[1].tile = -1;
  }
  __syncthreads();

  TileDesc cur = s_desc[0], nxt = s_desc[1];
  if (cur.b >= 0 && cur.mode == kModeAsync)
    prefetch_signal(sigbuf, reinterpret_cast<const float*>(p.pcm) + cur.off + cur.tile * kTileF * kHop - kNFft / 2, tid);
  cp_async_commit();

  // scheduler-lane state: pending ring; the previous tile (its clip max is published one tile late) and the one before
  // (its ticket is published two tiles late, after the fence that opens the next scheduler block: no fence ever waits
  // on an operation issued in the same block); ticket values of the two oldest pending tiles, loaded one tile ago
  int ring_head = 0, ring_count = 0;
  int prev_b = -1, prev_tile = 0, prev_silent = 0, prev2_b = -1;
  int chk0 = -1, chk1 = -1;  // clips whose tickets were requested last tile (-1: none)
  uint32_t tk0 = 0, tk1 = 0;
  int it = 0;

  while (cur.b >= 0) {
    const int b = cur.b, tile = cur.tile, len = cur.len;
    const int t0 = tile * kTileF;
    const int s_begin = t0 * kHop - kNFft / 2;  // unpadded sample index of sig[0]
    const int nvalid = min(kTileF, p.n_frames - t0);
    const bool silent = cur.mode == kModeSilent;
    float* const sig = sigbuf + (it & 1) * kSigBuf;

    // ---- top: start the NEXT tile's loads, then make sure this tile's signal has landed ----
    uint32_t id2 = 0;
    if (sched) id2 = atomicAdd(p.tile_counter, 1u);  // id of tile it+2, consumed in stage 2
    if (nxt.b >= 0 && nxt.mode == kModeAsync)
      prefetch_signal(sigbuf + ((it + 1) & 1) * kSigBuf,
                      reinterpret_cast<const float*>(p.pcm) + nxt.off + nxt.tile * kTileF * kHop - kNFft / 2, tid);
    cp_async_commit();
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
    cp_async_wait<1>();  // everything but the group just committed (= this tile's signal) is complete
    __syncthreads();     // S1: signal visible to all warps; s_fix / s_desc from the previous stage 2 published

    // ---- clamp fix-ups decided during the previous tile (own tiles, L2-resident) ----
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const FixEntry fx = s_fix[f];
      if (fx.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fx, warp, lane);
    }

    float tmax_y = -1.5f, tmin_y = 3.0e38f;
    if (!silent) {
      // ---- stage 1: warp w owns n1 = 2w, 2w+1 ----
      stage1_pair(sig + kSigStride * lane, s_cst + warp * kS1ConstVec, 2 * warp, zbuf + lane);
      __syncthreads();  // S2
    }

    // ---- stage 2 (compute half): warps 0..5 own (k2, k2+1) = (1,2)..(11,12); warp 6 owns k2 = 0;
    //      warp 7 lane 0 runs the scheduler: bookkeeping of the PREVIOUS tile + descriptor of tile it+2 ----
    f2 pw[16];
    if (!silent) {
      if (warp < 6)
        stage2_pair_compute(zbuf + lane, 2 * warp + 1, pw);
      else if (warp == 6)
        stage2_k0_compute(zbuf + lane, pw);
    }
    if (sched) {
      // (1) one fence per tile orders last tile's relaxed operations before this tile's: ticket loads -> key loads
      //     (acquire side), clip-max RED -> ticket RED (release side).  Nothing issued in this block is waited on.
      __threadfence();
      if (prev2_b >= 0) red_add_u32(p.clip_ticket + prev2_b, 1u);
      prev2_b = prev_b;
      // (2) fix-ups for the next tile's S1, decided from the tickets requested one tile ago
      int nfix = 0;
      if (chk0 >= 0 && tk0 == (uint32_t)p.ntiles) {
        const float floor_y = key2f(__ldcg(p.clip_key + chk0)) - 2.0f;
        const int bt = s_pend_bt[ring_head];
        const float pm = s_pend_min[ring_head];
        ring_head = (ring_head + 1) & (kRing - 1);
        --ring_count;
        if (pm < floor_y) s_fix[nfix++] = FixEntry{chk0, bt - chk0 * p.ntiles, floor_y, pm == kNegInf};
        if (chk1 >= 0 && tk1 == (uint32_t)p.ntiles) {
          const float floor1 = key2f(__ldcg(p.clip_key + chk1)) - 2.0f;
          const int bt1 = s_pend_bt[ring_head];
          const float pm1 = s_pend_min[ring_head];
          ring_head = (ring_head + 1) & (kRing - 1);
          --ring_count;
          if (pm1 < floor1) s_fix[nfix++] = FixEntry{chk1, bt1 - chk1 * p.ntiles, floor1, pm1 == kNegInf};
        }
      }
      for (int f = nfix; f < 2; ++f) s_fix[f].tile = -1;
      // (3) previous tile: publish its clip max, remember it in the ring
      if (prev_b >= 0) {
        float mx = -1.5f, mn = kNegInf;  // silent: max = (log10(1e-10)+4)/4, min marker = -inf
        if (!prev_silent) {
          const int pp = (it + 1) & 1;
          mx = s_red[pp][0][0];
          mn = s_red[pp][1][0];
#pragma unroll
          for (int w = 1; w < kWarps; ++w) {
            mx = fmaxf(mx, s_red[pp][0][w]);
            mn = fminf(mn, s_red[pp][1][w]);
          }
        }
        red_max_u32(p.clip_key + prev_b, f2key(mx));
        // ring full: cannot happen while ntiles <= kRing unless other CTAs lag a whole clip behind; the oldest entry's
        // clip then has every tile assigned to a RUNNING CTA (ids are handed out in order) whose scheduler publishes
        // before it ever waits, so this wait terminates
        if (ring_count == kRing) {
          const int bt = s_pend_bt[ring_head];
          const float pm = s_pend_min[ring_head];
          const int ob = bt / p.ntiles;
          ring_head = (ring_head + 1) & (kRing - 1);
          --ring_count;
          while (ld_acquire_u32(p.clip_ticket + ob) != (uint32_t)p.ntiles) __nanosleep(200);
          const float fl = key2f(__ldcg(p.clip_key + ob)) - 2.0f;
          if (pm < fl) {
            const FixEntry fx{ob, bt - ob * p.ntiles, fl, pm == kNegInf};
            for (int l = 0; l < 32; ++l)
              for (int w = 0; w < kWarps; ++w) fix_tile(p.out, p.n_mel, p.n_frames, fx, w, l);
          }
        }
        const int slot = (ring_head + ring_count) & (kRing - 1);
        s_pend_bt[slot] = prev_b * p.ntiles + prev_tile;
        s_pend_min[slot] = mn;
        ++ring_count;
      }
      prev_b = b;
      prev_tile = tile;
      prev_silent = silent;
      // (4) request the tickets of the two oldest pending tiles; looked at one tile from now
      chk0 = chk1 = -1;
      if (ring_count > 0) {
        chk0 = s_pend_bt[ring_head] / p.ntiles;
        tk0 = ld_relaxed_u32(p.clip_ticket + chk0);
      }
      if (ring_count > 1) {
        chk1 = s_pend_bt[(ring_head + 1) & (kRing - 1)] / p.ntiles;
        tk1 = ld_relaxed_u32(p.clip_ticket + chk1);
      }
      // (5) descriptor of tile it+2 (slot of tile it, whose descriptor is in registers)
      s_desc[it & 1] = make_desc<T>(p, id2);
    }

    if (!silent) {
      if (warp < 7) {
        bar_sync_named(1, 7 * 32);  // S2b (warps 0..6): all z planes have been read; the power buffer may overwrite them
        if (warp < 6)
          stage2_pair_store(pw, 2 * warp + 1, zbuf + lane);
        else
          stage2_k0_store(pw, zbuf + lane);
      }
      __syncthreads();  // S3
      // ---- stage 3: banded mel projection with mma.sync TF32, epilogue log10 / scale / store ----
      {
        const int g = lane >> 2, t = lane & 3;
        const int mt = (warp & 1) * 16;  // units come in (frames 0..15, frames 16..31) pairs per 8-mel tile
        // this thread's outputs per unit: frames mt+g, mt+g+8 (columns) x mels nb+2t, nb+2t+1 (rows); a warp's units
        // are 8 apart, i.e. 4 mel tiles = 32 rows apart: two running row pointers, no per-unit multiplies
        float* q0 = p.out + ((size_t)b * p.n_mel + 4 * (warp >> 1) * 2 + 2 * t) * p.n_frames + t0 + mt + g;
        const size_t row = (size_t)p.n_frames, step = 32 * row;
        const bool full = nvalid == kTileF;
        const int4* up = reinterpret_cast<const int4*>(s_units) + warp;
        const float* abase = zbuf + t * kPStride + mt + g;
        for (int u = warp; u < p.n_units; u += kWarps, up += kWarps, q0 += step) {
          const int4 mu = *up;  // kstep0, ks, kb, nb
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          const float* arow = abase + mu.z * kPStride;
          const float2* brow = s_btab + mu.x * 32 + lane;
#pragma unroll 2
          for (int s = 0; s < mu.y; ++s) {
            // fp32 bit patterns go in as-is: the tensor core reads the top 19 bits (truncation); the filter weights
            // carry a (1 + 2^-11) factor that centres the truncation error (see wfe_api.cu)
            uint32_t a[4];
            a[0] = __float_as_uint(arow[0]);
            a[1] = __float_as_uint(arow[8]);
            a[2] = __float_as_uint(arow[4 * kPStride]);
            a[3] = __float_as_uint(arow[4 * kPStride + 8]);
            const float2 bw = *brow;
            mma_tf32_16x8x8(acc, a, __float_as_uint(bw.x), __float_as_uint(bw.y));
            arow += 8 * kPStride;
            brow += 32;
          }
          // c0: (frame mt+g, mel nb+2t)  c1: (mt+g, nb+2t+1)  c2: (mt+g+8, nb+2t)  c3: (mt+g+8, nb+2t+1)
          const float y0 = logmel_feature(acc[0]), y1 = logmel_feature(acc[1]);
          const float y2 = logmel_feature(acc[2]), y3 = logmel_feature(acc[3]);
          float* q1 = q0 + row;
          if (full && mu.w + 8 <= p.n_mel) {
            q0[0] = y0;
            q1[0] = y1;
            q0[8] = y2;
            q1[8] = y3;
            tmax_y = fmaxf(tmax_y, fmaxf(fmaxf(y0, y1), fmaxf(y2, y3)));
            tmin_y = fminf(tmin_y, fminf(fminf(y0, y1), fminf(y2, y3)));
          } else {
            const bool f0 = mt + g < nvalid, f1 = mt + g + 8 < nvalid;
            const bool m0 = mu.w + 2 * t < p.n_mel, m1 = mu.w + 2 * t + 1 < p.n_mel;
            if (f0 && m0) { q0[0] = y0; tmax_y = fmaxf(tmax_y, y0); tmin_y = fminf(tmin_y, y0); }
            if (f0 && m1) { q1[0] = y1; tmax_y = fmaxf(tmax_y, y1); tmin_y = fminf(tmin_y, y1); }
            if (f1 && m0) { q0[8] = y2; tmax_y = fmaxf(tmax_y, y2); tmin_y = fminf(tmin_y, y2); }
            if (f1 && m1) { q1[8] = y3; tmax_y = fmaxf(tmax_y, y3); tmin_y = fminf(tmin_y, y3); }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        tmax_y = fmaxf(tmax_y, __shfl_xor_sync(0xffffffffu, tmax_y, o));
        tmin_y = fminf(tmin_y, __shfl_xor_sync(0xffffffffu, tmin_y, o));
      }
      if (lane == 0) {
        s_red[it & 1][0][warp] = tmax_y;
        s_red[it & 1][1][warp] = tmin_y;
      }
    }
    __syncthreads();  // S4: tile written (visible to this CTA); s_desc / s_fix / s_red published; smem free
    cur = nxt;
    nxt = s_desc[it & 1];
    ++it;
  }
  cp_async_wait<0>();

  // ---- epilogue: publish the last tile, then drain the tiles this CTA still has pending (every remaining tile of
  //      their clips is owned by a running CTA, so the waits terminate) ----
  if (sched) {
    __threadfence();
    if (prev2_b >= 0) red_add_u32(p.clip_ticket + prev2_b, 1u);
    if (prev_b >= 0) {
      float mx = -1.5f, mn = kNegInf;
      if (!prev_silent) {
        const int pp = (it + 1) & 1;
        mx = s_red[pp][0][0];
        mn = s_red[pp][1][0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
          mx = fmaxf(mx, s_red[pp][0][w]);
          mn = fminf(mn, s_red[pp][1][w]);
        }
      }
      red_max_u32(p.clip_key + prev_b, f2key(mx));
      __threadfence();
      red_add_u32(p.clip_ticket + prev_b, 1u);
      if (ring_count < kRing) {
        const int slot = (ring_head + ring_count) & (kRing - 1);
        s_pend_bt[slot] = prev_b * p.ntiles + prev_tile;
        s_pend_min[slot] = mn;
        ++ring_count;
      } else {  // ring full (see above): fix this one serially once its clip completes
        while (ld_acquire_u32(p.clip_ticket + prev_b) != (uint32_t)p.ntiles) __nanosleep(200);
        const float fl = key2f(__ldcg(p.clip_key + prev_b)) - 2.0f;
        if (mn < fl) {
          const FixEntry fx{prev_b, prev_tile, fl, mn == kNegInf};
          for (int l = 0; l < 32; ++l)
            for (int w = 0; w < kWarps; ++w) fix_tile(p.out, p.n_mel, p.n_frames, fx, w, l);
        }
      }
    }
  }
  // the fix-ups decided during the last tile's stage 2 were published by its S4
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const FixEntry fx = s_fix[f];
    if (fx.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fx, warp, lane);
  }
  __syncthreads();
  for (;;) {
    if (sched) {
      s_fix[0].tile = -2;  // -2: ring empty
      if (ring_count > 0) {
        const int bt = s_pend_bt[ring_head];
        const float pm = s_pend_min[ring_head];
        const int ob = bt / p.ntiles;
        ring_head = (ring_head + 1) & (kRing - 1);
        --ring_count;
        while (ld_acquire_u32(p.clip_ticket + ob) != (uint32_t)p.ntiles) __nanosleep(100);
        const float fl = key2f(__ldcg(p.clip_key + ob)) - 2.0f;
        s_fix[0] = FixEntry{ob, pm < fl ? bt - ob * p.ntiles : -1, fl, pm == kNegInf};
      }
    }
    __syncthreads();
    const FixEntry fx = s_fix[0];
    if (fx.tile == -2) break;
    if (fx.tile >= 0) fix_tile(p.out, p.n_mel, p.n_frames, fx, warp, lane);
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
