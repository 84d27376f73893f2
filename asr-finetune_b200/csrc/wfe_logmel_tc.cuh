// Whisper log-mel frontend on the 5th-generation tensor cores (tcgen05 + TMEM) of sm_100a: the throughput kernel.
//
// Why tensor cores at all: the CUDA-core kernel (wfe_logmel.cuh) is issue-bound at 8.8 k warp-instructions per 32 frames
// (r01 profile).  Here the bulk of the DFT arithmetic moves to tcgen05.mma and the CUDA cores keep ~4 k.
//
// Arithmetic (restated from HF:models/whisper/feature_extraction_whisper.py:135-164):
//   decimation in time, 400 = 4 x 100:  n = n1 + 4 n2,  k = k2 + 100 k1
//     Y_n1[k2] = sum_n2 (w x)[n1 + 4 n2] W100^(n2 k2)          four REAL-input 100-point DFTs per frame, k2 = 0..50
//     X[k2 + 100 k1] = sum_n1 W4^(n1 k1) W400^(n1 k2) Y_n1[k2]  twiddle + 4-point DFT, on the CUDA cores
//   The 100-point DFTs are GEMMs: D_n1 (128 frames x 104) = A_n1 (128 x 100) . B (100 x 104), B = [cos | -sin] of the
//   DFT-100 matrix, the same for every n1.  fp32 accuracy on fp16 tensor cores by operand splitting: a = a_hi + a_lo,
//   b = b_hi + b_lo (fp16 each), D = a_hi b_hi + a_hi b_lo + a_lo b_hi with fp32 accumulation in TMEM; the frame tile is
//   pre-scaled by a power of two so that its largest sample sits in [2^14, 2^15) (undone exactly in the log domain).
//   Measured (tools/ubench_tcgen05.cu, profiles/r02_ubench_tcgen05.txt): 2^-19.6 of the row maximum, 60.8 cycles per
//   128x112x16 MMA.
//
// One persistent CTA per SM, 11 warps, tile = 128 consecutive frames of one clip (24 tiles per 30-s clip), tile ids
// strided statically over the CTAs (every role derives the same sequence):
//   warps 0-3   PREP      thread = frame: raw samples (smem) -> window, scale, split hi/lo -> A operand ring (smem,
//                         K-major no-swizzle core matrices, one stage = one k-step of 16 n2 for all four n1)
//   warps 4-7   EPILOGUE  thread = frame = TMEM lane: tcgen05.ld, twiddle + DFT-4 + power (packed f32x2), banded mel with
//                         immediate weights (generated straight-line code), log10, (x+4)/4, store, tile min/max
//   warp  8     MMA       lane 0 issues 12 tcgen05.mma per k-step (4 n1 x 3 passes), commits to mbarriers
//   warp  9     LOADER    raw tile: 130 cp.async.bulk row copies (TMA) into a padded hop-row layout
//   warp 10     CLAMP     per-clip max-8 clamp books (same scheme as wfe_logmel.cuh: one published key per tile, fix-ups
//                         of this CTA's own tiles from L2 once their clip is complete)
// Specialised for n_samples = 480000 (3000 frames), fp32 accumulate, n_mel in {80, 128} with the slaney structure baked by
// tools/gen_tc_epilogue.py; everything else runs on the CUDA-core kernel.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "wfe_logmel.cuh"

namespace wfe {
namespace tc {

constexpr int kTileM = 128;                       // frames per tile = MMA M = TMEM lanes
constexpr int kNFrames = 3000;                    // compile-time: store offsets become immediates
constexpr int kNSamples = kNFrames * kHop;        // 480000
constexpr int kNTiles = (kNFrames + kTileM - 1) / kTileM;  // 24
constexpr int kRawLen = (kTileM - 1) * kHop + kNFft;       // 20720 samples per tile
constexpr int kRawRows = (kRawLen + kHop - 1) / kHop;      // 130 hop rows
constexpr int kRawPitch = kHop + 4;               // floats: lane stride 164 = 4 (mod 32) -> conflict-free LDS.128 by frame
constexpr int kRawFloats = kRawRows * kRawPitch;  // 21320
constexpr int kKSteps = 7;                        // 100 n2 padded to 112 = 7 x 16
constexpr int kN = 112;                           // MMA N: 52 k2 x (re, im) = 104, padded to a multiple of 16
constexpr int kChunkBytes = kTileM * 16;          // one 8-element K chunk of 128 rows
constexpr int kStageBytes = 4 * 2 * 2 * kChunkBytes;  // [n1][hi/lo][chunk][row][16 B] = 32768
constexpr int kStages = 2;
constexpr int kBChunkBytes = kN * 16;             // 1792
constexpr int kBBytes = 2 * 14 * kBChunkBytes;    // [hi/lo][chunk 14][n 112][16 B] = 50176
constexpr int kTwBytes = 26 * 3 * 16;             // [pair][n1-1] (cos_k2, cos_k2+1, sin_k2, sin_k2+1)
constexpr int kThreads = 11 * 32;
constexpr int kTmemCols = 512;
constexpr int kRing = 128;

constexpr size_t kSmemRaw = 0;
constexpr size_t kSmemA = kSmemRaw + (size_t)kRawFloats * 4;          // 85280
constexpr size_t kSmemB = kSmemA + (size_t)kStages * kStageBytes;     // +65536
constexpr size_t kSmemWs = kSmemB + kBBytes;                          // +50176
constexpr size_t kSmemTw = kSmemWs + 400 * 4;
constexpr size_t kSmemBytes = kSmemTw + kTwBytes;                     // 203840

#ifdef WFE_TC_TRACE
// timing-trace build (diagnostics only): CTA 0 stamps clock64() at role milestones of its tile iterations 4..11
constexpr int kTrTiles = 8, kTrRoles = 5, kTrPts = 16;
__device__ unsigned long long g_tc_trace[kTrTiles * kTrRoles * kTrPts];
#define TCT(role, it, pt)                                                                      \
  do {                                                                                         \
    if (blockIdx.x == 0 && (it) >= 4 && (it) < 4 + kTrTiles)                                   \
      g_tc_trace[(((it) - 4) * kTrRoles + (role)) * kTrPts + (pt)] = clock64();                \
  } while (0)
#else
#define TCT(role, it, pt) \
  do {                    \
  } while (0)
#endif

struct TcParams {
  const void* pcm;
  const int64_t* offsets;
  const int64_t* lengths;
  const float2* norm;
  void* out;                // (B, n_mel, 3000), element type = template OutT
  int32_t* mask;
  uint32_t* tile_key;       // [B][24]
  const uint4* b_mat;       // kBBytes: DFT-100 operand, canonical layout, hi then lo
  const float4* tw;         // kTwBytes: twiddles W400^(n1 k2)
  const float* win;         // [400] periodic Hann, fp32
  float pcm_scale;
  int pcm_dtype;            // 0 = float32, 1 = int16, 2 = float16 (wfe_pcm_dtype)
  int n_mel;
  uint32_t total_tiles;
};

// the generic staging path (tile edges, 2-byte PCM, normalisation) is the only place the PCM element type matters
__device__ __forceinline__ float load_pcm(const void* pcm, int dtype, int64_t i, float scale) {
  if (dtype == 0) return reinterpret_cast<const float*>(pcm)[i];
  if (dtype == 1) return (float)reinterpret_cast<const int16_t*>(pcm)[i] * scale;
  return __half2float(reinterpret_cast<const __half*>(pcm)[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifndef WFE_TC_WAIT
#define WFE_TC_WAIT 1  // 0: bare try_wait spin, 1: try_wait with a suspend-time hint, 2: nanosleep back-off between polls
#endif
__device__ __forceinline__ bool mbar_try_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Waits for the phase with the given parity.  Gives up after a few seconds so that a protocol bug turns into a wrong
// answer + error flag instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t* err_flag) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
#pragma unroll 1
  for (uint32_t spin = 0;; ++spin) {
#if WFE_TC_WAIT == 2
    __nanosleep(40);
    if (mbar_try(bar, parity)) return;
#elif WFE_TC_WAIT == 1
    if (mbar_try_hint(bar, parity, 4000u)) return;
#else
    if (mbar_try(bar, parity)) return;
#endif
    if ((spin & 1023u) == 1023u && clock64() - t0 > 6000000000ll) break;
  }
  if (err_flag != nullptr) atomicExch(err_flag, 0xDEADu);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): LBO = byte distance between the two
// 8-element K chunks of one MMA, SBO = byte distance between 8-row core matrices (128: rows are contiguous 16-byte units)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A = B = fp16, both K-major, M = 128
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// tile geometry: every role derives it from the tile id alone
// ---------------------------------------------------------------------------------------------------------------
struct Tile {
  int b, tile, len, mode;  // mode: kModeSilent / kModeAsync (TMA bulk) / kModeSync (generic staging)
  int64_t off;
};
__device__ __forceinline__ Tile tile_info(const TcParams& p, uint32_t id) {
  Tile t;
  t.b = (int)(id / (uint32_t)kNTiles);
  t.tile = (int)(id - (uint32_t)t.b * (uint32_t)kNTiles);
  t.off = __ldg(p.offsets + t.b);
  const int64_t avail = p.lengths != nullptr ? __ldg(p.lengths + t.b) : __ldg(p.offsets + t.b + 1) - t.off;
  t.len = (int)(avail < (int64_t)kNSamples ? avail : (int64_t)kNSamples);
  const int s_begin = t.tile * kTileM * kHop - kNFft / 2;
  // lowest source sample any VALID frame of the tile touches (frames beyond 3000 are never stored)
  const int nvalid = min(kTileM, kNFrames - t.tile * kTileM);
  const int s_hi = s_begin + (nvalid - 1) * kHop + kNFft - 1;
  int lowest = s_begin < 0 ? 0 : s_begin;
  if (s_hi >= kNSamples) lowest = min(lowest, 2 * (kNSamples - 1) - s_hi);
  if (lowest >= t.len) {
    t.mode = kModeSilent;
  } else {
    const float* src = reinterpret_cast<const float*>(p.pcm) + t.off + s_begin;
    const bool bulk = p.pcm_dtype == 0 && p.norm == nullptr && s_begin >= 0 && s_begin + kRawLen <= t.len &&
                      (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    t.mode = bulk ? kModeAsync : kModeSync;
  }
  return t;
}

template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename OutT>
__device__ __forceinline__ float from_out(OutT v);
template <>
__device__ __forceinline__ float from_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ float from_out<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float from_out<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// per-clip clamp applied to one 128-frame tile of this CTA by ONE warp: lane owns 4 consecutive frames of a mel row,
// 8 rows in flight.  silent: store the constant without reading.
template <typename OutT>
__device__ __forceinline__ void fix_tile_tc(OutT* __restrict__ out, int n_mel, int b, int tile, float fl, bool silent,
                                            int lane) {
  const int t0 = tile * kTileM;
  const int nvalid = min(kTileM, kNFrames - t0);  // multiple of 4 (3000 = 23 * 128 + 56)
  if (4 * lane >= nvalid) return;
  using Vec = typename std::conditional<sizeof(OutT) == 4, float4, uint2>::type;
  OutT* const base = out + (size_t)b * n_mel * kNFrames + t0 + 4 * lane;
  OutT cv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) cv[e] = to_out<OutT>(fl);
  const Vec cvec = *reinterpret_cast<const Vec*>(cv);
  constexpr int kDeep = 8;
  for (int m0 = 0; m0 < n_mel; m0 += kDeep) {
    if (silent) {
#pragma unroll
      for (int j = 0; j < kDeep; ++j)
        if (m0 + j < n_mel) *reinterpret_cast<Vec*>(base + (size_t)(m0 + j) * kNFrames) = cvec;
    } else {
      Vec v[kDeep];
#pragma unroll
      for (int j = 0; j < kDeep; ++j)
        if (m0 + j < n_mel) v[j] = __ldcg(reinterpret_cast<const Vec*>(base + (size_t)(m0 + j) * kNFrames));
#pragma unroll
      for (int j = 0; j < kDeep; ++j) {
        if (m0 + j >= n_mel) continue;
        OutT e[4];
        *reinterpret_cast<Vec*>(e) = v[j];
        bool need = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float f = from_out<OutT>(e[k]);
          if (f < fl) {  // (-inf, the log of a zero mel power, is below every floor)
            need = true;
            e[k] = cv[k];
          }
        }
        if (need) *reinterpret_cast<Vec*>(base + (size_t)(m0 + j) * kNFrames) = *reinterpret_cast<const Vec*>(e);
      }
    }
  }
}

__device__ __forceinline__ float wait_clip_floor_tc(const uint32_t* tile_key, int b, int lane) {
  const uint32_t* row = tile_key + (size_t)b * kNTiles;
  for (;;) {
    uint32_t k = lane < kNTiles ? ld_relaxed_u32(row + lane) : 1u;
    const bool zero = __any_sync(0xffffffffu, k == 0);
    k = __reduce_max_sync(0xffffffffu, k);
    if (!zero) return fmaxf(key2f(k) - 2.0f, -1.5f);
    __nanosleep(200);
  }
}

// named barrier among the 128 prep threads
__device__ __forceinline__ void prep_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <typename OutT, int kNMel>
__global__ void __launch_bounds__(kThreads, 1) logmel_tc_kernel(const TcParams p, uint32_t* err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* const raw = reinterpret_cast<float*>(smem + kSmemRaw);
  uint8_t* const a_ring = smem + kSmemA;
  uint8_t* const b_sm = smem + kSmemB;
  float* const ws = reinterpret_cast<float*>(smem + kSmemWs);
  const float4* const tw_sm = reinterpret_cast<const float4*>(smem + kSmemTw);

  __shared__ uint64_t bar_raw_full, bar_raw_empty, bar_a_full[kStages], bar_a_empty[kStages], bar_d_full, bar_d_empty,
      bar_st_full[2], bar_st_empty[2];
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_pmax[4];        // prep: per-warp max |x| bits
  __shared__ float s_tilek[2];          // per tile parity: additive constant of the log-domain un-scaling
  __shared__ float s_red[2][2][4];      // [tile parity][max, min][epilogue warp] of y over the warp's 32 frames
  __shared__ int2 s_pend_bt[kRing];
  __shared__ float2 s_pend_mm[kRing];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- one-time set-up ----
  for (int i = tid; i < kBBytes / 16; i += kThreads) reinterpret_cast<uint4*>(b_sm)[i] = p.b_mat[i];
  for (int i = tid; i < kTwBytes / 16; i += kThreads) reinterpret_cast<float4*>(smem + kSmemTw)[i] = p.tw[i];
  fence_async_smem();  // B is read by the tensor core (async proxy)
  if (tid == 0) {
    mbar_init(&bar_raw_full, 1);
    mbar_init(&bar_raw_empty, 128);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_a_full[s], 128);
      mbar_init(&bar_a_empty[s], 1);
    }
    mbar_init(&bar_d_full, 1);
    mbar_init(&bar_d_empty, 128);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_st_full[s], 4);
      mbar_init(&bar_st_empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(&s_tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp < 4) {
    // =========================================== PREP ===========================================
    const int m = tid;  // frame within the tile
    const float* const xrow = raw + m * kRawPitch;
    uint32_t ks = 0;   // running k-step count (A ring position)
    uint32_t nt = 0;   // running count of non-silent tiles
    for (uint32_t id = blockIdx.x; id < p.total_tiles; id += gridDim.x) {
      const Tile t = tile_info(p, id);
      if (t.mode == kModeSilent) continue;
      if (tid == 0) TCT(0, nt, 0);
      mbar_wait(&bar_raw_full, nt & 1, err_flag);
      if (tid == 0) TCT(0, nt, 1);
      if (t.mode == kModeSync) {
        // generic staging: truncate / right-zero-pad to 30 s, centred reflect pad, dtype conversion, normalisation
        const int s_begin = t.tile * kTileM * kHop - kNFft / 2;
        float mean = 0.f, rstd = 1.f;
        if (p.norm != nullptr) {
          const float2 st = __ldg(p.norm + t.b);
          mean = st.x;
          rstd = st.y;
        }
        for (int i = tid; i < kRawRows * kHop; i += 128) {
          int s = s_begin + i;
          if (s < 0) s = -s;
          if (s >= kNSamples) s = 2 * (kNSamples - 1) - s;
          float v = 0.f;
          if (s >= 0 && s < t.len) {
            v = load_pcm(p.pcm, p.pcm_dtype, t.off + s, p.pcm_scale);
            if (p.norm != nullptr) v = (v - mean) * rstd;
          }
          const int r = i / kHop;
          raw[r * kRawPitch + (i - r * kHop)] = v;
        }
        prep_bar();
      }
      if (tid == 0) TCT(0, nt, 2);
      // ---- tile maximum -> power-of-two scale, scaled window table ----
      uint32_t mx = 0;
      for (int i = tid; i < kRawRows * (kHop / 4); i += 128) {
        const int r = i / (kHop / 4), c = i - r * (kHop / 4);
        float4 v = *reinterpret_cast<const float4*>(raw + r * kRawPitch + 4 * c);
        if (r == kRawRows - 1 && 4 * c >= kRawLen - (kRawRows - 1) * kHop) v = make_float4(0.f, 0.f, 0.f, 0.f);  // beyond the tile
        const float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
        mx = max(mx, __float_as_uint(a));
      }
      mx = __reduce_max_sync(0xffffffffu, mx);
      if (lane == 0) s_pmax[warp] = mx;
      prep_bar();
      mx = max(max(s_pmax[0], s_pmax[1]), max(s_pmax[2], s_pmax[3]));
      // scale = 2^(14 - e) with e = unbiased exponent of the maximum (clamped so that the scale stays a normal float)
      int e = (int)((mx >> 23) & 0xffu) - 127;
      if (mx == 0u) e = 14;
      e = max(-100, min(e, 100));
      const float scale = __uint_as_float((uint32_t)(127 + 14 - e) << 23);
      if (tid < 100) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.win) + tid);
        reinterpret_cast<float4*>(ws)[tid] = make_float4(w.x * scale, w.y * scale, w.z * scale, w.w * scale);
      }
      if (tid == 0) {
        // y = (log10(mel_scaled * 2^(-2 (14 - e))) + 4) / 4 = lg2(mel_scaled) * C + (1 - 2 (14 - e) C)
        s_tilek[nt & 1] = 1.0f - (float)(2 * (14 - e)) * (0.25f * kLog10_2);
      }
      prep_bar();
      if (tid == 0) TCT(0, nt, 3);

      // ---- k-steps: 16 n2 (= 64 consecutive samples) for all four n1 ----
#pragma unroll
      for (int j = 0; j < kKSteps; ++j, ++ks) {  // unrolled: every smem offset below is an immediate
        const uint32_t stage = ks & 1u;
        mbar_wait(&bar_a_empty[stage], ((ks >> 1) & 1u) ^ 1u, err_flag);
        uint8_t* const st_base = a_ring + stage * kStageBytes + m * 16;
        const int nq = (j == kKSteps - 1) ? 4 : 16;  // valid n2 in this k-step (n2 < 100)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float h[4][8], l[4][8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int qq = 8 * c + q;
            if (qq < nq) {
              const int n = 64 * j + 4 * qq;  // sample index within the frame
              const int roff = (n / kHop) * kRawPitch + (n % kHop);
              const float4 x = *reinterpret_cast<const float4*>(xrow + roff);
              const float4 w = *reinterpret_cast<const float4*>(ws + n);
              const float y[4] = {x.x * w.x, x.y * w.y, x.z * w.z, x.w * w.w};
#pragma unroll
              for (int n1 = 0; n1 < 4; ++n1) {
                const float hh = __uint_as_float(__float_as_uint(y[n1]) & 0xFFFFE000u);  // 11 significant bits: exact in fp16
                h[n1][q] = hh;
                l[n1][q] = y[n1] - hh;
              }
            } else {
#pragma unroll
              for (int n1 = 0; n1 < 4; ++n1) {
                h[n1][q] = 0.f;
                l[n1][q] = 0.f;
              }
            }
          }
#pragma unroll
          for (int n1 = 0; n1 < 4; ++n1) {
            const uint4 hv = make_uint4(pack_h2(h[n1][0], h[n1][1]), pack_h2(h[n1][2], h[n1][3]),
                                        pack_h2(h[n1][4], h[n1][5]), pack_h2(h[n1][6], h[n1][7]));
            const uint4 lv = make_uint4(pack_h2(l[n1][0], l[n1][1]), pack_h2(l[n1][2], l[n1][3]),
                                        pack_h2(l[n1][4], l[n1][5]), pack_h2(l[n1][6], l[n1][7]));
            *reinterpret_cast<uint4*>(st_base + ((n1 * 2 + 0) * 2 + c) * kChunkBytes) = hv;
            *reinterpret_cast<uint4*>(st_base + ((n1 * 2 + 1) * 2 + c) * kChunkBytes) = lv;
          }
        }
        fence_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(&bar_a_full[stage]);
        if (tid == 0) TCT(0, nt, 4 + j);
      }
      mbar_arrive(&bar_raw_empty);  // this thread is done reading the raw tile
      ++nt;
    }
  } else if (warp < 8) {
    // =========================================== EPILOGUE ===========================================
    const int ew = warp - 4;                 // == warp % 4: TMEM lanes 32 ew .. 32 ew + 31
    const int m = ew * 32 + lane;            // frame within the tile
    const uint32_t tlane = tmem + ((uint32_t)(ew * 32) << 16);
    uint32_t nt = 0;
    for (uint32_t id = blockIdx.x; id < p.total_tiles; id += gridDim.x) {
      const Tile t = tile_info(p, id);
      if (t.mode == kModeSilent) continue;
      const int t0 = t.tile * kTileM;
      const bool valid = t0 + m < kNFrames;
      OutT* const obase = reinterpret_cast<OutT*>(p.out) + (size_t)t.b * kNMel * kNFrames + t0 + m;
      if (p.mask != nullptr && valid) p.mask[(size_t)t.b * kNFrames + t0 + m] = ((t0 + m) * kHop < t.len) ? 1 : 0;
      if (m == 0) TCT(1, nt, 0);
      mbar_wait(&bar_d_full, nt & 1u, err_flag);
      tc_fence_after();
      if (m == 0) TCT(1, nt, 1);
      const float tile_k = s_tilek[nt & 1u];
      uint32_t rmax = 0u, rmin = 0x7f800000u;
      uint32_t q0[16], q1[16], q2[16], q3[16];
      f2 P0, P1, P2, P3;

#define TC_LOAD(g)                                  \
  tmem_ld16(tlane + 0 * kN + 16 * (g), q0);         \
  tmem_ld16(tlane + 1 * kN + 16 * (g), q1);         \
  tmem_ld16(tlane + 2 * kN + 16 * (g), q2);         \
  tmem_ld16(tlane + 3 * kN + 16 * (g), q3);         \
  tmem_ld_wait();
      // columns of pair i within the group: (re k2, re k2+1, im k2, im k2+1).  T_n1 = Y_n1 * (cos - i sin):
      //   re = yr c + yi s, im = yi c - yr s; then the 4-point DFT over n1 and the four powers
#define TC_PAIR(i, pp)                                                                                     \
  {                                                                                                        \
    const f2 y0r = mk2(__uint_as_float(q0[4 * (i)]), __uint_as_float(q0[4 * (i) + 1]));                     \
    const f2 y0i = mk2(__uint_as_float(q0[4 * (i) + 2]), __uint_as_float(q0[4 * (i) + 3]));                 \
    f2 tr[3], ti[3];                                                                                       \
    {                                                                                                      \
      const float4 w = tw_sm[(pp) * 3 + 0];                                                                \
      const f2 yr = mk2(__uint_as_float(q1[4 * (i)]), __uint_as_float(q1[4 * (i) + 1]));                    \
      const f2 yi = mk2(__uint_as_float(q1[4 * (i) + 2]), __uint_as_float(q1[4 * (i) + 3]));                \
      tr[0] = vfma(yr, mk2(w.x, w.y), vmul(yi, mk2(w.z, w.w)));                                            \
      ti[0] = vfma(yi, mk2(w.x, w.y), -vmul(yr, mk2(w.z, w.w)));                                           \
    }                                                                                                      \
    {                                                                                                      \
      const float4 w = tw_sm[(pp) * 3 + 1];                                                                \
      const f2 yr = mk2(__uint_as_float(q2[4 * (i)]), __uint_as_float(q2[4 * (i) + 1]));                    \
      const f2 yi = mk2(__uint_as_float(q2[4 * (i) + 2]), __uint_as_float(q2[4 * (i) + 3]));                \
      tr[1] = vfma(yr, mk2(w.x, w.y), vmul(yi, mk2(w.z, w.w)));                                            \
      ti[1] = vfma(yi, mk2(w.x, w.y), -vmul(yr, mk2(w.z, w.w)));                                           \
    }                                                                                                      \
    {                                                                                                      \
      const float4 w = tw_sm[(pp) * 3 + 2];                                                                \
      const f2 yr = mk2(__uint_as_float(q3[4 * (i)]), __uint_as_float(q3[4 * (i) + 1]));                    \
      const f2 yi = mk2(__uint_as_float(q3[4 * (i) + 2]), __uint_as_float(q3[4 * (i) + 3]));                \
      tr[2] = vfma(yr, mk2(w.x, w.y), vmul(yi, mk2(w.z, w.w)));                                            \
      ti[2] = vfma(yi, mk2(w.x, w.y), -vmul(yr, mk2(w.z, w.w)));                                           \
    }                                                                                                      \
    const f2 s02r = y0r + tr[1], s02i = y0i + ti[1], d02r = y0r - tr[1], d02i = y0i - ti[1];               \
    const f2 s13r = tr[0] + tr[2], s13i = ti[0] + ti[2], d13r = tr[0] - tr[2], d13i = ti[0] - ti[2];       \
    const f2 x0r = s02r + s13r, x0i = s02i + s13i, x2r = s02r - s13r, x2i = s02i - s13i;                   \
    const f2 x1r = d02r + d13i, x1i = d02i - d13r, x3r = d02r - d13i, x3i = d02i + d13r;                   \
    P0 = vfma(x0r, x0r, vmul(x0i, x0i));                                                                   \
    P1 = vfma(x1r, x1r, vmul(x1i, x1i));                                                                   \
    P2 = vfma(x2r, x2r, vmul(x2i, x2i));                                                                   \
    P3 = vfma(x3r, x3r, vmul(x3i, x3i));                                                                   \
  }
#define TC_RELEASE()         \
  tc_fence_before();         \
  mbar_arrive(&bar_d_empty); \
  if (m == 0) TCT(1, nt, 2);
#define TC_ACC_SET(mm, pexpr, wbits) float a_##mm = (pexpr) * __uint_as_float(wbits);
#define TC_ACC(mm, pexpr, wbits) a_##mm = fmaf((pexpr), __uint_as_float(wbits), a_##mm);
#define TC_FIN_ZERO(mm) \
  float a_##mm = 0.f;   \
  TC_FIN(mm)
#define TC_FIN(mm)                                                                            \
  if (valid) {                                                                                \
    const uint32_t u_ = __float_as_uint(a_##mm);                                              \
    rmax = max(rmax, u_);                                                                     \
    rmin = min(rmin, u_);                                                                     \
    obase[(mm) * kNFrames] = to_out<OutT>(fmaf(lg2_approx(a_##mm), 0.25f * kLog10_2, tile_k)); \
  }
      if constexpr (kNMel == 128) {
#define WFE_TC_GEN_NMEL 128
#include "wfe_tc_epilogue_gen.inc"
      } else {
#undef WFE_TC_GEN_NMEL
#define WFE_TC_GEN_NMEL 80
#include "wfe_tc_epilogue_gen.inc"
      }
#undef WFE_TC_GEN_NMEL
#undef TC_LOAD
#undef TC_PAIR
#undef TC_RELEASE
#undef TC_ACC_SET
#undef TC_ACC
#undef TC_FIN
#undef TC_FIN_ZERO
      // ---- tile extrema (raw scaled mel powers, >= 0: uint order == float order) -> clamp warp ----
      if (m == 0) TCT(1, nt, 3);
      rmax = __reduce_max_sync(0xffffffffu, rmax);
      rmin = __reduce_min_sync(0xffffffffu, rmin);
      mbar_wait(&bar_st_empty[nt & 1u], ((nt >> 1) & 1u) ^ 1u, err_flag);
      __syncwarp();
      if (lane == 0) {
        s_red[nt & 1u][0][ew] = fmaf(lg2_approx(__uint_as_float(rmax)), 0.25f * kLog10_2, tile_k);
        s_red[nt & 1u][1][ew] = fmaf(lg2_approx(__uint_as_float(rmin)), 0.25f * kLog10_2, tile_k);
        __threadfence_block();
        mbar_arrive(&bar_st_full[nt & 1u]);  // release: the tile's global stores (ordered by __syncwarp) and s_red
      }
      ++nt;
    }
  } else if (warp == 8) {
    // =========================================== MMA ISSUER ===========================================
    if (lane == 0) {
      // descriptors differ only in the start address: add (byte offset >> 4) to the low word (addresses < 256 KB)
      const uint64_t a_desc0 = smem_desc(smem_u32(a_ring), kChunkBytes, 128);
      const uint64_t b_desc0 = smem_desc(smem_u32(b_sm), kBChunkBytes, 128);
      uint32_t ks = 0, nt = 0;
      for (uint32_t id = blockIdx.x; id < p.total_tiles; id += gridDim.x) {
        const Tile t = tile_info(p, id);
        if (t.mode == kModeSilent) continue;
        TCT(2, nt, 0);
        mbar_wait(&bar_d_empty, (nt & 1u) ^ 1u, err_flag);  // epilogue has drained the previous tile's accumulators
        tc_fence_after();
        TCT(2, nt, 1);
#pragma unroll 1
        for (int j = 0; j < kKSteps; ++j, ++ks) {
          const uint32_t stage = ks & 1u;
          mbar_wait(&bar_a_full[stage], (ks >> 1) & 1u, err_flag);
          tc_fence_after();
          const uint64_t sa = a_desc0 + (uint64_t)((stage * kStageBytes) >> 4);
          const uint64_t bh = b_desc0 + (uint64_t)(((uint32_t)(2 * j) * kBChunkBytes) >> 4);
          const uint64_t bl = bh + (uint64_t)((kBBytes / 2) >> 4);
          const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
          for (int n1 = 0; n1 < 4; ++n1) {
            const uint64_t ah = sa + (uint64_t)((((n1 * 2 + 0) * 2) * kChunkBytes) >> 4);
            const uint64_t al = sa + (uint64_t)((((n1 * 2 + 1) * 2) * kChunkBytes) >> 4);
            const uint32_t d = tmem + (uint32_t)(n1 * kN);
            mma_f16(d, ah, bh, kIdesc, acc);
            mma_f16(d, ah, bl, kIdesc, 1u);
            mma_f16(d, al, bh, kIdesc, 1u);
          }
          mma_commit(&bar_a_empty[stage]);  // implies tcgen05.fence::before_thread_sync
          TCT(2, nt, 2 + j);
        }
        mma_commit(&bar_d_full);
        ++nt;
      }
    }
  } else if (warp == 9) {
    // =========================================== LOADER ===========================================
    uint32_t nt = 0;
    for (uint32_t id = blockIdx.x; id < p.total_tiles; id += gridDim.x) {
      const Tile t = tile_info(p, id);
      if (t.mode == kModeSilent) continue;
      if (lane == 0) TCT(3, nt, 0);
      mbar_wait(&bar_raw_empty, (nt & 1u) ^ 1u, err_flag);  // prep has finished reading the previous raw tile
      if (lane == 0) TCT(3, nt, 1);
      if (t.mode == kModeAsync) {
        const float* src = reinterpret_cast<const float*>(p.pcm) + t.off + (t.tile * kTileM * kHop - kNFft / 2);
        if (lane == 0) mbar_arrive_expect_tx(&bar_raw_full, kRawLen * 4);
        __syncwarp();
        for (int r = lane; r < kRawRows; r += 32) {
          const uint32_t bytes = (r < kRawRows - 1) ? kHop * 4 : (kRawLen - (kRawRows - 1) * kHop) * 4;
          bulk_g2s(raw + r * kRawPitch, src + r * kHop, bytes, &bar_raw_full);
        }
      } else if (lane == 0) {
        mbar_arrive(&bar_raw_full);  // generic staging: the prep warps fill the buffer themselves
      }
      if (lane == 0) TCT(3, nt, 2);
      ++nt;
    }
  } else {
    // =========================================== CLAMP BOOKS (warp 10) ===========================================
    OutT* const out = reinterpret_cast<OutT*>(p.out);
    int ring_head = 0, ring_count = 0;
    uint32_t nt = 0;
    for (uint32_t id = blockIdx.x; id < p.total_tiles; id += gridDim.x) {
      const Tile t = tile_info(p, id);
      float mx, mn;
      int silent = 0;
      if (t.mode == kModeSilent) {
        silent = 1;
        mx = -1.5f;
        mn = -__int_as_float(0x7f800000);
        if (p.mask != nullptr) {
          const int t0 = t.tile * kTileM;
          for (int f = lane; f < kTileM && t0 + f < kNFrames; f += 32) p.mask[(size_t)t.b * kNFrames + t0 + f] = 0;
        }
      } else {
        mbar_wait(&bar_st_full[nt & 1u], (nt >> 1) & 1u, err_flag);
        // an epilogue warp whose 32 frames lie beyond frame 3000 reports the identities (lg2(0) = -inf, lg2(inf) = +inf)
        mx = fmaxf(fmaxf(s_red[nt & 1u][0][0], s_red[nt & 1u][0][1]), fmaxf(s_red[nt & 1u][0][2], s_red[nt & 1u][0][3]));
        mn = fminf(fminf(s_red[nt & 1u][1][0], s_red[nt & 1u][1][1]), fminf(s_red[nt & 1u][1][2], s_red[nt & 1u][1][3]));
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_st_empty[nt & 1u]);
        ++nt;
      }
      // ring full (a clip whose other tiles lag far behind): resolve the oldest entry by waiting for its clip
      if (ring_count == kRing) {
        const int2 bt = s_pend_bt[ring_head];
        const float2 pm = s_pend_mm[ring_head];
        ring_head = (ring_head + 1) & (kRing - 1);
        --ring_count;
        const float fl = wait_clip_floor_tc(p.tile_key, bt.x, lane);
        if (pm.x < fl) fix_tile_tc<OutT>(out, kNMel, bt.x, bt.y & ~kSilentBit, fl, (bt.y & kSilentBit) != 0 || pm.y <= fl, lane);
      }
      // publish this tile's maximum, remember the tile
      if (lane == 0) {
        st_relaxed_u32(p.tile_key + (size_t)t.b * kNTiles + t.tile, f2key(mx));
        const int slot = (ring_head + ring_count) & (kRing - 1);
        s_pend_bt[slot] = make_int2(t.b, t.tile | (silent ? kSilentBit : 0));
        s_pend_mm[slot] = make_float2(mn, mx);
      }
      __syncwarp();
      ++ring_count;
      // retire every pending tile whose clip is complete, oldest first (bounded: at most 3 per visit)
      for (int tries = 0; tries < 3 && ring_count > 0; ++tries) {
        const int2 bt = s_pend_bt[ring_head];
        uint32_t k = lane < kNTiles ? ld_relaxed_u32(p.tile_key + (size_t)bt.x * kNTiles + lane) : 1u;
        const bool zero = __any_sync(0xffffffffu, k == 0);
        if (zero) break;
        k = __reduce_max_sync(0xffffffffu, k);
        const float fl = fmaxf(key2f(k) - 2.0f, -1.5f);
        const float2 pm = s_pend_mm[ring_head];
        ring_head = (ring_head + 1) & (kRing - 1);
        --ring_count;
        if (pm.x < fl) fix_tile_tc<OutT>(out, kNMel, bt.x, bt.y & ~kSilentBit, fl, (bt.y & kSilentBit) != 0 || pm.y <= fl, lane);
      }
    }
    // drain: every remaining tile of these clips belongs to a running CTA whose clamp warp publishes without waiting
    while (ring_count > 0) {
      const int2 bt = s_pend_bt[ring_head];
      const float2 pm = s_pend_mm[ring_head];
      ring_head = (ring_head + 1) & (kRing - 1);
      --ring_count;
      const float fl = wait_clip_floor_tc(p.tile_key, bt.x, lane);
      if (pm.x < fl) fix_tile_tc<OutT>(out, kNMel, bt.x, bt.y & ~kSilentBit, fl, (bt.y & kSilentBit) != 0 || pm.y <= fl, lane);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace tc
}  // namespace wfe
