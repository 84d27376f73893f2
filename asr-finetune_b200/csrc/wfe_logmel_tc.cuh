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
//   Measured (tools/ubench_tcgen05.cu, profiles/r02_ubench_tcgen05.txt): 2^-19.6 of the row maximum.
//
// Data movement.  A first version kept the A operand in a shared-memory ring: the role trace (tools/tc_trace.py) showed
// the kernel bound by SHARED-MEMORY BANDWIDTH -- an SS-mode 128x112x16 MMA fetches 7.5 KB of operands (60.8 cycles at
// 128 B/clk, the whole pipe) and the prep warps' own LDS/STS came on top: 1400 cycles per k-step.  So:
//   * A lives in TENSOR MEMORY: the prep threads (thread = frame = TMEM lane) write their fp16 hi/lo rows with tcgen05.st
//     into the 64 columns the accumulators leave free (4 slots, one per n1, one k-step each); the MMA reads A from TMEM
//     (57 cycles per MMA) and only B (3.5 KB) from shared memory;
//   * the raw tile arrives by ONE 2-D TMA (cp.async.bulk.tensor, box 164 x 130 floats over a 640-byte-pitch view of the
//     clip: the 4 extra floats per row pad each hop row to 164 words, so LDS.128 by frame is bank-conflict-free); the
//     first version issued 130 row copies and spent 9.4 k cycles just issuing them;
//   * raw tiles are double-buffered, so the load of tile t+1 overlaps everything of tile t.
//
//   * TMEM is full (4 x 112 accumulator columns + 64 operand columns of 512), so the accumulators cannot be double-
//     buffered: a tile's MMA phase and its epilogue phase alternate.  Both phases are latency-bound per warp, so the
//     eight WORKER warps (two per SM sub-partition and TMEM lane quarter) take part in BOTH: in the MMA phase the two
//     warps of a quarter prepare alternate k-steps; in the epilogue each takes half of the k2 range, and the few mel
//     filters fed by both halves are completed through a 2.5 KB shared-memory exchange.
//
// One persistent CTA per SM, 11 warps, tile = 128 consecutive frames of one clip (24 tiles per 30-s clip), tile ids
// strided statically over the CTAs (every role derives the same sequence):
//   warps 0-7   WORKERS   thread = frame = TMEM lane (quarter = warp % 4, half = warp / 4)
//               prep:     raw samples (smem) -> scale, window (FMUL immediates), split hi/lo -> TMEM (tcgen05.st)
//               epilogue: tcgen05.ld, twiddle + DFT-4 + power (packed f32x2), banded mel with immediate weights
//                         (generated straight-line code), log10, (x+4)/4, store, tile min/max
//   warp  8     MMA       one elected lane issues 3 tcgen05.mma per (k-step, n1) slot, commits to mbarriers
//   warp  9     LOADER    one TMA per tile
//   warp 10     CLAMP     per-clip max-8 clamp books (same scheme as wfe_logmel.cuh: one published key per tile, fix-ups
//                         of this CTA's own tiles from L2 once their clip is complete)
// Specialised for n_samples = 480000 (3000 frames), n_mel in {80, 128} with the slaney structure baked by
// tools/gen_tc_epilogue.py; everything else runs on the CUDA-core kernel.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
constexpr int kBChunkBytes = kN * 16;             // 1792
constexpr int kBBytes = 2 * 14 * kBChunkBytes;    // [hi/lo][chunk 14][n 112][16 B] = 50176
constexpr int kTwBytes = 26 * 3 * 16;             // [pair][n1-1] (cos_k2, cos_k2+1, sin_k2, sin_k2+1)
constexpr int kThreads = 11 * 32;
constexpr int kTmemCols = 512;
constexpr int kTmemA = 4 * kN;                    // columns 448..511: A slots, 16 columns per n1 (8 hi + 8 lo)
constexpr int kRing = 16;   // pending-tile ring of the clamp warp (a clip completes within about one tile iteration)
constexpr int kRawBoxBytes = kRawRows * kRawPitch * 4;              // 85280: what one TMA delivers
constexpr int kRawBufBytes = (kRawBoxBytes + 127) & ~127;           // 85376: TMA destinations are 128-byte aligned

constexpr size_t kSmemRaw = 0;
constexpr size_t kSmemB = kSmemRaw + 2 * (size_t)kRawBufBytes;       // 170752
constexpr size_t kSmemTw = kSmemB + kBBytes;                         // +50176
constexpr size_t kSmemWs = kSmemTw + kTwBytes;                       // scaled window, 7 x 64 entries (zero beyond 400)
constexpr int kWsFloats = kKSteps * 64;                              // 448
constexpr size_t kSmemBytes = kSmemWs + 2 * kWsFloats * 4;           // 225760 (+ ~4.2 KB static: the 227 KB of an SM)

#define WFE_TC_GEN_WINDOW 1
#include "wfe_tc_epilogue_gen.inc"
#undef WFE_TC_GEN_WINDOW
#define WFE_TC_GEN_COUNTS 1
#define WFE_TC_GEN_NMEL 80
#include "wfe_tc_epilogue_gen.inc"
#undef WFE_TC_GEN_NMEL
#define WFE_TC_GEN_NMEL 128
#include "wfe_tc_epilogue_gen.inc"
#undef WFE_TC_GEN_NMEL
#undef WFE_TC_GEN_COUNTS
constexpr int kMaxShared = kTcShared80 > kTcShared128 ? kTcShared80 : kTcShared128;  // filters fed by both epilogue halves

#ifdef WFE_TC_TRACE
// timing-trace build (diagnostics only): CTA 0 stamps clock64() at role milestones of its tile iterations 4..11
constexpr int kTrTiles = 8, kTrRoles = 5, kTrPts = 32;
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
  float pcm_scale;
  int pcm_dtype;            // 0 = float32, 1 = int16, 2 = float16 (wfe_pcm_dtype); 3 = float32 the TMA cannot address
  int n_mel;
  uint32_t total_tiles;
};

// the generic staging path (tile edges, 2-byte PCM, normalisation) is the only place the PCM element type matters
__device__ __forceinline__ float load_pcm(const void* pcm, int dtype, int64_t i, float scale) {
  if (dtype == 0 || dtype == 3) return reinterpret_cast<const float*>(pcm)[i];
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
#ifndef WFE_TC_DEPOSIT
#define WFE_TC_DEPOSIT 0  // operand slots handed over one by one (0) or in two groups of two (1)
#endif
#ifndef WFE_TC_WAIT
#define WFE_TC_WAIT 1  // 0: bare try_wait spin, 1: try_wait with a suspend-time hint, 2: nanosleep back-off between polls
#endif
// Waits for the phase with the given parity.  The fast path is one try_wait; the slow path is out of line (the kernel's
// straight-line code is instruction-cache bound: ~40 inlined copies of the loop cost 25 KB) and gives up after a few
// seconds, so that a protocol bug turns into a wrong answer + error flag instead of a hung GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity, uint32_t* err_flag) {
  const long long t0 = clock64();
#pragma unroll 1
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
#if WFE_TC_WAIT == 2
    __nanosleep(40);
#endif
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
#if WFE_TC_WAIT == 1
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 4000;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u && clock64() - t0 > 6000000000ll) break;
  }
  if (err_flag != nullptr) atomicExch(err_flag, 0xDEADu);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t* err_flag) {
  if (mbar_try(bar, parity)) return;
  mbar_wait_slow(smem_u32(bar), parity, err_flag);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// one 2-D tiled TMA: box (164 floats x 130 rows) at element coordinates (c0, c1) of the tensor map -> shared memory
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int32_t c0, int32_t c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]: A is 128 lanes x 8 columns (16 fp16, two per 32-bit column)
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// true in exactly one (elected) lane of the converged warp: tcgen05.mma / commit / TMA are issued from inside
// `if (elect_one())` so that the compiler sees a single active thread and emits no per-lane serialisation loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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
constexpr int kModeAsyncHead = 3, kModeDone = -1;
struct Tile {
  int b, tile, len, mode;  // mode: kModeSilent / kModeAsync (TMA) / kModeAsyncHead (TMA + reflect patch) / kModeSync (generic)
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
    // TMA: the whole 164 x 130 box (row pitch 160 floats) must lie inside the clip, start on a 16-byte boundary and
    // be addressable with an int32 element coordinate
    const float* src = reinterpret_cast<const float*>(p.pcm) + t.off + s_begin;
    // (the first tile of a clip starts 200 samples early: those land as whatever precedes the clip -- zeros if nothing
    //  does, the TMA fills out-of-range coordinates with zeros -- and are overwritten by the reflect pad afterwards)
    //  -- provided the clip does not start the buffer: the TMA bounds-checks the COORDINATE, not the address, and would
    //  zero-fill the head of every row whose column coordinate is negative)
    const bool bulk = p.pcm_dtype == 0 && p.norm == nullptr && (s_begin >= 0 || t.tile == 0) &&
                      s_begin + (kRawRows - 1) * kHop + kRawPitch <= t.len &&
                      (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && t.off + s_begin < (int64_t)0x7fff0000 &&
                      t.off + s_begin >= 0;
    t.mode = bulk ? (s_begin >= 0 ? kModeAsync : kModeAsyncHead) : kModeSync;
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
  constexpr int kDeep = 16;  // rows in flight: one warp has to keep up with a fix-up per tile on speech-like audio
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

// what the loader warp hands to the workers with each raw tile
struct alignas(16) TileMeta {
  int32_t b, tile, len, mode;  // mode kModeDone: no more tiles
  int64_t off;
  float scale, tile_k;         // power-of-two scale of the tile; y = lg2(mel_scaled) * C + tile_k
};
// scale = 2^(14 - e), e = unbiased exponent of the tile maximum (clamped so that the scale stays a normal float);
// y = (log10(mel_scaled * 2^(-2 (14 - e))) + 4) / 4 = lg2(mel_scaled) * C + (1 - 2 (14 - e) C)
__device__ __forceinline__ void scale_from_max(uint32_t mx_bits, float& scale, float& tile_k) {
  int e = (int)((mx_bits >> 23) & 0xffu) - 127;
  if (mx_bits == 0u) e = 14;
  e = max(-100, min(e, 100));
  scale = __uint_as_float((uint32_t)(127 + 14 - e) << 23);
  tile_k = 1.0f - (float)(2 * (14 - e)) * (0.25f * kLog10_2);
}
// max |x| over a staged raw tile, as float bits, by `nthreads` cooperating threads (thread index t): thread t scans hop
// rows t, t + nthreads, ... (row pitch 164 words: conflict-free LDS.128 across lanes), ten loads in flight
__device__ __forceinline__ uint32_t raw_absmax(const float* raw, int t, int nthreads) {
  float mx = 0.f;
  for (int r = t; r < kRawRows; r += nthreads) {
    const float4* row = reinterpret_cast<const float4*>(raw + r * kRawPitch);
    const int n4 = r == kRawRows - 1 ? (kRawLen - (kRawRows - 1) * kHop) / 4 : kHop / 4;  // the last row is half a row
#pragma unroll
    for (int c0 = 0; c0 < kHop / 4; c0 += 10) {
      float4 v[10];
#pragma unroll
      for (int u = 0; u < 10; ++u) v[u] = row[c0 + u];
#pragma unroll
      for (int u = 0; u < 10; ++u)
        if (c0 + u < n4) mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
    }
  }
  return __float_as_uint(mx);
}
// the same for float4 elements i = t + nthreads * u, u in [u0, u1): slices that the workers fit between their k-steps
__device__ __forceinline__ float raw_absmax_slice(const float* raw, int t, int nthreads, int u0, int u1) {
  float mx = 0.f;
  constexpr int kQuads = ((kRawRows - 1) * kHop + (kRawLen - (kRawRows - 1) * kHop)) / 4;  // 5180 float4 of real samples
#pragma unroll
  for (int u = u0; u < u1; ++u) {
    const int i = t + nthreads * u;
    if (i < kQuads) {
      const int r = i / (kHop / 4);
      const float4 v = *reinterpret_cast<const float4*>(raw + r * kRawPitch + 4 * (i - r * (kHop / 4)));
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  }
  return mx;
}
// scaled window table (exact: the scale is a power of two), zero beyond the 400-point frame
__device__ __forceinline__ void write_ws(float* ws, float scale, int t, int nthreads) {
  for (int i = t; i < kWsFloats / 4; i += nthreads) {
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (4 * i < kNFft) w = __ldg(reinterpret_cast<const uint4*>(kWinBits) + i);  // (no shared memory left for a copy)
    reinterpret_cast<float4*>(ws)[i] = make_float4(__uint_as_float(w.x) * scale, __uint_as_float(w.y) * scale,
                                                   __uint_as_float(w.z) * scale, __uint_as_float(w.w) * scale);
  }
}

// named barrier among the 256 worker threads
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <typename OutT, int kNMel>
__global__ void __launch_bounds__(kThreads, 1)
    logmel_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap, uint32_t* err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const b_sm = smem + kSmemB;
  const float4* const tw_sm = reinterpret_cast<const float4*>(smem + kSmemTw);
  float* const ws_sm = reinterpret_cast<float*>(smem + kSmemWs);  // [2 raw buffers][448]

  __shared__ uint64_t bar_raw_full[2], bar_raw_empty[2], bar_meta_full[2], bar_a_full[4], bar_a_empty[2][4], bar_d_full,
      bar_d_empty, bar_st_full[2], bar_st_empty[2];
  __shared__ TileMeta s_meta[2];        // loader -> workers: geometry, scale and log-domain constant of the tile in buffer rb
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_pmax[8];        // generic staging: per-warp max |x| bits
  __shared__ uint32_t s_tilemax[2];     // per raw buffer: max |x| bits of the tile (atomicMax by the worker warps)
  __shared__ float s_red[2][2][8];      // [tile parity][max, min][worker warp] of y over the warp's share of the tile
  __shared__ float s_part[(kNMel == 128 ? kTcShared128 : kTcShared80) * kTileM];  // epilogue half 1 -> half 0: partial sums of the shared mel filters
  __shared__ int2 s_pend_bt[kRing];
  __shared__ float2 s_pend_mm[kRing];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: role branches do not diverge
#ifdef WFE_TC_TRACE
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) {  // whole-kernel span of two CTAs: SM clock and wall clock
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_tc_trace[(4 + (blockIdx.x ? 5 : 0)) * kTrPts + 0] = clock64();
    g_tc_trace[(4 + (blockIdx.x ? 5 : 0)) * kTrPts + 1] = gt;
  }
#endif

  // ---- one-time set-up ----
  for (int i = tid; i < kBBytes / 16; i += kThreads) reinterpret_cast<uint4*>(b_sm)[i] = p.b_mat[i];
  for (int i = tid; i < kTwBytes / 16; i += kThreads) reinterpret_cast<float4*>(smem + kSmemTw)[i] = p.tw[i];
  fence_async_smem();  // B is read by the tensor core (async proxy)
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_raw_full[s], 1);
      mbar_init(&bar_raw_empty[s], 256);
      mbar_init(&bar_meta_full[s], 1);
      mbar_init(&bar_st_full[s], 8);
      mbar_init(&bar_st_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bar_a_full[s], 128);
      mbar_init(&bar_a_empty[0][s], 1);
      mbar_init(&bar_a_empty[1][s], 1);
    }
    mbar_init(&bar_d_full, 1);
    mbar_init(&bar_d_empty, 256);
    s_tilemax[0] = s_tilemax[1] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(&s_tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp < 8) {
    // =========================================== WORKERS ===========================================
    const int qt = warp & 3;                 // TMEM lane quarter: lanes 32 qt .. 32 qt + 31
    const int hh = warp >> 2;                // which of the quarter's two warps
    const int m = qt * 32 + lane;            // frame within the tile == TMEM lane
    const int wt = tid;                      // 0..255 among the workers
    const uint32_t tlane = tmem + ((uint32_t)(qt * 32) << 16);
    const uint32_t a_slot0 = tlane + kTmemA;

    // receive the tile in raw buffer (n & 1): geometry + scale from the loader; tiles that need the generic staging
    // path (clip edges, 2-byte PCM, normalisation) are staged, scanned and scaled here, by all workers
    auto fetch = [&](uint32_t n, TileMeta& tm) {
      const uint32_t rb = n & 1u;
      mbar_wait(&bar_meta_full[rb], (n >> 1) & 1u, err_flag);
      tm = s_meta[rb];
      if (tm.mode != kModeSync) return;
      float* const raw = reinterpret_cast<float*>(smem + kSmemRaw + rb * kRawBufBytes);
      const int t0 = tm.tile * kTileM;
      const int s_begin = t0 * kHop - kNFft / 2;
      const int nvalid = min(kTileM, kNFrames - t0);
      const int n_quads = ((nvalid - 1) * kHop + kNFft + 3) / 4;  // only the rows the tile's valid frames read
      float mean = 0.f, rstd = 1.f;
      if (p.norm != nullptr) {
        const float2 st = __ldg(p.norm + tm.b);
        mean = st.x;
        rstd = st.y;
      }
      const bool vec_ok = (p.pcm_dtype == 0 || p.pcm_dtype == 3) &&
                          ((reinterpret_cast<uintptr_t>(reinterpret_cast<const float*>(p.pcm) + tm.off + s_begin) & 15u) == 0);
      constexpr int kBatch = 8;  // quads of samples in flight per thread
      for (int g0 = wt; g0 < n_quads; g0 += 256 * kBatch) {
        float4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int s = s_begin + 4 * (g0 + 256 * u);
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g0 + 256 * u < n_quads) {
            if (vec_ok && s >= 0 && s + 3 < tm.len) {  // interior (len <= n_samples: no reflection either)
              v[u] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.pcm) + tm.off + s));
            } else {
              float e[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                int sk = s + k;
                if (sk < 0) sk = -sk;
                if (sk >= kNSamples) sk = 2 * (kNSamples - 1) - sk;
                e[k] = (sk >= 0 && sk < tm.len) ? load_pcm(p.pcm, p.pcm_dtype, tm.off + sk, p.pcm_scale) : 0.f;
              }
              v[u] = make_float4(e[0], e[1], e[2], e[3]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int g = g0 + 256 * u;
          if (g < n_quads) {
            float4 x = v[u];
            if (p.norm != nullptr) {
              const int s = s_begin + 4 * g;
              float* e = reinterpret_cast<float*>(&x);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                int sk = s + k;
                if (sk < 0) sk = -sk;
                if (sk >= kNSamples) sk = 2 * (kNSamples - 1) - sk;
                if (sk >= 0 && sk < tm.len) e[k] = (e[k] - mean) * rstd;
              }
            }
            const int r = g / (kHop / 4);
            *reinterpret_cast<float4*>(raw + r * kRawPitch + 4 * (g - r * (kHop / 4))) = x;
          }
        }
      }
      // rows beyond the valid frames feed the idle lanes' arithmetic: make them finite
      for (int g = n_quads + wt; g < kRawRows * (kHop / 4); g += 256) {
        const int r = g / (kHop / 4);
        *reinterpret_cast<float4*>(raw + r * kRawPitch + 4 * (g - r * (kHop / 4))) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      worker_bar();
      uint32_t mx = __reduce_max_sync(0xffffffffu, raw_absmax(raw, wt, 256));
      if (lane == 0) s_pmax[warp] = mx;
      worker_bar();
      mx = max(max(max(s_pmax[0], s_pmax[1]), max(s_pmax[2], s_pmax[3])), max(max(s_pmax[4], s_pmax[5]), max(s_pmax[6], s_pmax[7])));
      scale_from_max(mx, tm.scale, tm.tile_k);
      write_ws(ws_sm + rb * kWsFloats, tm.scale, wt, 256);
      worker_bar();  // ws complete; s_pmax free for the next staged tile
    };

    // one k-step (16 n2 = 64 consecutive samples, all four n1) of the tile in raw buffer rb: window, split into fp16
    // hi / lo, and hand the four operand slots to the tensor core as they become free.  kj = running k-step index.
    auto prep_kstep = [&](uint32_t rb, int j, uint32_t kj) {
      const float* const xrow = reinterpret_cast<const float*>(smem + kSmemRaw + rb * kRawBufBytes) + m * kRawPitch;
      const float* const ws = ws_sm + rb * kWsFloats;
      uint32_t hv[4][8], lv[4][8];  // per n1: 16 fp16 hi (K order), 16 fp16 lo
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        // the 32 samples of a half k-step never straddle a hop row (160 = 5 x 32)
        const int n0 = 64 * j + 32 * c;
        const float* xp = xrow + n0 + (kRawPitch - kHop) * ((n0 >= kHop ? 1 : 0) + (n0 >= 2 * kHop ? 1 : 0));
        const float* wp = ws + n0;
        float h[4][8], l[4][8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x = *reinterpret_cast<const float4*>(xp + 4 * q);
          const float4 w = *reinterpret_cast<const float4*>(wp + 4 * q);
          const float y[4] = {x.x * w.x, x.y * w.y, x.z * w.z, x.w * w.w};
#pragma unroll
          for (int n1 = 0; n1 < 4; ++n1) {
            const float hb = __uint_as_float(__float_as_uint(y[n1]) & 0xFFFFE000u);  // 11 significant bits: exact in fp16
            h[n1][q] = hb;
            l[n1][q] = y[n1] - hb;
          }
        }
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            hv[n1][4 * c + u] = pack_h2(h[n1][2 * u], h[n1][2 * u + 1]);
            lv[n1][4 * c + u] = pack_h2(l[n1][2 * u], l[n1][2 * u + 1]);
          }
      }
      // The MMAs of k-step kj - 1 have read a slot when their commit arrives.  Commits go to the barrier set of THAT
      // k-step's parity: a parity wait cannot tell phase i from phase i + 2, and with the quarter's two warps
      // alternating k-steps a single set would let a warp run two phases ahead.  This way each warp consumes every
      // phase of "its" set.
      const uint32_t par = hh ? ((kj >> 1) & 1u) : (((kj >> 1) & 1u) ^ 1u);
#if WFE_TC_DEPOSIT == 1
      // slots in two groups of two: one tcgen05.wait::st per group instead of per slot
#pragma unroll
      for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int n1 = 2 * g; n1 < 2 * g + 2; ++n1) {
          mbar_wait(&bar_a_empty[hh ^ 1][n1], par, err_flag);
          tc_fence_after();
          const uint32_t r[16] = {hv[n1][0], hv[n1][1], hv[n1][2], hv[n1][3], hv[n1][4], hv[n1][5], hv[n1][6], hv[n1][7],
                                  lv[n1][0], lv[n1][1], lv[n1][2], lv[n1][3], lv[n1][4], lv[n1][5], lv[n1][6], lv[n1][7]};
          tmem_st16(a_slot0 + 16 * n1, r);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bar_a_full[2 * g]);
        mbar_arrive(&bar_a_full[2 * g + 1]);
      }
#else
#pragma unroll
      for (int n1 = 0; n1 < 4; ++n1) {
        mbar_wait(&bar_a_empty[hh ^ 1][n1], par, err_flag);
        tc_fence_after();
        const uint32_t r[16] = {hv[n1][0], hv[n1][1], hv[n1][2], hv[n1][3], hv[n1][4], hv[n1][5], hv[n1][6], hv[n1][7],
                                lv[n1][0], lv[n1][1], lv[n1][2], lv[n1][3], lv[n1][4], lv[n1][5], lv[n1][6], lv[n1][7]};
        tmem_st16(a_slot0 + 16 * n1, r);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bar_a_full[n1]);
      }
#endif
    };

    // scale + scaled window of a TMA tile whose maximum the workers have accumulated in s_tilemax[rb] (all workers,
    // after a worker_bar that ordered the atomics); resets the accumulator for the buffer's next tile
    auto finish_scale = [&](uint32_t rb, TileMeta& tm) {
      if (tm.mode == kModeAsync || tm.mode == kModeAsyncHead) {
        scale_from_max(s_tilemax[rb], tm.scale, tm.tile_k);
        write_ws(ws_sm + rb * kWsFloats, tm.scale, wt, 256);
      }
    };
    constexpr int kScanU = (5180 + 255) / 256;  // float4 per worker thread for a whole tile: 21

    uint32_t ks = 0;   // running k-step index of the current tile's first k-step
    uint32_t nt = 0;   // running count of non-silent tiles
    TileMeta cur, nxt;
    fetch(0, cur);
    if (cur.mode == kModeAsync || cur.mode == kModeAsyncHead) {  // first tile: nobody has scanned it yet
      const float mxf = raw_absmax_slice(reinterpret_cast<const float*>(smem + kSmemRaw), wt, 256, 0, kScanU);
      const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(mxf));
      if (lane == 0) atomicMax(&s_tilemax[0], mx);
    }
    worker_bar();
    finish_scale(0, cur);
    worker_bar();
    if (wt == 0) s_tilemax[0] = 0u;  // next accumulated during tile 1's MMA phase (for tile 2), after more barriers
    while (cur.mode != kModeDone) {
      const uint32_t rb = nt & 1u;
      const int t0 = cur.tile * kTileM;
      const float tile_k = cur.tile_k;
      if (wt == 0) TCT(0, nt, 3);
      // the next tile is already in the other raw buffer (its TMA was issued one epilogue ago)
      fetch(nt + 1, nxt);
      const bool scan_next = nxt.mode == kModeAsync || nxt.mode == kModeAsyncHead;
      const float* const raw_next = reinterpret_cast<const float*>(smem + kSmemRaw + (rb ^ 1u) * kRawBufBytes);
      // ---- MMA phase: the quarter's two warps take alternate k-steps (running index parity == hh) ----
#pragma unroll 1
      for (int j = (int)((ks & 1u) ^ (uint32_t)hh); j < kKSteps; j += 2) {
        prep_kstep(rb, j, ks + (uint32_t)j);
        if (lane == 0 && qt == 0) TCT(0, nt, 4 + j);
      }
      mbar_arrive(&bar_raw_empty[rb]);  // this thread is done reading the raw tile
      // ---- maximum of the NEXT tile, in the shadow of the tensor core's last k-steps (interleaving it with the k-step
      //      loop slowed the loop down by as much as it saved) ----
      if (scan_next) {
        const float mx_next = raw_absmax_slice(raw_next, wt, 256, 0, kScanU);
        const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(mx_next));
        if (lane == 0) atomicMax(&s_tilemax[rb ^ 1u], mx);
      }
      worker_bar();  // every warp's share of the next tile's maximum is in
      finish_scale(rb ^ 1u, nxt);

      // ---- epilogue phase ----
      const bool valid = t0 + m < kNFrames;
      OutT* const obase = reinterpret_cast<OutT*>(p.out) + (size_t)cur.b * kNMel * kNFrames + t0 + m;
      if (hh == 0 && p.mask != nullptr && valid) p.mask[(size_t)cur.b * kNFrames + t0 + m] = ((t0 + m) * kHop < cur.len) ? 1 : 0;
      if (wt == 0) TCT(1, nt, 0);
      mbar_wait(&bar_d_full, nt & 1u, err_flag);
      tc_fence_after();
      if (wt == 0) TCT(1, nt, 1);
      uint32_t rmax = 0u, rmin = 0x7f800000u;
      uint32_t qb0_[4][8], qb1_[4][8];  // two register buffers of TMEM columns: [n1][2 pairs x (re, re, im, im)]
      f2 P0, P1, P2, P3;

#define TC_LOAD(buf, col)                                  \
  tmem_ld8(tlane + 0 * kN + (col), qb##buf##_[0]);         \
  tmem_ld8(tlane + 1 * kN + (col), qb##buf##_[1]);         \
  tmem_ld8(tlane + 2 * kN + (col), qb##buf##_[2]);         \
  tmem_ld8(tlane + 3 * kN + (col), qb##buf##_[3]);
#define TC_LOAD_WAIT() tmem_ld_wait();
      // columns of pair i within the group: (re k2, re k2+1, im k2, im k2+1).  T_n1 = Y_n1 * (cos - i sin):
      //   re = yr c + yi s, im = yi c - yr s; then the 4-point DFT over n1 and the four powers
#define TC_TWID(qq, i, pp, n1m1, outr, outi)                                                                \
  {                                                                                                         \
    const float4 w = tw_sm[(pp) * 3 + (n1m1)];                                                              \
    const f2 yr = mk2(__uint_as_float(qq[4 * (i)]), __uint_as_float(qq[4 * (i) + 1]));                       \
    const f2 yi = mk2(__uint_as_float(qq[4 * (i) + 2]), __uint_as_float(qq[4 * (i) + 3]));                   \
    outr = vfma(yr, mk2(w.x, w.y), vmul(yi, mk2(w.z, w.w)));                                                \
    outi = vfma(yi, mk2(w.x, w.y), -vmul(yr, mk2(w.z, w.w)));                                               \
  }
#define TC_PAIR(buf, i, pp)                                                                                \
  {                                                                                                        \
    const f2 y0r = mk2(__uint_as_float(qb##buf##_[0][4 * (i)]), __uint_as_float(qb##buf##_[0][4 * (i) + 1]));     \
    const f2 y0i = mk2(__uint_as_float(qb##buf##_[0][4 * (i) + 2]), __uint_as_float(qb##buf##_[0][4 * (i) + 3])); \
    f2 t1r, t1i, t2r, t2i, t3r, t3i;                                                                       \
    TC_TWID(qb##buf##_[1], i, pp, 0, t1r, t1i)                                                             \
    TC_TWID(qb##buf##_[2], i, pp, 1, t2r, t2i)                                                             \
    TC_TWID(qb##buf##_[3], i, pp, 2, t3r, t3i)                                                             \
    const f2 s02r = y0r + t2r, s02i = y0i + t2i, d02r = y0r - t2r, d02i = y0i - t2i;                       \
    const f2 s13r = t1r + t3r, s13i = t1i + t3i, d13r = t1r - t3r, d13i = t1i - t3i;                       \
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
  if (wt == 0) TCT(1, nt, 2);
#define TC_ACC_SET(mm, pexpr, wbits) float a_##mm = (pexpr) * __uint_as_float(wbits);
#define TC_ACC(mm, pexpr, wbits) a_##mm = fmaf((pexpr), __uint_as_float(wbits), a_##mm);
#define TC_PART_PUT(slot, mm) s_part[(slot) * kTileM + m] = a_##mm;
#define TC_PART_SIGNAL() asm volatile("bar.arrive %0, 64;" ::"r"(2 + qt) : "memory");
#define TC_PART_WAIT() asm volatile("bar.sync %0, 64;" ::"r"(2 + qt) : "memory");
#define TC_PART_GET(slot, mm) a_##mm += s_part[(slot) * kTileM + m];
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
#undef TC_LOAD_WAIT
#undef TC_TWID
#undef TC_PAIR
#undef TC_RELEASE
#undef TC_ACC_SET
#undef TC_ACC
#undef TC_PART_PUT
#undef TC_PART_SIGNAL
#undef TC_PART_WAIT
#undef TC_PART_GET
#undef TC_FIN
#undef TC_FIN_ZERO
      // ---- tile extrema (raw scaled mel powers, >= 0: uint order == float order) -> clamp warp ----
      if (wt == 0) TCT(1, nt, 3);
      rmax = __reduce_max_sync(0xffffffffu, rmax);
      rmin = __reduce_min_sync(0xffffffffu, rmin);
      mbar_wait(&bar_st_empty[nt & 1u], ((nt >> 1) & 1u) ^ 1u, err_flag);
      __syncwarp();
      if (lane == 0) {
        s_red[nt & 1u][0][warp] = fmaf(lg2_approx(__uint_as_float(rmax)), 0.25f * kLog10_2, tile_k);
        s_red[nt & 1u][1][warp] = fmaf(lg2_approx(__uint_as_float(rmin)), 0.25f * kLog10_2, tile_k);
        __threadfence_block();
        mbar_arrive(&bar_st_full[nt & 1u]);  // release: the tile's global stores (ordered by __syncwarp) and s_red
      }
      worker_bar();  // the next tile's window table is complete (and everybody has read its maximum)
      if (wt == 0) s_tilemax[rb ^ 1u] = 0u;  // next written two tiles from now, after another worker_bar
      cur = nxt;
      ks += kKSteps;
      ++nt;
    }
  } else if (warp == 8) {
    // =========================================== MMA ISSUER ===========================================
    // The whole warp walks the loop (waits included); one elected lane issues the tensor-core instructions.
    // B descriptors differ only in the start address: add (byte offset >> 4) to the low word (addresses < 256 KB)
    const uint64_t b_desc0 = smem_desc(smem_u32(b_sm), kBChunkBytes, 128);
    uint32_t ks = 0, nt = 0;
    for (uint32_t id = blockIdx.x; id < p.total_tiles; id += gridDim.x) {
      const Tile t = tile_info(p, id);
      if (t.mode == kModeSilent) continue;
      if (lane == 0) TCT(2, nt, 0);
      mbar_wait(&bar_d_empty, (nt & 1u) ^ 1u, err_flag);  // epilogue has drained the previous tile's accumulators
      tc_fence_after();
      if (lane == 0) TCT(2, nt, 1);
#pragma unroll 1
      for (int j = 0; j < kKSteps; ++j, ++ks) {
        const uint64_t bh = b_desc0 + (uint64_t)(((uint32_t)(2 * j) * kBChunkBytes) >> 4);
        const uint64_t bl = bh + (uint64_t)((kBBytes / 2) >> 4);
        const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1) {
          mbar_wait(&bar_a_full[n1], ks & 1u, err_flag);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d = tmem + (uint32_t)(n1 * kN);
            const uint32_t ah = tmem + (uint32_t)(kTmemA + 16 * n1), al = ah + 8;
            mma_f16_ts(d, ah, bh, kIdesc, acc);
            mma_f16_ts(d, ah, bl, kIdesc, 1u);
            mma_f16_ts(d, al, bh, kIdesc, 1u);
            mma_commit(&bar_a_empty[ks & 1u][n1]);  // implies tcgen05.fence::before_thread_sync
          }
          __syncwarp();
        }
        if (lane == 0) TCT(2, nt, 2 + j);
      }
      if (elect_one()) mma_commit(&bar_d_full);
      __syncwarp();
      ++nt;
    }
  } else if (warp == 9) {
    // =========================================== LOADER ===========================================
    // One TMA per tile into the free raw buffer, the reflect-pad patch of a clip's first tile, and the tile's geometry:
    // handed to the workers through bar_meta_full.  (Scanning the tile for its maximum here too was tried: one warp
    // needs 9 k cycles for it and slows the two workers of its sub-partition -- the workers do it in their idle slots.)
    uint32_t nt = 0;
    for (uint32_t id = blockIdx.x;; id += gridDim.x) {
      Tile t;
      t.mode = kModeDone;
      t.b = t.tile = t.len = 0;
      t.off = 0;
      if (id < p.total_tiles) {
        t = tile_info(p, id);
        if (t.mode == kModeSilent) continue;
      }
      const uint32_t rb = nt & 1u;
      float* const raw = reinterpret_cast<float*>(smem + kSmemRaw + rb * kRawBufBytes);
      if (lane == 0) TCT(3, nt, 0);
      mbar_wait(&bar_raw_empty[rb], ((nt >> 1) & 1u) ^ 1u, err_flag);  // the workers have finished with this buffer's previous tile
      if (lane == 0) TCT(3, nt, 1);
      float scale = 1.f, tile_k = 1.f;
      if (t.mode == kModeAsync || t.mode == kModeAsyncHead) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_raw_full[rb], kRawBoxBytes);
          tma_load_2d(raw, &tmap, (int32_t)(t.off + (int64_t)(t.tile * kTileM * kHop - kNFft / 2)), 0, &bar_raw_full[rb]);
        }
        __syncwarp();
        mbar_wait(&bar_raw_full[rb], (nt >> 1) & 1u, err_flag);
        if (lane == 0) TCT(3, nt, 2);
        if (t.mode == kModeAsyncHead) {
          // first tile of a clip: the TMA started 200 samples before the clip; replace them by the centred reflect pad
          for (int i = lane; i < kNFft / 2; i += 32) {
            const int s = kNFft / 2 - i;  // raw[i] = x[200 - i]
            raw[i + (i >= kHop ? kRawPitch - kHop : 0)] = s < t.len ? load_pcm(p.pcm, p.pcm_dtype, t.off + s, p.pcm_scale) : 0.f;
          }
          __syncwarp();
        }
      }
      if (lane == 0) {
        TileMeta tm;
        tm.b = t.b;
        tm.tile = t.tile;
        tm.len = t.len;
        tm.mode = t.mode;
        tm.off = t.off;
        tm.scale = scale;
        tm.tile_k = tile_k;
        s_meta[rb] = tm;
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        mbar_arrive(&bar_meta_full[rb]);
        TCT(3, nt, 3);
      }
      if (t.mode == kModeDone) break;
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
        // a worker warp whose 32 frames lie beyond frame 3000 reports the identities (lg2(0) = -inf, lg2(inf) = +inf)
        mx = -__int_as_float(0x7f800000);
        mn = __int_as_float(0x7f800000);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          mx = fmaxf(mx, s_red[nt & 1u][0][w]);
          mn = fminf(mn, s_red[nt & 1u][1][w]);
        }
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
#ifdef WFE_TC_TRACE
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_tc_trace[(4 + (blockIdx.x ? 5 : 0)) * kTrPts + 2] = clock64();
    g_tc_trace[(4 + (blockIdx.x ? 5 : 0)) * kTrPts + 3] = gt;
  }
#endif
  if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace tc
}  // namespace wfe
