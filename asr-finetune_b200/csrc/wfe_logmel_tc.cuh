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
//     eight EPILOGUE warps (two per SM sub-partition and TMEM lane quarter) take part in BOTH: in the MMA phase they
//     prepare k-steps 3..6 (the prep warps 0..2); in the epilogue each takes half of the k2 range, and the few mel
//     filters fed by both halves are completed through a 2.5 KB shared-memory exchange.
//
// One persistent CTA per SM, 16 warps, tile = 128 consecutive frames of one clip (24 tiles per 30-s clip); tile ids come
// from a global counter (the SMs do not run at the same pace), the loader hands each tile's geometry to everybody:
//   warps 0-7   EPILOGUE WORKERS   thread = frame = TMEM lane (quarter = warp % 4, half = warp / 4); 168 registers
//               MMA phase: two k-steps each (half 1: k3, k5; half 0: k4, k6): raw samples (smem) x scaled window, split
//                         hi / lo (packed f32x2) -> TMEM (tcgen05.st)
//               epilogue: tcgen05.ld, twiddle + DFT-4 + power (packed f32x2), banded mel with immediate weights
//                         (generated straight-line code), log10, (x+4)/4, store, tile maximum + per-block minima
//   warps 8-11  PREP WORKERS       k-steps 0, 1, 2 of every tile (k0, k1 of the NEXT tile in the epilogue's shadow), then
//                                  the next tile's max |x| -> power-of-two scale -> scaled window table; 104 registers
//   warp 12     MMA ISSUER         one elected lane issues 3 tcgen05.mma per (k-step, n1) slot, commits to mbarriers
//   warp 13     LOADER             next tile id, one TMA per tile, clip-edge patches from the landed tile, tile meta data
//   warp 14     BOOKS              publishes each tile's maximum and block minima for the clamp pass (clamp_kernel)
//   warp 15     idle               (every scheduler must start with four warps' worth of registers); 12-14: 72 registers
// Specialised for n_samples = 480000 (3000 frames), n_mel in {80, 128} with the slaney structure baked by
// tools/gen_tc_epilogue.py; everything else runs on the CUDA-core kernel.  The kernel is instruction-fetch sensitive:
// code that rarely runs is kept small and rolled.
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
constexpr int kThreads = 16 * 32;                // 8 epilogue workers, 4 prep workers, MMA issuer, loader, clamp books, 1 idle
// register budget (setmaxnreg).  The kernel launches with 128 registers per thread (16 warp slots x 32 x 128 = the whole
// file; the host checks the compiled count); a warp can only grow into what the warps of its OWN scheduler (warp index
// mod 4) gave back: per scheduler 2 x kRegsE + kRegsP + kRegsH <= 4 x 128.
#ifndef WFE_TC_REGS_E
#define WFE_TC_REGS_E 168
#define WFE_TC_REGS_P 104
#define WFE_TC_REGS_H 72
#endif
constexpr int kRegsLaunch = 128, kRegsE = WFE_TC_REGS_E, kRegsP = WFE_TC_REGS_P, kRegsH = WFE_TC_REGS_H;
#ifndef WFE_TC_REGS_PROBE
static_assert(2 * kRegsE + kRegsP + kRegsH <= 4 * kRegsLaunch, "register budget of an SM sub-partition");
#endif
constexpr int kTmemCols = 512;
constexpr int kTmemA = 4 * kN;                    // columns 448..511: A slots, 16 columns per n1 (8 hi + 8 lo)
constexpr int kWarpMma = 12, kWarpLoad = 13, kWarpClamp = 14;
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

#ifndef WFE_TC_TRACE_CTA
#define WFE_TC_TRACE_CTA 0
#endif
#ifdef WFE_TC_TRACE
// timing-trace build (diagnostics only): CTA 0 stamps clock64() at role milestones of its tile iterations 4..11
constexpr int kTrTiles = 8, kTrRoles = 8, kTrPts = 64;
__device__ unsigned long long g_tc_trace[kTrTiles * kTrRoles * kTrPts];
__device__ unsigned long long g_tc_tiles[3][64];  // tile start stamps of CTA 0, 77, 147 ([63] kernel start, [62] workers done, [61..59] helpers done)
__device__ unsigned long long g_tc_warps[16][8];
__device__ unsigned long long g_tc_cta[160][4];  // per CTA: smid, tiles done, first tile start, workers done
#define TCT(role, it, pt)                                                                      \
  do {                                                                                         \
    if (blockIdx.x == 0 && (it) >= 4 && (it) < 4 + kTrTiles)                                   \
      g_tc_trace[(((it) - 4) * kTrRoles + (role)) * kTrPts + (pt)] = clock64();                \
  } while (0)
#define TCW(i)                                                                     \
  do {                                                                             \
    if (blockIdx.x == WFE_TC_TRACE_CTA && nt == 6 && lane == 0) g_tc_warps[warp][i] = clock64();  \
  } while (0)
// start-up stamps of the traced CTA: row 12 = kernel, 13 = loader, 14 = epilogue warp 0, 15 = prep warp 8
#define TCS(row, i)                                                                       \
  do {                                                                                    \
    if (blockIdx.x == WFE_TC_TRACE_CTA && (threadIdx.x & 31) == 0) g_tc_warps[row][i] = clock64(); \
  } while (0)
#else
#define TCS(row, i) \
  do {              \
  } while (0)
#define TCT(role, it, pt) \
  do {                    \
  } while (0)
#define TCW(i) \
  do {         \
  } while (0)
#endif

struct TcParams {
  const void* pcm;
  const int64_t* offsets;
  const int64_t* lengths;
  const float2* norm;
  void* out;                // (B, n_mel, 3000), element type = template OutT
  int32_t* mask;
  uint32_t* tile_key;       // [B][24] key of each tile's maximum (f2key)
  uint32_t* tile_counter;   // [1] dynamic tile scheduler (zero-initialised)
  uint32_t* tile_min;       // [B][24][16] float bits of the minimum of each (32 frames x 32 mels) block of a tile, or kMinSilent
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
// Waits for the phase with the given parity.  Inline on purpose: a real call inside a role whose register count was
// changed by setmaxnreg makes ptxas give up ("register allocation failed").  The loop is a handful of instructions:
// try_wait suspends the thread for a hardware time slice (hint: 4 us) and is woken by the phase completing.  It gives
// up after ~4 M probes, so that a protocol bug turns into a wrong answer + error flag instead of a hung GPU.
#ifndef WFE_TC_SLEEP_SHORT
#define WFE_TC_SLEEP_SHORT 32   // ns between probes of a latency-critical wait (operand slots, accumulators)
#define WFE_TC_SLEEP_LONG 256   // ns between probes of a wait that is a tile long (raw buffers, clamp books)
#endif
// (One tight PTX loop per call site, ~8 SASS instructions: the kernel is instruction-fetch sensitive even to code it
//  rarely runs, and the C++ version of this loop -- probe, sleep, count, report through an atomic -- came to ~20 per site
//  at ~70 sites.)  `err_word` = shared-memory address of the CTA's time-out word, copied to global memory at the end.
template <int kSleepNs = WFE_TC_SLEEP_SHORT>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t err_word) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n"
      "WFE_W:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WFE_D;\n\t"
      "nanosleep.u32 %2;\n\t"  // a probing warp takes issue slots from the warps of its scheduler; a sleeping one none
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 4194304;\n\t"
      "@p bra WFE_W;\n\t"
      "st.shared.u32 [%3], 57005;\n"  // 0xDEAD: gave up
      "WFE_D:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "n"(kSleepNs > 0 ? kSleepNs : 1), "r"(err_word)
      : "memory");
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
constexpr int kModeAsyncHead = 3, kModeAsyncTail = 4, kModeDone = -1;
__device__ __forceinline__ bool is_tma_mode(int mode) { return mode == kModeAsync || mode == kModeAsyncHead || mode == kModeAsyncTail; }
struct Tile {
  int b, tile, len, mode;  // kModeSilent / kModeAsync (TMA) / kModeAsyncHead, kModeAsyncTail (TMA + patch by the loader) / kModeSync (generic)
  int64_t off;
};
// `extent` = one past the last PCM element of the whole batch (only the loader knows it; 0 = do not classify tail tiles)
__device__ __forceinline__ Tile tile_info(const TcParams& p, uint32_t id, int64_t extent = 0) {
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
    const uintptr_t src = reinterpret_cast<uintptr_t>(reinterpret_cast<const float*>(p.pcm) + t.off + s_begin);
    // (the first tile of a clip starts 200 samples early: those land as whatever precedes the clip -- zeros if nothing
    //  does, the TMA fills out-of-range coordinates with zeros -- and are overwritten by the reflect pad afterwards)
    //  -- provided the clip does not start the buffer: the TMA bounds-checks the COORDINATE, not the address, and would
    //  zero-fill the head of every row whose column coordinate is negative)
    // (2-byte PCM by TMA -- a box of 168 x 130 two-byte elements into the tail of the raw buffer, widened in place -- was
    //  built and measured: ONE warp widening a tile needs 20-30 k cycles, twice the tile period, and the conversions cost
    //  ~3.3 k warp-instructions per tile wherever they run; int16 / float16 PCM therefore stays on the generic path.)
    const bool base_ok = p.pcm_dtype == 0 && p.norm == nullptr && (s_begin >= 0 || t.tile == 0) && (src & 15u) == 0 &&
                         t.off + s_begin < (int64_t)0x7fff0000 && t.off + s_begin >= 0;
    const int box_end = s_begin + (kRawRows - 1) * kHop + kRawPitch;
    if (base_ok && box_end <= t.len)
      t.mode = s_begin >= 0 ? kModeAsync : kModeAsyncHead;
    else if (base_ok && s_begin >= 0 && (extent < 0 || t.off + box_end <= extent))  // (extent < 0: the caller checks)
      // the tile that straddles the end of its clip: the box is still inside the caller's buffer (it reads into whatever
      // follows the clip), and the loader overwrites everything from the clip's end on (zeros / the reflect pad)
      t.mode = kModeAsyncTail;
    else
      t.mode = kModeSync;
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

// ---------------------------------------------------------------------------------------------------------------
// Per-clip clamp max(x, max - 8) as a second, HBM-speed pass over the tiles that need it.
// The main kernel stores unclamped values and publishes, per tile, the key of its maximum (tile_key) and the bits of its
// minimum (tile_min; kMinSilent = the tile lies wholly in the zero padding and was not written at all).  A clip's floor
// is known once all its tiles are: this kernel takes it from the 24 keys, skips every tile whose minimum is already above
// it (white noise: all of them), rewrites only the elements below it elsewhere, and fills padding tiles with the constant.
// (Round 1 and the first tensor-core versions fixed tiles up inside the main kernel, from L2, by one warp per CTA: that
//  couples the CTAs -- a pending ring that fills up stalls a fast CTA behind a slow one -- and on speech-like audio the
//  one warp became the bottleneck: 15.0 M audio-s/s against 21.1 M on noise.)
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kMinSilent = 0x7fc00001u;
constexpr uint32_t kBooksDone = 0xffffffffu;  // s_red_id: no more tiles
constexpr int kClampThreads = 256;
// a tile's minima are published per BLOCK of 32 frames (one epilogue warp's quarter) x 32 mel rows: on speech-like audio
// nearly every 128-frame tile holds a few elements below its clip's floor (0.3 % of all elements, 95 % of the tiles),
// but only 30 % of the blocks do -- and those are what the pass reads back
constexpr int kMinGroups = 4;                        // mel groups of 32 rows
constexpr int kMinBlocks = 4 * kMinGroups;           // per tile: [frame quarter][mel group]
template <typename OutT>
__global__ void __launch_bounds__(kClampThreads)
    clamp_kernel(OutT* __restrict__ out, const uint32_t* __restrict__ tile_key, const uint32_t* __restrict__ tile_min,
                 int n_mel, uint32_t total_tiles) {
  using Vec = typename std::conditional<sizeof(OutT) == 4, float4, uint2>::type;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fq = lane & 7, r0 = lane >> 3;  // 4 frames per thread, 4 rows per pass, 8 passes = the block's 32 rows
  // newest tiles first: the tail of the main kernel's output is still in L2
  for (uint32_t it = blockIdx.x; it < total_tiles; it += gridDim.x) {
    const uint32_t id = total_tiles - 1u - it;
    const int b = (int)(id / (uint32_t)kNTiles), tile = (int)(id - (uint32_t)b * (uint32_t)kNTiles);
    uint32_t k = lane < kNTiles ? __ldg(tile_key + (size_t)b * kNTiles + lane) : 0u;
    k = __reduce_max_sync(0xffffffffu, k);
    const float fl = fmaxf(key2f(k) - 2.0f, -1.5f);
    OutT cv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) cv[e] = to_out<OutT>(fl);
    const Vec cvec = *reinterpret_cast<const Vec*>(cv);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int blk = warp + 8 * half, q = blk >> 2, g = blk & 3;  // frame quarter, mel group
      const uint32_t mb = __ldg(tile_min + (size_t)id * kMinBlocks + blk);
      const bool silent = mb == kMinSilent;
      if (!silent && !(__uint_as_float(mb) < fl)) continue;  // nothing below the floor in this block
      const int t0 = tile * kTileM + 32 * q;
      if (t0 + 4 * fq >= kNFrames) continue;  // (3000 = 23 * 128 + 56: the valid frames of a block are a multiple of 4)
      const int rows = min(32, n_mel - 32 * g);
      OutT* const base = out + ((size_t)b * n_mel + 32 * g + r0) * kNFrames + t0 + 4 * fq;
      if (silent) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (r0 + 4 * j < rows) *reinterpret_cast<Vec*>(base + (size_t)(4 * j) * kNFrames) = cvec;
        continue;
      }
      Vec v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + 4 * j < rows) v[j] = __ldcs(reinterpret_cast<const Vec*>(base + (size_t)(4 * j) * kNFrames));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (r0 + 4 * j >= rows) continue;
        OutT e[4];
        *reinterpret_cast<Vec*>(e) = v[j];
        bool need = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (from_out<OutT>(e[u]) < fl) {  // (-inf, the log of a zero mel power, is below every floor)
            need = true;
            e[u] = cv[u];
          }
        }
        if (need) *reinterpret_cast<Vec*>(base + (size_t)(4 * j) * kNFrames) = *reinterpret_cast<const Vec*>(e);
      }
    }
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
// max |x| over a raw tile, as float bits, by rows with 64-bit loads, nine in flight: warp `w4` of four takes hop rows w4, w4 + 4, ...; lanes read the
// row's 80 float2 in three coalesced loads.  (tools/ubench_lds.cu: a conflict-free LDS.64 moves 256 B in 2 cycles, an
// LDS.128 512 B in 8 -- per byte the 64-bit load costs the shared-memory pipe half as much.)
__device__ __forceinline__ uint32_t raw_absmax_rows(const float* raw, int w4, int lane) {
  constexpr int kPitch2 = kRawPitch / 2;  // 82 float2 per row
  const float2* const p = reinterpret_cast<const float2*>(raw) + lane;
  float mx = 0.f;
#pragma unroll 1
  for (int r = w4; r < kRawRows - 1; r += 12) {
    float2 v[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float2* q = p + min(r + 4 * k, kRawRows - 2) * kPitch2;  // (re-reading a row is harmless)
      v[3 * k] = q[0];
      v[3 * k + 1] = q[32];
      v[3 * k + 2] = lane < 16 ? q[64] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) mx = fmaxf(mx, fmaxf(fabsf(v[k].x), fabsf(v[k].y)));
  }
  if (w4 == ((kRawRows - 1) & 3)) {  // the last row is half a row: 80 samples
    const float2* q = p + (kRawRows - 1) * kPitch2;
    const float2 a = q[0], b = lane < 8 ? q[32] : make_float2(0.f, 0.f);
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y))));
  }
  return __float_as_uint(mx);
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

// One whole k-step (16 n2 = 64 consecutive samples, all four n1) of this thread's frame: window, split into fp16 hi / lo.
// xrow = the frame's first sample in the raw buffer, ws = the buffer's scaled window.  Two float4 of samples give, for
// every n1, one packed fp16x2 word of hi parts and one of lo parts (the 32 samples of a half k-step never straddle a
// hop row: 160 = 5 x 32).
__device__ __forceinline__ void prep_kstep_regs(const float* xrow, const float* ws, int j, uint32_t (&hv)[4][8],
                                                uint32_t (&lv)[4][8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int n0 = 64 * j + 32 * c;
    const float* xp = xrow + n0 + (kRawPitch - kHop) * ((n0 >= kHop ? 1 : 0) + (n0 >= 2 * kHop ? 1 : 0));
    const float* wp = ws + n0;
#pragma unroll
    for (int q2 = 0; q2 < 4; ++q2) {
      float h[4][2], l[4][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(xp + 8 * q2 + 4 * q);
        const float4 w = *reinterpret_cast<const float4*>(wp + 8 * q2 + 4 * q);
        // packed f32x2: the same roundings as four FMUL / FADD, half the issue slots
        const float2 y01 = __fmul2_rn(make_float2(x.x, x.y), make_float2(w.x, w.y));
        const float2 y23 = __fmul2_rn(make_float2(x.z, x.w), make_float2(w.z, w.w));
        const float y[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1)
          h[n1][q] = __uint_as_float(__float_as_uint(y[n1]) & 0xFFFFE000u);  // 11 significant bits: exact in fp16
        const float2 l01 = __fadd2_rn(y01, make_float2(-h[0][q], -h[1][q]));
        const float2 l23 = __fadd2_rn(y23, make_float2(-h[2][q], -h[3][q]));
        l[0][q] = l01.x;
        l[1][q] = l01.y;
        l[2][q] = l23.x;
        l[3][q] = l23.y;
      }
#pragma unroll
      for (int n1 = 0; n1 < 4; ++n1) {
        hv[n1][4 * c + q2] = pack_h2(h[n1][0], h[n1][1]);
        lv[n1][4 * c + q2] = pack_h2(l[n1][0], l[n1][1]);
      }
    }
  }
}
// Hand one k-step to the tensor core, slot by slot (slot = n1).  A slot is free once the MMAs of the PREVIOUS k-step
// have read it: their commit arrives on `empty[n1]`, a barrier that completes exactly once per tile, so `parity` (the
// tile's) is unambiguous whichever warp deposits.
__device__ __forceinline__ void deposit_kstep(uint64_t* empty, uint32_t parity, uint64_t* full, uint32_t a_slot0,
                                              const uint32_t (&hv)[4][8], const uint32_t (&lv)[4][8], uint32_t err_flag) {
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    mbar_wait(&empty[n1], parity, err_flag);
    tc_fence_after();
    const uint32_t r[16] = {hv[n1][0], hv[n1][1], hv[n1][2], hv[n1][3], hv[n1][4], hv[n1][5], hv[n1][6], hv[n1][7],
                            lv[n1][0], lv[n1][1], lv[n1][2], lv[n1][3], lv[n1][4], lv[n1][5], lv[n1][6], lv[n1][7]};
    tmem_st16(a_slot0 + 16 * n1, r);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&full[n1]);
  }
}

// named barriers: 1 the epilogue workers' own; 2..5 the epilogue's partial-sum exchange (one per lane quarter); 8 "tile
// 0's window table is ready" (epilogue workers arrive, prep workers wait); 9 the prep workers' own.  (The per-tile hand-
// shakes between the two groups are mbarriers: a named barrier would make the early epilogue warps wait for the late ones.)
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int kBarWs0 = 8, kBarPrep = 9, kBarBoth = 384;

// named barrier among the 256 worker threads
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <typename OutT, int kNMel>
__global__ void __launch_bounds__(kThreads, 1)
    logmel_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap, uint32_t* err_global) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const b_sm = smem + kSmemB;
  const float4* const tw_sm = reinterpret_cast<const float4*>(smem + kSmemTw);
  float* const ws_sm = reinterpret_cast<float*>(smem + kSmemWs);  // [2 raw buffers][448]

  __shared__ uint64_t bar_raw_full[2], bar_raw_empty[2], bar_meta_full[2], bar_a_full[4], bar_a_empty[kKSteps][4], bar_d_full,
      bar_d_empty, bar_st_full[2], bar_st_empty[2], bar_ws, bar_staged, bar_tab;
  __shared__ TileMeta s_meta[2];        // loader -> workers: geometry, scale and log-domain constant of the tile in buffer rb
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_err;            // set by a wait that gave up (mbar_wait); reported through err_global at the end
  __shared__ uint32_t s_pmax[8];        // generic staging: per-warp max |x| bits
  __shared__ uint32_t s_tilemax[2];     // per raw buffer: max |x| bits of the tile (atomicMax by the worker warps)
  __shared__ float2 s_scale[2];         // per raw buffer: (power-of-two scale, log-domain constant) of the tile
  __shared__ uint32_t s_red_id[2];      // [tile parity] id (clip * 24 + tile) of the tile whose extrema are in s_red, or kBooksDone
  __shared__ float s_red[2][1 + kMinGroups][8];  // [tile parity][max, min of mel group 0..3][worker warp] of y over the warp's share of the tile
  __shared__ float s_part[(kNMel == 128 ? kTcShared128 : kTcShared80) * kTileM];  // epilogue half 1 -> half 0: partial sums of the shared mel filters

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: role branches do not diverge
  const uint32_t err_flag = smem_u32(&s_err);
#ifdef WFE_TC_TRACE
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77 || blockIdx.x == 147))
    g_tc_tiles[blockIdx.x == 0 ? 0 : (blockIdx.x == 77 ? 1 : 2)][63] = clock64();
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) {  // whole-kernel span of two CTAs: SM clock and wall clock
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_tc_trace[(4 + (blockIdx.x ? 5 : 0)) * kTrPts + 0] = clock64();
    g_tc_trace[(4 + (blockIdx.x ? 5 : 0)) * kTrPts + 1] = gt;
  }
#endif

  // ---- one-time set-up ----
  if (tid == 0) TCS(12, 0);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_raw_full[s], 1);
      mbar_init(&bar_raw_empty[s], 384);
      mbar_init(&bar_meta_full[s], 1);
      mbar_init(&bar_st_full[s], 8);
      mbar_init(&bar_st_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bar_a_full[s], 128);
      for (int j = 0; j < kKSteps; ++j) mbar_init(&bar_a_empty[j][s], 1);
    }
    mbar_init(&bar_ws, 128);      // prep workers: the next tile's window table and scale are ready
    mbar_init(&bar_staged, 256);  // epilogue workers: an edge tile (if any) has been staged
    mbar_init(&bar_d_full, 1);
    mbar_init(&bar_d_empty, 256);
    s_tilemax[0] = s_tilemax[1] = 0u;
    s_err = 0u;
    mbar_init(&bar_tab, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the constant tables (B operand 49 KB, twiddles) arrive by two bulk copies behind everything else of the start-up;
    // their first users -- the MMA issuer, the epilogue -- wait for bar_tab.  (Copied by the threads, they cost the
    // kernel ~5 k cycles before any role could start.)
    mbar_arrive_expect_tx(&bar_tab, kBBytes + kTwBytes);
    bulk_g2s(b_sm, p.b_mat, kBBytes, &bar_tab);
    bulk_g2s(smem + kSmemTw, p.tw, kTwBytes, &bar_tab);
  }
  if (warp == kWarpLoad) tmem_alloc(&s_tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) TCS(12, 1);

  if (warp < 8) {
    // ====================================== EPILOGUE WORKERS ======================================
    reg_grow<kRegsE>();
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
      // Quads [q_lo, q_hi) lie wholly inside the clip (no reflection, no zero padding) and, when the clip starts on a
      // 16-byte (float32) / 8-byte (2-byte PCM) boundary, are one vector load each: eight per thread in flight.  The few
      // quads around the clip's edges -- and everything of a clip at an odd address -- go sample by sample in a rolled
      // loop: small code matters more here than speed (the kernel is instruction-fetch sensitive).
      const int es = (p.pcm_dtype == 0 || p.pcm_dtype == 3) ? 4 : 2;
      const uintptr_t first = reinterpret_cast<uintptr_t>(p.pcm) + (uintptr_t)((tm.off + s_begin) * es);
      const bool vec_ok = (first & (es == 4 ? 15u : 7u)) == 0;
      const int q_lo = vec_ok ? min(n_quads, s_begin < 0 ? (-s_begin + 3) / 4 : 0) : n_quads;
      const int q_hi = vec_ok ? max(q_lo, min(n_quads, (tm.len - s_begin) / 4)) : n_quads;
      auto put = [&](int g, float4 x) {
        const int r = g / (kHop / 4);
        *reinterpret_cast<float4*>(raw + r * kRawPitch + 4 * (g - r * (kHop / 4))) = x;
      };
      constexpr int kBatch = 8;
      for (int g0 = q_lo + wt; g0 < q_hi; g0 += 256 * kBatch) {
        uint4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int g = min(g0 + 256 * u, q_hi - 1);  // (re-reading the last quad is harmless)
          const uint8_t* src = reinterpret_cast<const uint8_t*>(first) + (size_t)(4 * g) * es;
          if (es == 4) {
            v[u] = __ldg(reinterpret_cast<const uint4*>(src));
          } else {
            const uint2 w = __ldg(reinterpret_cast<const uint2*>(src));
            v[u] = make_uint4(w.x, w.y, 0u, 0u);
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int g = g0 + 256 * u;
          if (g >= q_hi) continue;
          float4 x;
          if (es == 4) {
            x = make_float4(__uint_as_float(v[u].x), __uint_as_float(v[u].y), __uint_as_float(v[u].z), __uint_as_float(v[u].w));
          } else if (p.pcm_dtype == 1) {
            x = make_float4((float)(int16_t)(v[u].x & 0xffffu) * p.pcm_scale, (float)(int16_t)(v[u].x >> 16) * p.pcm_scale,
                            (float)(int16_t)(v[u].y & 0xffffu) * p.pcm_scale, (float)(int16_t)(v[u].y >> 16) * p.pcm_scale);
          } else {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v[u].x));
            const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v[u].y));
            x = make_float4(a.x, a.y, c.x, c.y);
          }
          put(g, x);
        }
      }
      // the edges: quads [0, q_lo) and [q_hi, n_quads)
#pragma unroll 1
      for (int g = wt; g < q_lo + (n_quads - q_hi); g += 256) {
        const int gq = g < q_lo ? g : g - q_lo + q_hi;
        float e[4];
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          int sk = s_begin + 4 * gq + k;
          if (sk < 0) sk = -sk;
          if (sk >= kNSamples) sk = 2 * (kNSamples - 1) - sk;
          e[k] = (sk >= 0 && sk < tm.len) ? load_pcm(p.pcm, p.pcm_dtype, tm.off + sk, p.pcm_scale) : 0.f;
        }
        put(gq, make_float4(e[0], e[1], e[2], e[3]));
      }
      if (p.norm != nullptr) {
        // zero-mean / unit-variance (do_normalize): every sample of the clip, i.e. not the zero padding.  Thread wt takes
        // the quads g = wt (mod 256), which OTHER threads wrote above whenever q_lo or q_hi is not a multiple of 256 (a
        // clip's first tile: q_lo = 50): without the barrier a fast warp normalised quads a slow one had not written yet
        // and the raw samples landed on top (found by tools/fuzz_tc_vs_cc.py: 3 of 3500 normalised batches).
        worker_bar();
#pragma unroll 1
        for (int g = wt; g < n_quads; g += 256) {
          const int r = g / (kHop / 4);
          float* e = raw + r * kRawPitch + 4 * (g - r * (kHop / 4));
#pragma unroll 1
          for (int k = 0; k < 4; ++k) {
            int sk = s_begin + 4 * g + k;
            if (sk < 0) sk = -sk;
            if (sk >= kNSamples) sk = 2 * (kNSamples - 1) - sk;
            if (sk >= 0 && sk < tm.len) e[k] = (e[k] - mean) * rstd;
          }
        }
      }
      // rows beyond the valid frames feed the idle lanes' arithmetic: make them finite
      for (int g = n_quads + wt; g < kRawRows * (kHop / 4); g += 256) {
        const int r = g / (kHop / 4);
        *reinterpret_cast<float4*>(raw + r * kRawPitch + 4 * (g - r * (kHop / 4))) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      worker_bar();
      uint32_t mx = __reduce_max_sync(0xffffffffu, raw_absmax_rows(raw, warp & 3, lane));  // (both warps of a quarter scan the same rows: rare path, one code copy)
      if (lane == 0) s_pmax[warp] = mx;
      worker_bar();
      mx = max(max(max(s_pmax[0], s_pmax[1]), max(s_pmax[2], s_pmax[3])), max(max(s_pmax[4], s_pmax[5]), max(s_pmax[6], s_pmax[7])));
      scale_from_max(mx, tm.scale, tm.tile_k);
      write_ws(ws_sm + rb * kWsFloats, tm.scale, wt, 256);
      if (wt == 0) s_scale[rb] = make_float2(tm.scale, tm.tile_k);
      worker_bar();  // ws complete; s_pmax free for the next staged tile
    };

    // scale + scaled window of a TMA tile whose maximum the workers have accumulated in s_tilemax[rb] (all workers,
    // after a worker_bar that ordered the atomics); resets the accumulator for the buffer's next tile
    auto finish_scale = [&](uint32_t rb, TileMeta& tm) {
      if (is_tma_mode(tm.mode)) {
        scale_from_max(s_tilemax[rb], tm.scale, tm.tile_k);
        write_ws(ws_sm + rb * kWsFloats, tm.scale, wt, 256);
        if (wt == 0) s_scale[rb] = make_float2(tm.scale, tm.tile_k);
      }
    };

    uint32_t nt = 0;   // running count of non-silent tiles
    TileMeta cur, nxt;
    // (Every piece of straight-line code below has ONE call site: the kernel is instruction-fetch sensitive -- the two
    //  SMs of a TPC share an instruction cache, and code copies or a rarely used path running on one SM slow down both.)
    bool have_cur = false;
    uint32_t hv[4][8], lv[4][8];
    mbar_wait<WFE_TC_SLEEP_LONG>(&bar_tab, 0u, err_flag);  // twiddles
    for (;;) {
      fetch(have_cur ? nt + 1 : 0u, nxt);
      if (warp == 0 && nt == 0) TCS(14, have_cur ? 4 : 0);
      if (!have_cur) {  // tile 0: nobody has scanned it yet; its scale and window table are ours to make
        cur = nxt;
        have_cur = true;
        if (is_tma_mode(cur.mode)) {
          const uint32_t mx = __reduce_max_sync(0xffffffffu, raw_absmax_rows(reinterpret_cast<const float*>(smem + kSmemRaw), warp & 3, lane));
          if (lane == 0) atomicMax(&s_tilemax[0], mx);
        }
        worker_bar();
        if (warp == 0) TCS(14, 1);
        finish_scale(0, cur);
        worker_bar();
        if (warp == 0) TCS(14, 2);
        if (wt == 0) s_tilemax[0] = 0u;
        bar_arrive_n(kBarWs0, kBarBoth);  // the prep workers may start on tile 0
        if (cur.mode == kModeDone) break;
        continue;
      }
      const uint32_t rb = nt & 1u;
      const int t0 = cur.tile * kTileM;
#ifdef WFE_TC_TRACE
      if (wt == 0 && nt < 58 && (blockIdx.x == 0 || blockIdx.x == 77 || blockIdx.x == 147))
        g_tc_tiles[blockIdx.x == 0 ? 0 : (blockIdx.x == 77 ? 1 : 2)][nt] = clock64();
      if (wt == 0 && blockIdx.x < 160) {
        if (nt == 1) g_tc_cta[blockIdx.x][2] = clock64();
        g_tc_cta[blockIdx.x][1] = nt;
        g_tc_cta[blockIdx.x][3] = clock64();
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_tc_cta[blockIdx.x][0] = smid;
      }
#endif
      if (wt == 0) TCT(0, nt, 3);
      TCW(0);
      // this tile's scaled window and log-domain constant come from the prep workers (tile 0: from ourselves, above)
      if (nt > 0) mbar_wait(&bar_ws, (nt - 1u) & 1u, err_flag);
      const float tile_k = s_scale[rb].y;
      // (the next tile is already in the other raw buffer: its TMA was issued one epilogue ago; edge tiles were staged above)
      const float* const xrow = reinterpret_cast<const float*>(smem + kSmemRaw + rb * kRawBufBytes) + m * kRawPitch;
      const float* const ws = ws_sm + rb * kWsFloats;
      const uint32_t par = nt & 1u;
      mbar_arrive(&bar_staged);  // the prep workers may look at the next tile (an edge tile has been staged above)
      // ---- MMA phase.  Static k-step assignment per lane quarter: the prep worker takes k0, k1 (prepared while we were
      //      in the previous epilogue) and k2 (it is free the moment k1 is handed over; we are still finishing the
      //      epilogue then); the half-1 worker, which leaves the epilogue first, k3 and k5; the half-0 worker k4 and k6. ----
#pragma unroll 1
      for (int r = 0; r < 2; ++r) {
        const int j = 4 - hh + 2 * r;
        prep_kstep_regs(xrow, ws, j, hv, lv);
        if (r == 1) mbar_arrive(&bar_raw_empty[rb]);  // this thread is done reading the raw tile
        TCW(1 + 2 * r);
        deposit_kstep(bar_a_empty[j - 1], par, bar_a_full, a_slot0, hv, lv, err_flag);
        TCW(2 + 2 * r);
      }

      // ---- epilogue phase ----
      const bool valid = t0 + m < kNFrames;
      OutT* const obase = reinterpret_cast<OutT*>(p.out) + (size_t)cur.b * kNMel * kNFrames + t0 + m;
      if (hh == 0 && p.mask != nullptr && valid) p.mask[(size_t)cur.b * kNFrames + t0 + m] = ((t0 + m) * kHop < cur.len) ? 1 : 0;
      if (wt == 0) TCT(1, nt, 0);
      mbar_wait(&bar_d_full, nt & 1u, err_flag);
      tc_fence_after();
      if (warp == 0 && nt == 0) TCS(14, 5);
      if (wt == 0) TCT(1, nt, 1);
      TCW(6);
      uint32_t rmax = 0u, rming[kMinGroups] = {0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u};
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
    rming[(mm) >> 5] = min(rming[(mm) >> 5], u_);                                             \
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
#pragma unroll
      for (int g = 0; g < kMinGroups; ++g) rming[g] = __reduce_min_sync(0xffffffffu, rming[g]);
      mbar_wait(&bar_st_empty[nt & 1u], ((nt >> 1) & 1u) ^ 1u, err_flag);
      __syncwarp();
      if (lane == 0) {
        s_red[nt & 1u][0][warp] = fmaf(lg2_approx(__uint_as_float(rmax)), 0.25f * kLog10_2, tile_k);
#pragma unroll
        for (int g = 0; g < kMinGroups; ++g)
          s_red[nt & 1u][1 + g][warp] = fmaf(lg2_approx(__uint_as_float(rming[g])), 0.25f * kLog10_2, tile_k);
        if (warp == 0) s_red_id[nt & 1u] = (uint32_t)cur.b * (uint32_t)kNTiles + (uint32_t)cur.tile;
        __threadfence_block();
        mbar_arrive(&bar_st_full[nt & 1u]);  // release: the tile's global stores (ordered by __syncwarp) and s_red
      }
      TCW(7);
      cur = nxt;
      ++nt;
      if (cur.mode == kModeDone) break;
    }
    // no more tiles: tell the books warp through the same channel (it never looks at the loader's meta data, whose slot
    // the loader may already have refilled by the time the books warp gets to it)
    mbar_wait(&bar_st_empty[nt & 1u], ((nt >> 1) & 1u) ^ 1u, err_flag);
    __syncwarp();
    if (lane == 0) {
      if (warp == 0) s_red_id[nt & 1u] = kBooksDone;
      __threadfence_block();
      mbar_arrive(&bar_st_full[nt & 1u]);
    }
  } else if (warp < 12) {
    // ======================================== PREP WORKERS ========================================
    // One warp per lane quarter that only prepares operands: k-steps 0, 1 and 2 of every tile, k0 and k1 of the NEXT
    // tile while the epilogue workers are busy with this one -- the tensor core restarts the moment the accumulators
    // are drained.  They also turn the next tile's maximum into its scale and scaled window table.
    reg_shrink<kRegsP>();
    const int qt = warp - 8;
    const int m = qt * 32 + lane;
    const int pt = qt * 32 + lane;  // 0..127 among the prep workers
    const uint32_t a_slot0 = tmem + ((uint32_t)(qt * 32) << 16) + kTmemA;
    uint32_t hv[4][8], lv[4][8];
    uint32_t nt = 0;
    mbar_wait(&bar_meta_full[0], 0u, err_flag);
    int mode = s_meta[0].mode;
    if (warp == 8) TCS(15, 0);
    if (mode != kModeDone) bar_sync_n(kBarWs0, kBarBoth);  // tile 0's window table is ready (later tiles: we made it)
    if (warp == 8) TCS(15, 1);
    while (mode != kModeDone) {
      const uint32_t rb = nt & 1u, par = nt & 1u;
      const float* const xrow = reinterpret_cast<const float*>(smem + kSmemRaw + rb * kRawBufBytes) + m * kRawPitch;
      const float* const ws = ws_sm + rb * kWsFloats;
      TCW(0);
#pragma unroll 1
      for (int jj = 0; jj < 3; ++jj) {
        const int j = jj;
        prep_kstep_regs(xrow, ws, j, hv, lv);
        if (jj == 2) mbar_arrive(&bar_raw_empty[rb]);  // k-steps 0, 1, 2: this thread is done reading the raw tile
        TCW(1 + 2 * jj);
        // k0 follows the previous tile's last MMAs (for tile 0: nothing -- parity 1 of a fresh barrier has "completed")
        deposit_kstep(bar_a_empty[j == 0 ? kKSteps - 1 : j - 1], j == 0 ? par ^ 1u : par, bar_a_full, a_slot0, hv, lv, err_flag);
        TCW(2 + 2 * jj);
      }
      // the next tile: geometry from the loader, maximum from the epilogue workers
      mbar_wait(&bar_meta_full[rb ^ 1u], ((nt + 1u) >> 1) & 1u, err_flag);
      mode = s_meta[rb ^ 1u].mode;
      mbar_wait(&bar_staged, par, err_flag);
      if (is_tma_mode(mode)) {
        // the tile's maximum -> power-of-two scale -> scaled window table (we have a tensor-core k-step or two to spare)
        const uint32_t mx = __reduce_max_sync(
            0xffffffffu, raw_absmax_rows(reinterpret_cast<const float*>(smem + kSmemRaw + (rb ^ 1u) * kRawBufBytes), qt, lane));
        if (lane == 0) atomicMax(&s_tilemax[rb ^ 1u], mx);
        bar_sync_n(kBarPrep, 128);
        float scale, tile_k;
        scale_from_max(s_tilemax[rb ^ 1u], scale, tile_k);
        write_ws(ws_sm + (rb ^ 1u) * kWsFloats, scale, pt, 128);
        if (pt == 0) s_scale[rb ^ 1u] = make_float2(scale, tile_k);
      }
      bar_sync_n(kBarPrep, 128);  // the table is complete and everybody has read the maximum
      if (pt == 0) s_tilemax[rb ^ 1u] = 0u;  // next accumulated two tiles from now
      mbar_arrive(&bar_ws);
      TCW(7);
      ++nt;
    }
  } else if (warp == kWarpMma) {
    reg_shrink<kRegsH>();
    // =========================================== MMA ISSUER ===========================================
    // The whole warp walks the loop (waits included); one elected lane issues the tensor-core instructions.
    // B descriptors differ only in the start address: add (byte offset >> 4) to the low word (addresses < 256 KB)
    const uint64_t b_desc0 = smem_desc(smem_u32(b_sm), kBChunkBytes, 128);
    uint32_t ks = 0, nt = 0;
    mbar_wait<WFE_TC_SLEEP_LONG>(&bar_tab, 0u, err_flag);  // B
    for (;;) {
      mbar_wait<WFE_TC_SLEEP_LONG>(&bar_meta_full[nt & 1u], (nt >> 1) & 1u, err_flag);  // the loader has handed out another tile (or none)
      if (s_meta[nt & 1u].mode == kModeDone) break;
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
            mma_commit(&bar_a_empty[j][n1]);  // implies tcgen05.fence::before_thread_sync
          }
          __syncwarp();
        }
        if (lane == 0) TCT(2, nt, 2 + j);
      }
      if (elect_one()) mma_commit(&bar_d_full);
      __syncwarp();
      ++nt;
    }
  } else if (warp == kWarpLoad) {
    reg_shrink<kRegsH>();
    // =========================================== LOADER ===========================================
    // One TMA per tile into the free raw buffer, the reflect-pad patch of a clip's first tile, and the tile's geometry:
    // handed to the workers through bar_meta_full.  (Scanning the tile for its maximum here too was tried: one warp
    // needs 9 k cycles for it and slows the two workers of its sub-partition -- the workers do it in their idle slots.)
    // one past the last PCM element of the batch: a box that ends below it reads the caller's buffer, whatever clip it is in
    int64_t extent = -1;  // computed when the first clip tail asks for it: at the start it would delay the first tile by 5 k cycles
    TCS(13, 0);
    uint32_t nt = 0;
    uint32_t tma_loads[2] = {0u, 0u};
    for (uint32_t i = 0;; ++i) {
      Tile t;
      t.mode = kModeDone;
      t.b = t.tile = t.len = 0;
      t.off = 0;
      // Tiles are handed out in order from a global counter: the SMs do not run at the same pace (the first TPC of every
      // GPC is ~12 % faster than the last ones -- instruction supply), and with a static split the slowest CTA sets the
      // launch time.  The round trip of the atomic hides behind the wait for the raw buffer below.
      uint32_t id = blockIdx.x;  // (the first tile without the atomic's round trip)
      if (i > 0) {
        if (lane == 0) id = atomicAdd(p.tile_counter, 1u) + gridDim.x;
        id = __shfl_sync(0xffffffffu, id, 0);
      }
      if (id < p.total_tiles) {
        t = tile_info(p, id, extent);
        if (t.mode == kModeAsyncTail) {
          // the box reads past the clip: fine as long as it stays inside the caller's buffer, i.e. below the end of the
          // last clip (known after one pass over the batch's offsets, made when the first tail shows up)
          if (extent < 0) {
            extent = 0;
            // (eight clips per lane in flight: a rolled loop pays one global-memory latency per 32 clips -- 8 k cycles of the
            //  kernel's start for a batch of 256)
            const int n_clips = (int)(p.total_tiles / (uint32_t)kNTiles);
            for (int b0 = lane; b0 < n_clips; b0 += 32 * 8) {
              int64_t off[8], end[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int b = min(b0 + 32 * u, n_clips - 1);  // (re-reading the last clip is harmless)
                off[u] = __ldg(p.offsets + b);
                end[u] = p.lengths != nullptr ? __ldg(p.lengths + b) : __ldg(p.offsets + b + 1);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) extent = max(extent, p.lengths != nullptr ? off[u] + end[u] : end[u]);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) extent = max(extent, __shfl_xor_sync(0xffffffffu, extent, d));
          }
          if (t.off + (int64_t)(t.tile * kTileM * kHop - kNFft / 2 + (kRawRows - 1) * kHop + kRawPitch) > extent) t.mode = kModeSync;
        }
        if (t.mode == kModeSilent) {
          // a tile wholly in the zero padding is never computed: published for the clamp pass (which fills it) right here
          if (lane == 0) p.tile_key[id] = f2key(-1.5f);  // (log10(1e-10) + 4) / 4: what the reference computes for zero padding
          if (lane < kMinBlocks) p.tile_min[(size_t)id * kMinBlocks + lane] = kMinSilent;
          if (p.mask != nullptr) {
            const int t0 = t.tile * kTileM;
            for (int f = lane; f < kTileM && t0 + f < kNFrames; f += 32) p.mask[(size_t)t.b * kNFrames + t0 + f] = 0;
          }
          continue;
        }
      }
      const uint32_t rb = nt & 1u;
      float* const raw = reinterpret_cast<float*>(smem + kSmemRaw + rb * kRawBufBytes);
      if (nt < 2) TCS(13, 1 + 3 * nt);
      if (lane == 0) TCT(3, nt, 0);
      mbar_wait<WFE_TC_SLEEP_LONG>(&bar_raw_empty[rb], ((nt >> 1) & 1u) ^ 1u, err_flag);  // the workers have finished with this buffer's previous tile
      if (lane == 0) TCT(3, nt, 1);
      float scale = 1.f, tile_k = 1.f;
      if (is_tma_mode(t.mode)) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_raw_full[rb], kRawBoxBytes);
          tma_load_2d(raw, &tmap, (int32_t)(t.off + (int64_t)(t.tile * kTileM * kHop - kNFft / 2)), 0, &bar_raw_full[rb]);
        }
        __syncwarp();
        // (the phase of bar_raw_full[rb] counts the TMA loads into this buffer, not its uses: staged tiles never touch it.
        //  Round-2 versions up to v16 took the parity from the use count and, after the first staged tile, stopped
        //  waiting for the data -- a rare run-to-run difference at full size was the symptom.)
        mbar_wait<WFE_TC_SLEEP_LONG>(&bar_raw_full[rb], tma_loads[rb] & 1u, err_flag);
        if (nt < 2) TCS(13, 2 + 3 * nt);
        ++tma_loads[rb];
        if (lane == 0) TCT(3, nt, 2);
        // Edge patches, from the tile itself (the samples a reflection needs lie within 200 samples of the edge):
        if (t.mode == kModeAsyncHead) {
          // first tile of a clip: the TMA started 200 samples before the clip; replace them by the centred reflect pad,
          // raw[i] = x[200 - i] = raw[400 - i] (zero where the clip is shorter than that)
          float v[7];
#pragma unroll
          for (int u = 0; u < 7; ++u) {
            const int i = lane + 32 * u, j = kNFft - i;  // source index 201..400, rows 1 and 2
            const int jr = j / kHop;
            v[u] = (i < kNFft / 2 && kNFft / 2 - i < t.len) ? raw[jr * kRawPitch + (j - jr * kHop)] : 0.f;
          }
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 7; ++u) {
            const int i = lane + 32 * u;
            if (i < kNFft / 2) raw[i + (i >= kHop ? kRawPitch - kHop : 0)] = v[u];
          }
          __syncwarp();
        }
        if (t.mode == kModeAsyncTail) {
          // the clip ends inside this tile: everything from its end on is zero padding, or -- beyond sample 480000 -- the
          // centred reflect pad; rows that no valid frame reads are zeroed too (the maximum is taken over the whole tile)
          const int s_begin = t.tile * kTileM * kHop - kNFft / 2;
          const int i0 = t.len - s_begin;  // first raw index past the clip (> 0: the tile is not silent)
          // stores only: the rest of the clip's last row, then whole rows, 16 bytes at a time
          // (a clip that ends inside the four pad columns of the box's last row has i0 = 130 * 160 .. + 3: every sample a
          //  frame reads is real, nothing to zero -- and row 130 would be the first row of the NEXT buffer, or of B)
          const int r0 = i0 / kHop, c0 = i0 - r0 * kHop;
          if (r0 < kRawRows)
            for (int c = c0 + lane; c < kHop; c += 32) raw[r0 * kRawPitch + c] = 0.f;
          for (int q = lane; q < (kRawRows - 1 - r0) * (kHop / 4); q += 32) {
            const int r = r0 + 1 + q / (kHop / 4);
            *reinterpret_cast<float4*>(raw + r * kRawPitch + 4 * (q % (kHop / 4))) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          __syncwarp();
          // reflect pad: raw index ir + k <-> sample 480000 + k <-> source sample 479998 - k = raw index ir - 2 - k (zero
          // already where the clip is shorter); only the clip's last tile reaches sample 480000
          const int ir = kNSamples - s_begin;  // raw index of sample 480000
          if (ir < kRawRows * kHop) {
            float v[7];
#pragma unroll
            for (int u = 0; u < 7; ++u) {
              const int k = lane + 32 * u, j = ir - 2 - k;
              const int jr = j / kHop;
              v[u] = (k < kNFft / 2 && j >= 0) ? raw[jr * kRawPitch + (j - jr * kHop)] : 0.f;
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 7; ++u) {
              const int i = ir + lane + 32 * u;
              if (lane + 32 * u < kNFft / 2 && i < kRawRows * kHop) {
                const int r = i / kHop;
                raw[r * kRawPitch + (i - r * kHop)] = v[u];
              }
            }
            __syncwarp();
          }
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
      if (nt < 2) TCS(13, 3 + 3 * nt);
      if (t.mode == kModeDone) break;
      ++nt;
    }
  } else if (warp == kWarpClamp) {
    reg_shrink<kRegsH>();
    // =========================================== CLAMP BOOKS ===========================================
    // Publishes every computed tile's extrema for the clamp pass (clamp_kernel).  Tile id and extrema both come from the
    // epilogue workers (s_red_id, s_red), handed over by bar_st_full / bar_st_empty.  (Up to round-2 v22 the id was read
    // from the loader's s_meta slot, which the loader refills as soon as the workers have finished READING the tile's raw
    // samples -- immediately when no TMA is needed: the end-of-work marker, a staged tile.  A books warp that was a
    // thousand cycles late then published under the wrong id, or left early, and the clamp pass read stale words for
    // that tile: wrong clamping in a few blocks, once in a few thousand ragged batches -- tools/fuzz_tc_vs_cc.py.)
    uint32_t nt = 0;
    for (;;) {
      mbar_wait<WFE_TC_SLEEP_LONG>(&bar_st_full[nt & 1u], (nt >> 1) & 1u, err_flag);
      const uint32_t id = s_red_id[nt & 1u];
      if (id == kBooksDone) break;
      // a worker warp whose 32 frames lie beyond frame 3000 reports the identities (lg2(0) = -inf, lg2(inf) = +inf)
      float mx = -__int_as_float(0x7f800000);
#pragma unroll
      for (int w = 0; w < 8; ++w) mx = fmaxf(mx, s_red[nt & 1u][0][w]);
      // the two epilogue warps of a frame quarter (w = quarter, quarter + 4) each finish part of every mel group;
      // lanes 0..15: minimum of block (frame quarter lane / 4, mel group lane % 4)
      const int q = (lane >> 2) & 3, g = lane & 3;
      const uint32_t mnb = __float_as_uint(fminf(s_red[nt & 1u][1 + g][q], s_red[nt & 1u][1 + g][q + 4]));
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_st_empty[nt & 1u]);
      ++nt;
      if (lane == 0) p.tile_key[id] = f2key(mx);
      if (lane < kMinBlocks) p.tile_min[(size_t)id * kMinBlocks + lane] = mnb;
    }
  }

  else {
    reg_shrink<24>();  // sixteenth warp: only there so that every scheduler starts with four warps' worth of registers
  }

  // ---- teardown ----
#ifdef WFE_TC_TRACE
  if (lane == 0 && warp >= kWarpMma && warp <= kWarpClamp && (blockIdx.x == 0 || blockIdx.x == 77 || blockIdx.x == 147))
    g_tc_tiles[blockIdx.x == 0 ? 0 : (blockIdx.x == 77 ? 1 : 2)][61 - (warp - kWarpMma)] = clock64();
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77 || blockIdx.x == 147))
    g_tc_tiles[blockIdx.x == 0 ? 0 : (blockIdx.x == 77 ? 1 : 2)][62] = clock64();
#endif
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
  if (warp == kWarpLoad) tmem_dealloc(tmem, kTmemCols);
  if (tid == 0 && s_err != 0u) atomicExch(err_global, 0xDEADu);
}

}  // namespace tc
}  // namespace wfe
