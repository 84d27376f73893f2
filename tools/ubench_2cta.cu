// tcgen05 CTA-pair micro-benchmark: can the log-mel kernel's DFT stage (M = 128 frames per SM, N = 112, K = 16, A operand
// in TMEM) run as ONE cta_group::2 instruction per SM pair (M = 256), so that every SM fetches only HALF of the B operand
// from shared memory?  (The tensor-core kernel of wfe_logmel_tc.cuh is shared-memory bound in its MMA phase: 3.5 KB of B
// per 57-cycle MMA is 47 % of the pipe, the operand preparation needs the other half.)
//   1. correctness: D = A . B^T with A (2 x 128 x 16 fp16, per-CTA rows, written to TMEM by tcgen05.st) and B (112 x 16
//      fp16, rows [0, 56) in CTA 0's shared memory, rows [56, 112) in CTA 1's), N = 112 (not a multiple of 32), checked
//      against the host product for both CTAs' accumulators;
//   2. rate: cycles per MMA, back to back, cta_group::1 (every CTA on its own, whole B) vs cta_group::2;
//   3. shared-memory relief: the same with eight warps per CTA streaming conflict-free LDS.128 beside the MMAs.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_2cta tools/ubench_2cta.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

constexpr int kM = 128, kN = 112, kK = 16, kACol = 448;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__host__ __device__ inline uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
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

// kPair = true: one cta_group::2 MMA per SM pair (leader issues, B split over the two CTAs);
// kPair = false: every CTA issues its own cta_group::1 MMAs with the whole B in its shared memory.
template <bool kPair>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(288, 1)
    pair_kernel(const __half* __restrict__ a_g, const __half* __restrict__ b_g, float* __restrict__ d_g, int reps,
                int lds_iters, long long* __restrict__ cycles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];  // [0, 4 KB) B (half or whole), [8 KB, 8 KB + 64 KB) LDS stream
  __shared__ uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int n_rows = kPair ? kN / 2 : kN, n0 = kPair ? (int)rank * (kN / 2) : 0;
  // B, canonical K-major no-swizzle layout: [K chunk of 8][row][8 x fp16]
  for (int i = tid; i < 2 * n_rows; i += 288) {
    const int c = i / n_rows, n = i - c * n_rows;
    reinterpret_cast<uint4*>(smem)[i] = *reinterpret_cast<const uint4*>(b_g + (size_t)(n0 + n) * kK + 8 * c);
  }
  for (int i = tid; i < 4096; i += 288) reinterpret_cast<float4*>(smem + 8192)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  // A: thread = row = TMEM lane; 16 fp16 = 8 columns at kACol
  if (tid < 128) {
    uint32_t r[8];
    const uint4* src = reinterpret_cast<const uint4*>(a_g + ((size_t)rank * kM + tid) * kK);
    const uint4 v0 = src[0], v1 = src[1];
    r[0] = v0.x, r[1] = v0.y, r[2] = v0.z, r[3] = v0.w, r[4] = v1.x, r[5] = v1.y, r[6] = v1.z, r[7] = v1.w;
    tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + kACol, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  long long t0 = clock64();
  if (warp == 0 && (!kPair || rank == 0)) {
    if (lane == 0) {
      const uint64_t bdesc = smem_desc(smem_u32(smem), (uint32_t)n_rows * 16, 128);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)((kPair ? 2 * kM : kM) >> 4) << 24);
      for (int r = 0; r < reps; ++r) {
        if (kPair)
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
              "r"(tmem + kACol), "l"(bdesc), "r"(idesc), "r"(0u)
              : "memory");
        else
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
              "r"(tmem + kACol), "l"(bdesc), "r"(idesc), "r"(0u)
              : "memory");
      }
      if (kPair)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(&bar_done)),
                     "h"((uint16_t)3)
                     : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done))
                     : "memory");
    }
    __syncwarp();
  }
  float sink = 0.f;
  long long t_lds = 0;
  if (warp > 0 && lds_iters > 0) {
    // conflict-free LDS.128 stream (lane stride 16 B), 8 loads in flight
    const float4* base = reinterpret_cast<const float4*>(smem + 8192) + lane;
    const long long s0 = clock64();
    for (int it = 0; it < lds_iters; ++it) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = base[((it * 8 + u) & 127) * 32];
#pragma unroll
      for (int u = 0; u < 8; ++u) sink += v[u].x + v[u].w;
    }
    t_lds = clock64() - s0;
  }
  const bool ok = mbar_wait(&bar_done, 0);
  const long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && lane == 0) atomicExch(err, 1);
  if (tid == 0) cycles[blockIdx.x * 2] = t1 - t0;
  if (tid == 32) cycles[blockIdx.x * 2 + 1] = t_lds;
  if (sink == 12345.678f) d_g[0] = sink;
  // accumulators: lanes 32 warp .. 32 warp + 31, columns 0..111
  if (blockIdx.x < 2 && tid < 128) {
    for (int c = 0; c < kN; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int u = 0; u < 16; ++u) d_g[((size_t)blockIdx.x * kM + tid) * kN + c + u] = __uint_as_float(r[u]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0) {
    if (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <bool kPair>
void run(const char* name, const __half* d_a, const __half* d_b, float* d_d, const std::vector<float>& ref, int grid) {
  long long* d_cyc;
  int* d_err;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 2 * grid));
  CK(cudaMalloc(&d_err, sizeof(int)));
  const size_t smem = 8192 + 65536;
  CK(cudaFuncSetAttribute(pair_kernel<kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int lds_iters = cfg == 0 ? 0 : 4000;
    const int reps = cfg == 1 ? 1 : 12000;  // cfg 1: the LDS stream (almost) alone
    CK(cudaMemset(d_err, 0, sizeof(int)));
    CK(cudaMemset(d_d, 0, sizeof(float) * 2 * kM * kN));
    pair_kernel<kPair><<<grid, 288, smem>>>(d_a, d_b, d_d, reps, lds_iters, d_cyc, d_err);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int err = 0;
    CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<long long> cyc(2 * grid);
    CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
    std::vector<float> d(2 * kM * kN);
    CK(cudaMemcpy(d.data(), d_d, sizeof(float) * d.size(), cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (size_t i = 0; i < d.size(); ++i) maxerr = fmax(maxerr, fabs((double)d[i] - ref[i]));
    double mma = 0, lds = 0;
    for (int b = 0; b < grid; ++b) mma += (double)cyc[2 * b] / grid, lds += (double)cyc[2 * b + 1] / grid;
    // LDS stream per CTA: 8 warps x lds_iters x 8 LDS.128 x 4 wavefronts
    printf("%-28s grid %3d  lds_iters %5d: %7.1f cycles per MMA (wall of %d back-to-back)  max |D - ref| = %.3g  timeout=%d", name,
           grid, lds_iters, mma / reps, reps, maxerr, err);
    if (lds_iters) printf("   LDS stream: %.0f cycles for %d wavefronts per CTA = %.2f wavefronts/cycle", lds, 8 * lds_iters * 32,
                          8.0 * lds_iters * 32 / lds);
    printf("\n");
  }
  CK(cudaFree(d_cyc));
  CK(cudaFree(d_err));
}

int main() {
  std::vector<__half> a(2 * kM * kK), b(kN * kK);
  std::vector<float> ref(2 * kM * kN);
  srand(7);
  for (auto& v : a) v = __float2half((float)(rand() % 2001 - 1000) / 500.0f);
  for (auto& v : b) v = __float2half((float)(rand() % 2001 - 1000) / 1000.0f);
  for (int r = 0; r < 2 * kM; ++r)
    for (int n = 0; n < kN; ++n) {
      double s = 0;
      for (int k = 0; k < kK; ++k) s += (double)__half2float(a[r * kK + k]) * (double)__half2float(b[n * kK + k]);
      ref[(size_t)r * kN + n] = (float)s;
    }
  __half *d_a, *d_b;
  float* d_d;
  CK(cudaMalloc(&d_a, a.size() * 2));
  CK(cudaMalloc(&d_b, b.size() * 2));
  CK(cudaMalloc(&d_d, ref.size() * 4));
  CK(cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int grid = prop.multiProcessorCount & ~1;
  printf("%s, %d SMs; M = 128 rows per CTA, N = %d, K = %d, A in TMEM, kind::f16\n", prop.name, prop.multiProcessorCount, kN, kK);
  run<false>("cta_group::1 (whole B per SM)", d_a, d_b, d_d, ref, grid);
  run<true>("cta_group::2 (half B per SM)", d_a, d_b, d_d, ref, grid);
  return 0;
}
