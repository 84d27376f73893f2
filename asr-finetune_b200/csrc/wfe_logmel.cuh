// Whisper log-mel frontend kernels for sm_100a (B200).
//
// One CTA = one tile of 32 consecutive STFT frames of one clip; LANE == FRAME everywhere, so every
// shared-memory access is [row][lane] (stride-1 across the warp, conflict-free), every twiddle /
// window / mel weight is warp-uniform, and no shuffles or divergent branches are needed.
//
//   stage 0  coalesced 128-bit loads of the tile's 5360 PCM samples -> smem, applying truncate /
//            right-zero-pad to n_samples and the centred reflect pad            (Appendix A steps 2-3)
//   stage 1  16 tasks (n1): Hann window + real 25-point DFT (5x5) over samples n1+16*n2, then the
//            W400^(n1*k2) twiddle                                                (steps 5-6)
//   stage 2  13 tasks (k2): complex 16-point DFT (4x4) over n1 -> bins 25*k1+k2 (folded to 0..200 by
//            conjugate symmetry), power |X|^2 stored in place                    (steps 6-7)
//   stage 3  n_mel tasks: banded slaney mel projection (fp32 FFMA, <=2 non-zeros per bin), log10,
//            (x+4)/4, coalesced store, per-clip running max                      (steps 8, 9, 11)
//   stage 4  the LAST tile of a clip to finish applies the per-clip max-8 clamp to the (L2-resident)
//            tiles that need it                                                  (step 10)
//
// Arithmetic restated from HF:models/whisper/feature_extraction_whisper.py:135-164 (see SURVEY.md
// Appendix A); 400 = 16 x 25 Cooley-Tukey: n = n1 + 16*n2, k = k2 + 25*k1,
//   X[k2+25k1] = sum_n1 W16^(n1 k1) * W400^(n1 k2) * sum_n2 x[n1+16 n2] W25^(n2 k2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wfe_codelets.cuh"

namespace wfe {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSigLen = (kTileF - 1) * kHop + kNFft;      // 5360 padded-signal samples per tile
constexpr int kSigSm = kSigLen + kSigLen / kHop + 2;      // +1 pad word per 160 samples (bank skew)
constexpr int kZRows = 400;                               // intermediate rows per frame
constexpr size_t kSmemBytes = (size_t)(kSigSm + kZRows * kTileF) * sizeof(float);

// order-preserving float <-> uint32 key (for atomicMax on floats of either sign); key 0 < every float
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

struct LogmelParams {
  const void* pcm;
  const int64_t* offsets;
  const float2* norm;      // (mean, rstd) per clip or nullptr
  float* out;              // (B, n_mel, n_frames)
  int32_t* mask;           // (B, n_frames) or nullptr
  uint32_t* clip_key;      // [B] running max of log2(mel) as ordered key (zero-initialised)
  uint32_t* clip_ticket;   // [B] finished-tile counter (zero-initialised)
  float* tile_min;         // [B * ntiles]
  const int2* mel_tab;     // nnz entries: (row*32, float bits of weight), grouped by mel
  const int32_t* mel_start;  // [n_mel + 1]
  float pcm_scale;
  int n_mel, n_samples, n_frames, ntiles;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) logmel_kernel(const LogmelParams p) {
  extern __shared__ __align__(16) float smem[];
  float* sig = smem;
  float* zbuf = smem + kSigSm;
  __shared__ float s_red[2][kWarps];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.ntiles, tile = blockIdx.x - b * p.ntiles;
  const int t0 = tile * kTileF;
  const int64_t off = p.offsets[b];
  const int64_t avail = p.offsets[b + 1] - off;
  const int len = (int)(avail < (int64_t)p.n_samples ? avail : (int64_t)p.n_samples);   // truncate to 30 s
  const int s_begin = t0 * kHop - kNFft / 2;   // unpadded sample index of sig[0]
  const int nvalid = min(kTileF, p.n_frames - t0);
  float* const out_tile = p.out + ((size_t)b * p.n_mel) * p.n_frames + t0;

  if (p.mask != nullptr && tid < nvalid) p.mask[(size_t)b * p.n_frames + t0 + tid] = ((t0 + tid) * kHop < len) ? 1 : 0;

  // lowest source sample this tile touches (right reflect maps s >= n_samples to 2(n-1)-s)
  const int s_hi = s_begin + kSigLen - 1;
  int lowest = s_begin < 0 ? 0 : s_begin;
  if (s_hi >= p.n_samples) lowest = min(lowest, 2 * (p.n_samples - 1) - s_hi);
  const bool silent = lowest >= len;   // every sample of every frame in the tile is zero padding

  float tmax_lg = -3.0e38f, tmin_y = 3.0e38f;
  if (silent) {
    // mel == 0 exactly -> log10(1e-10) path; no FFT needed
    const float lg = __log2f(1e-10f);
    const float y = fmaf(lg, 0.25f * kLog10_2, 1.0f);
    for (int i = tid; i < p.n_mel * kTileF; i += kThreads) {
      const int m = i >> 5, f = i & 31;
      if (f < nvalid) out_tile[(size_t)m * p.n_frames + f] = y;
    }
    tmax_lg = lg;
    tmin_y = y;
  } else {
    // ---- stage 0: PCM -> smem (skewed), with zero pad / reflect pad / optional normalisation ----
    const T* pcm = reinterpret_cast<const T*>(p.pcm) + off;
    float mean = 0.f, rstd = 1.f;
    if (p.norm != nullptr) {
      const float2 st = p.norm[b];
      mean = st.x;
      rstd = st.y;
    }
    const bool interior = (s_begin >= 0) && (s_begin + kSigLen <= len);
    constexpr int kVec = 16 / (int)sizeof(T);
    if (interior) {
      const T* src = pcm + s_begin;
      const int mis = (int)((reinterpret_cast<uintptr_t>(src) / sizeof(T)) % kVec);
      const int head = (kVec - mis) % kVec;
      const int nvec = (kSigLen - head) / kVec;
      for (int i = tid; i < head; i += kThreads) sig[i + i / kHop] = (pcm_to_float<T>(src[i], p.pcm_scale) - mean) * rstd;
      const uint4* src4 = reinterpret_cast<const uint4*>(src + head);
      for (int v = tid; v < nvec; v += kThreads) {
        const uint4 raw = __ldg(src4 + v);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
          const int i = head + v * kVec + j;
          sig[i + i / kHop] = (pcm_to_float<T>(e[j], p.pcm_scale) - mean) * rstd;
        }
      }
      for (int i = head + nvec * kVec + tid; i < kSigLen; i += kThreads)
        sig[i + i / kHop] = (pcm_to_float<T>(src[i], p.pcm_scale) - mean) * rstd;
    } else {
      for (int i = tid; i < kSigLen; i += kThreads) {
        int s = s_begin + i;
        if (s < 0) s = -s;
        if (s >= p.n_samples) s = 2 * (p.n_samples - 1) - s;
        float v = 0.f;
        if (s >= 0 && s < len) v = (pcm_to_float<T>(pcm[s], p.pcm_scale) - mean) * rstd;
        sig[i + i / kHop] = v;
      }
    }
    __syncthreads();

    // ---- stage 1 ----
    {
      const float* sig_lane = sig + (kHop + 1) * lane;
      float* zcol = zbuf + lane;
      for (int n1 = warp; n1 < 16; n1 += kWarps) stage1_task(sig_lane, n1, zcol);
    }
    __syncthreads();
    // ---- stage 2 ----
    for (int k2 = warp; k2 < 13; k2 += kWarps) stage2_task(zbuf + lane, k2);
    __syncthreads();
    // ---- stage 3: banded mel + log + scale ----
    const bool lane_ok = lane < nvalid;
    for (int m = warp; m < p.n_mel; m += kWarps) {
      const int e0 = __ldg(p.mel_start + m), e1 = __ldg(p.mel_start + m + 1);
      float acc = 0.f;
      for (int e = e0; e < e1; ++e) {
        const int2 ent = __ldg(p.mel_tab + e);
        acc = fmaf(__int_as_float(ent.y), zbuf[ent.x + lane], acc);
      }
      const float lg = __log2f(fmaxf(acc, 1e-10f));
      const float y = fmaf(lg, 0.25f * kLog10_2, 1.0f);
      if (lane_ok) {
        out_tile[(size_t)m * p.n_frames + lane] = y;
        tmax_lg = fmaxf(tmax_lg, lg);
        tmin_y = fminf(tmin_y, y);
      }
    }
  }

  // ---- tile max / min -> clip running max; last tile of the clip applies the clamp ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tmax_lg = fmaxf(tmax_lg, __shfl_xor_sync(0xffffffffu, tmax_lg, o));
    tmin_y = fminf(tmin_y, __shfl_xor_sync(0xffffffffu, tmin_y, o));
  }
  if (lane == 0) {
    s_red[0][warp] = tmax_lg;
    s_red[1][warp] = tmin_y;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    float mx = s_red[0][0], mn = s_red[1][0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) {
      mx = fmaxf(mx, s_red[0][w]);
      mn = fminf(mn, s_red[1][w]);
    }
    atomicMax(p.clip_key + b, f2key(mx));
    p.tile_min[(size_t)b * p.ntiles + tile] = mn;
    __threadfence();
    const uint32_t ticket = atomicAdd(p.clip_ticket + b, 1u);
    s_last = (ticket == (uint32_t)p.ntiles - 1u);
  }
  __syncthreads();
  if (!s_last) return;

  // ---- stage 4 (one CTA per clip): out = max(out, ((g - 8) + 4) / 4), g = log10 of the clip max ----
  __threadfence();
  const float g = key2f(__ldcg(p.clip_key + b)) * kLog10_2;
  const float floor_y = ((g - 8.0f) + 4.0f) * 0.25f;
  float* const out_clip = p.out + ((size_t)b * p.n_mel) * p.n_frames;
  for (int tt = 0; tt < p.ntiles; ++tt) {
    if (__ldcg(p.tile_min + (size_t)b * p.ntiles + tt) >= floor_y) continue;   // uniform across the CTA
    const int nv = min(kTileF, p.n_frames - tt * kTileF);
    for (int i = tid; i < p.n_mel * kTileF; i += kThreads) {
      const int m = i >> 5, f = i & 31;
      if (f < nv) {
        float* q = out_clip + (size_t)m * p.n_frames + tt * kTileF + f;
        if (__ldcg(q) < floor_y) *q = floor_y;
      }
    }
  }
}

// ---- per-clip mean / rstd for do_normalize (HF:...feature_extraction_whisper.py:168-187) ------------
template <typename T>
__global__ void __launch_bounds__(512) clip_stats_kernel(const void* pcm_, float scale, const int64_t* offsets,
                                                         int n_samples, float2* stats) {
  const int b = blockIdx.x;
  const int64_t off = offsets[b];
  const int64_t avail = offsets[b + 1] - off;
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
