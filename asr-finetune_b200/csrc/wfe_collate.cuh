// Speech seq2seq padding collator as ONE batched kernel (sm_100a).
//
// Restates ref:finetune/training/data_and_collator/datasets_and_collators.py:444-457
//   labels  = tokenizer.pad(ids).input_ids.masked_fill(attention_mask.ne(1), -100)   (length based)
//   bos     = (labels[:, 0] == decoder_start_token_id).all()
//   feats   = feature_extractor.pad(list of (n_mel, 3000), "longest")  == bit-exact stack
// Pure HBM-bound byte movement: 128-bit loads/stores, grid sized from the byte count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wfe {

constexpr int kCollateThreads = 256;

struct CollateParams {
  const int64_t* ids;        // ragged concat
  const int64_t* offsets;    // [B+1]
  int64_t* labels;           // (B, width)
  int32_t* bos_flag;         // [1] or nullptr
  const float* const* feat_srcs;   // device array of B device pointers, or nullptr
  float* feat_out;           // (B, feat_elems)
  int64_t feat_elems;
  int64_t dec_start, ignore_index;
  int32_t batch, width;
  int32_t label_blocks;      // blocks [0, label_blocks) do labels; the rest copy features
  int32_t feat_blocks_per_clip;
};

__global__ void __launch_bounds__(kCollateThreads) collate_kernel(const CollateParams p) {
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < p.label_blocks) {
    // ---- labels: one element per thread, grid-stride over B*width ----
    const int64_t total = (int64_t)p.batch * p.width;
    for (int64_t i = (int64_t)blockIdx.x * kCollateThreads + tid; i < total; i += (int64_t)p.label_blocks * kCollateThreads) {
      const int b = (int)(i / p.width), j = (int)(i - (int64_t)b * p.width);
      const int64_t o0 = __ldg(p.offsets + b), o1 = __ldg(p.offsets + b + 1);
      p.labels[i] = (j < o1 - o0) ? __ldg(p.ids + o0 + j) : p.ignore_index;
    }
    if (blockIdx.x == 0 && p.bos_flag != nullptr) {
      int ok = 1;
      for (int b = tid; b < p.batch; b += kCollateThreads) {
        const int64_t o0 = __ldg(p.offsets + b), o1 = __ldg(p.offsets + b + 1);
        // an empty row pads to ignore_index at column 0, which never equals the start token
        const int64_t first = (o1 > o0 && p.width > 0) ? __ldg(p.ids + o0) : p.ignore_index;
        ok &= (first == p.dec_start);
      }
      ok = __syncthreads_and(ok);
      if (tid == 0) *p.bos_flag = (p.batch > 0 && p.width > 0) ? ok : 0;
    }
    return;
  }
  // ---- feature stack: clip-major blocks, 128-bit copies when both sides are 16 B aligned ----
  const int fb = (int)blockIdx.x - p.label_blocks;
  const int b = fb / p.feat_blocks_per_clip, part = fb - b * p.feat_blocks_per_clip;
  const float* src = p.feat_srcs[b];
  float* dst = p.feat_out + (int64_t)b * p.feat_elems;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0 && (p.feat_elems & 3) == 0;
  if (vec_ok) {
    const int64_t n4 = p.feat_elems >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const int64_t stride = (int64_t)p.feat_blocks_per_clip * kCollateThreads;
    int64_t i = (int64_t)part * kCollateThreads + tid;
    // 4 independent 128-bit loads in flight per thread
    for (; i + 3 * stride < n4; i += 4 * stride) {
      const float4 a = __ldcs(s4 + i), b4 = __ldcs(s4 + i + stride), c = __ldcs(s4 + i + 2 * stride),
                   d = __ldcs(s4 + i + 3 * stride);
      d4[i] = a;
      d4[i + stride] = b4;
      d4[i + 2 * stride] = c;
      d4[i + 3 * stride] = d;
    }
    for (; i < n4; i += stride) d4[i] = __ldcs(s4 + i);
  } else {
    for (int64_t i = (int64_t)part * kCollateThreads + tid; i < p.feat_elems; i += (int64_t)p.feat_blocks_per_clip * kCollateThreads)
      dst[i] = src[i];
  }
}

}  // namespace wfe
