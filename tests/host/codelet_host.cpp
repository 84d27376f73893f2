// Host build of the device FFT codelets (asr-finetune_b200/csrc/wfe_codelets.cuh) for CPU-only tests:
// the same straight-line arithmetic the sm_100a kernel runs per thread (packed f32x2 pairs become two floats),
// driven lane by lane through the same shared-memory layouts.
#include <string.h>

#include "../../asr-finetune_b200/csrc/wfe_codelets.cuh"

extern "C" {

// sig: one tile of padded signal, kSigLen = 31*160+400 = 5360 samples (plain, unskewed)
// power: (201, 32) row-major — |STFT|^2 of the tile's 32 frames
void codelet_tile_power(const float* sig, float* power) {
  using namespace wfe;
  alignas(16) static float cst[8 * kS1ConstVec * 4];
  static bool init = false;
  if (!init) {
    fill_stage1_consts(cst);
    init = true;
  }
  alignas(16) static float skew[5360 + 2 * (5360 / 160) + 8];
  static float zbuf[kZPlanes * 16 * kTileF];
  static float pbuf[kBins * kPStride];
  memset(pbuf, 0, sizeof(pbuf));
  for (int i = 0; i < 5360; ++i) skew[i + 2 * (i / kHop)] = sig[i];
  for (int lane = 0; lane < 32; ++lane)
    for (int w = 0; w < 8; ++w)
      stage1_pair(skew + (kHop + 2) * lane, reinterpret_cast<const float4*>(cst) + w * kS1ConstVec, 2 * w, zbuf + lane);
  for (int lane = 0; lane < 32; ++lane) {
    f2 pw[16];
    for (int a = 1; a < 13; a += 2) {
      stage2_pair_compute(zbuf + lane, a, pw);
      stage2_pair_store(pw, a, pbuf + lane);
    }
    stage2_k0_compute(zbuf + lane, pw);
    stage2_k0_store(pw, pbuf + lane);
  }
  for (int k = 0; k < kBins; ++k) memcpy(power + k * 32, pbuf + k * kPStride, 32 * sizeof(float));
}
}
