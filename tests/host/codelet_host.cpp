// Host build of the device FFT codelets (asr-finetune_b200/csrc/wfe_codelets.cuh) for CPU-only tests:
// the same straight-line arithmetic the sm_100a kernel runs per thread, driven lane by lane.
#include <string.h>

#include "../../asr-finetune_b200/csrc/wfe_codelets.cuh"

extern "C" {

// sig: one tile of padded signal, kSigLen = 31*160+400 = 5360 samples (plain, unskewed)
// power: (201, 32) row-major — |STFT|^2 of the tile's 32 frames
void codelet_tile_power(const float* sig, float* power) {
  using namespace wfe;
  static bool init = false;
  if (!init) {
    fill_tables(c_win, c_tw400);
    init = true;
  }
  static float skew[5360 + 5360 / 160 + 2];
  static float zbuf[400 * 32];
  for (int i = 0; i < 5360; ++i) skew[i + i / kHop] = sig[i];
  for (int lane = 0; lane < 32; ++lane)
    for (int n1 = 0; n1 < 16; ++n1) stage1_task(skew + (kHop + 1) * lane, n1, zbuf + lane);
  for (int lane = 0; lane < 32; ++lane)
    for (int k2 = 0; k2 < 13; ++k2) stage2_task(zbuf + lane, k2);
  for (int k = 0; k < kBins; ++k) memcpy(power + k * 32, zbuf + bin_to_row(k) * 32, 32 * sizeof(float));
}
}
