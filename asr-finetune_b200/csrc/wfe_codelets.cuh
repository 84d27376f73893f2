// FFT codelets of the log-mel frontend: 400 = 16 x 25 Cooley-Tukey, one frame per thread column.
// Compiles as device code under nvcc and as plain C++ under g++ (tests/test_codelets_cpu.py runs the very
// same arithmetic on the host against numpy's FFT before any GPU time is spent).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define WFE_DEV __device__ __forceinline__
#define WFE_CONST __constant__
#define WFE_DEVCONST __device__ constexpr
#else
#define WFE_DEV static inline
#define WFE_CONST static
#define WFE_DEVCONST static constexpr
#define __restrict__
struct float2 { float x, y; };
#endif

namespace wfe {

constexpr int kNFft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kTileF = 32;                                // frames per tile (= warp width)

constexpr float kLog10_2 = 0.30102999566398120f;

WFE_CONST float c_win[kNFft];          // periodic Hann, fp64-computed
WFE_CONST float2 c_tw400[16 * 12];     // [n1][k2-1] = (cos, -sin)(2*pi*n1*k2/400)

struct cpx {
  float r, i;
};
WFE_DEV cpx operator+(cpx a, cpx b) { return {a.r + b.r, a.i + b.i}; }
WFE_DEV cpx operator-(cpx a, cpx b) { return {a.r - b.r, a.i - b.i}; }
// a * (wr + i wi)
WFE_DEV cpx cmul(cpx a, float wr, float wi) {
  return {fmaf(a.r, wr, -a.i * wi), fmaf(a.r, wi, a.i * wr)};
}

// DFT-5 constants
constexpr float kC1 = 0.30901699437494745f;   // cos(2pi/5)
constexpr float kC2 = -0.8090169943749473f;   // cos(4pi/5)
constexpr float kS1 = 0.9510565162951535f;    // sin(2pi/5)
constexpr float kS2 = 0.5877852522924732f;    // sin(4pi/5)

// real-input 5-point DFT: A[0] (real), A[1], A[2] (A[3] = conj A[2], A[4] = conj A[1])
WFE_DEV void dft5_real(float u0, float u1, float u2, float u3, float u4, float& a0, cpx& a1,
                                          cpx& a2) {
  const float t1 = u1 + u4, t2 = u2 + u3, t3 = u1 - u4, t4 = u2 - u3;
  a0 = u0 + t1 + t2;
  a1.r = fmaf(kC2, t2, fmaf(kC1, t1, u0));
  a2.r = fmaf(kC1, t2, fmaf(kC2, t1, u0));
  a1.i = -fmaf(kS2, t4, kS1 * t3);
  a2.i = fmaf(kS1, t4, -kS2 * t3);
}

// complex 5-point DFT, forward (e^{-2 pi i nk/5})
WFE_DEV void dft5_cpx(cpx v0, cpx v1, cpx v2, cpx v3, cpx v4, cpx& o0, cpx& o1, cpx& o2, cpx& o3,
                                         cpx& o4) {
  const cpx t1 = v1 + v4, t2 = v2 + v3, t3 = v1 - v4, t4 = v2 - v3;
  o0 = {v0.r + t1.r + t2.r, v0.i + t1.i + t2.i};
  const cpx m1 = {fmaf(kC2, t2.r, fmaf(kC1, t1.r, v0.r)), fmaf(kC2, t2.i, fmaf(kC1, t1.i, v0.i))};
  const cpx m2 = {fmaf(kC1, t2.r, fmaf(kC2, t1.r, v0.r)), fmaf(kC1, t2.i, fmaf(kC2, t1.i, v0.i))};
  const cpx n1 = {fmaf(kS2, t4.r, kS1 * t3.r), fmaf(kS2, t4.i, kS1 * t3.i)};
  const cpx n2 = {fmaf(-kS1, t4.r, kS2 * t3.r), fmaf(-kS1, t4.i, kS2 * t3.i)};
  // o1 = m1 - i n1, o4 = m1 + i n1, o2 = m2 - i n2, o3 = m2 + i n2
  o1 = {m1.r + n1.i, m1.i - n1.r};
  o4 = {m1.r - n1.i, m1.i + n1.r};
  o2 = {m2.r + n2.i, m2.i - n2.r};
  o3 = {m2.r - n2.i, m2.i + n2.r};
}

// W25^m = cos(2 pi m/25) - i sin(2 pi m/25), m = b*c <= 8
WFE_DEVCONST float kW25C[9] = {1.f, 0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f,
                                       0.53582679497899655f, 0.30901699437494745f, 0.062790519529313527f,
                                       -0.1873813145857246f, -0.42577929156507272f};
WFE_DEVCONST float kW25S[9] = {0.f, 0.24868988716485479f, 0.48175367410171532f, 0.68454710592868862f,
                                       0.84432792550201508f, 0.95105651629515353f, 0.99802672842827156f,
                                       0.98228725072868872f, 0.90482705246601947f};
// W16^m, m = n2*k1 <= 9
WFE_DEVCONST float kW16C[10] = {1.f, 0.92387953251128674f, 0.70710678118654757f, 0.38268343236508984f, 0.f,
                                        -0.38268343236508973f, -0.70710678118654746f, -0.92387953251128674f, -1.f,
                                        -0.92387953251128685f};
WFE_DEVCONST float kW16S[10] = {0.f, 0.38268343236508978f, 0.70710678118654746f, 0.92387953251128674f, 1.f,
                                        0.92387953251128674f, 0.70710678118654757f, 0.38268343236508989f, 0.f,
                                        -0.38268343236508967f};

// ---- stage 1: window + real DFT-25 + W400 twiddle for one n1, lane = frame -----------------------
// sig_lane = sig + 161*lane (161 = 160 samples + 1 skew word per hop); zcol = zbuf + lane.
WFE_DEV void stage1_task(const float* __restrict__ sig_lane, int n1, float* __restrict__ zcol) {
  float x[25];
#pragma unroll
  for (int n2 = 0; n2 < 25; ++n2) x[n2] = sig_lane[n1 + 16 * n2 + n2 / 10] * c_win[n1 + 16 * n2];

  float a0[5];
  cpx a1[5], a2[5];
#pragma unroll
  for (int b = 0; b < 5; ++b) dft5_real(x[b], x[5 + b], x[10 + b], x[15 + b], x[20 + b], a0[b], a1[b], a2[b]);
#pragma unroll
  for (int b = 1; b < 5; ++b) {
    a1[b] = cmul(a1[b], kW25C[b], -kW25S[b]);
    a2[b] = cmul(a2[b], kW25C[2 * b], -kW25S[2 * b]);
  }
  cpx y[13];
  cpx y16, y21, y17, y22;
  {
    float y0;
    dft5_real(a0[0], a0[1], a0[2], a0[3], a0[4], y0, y[5], y[10]);
    y[0] = {y0, 0.f};
  }
  dft5_cpx(a1[0], a1[1], a1[2], a1[3], a1[4], y[1], y[6], y[11], y16, y21);
  dft5_cpx(a2[0], a2[1], a2[2], a2[3], a2[4], y[2], y[7], y[12], y17, y22);
  y[9] = {y16.r, -y16.i};
  y[4] = {y21.r, -y21.i};
  y[8] = {y17.r, -y17.i};
  y[3] = {y22.r, -y22.i};

  zcol[n1 * kTileF] = y[0].r;
#pragma unroll
  for (int k2 = 1; k2 < 13; ++k2) {
    const float2 w = c_tw400[n1 * 12 + (k2 - 1)];
    const cpx z = cmul(y[k2], w.x, w.y);
    const int row = 16 + (k2 - 1) * 32 + 2 * n1;
    zcol[row * kTileF] = z.r;
    zcol[(row + 1) * kTileF] = z.i;
  }
}

WFE_DEV void dft4(cpx a0, cpx a1, cpx a2, cpx a3, cpx& o0, cpx& o1, cpx& o2, cpx& o3) {
  const cpx s0 = a0 + a2, s1 = a0 - a2, s2 = a1 + a3, s3 = a1 - a3;
  o0 = s0 + s2;
  o2 = s0 - s2;
  o1 = {s1.r + s3.i, s1.i - s3.r};
  o3 = {s1.r - s3.i, s1.i + s3.r};
}

// ---- stage 2: complex DFT-16 over n1 for one k2, power, stored in place ---------------------------
WFE_DEV void stage2_task(float* __restrict__ zcol, int k2) {
  cpx z[16];
  const int base = (k2 == 0) ? 0 : 16 + (k2 - 1) * 32;
  if (k2 == 0) {
#pragma unroll
    for (int n = 0; n < 16; ++n) z[n] = {zcol[n * kTileF], 0.f};
  } else {
#pragma unroll
    for (int n = 0; n < 16; ++n) z[n] = {zcol[(base + 2 * n) * kTileF], zcol[(base + 2 * n + 1) * kTileF]};
  }
  cpx g[4][4];
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(z[n2], z[4 + n2], z[8 + n2], z[12 + n2], g[n2][0], g[n2][1], g[n2][2], g[n2][3]);
#pragma unroll
  for (int n2 = 1; n2 < 4; ++n2)
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1) {
      const int m = n2 * k1;
      if (m == 4)
        g[n2][k1] = {g[n2][k1].i, -g[n2][k1].r};   // * (-i)
      else
        g[n2][k1] = cmul(g[n2][k1], kW16C[m], -kW16S[m]);
    }
  cpx X[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(g[0][k1], g[1][k1], g[2][k1], g[3][k1], X[k1], X[k1 + 4], X[k1 + 8], X[k1 + 12]);
  if (k2 == 0) {
#pragma unroll
    for (int k1 = 0; k1 < 9; ++k1) zcol[k1 * kTileF] = fmaf(X[k1].r, X[k1].r, X[k1].i * X[k1].i);
  } else {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) zcol[(base + k1) * kTileF] = fmaf(X[k1].r, X[k1].r, X[k1].i * X[k1].i);
  }
}


// bin k (0..200) -> row of the in-place power buffer written by stage2_task (host side: mel table build)
inline int bin_to_row(int k) {
  int k2 = k % 25, k1 = k / 25;
  if (k2 == 0) return k1;
  if (k2 <= 12) return 16 + (k2 - 1) * 32 + k1;
  k2 = 25 - k2;
  k1 = 15 - k1;
  return 16 + (k2 - 1) * 32 + k1;
}

// fp64-computed window and W400 twiddles, rounded once to fp32 (host side)
inline void fill_tables(float* win /*400*/, float2* tw /*16*12*/) {
  const double kPi = 3.14159265358979323846;
  for (int n = 0; n < kNFft; ++n) win[n] = (float)(0.5 - 0.5 * cos(2.0 * kPi * n / kNFft));
  for (int n1 = 0; n1 < 16; ++n1)
    for (int k2 = 1; k2 <= 12; ++k2) {
      const double a = 2.0 * kPi * (double)(n1 * k2) / 400.0;
      tw[n1 * 12 + (k2 - 1)].x = (float)cos(a);
      tw[n1 * 12 + (k2 - 1)].y = (float)(-sin(a));
    }
}

}  // namespace wfe
