// FFT codelets of the log-mel frontend: 400 = 16 x 25 Cooley-Tukey, one frame per thread column.
//
// Every codelet is a template over the lane type T:
//   T = float  one task per thread (scalar FADD/FMUL/FFMA)
//   T = f2     TWO tasks per thread, packed in a 64-bit register pair -> Blackwell FADD2/FMUL2/FFMA2 (f32x2).
//              The FMA pipe does the same work either way (measured: FFMA2 issues at half the FFMA rate, 70 TFLOP/s
//              both, tools/ubench_f32x2.cu); what packing buys is ISSUE SLOTS, which is what bounds this kernel.
//
// Compiles as device code under nvcc and as plain C++ under g++ (tests/test_codelets_cpu.py runs the very same
// arithmetic on the host against numpy's FFT before any GPU time is spent; f2 is then two floats).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define WFE_DEV __device__ __forceinline__
#define WFE_DEVCONST __device__ constexpr
#else
#define WFE_DEV static inline
#define WFE_DEVCONST static constexpr
#ifndef __restrict__
#define __restrict__
#endif
struct float2 {
  float x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#endif

namespace wfe {

constexpr int kNFft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kTileF = 32;  // frames per tile (= warp width)

constexpr float kLog10_2 = 0.30102999566398120f;

// ---- packed pair of floats ---------------------------------------------------------------------
struct f2 {
  float2 v;
};
WFE_DEV f2 mk2(float lo, float hi) { return f2{make_float2(lo, hi)}; }

#if defined(__CUDA_ARCH__)
WFE_DEV f2 operator+(f2 a, f2 b) { return f2{__fadd2_rn(a.v, b.v)}; }
WFE_DEV f2 operator-(f2 a, f2 b) { return f2{__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y))}; }
WFE_DEV f2 operator-(f2 a) { return f2{make_float2(-a.v.x, -a.v.y)}; }
WFE_DEV f2 vmul(f2 a, f2 b) { return f2{__fmul2_rn(a.v, b.v)}; }
WFE_DEV f2 vmul(f2 a, float s) { return f2{__fmul2_rn(a.v, make_float2(s, s))}; }
WFE_DEV f2 vfma(f2 a, f2 b, f2 c) { return f2{__ffma2_rn(a.v, b.v, c.v)}; }
WFE_DEV f2 vfma(f2 a, float s, f2 c) { return f2{__ffma2_rn(a.v, make_float2(s, s), c.v)}; }
#else
WFE_DEV f2 operator+(f2 a, f2 b) { return mk2(a.v.x + b.v.x, a.v.y + b.v.y); }
WFE_DEV f2 operator-(f2 a, f2 b) { return mk2(a.v.x - b.v.x, a.v.y - b.v.y); }
WFE_DEV f2 operator-(f2 a) { return mk2(-a.v.x, -a.v.y); }
WFE_DEV f2 vmul(f2 a, f2 b) { return mk2(a.v.x * b.v.x, a.v.y * b.v.y); }
WFE_DEV f2 vmul(f2 a, float s) { return mk2(a.v.x * s, a.v.y * s); }
WFE_DEV f2 vfma(f2 a, f2 b, f2 c) { return mk2(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y)); }
WFE_DEV f2 vfma(f2 a, float s, f2 c) { return mk2(fmaf(a.v.x, s, c.v.x), fmaf(a.v.y, s, c.v.y)); }
#endif
WFE_DEV float vmul(float a, float s) { return a * s; }
WFE_DEV float vfma(float a, float s, float c) { return fmaf(a, s, c); }

template <class T>
struct cx {
  T r, i;
};
template <class T>
WFE_DEV cx<T> operator+(cx<T> a, cx<T> b) {
  return {a.r + b.r, a.i + b.i};
}
template <class T>
WFE_DEV cx<T> operator-(cx<T> a, cx<T> b) {
  return {a.r - b.r, a.i - b.i};
}
// a * (wr + i wi); W = float (same constant for both packed tasks) or T (one constant per task)
template <class T, class W>
WFE_DEV cx<T> cmul(cx<T> a, W wr, W wi) {
  return {vfma(a.r, wr, -vmul(a.i, wi)), vfma(a.r, wi, vmul(a.i, wr))};
}
template <class T>
WFE_DEV cx<T> conj(cx<T> a) {
  return {a.r, -a.i};
}
// a * (-i)
template <class T>
WFE_DEV cx<T> mul_mi(cx<T> a) {
  return {a.i, -a.r};
}

// DFT-5 constants
constexpr float kC1 = 0.30901699437494745f;  // cos(2pi/5)
constexpr float kC2 = -0.8090169943749473f;  // cos(4pi/5)
constexpr float kS1 = 0.9510565162951535f;   // sin(2pi/5)
constexpr float kS2 = 0.5877852522924732f;   // sin(4pi/5)

// real-input 5-point DFT: A[0] (real), A[1], A[2] (A[3] = conj A[2], A[4] = conj A[1])
template <class T>
WFE_DEV void dft5_real(T u0, T u1, T u2, T u3, T u4, T& a0, cx<T>& a1, cx<T>& a2) {
  const T t1 = u1 + u4, t2 = u2 + u3, t3 = u1 - u4, t4 = u2 - u3;
  a0 = u0 + t1 + t2;
  a1.r = vfma(t2, kC2, vfma(t1, kC1, u0));
  a2.r = vfma(t2, kC1, vfma(t1, kC2, u0));
  a1.i = vfma(t4, -kS2, vmul(t3, -kS1));
  a2.i = vfma(t4, kS1, vmul(t3, -kS2));
}

// complex 5-point DFT, forward (e^{-2 pi i nk/5})
template <class T>
WFE_DEV void dft5_cpx(cx<T> v0, cx<T> v1, cx<T> v2, cx<T> v3, cx<T> v4, cx<T>& o0, cx<T>& o1, cx<T>& o2, cx<T>& o3,
                      cx<T>& o4) {
  const cx<T> t1 = v1 + v4, t2 = v2 + v3, t3 = v1 - v4, t4 = v2 - v3;
  o0 = {v0.r + t1.r + t2.r, v0.i + t1.i + t2.i};
  const cx<T> m1 = {vfma(t2.r, kC2, vfma(t1.r, kC1, v0.r)), vfma(t2.i, kC2, vfma(t1.i, kC1, v0.i))};
  const cx<T> m2 = {vfma(t2.r, kC1, vfma(t1.r, kC2, v0.r)), vfma(t2.i, kC1, vfma(t1.i, kC2, v0.i))};
  const cx<T> n1 = {vfma(t4.r, kS2, vmul(t3.r, kS1)), vfma(t4.i, kS2, vmul(t3.i, kS1))};
  const cx<T> n2 = {vfma(t4.r, -kS1, vmul(t3.r, kS2)), vfma(t4.i, -kS1, vmul(t3.i, kS2))};
  // o1 = m1 - i n1, o4 = m1 + i n1, o2 = m2 - i n2, o3 = m2 + i n2
  o1 = {m1.r + n1.i, m1.i - n1.r};
  o4 = {m1.r - n1.i, m1.i + n1.r};
  o2 = {m2.r + n2.i, m2.i - n2.r};
  o3 = {m2.r - n2.i, m2.i + n2.r};
}

// W25^m = cos(2 pi m/25) - i sin(2 pi m/25), m = b*c <= 8
WFE_DEVCONST float kW25C[9] = {1.f,
                               0.96858316112863108f,
                               0.87630668004386358f,
                               0.72896862742141155f,
                               0.53582679497899655f,
                               0.30901699437494745f,
                               0.062790519529313527f,
                               -0.1873813145857246f,
                               -0.42577929156507272f};
WFE_DEVCONST float kW25S[9] = {0.f,
                               0.24868988716485479f,
                               0.48175367410171532f,
                               0.68454710592868862f,
                               0.84432792550201508f,
                               0.95105651629515353f,
                               0.99802672842827156f,
                               0.98228725072868872f,
                               0.90482705246601947f};
// W16^m, m = n2*k1 <= 9
WFE_DEVCONST float kW16C[10] = {1.f,
                                0.92387953251128674f,
                                0.70710678118654757f,
                                0.38268343236508984f,
                                0.f,
                                -0.38268343236508973f,
                                -0.70710678118654746f,
                                -0.92387953251128674f,
                                -1.f,
                                -0.92387953251128685f};
WFE_DEVCONST float kW16S[10] = {0.f,
                                0.38268343236508978f,
                                0.70710678118654746f,
                                0.92387953251128674f,
                                1.f,
                                0.92387953251128674f,
                                0.70710678118654757f,
                                0.38268343236508989f,
                                0.f,
                                -0.38268343236508967f};

// ---- real 25-point DFT (5 x 5), outputs k2 = 0..12 (the rest follow by conjugate symmetry) ----------
// x[n2] = windowed samples n1 + 16*n2 of one frame (or of two n1 when T = f2).
template <class T>
WFE_DEV void dft25_real(const T (&x)[25], cx<T> (&y)[13]) {
  T a0[5];
  cx<T> a1[5], a2[5];
#pragma unroll
  for (int b = 0; b < 5; ++b) dft5_real<T>(x[b], x[5 + b], x[10 + b], x[15 + b], x[20 + b], a0[b], a1[b], a2[b]);
#pragma unroll
  for (int b = 1; b < 5; ++b) {
    a1[b] = cmul(a1[b], kW25C[b], -kW25S[b]);
    a2[b] = cmul(a2[b], kW25C[2 * b], -kW25S[2 * b]);
  }
  cx<T> y16, y21, y17, y22;
  {
    T y0;
    dft5_real<T>(a0[0], a0[1], a0[2], a0[3], a0[4], y0, y[5], y[10]);
    y[0].r = y0;
    y[0].i = y0;  // imaginary part of bin 0 is identically zero; never read
  }
  dft5_cpx<T>(a1[0], a1[1], a1[2], a1[3], a1[4], y[1], y[6], y[11], y16, y21);
  dft5_cpx<T>(a2[0], a2[1], a2[2], a2[3], a2[4], y[2], y[7], y[12], y17, y22);
  y[9] = conj(y16);
  y[4] = conj(y21);
  y[8] = conj(y17);
  y[3] = conj(y22);
}

template <class T>
WFE_DEV void dft4(cx<T> a0, cx<T> a1, cx<T> a2, cx<T> a3, cx<T>& o0, cx<T>& o1, cx<T>& o2, cx<T>& o3) {
  const cx<T> s0 = a0 + a2, s1 = a0 - a2, s2 = a1 + a3, s3 = a1 - a3;
  o0 = s0 + s2;
  o2 = s0 - s2;
  o1 = {s1.r + s3.i, s1.i - s3.r};
  o3 = {s1.r - s3.i, s1.i + s3.r};
}

// ---- complex 16-point DFT (4 x 4) over n1, then power |X|^2 ---------------------------------------
template <class T>
WFE_DEV void dft16_power(const cx<T> (&z)[16], T (&pw)[16]) {
  cx<T> g[4][4];
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<T>(z[n2], z[4 + n2], z[8 + n2], z[12 + n2], g[n2][0], g[n2][1], g[n2][2], g[n2][3]);
#pragma unroll
  for (int n2 = 1; n2 < 4; ++n2)
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1) {
      const int m = n2 * k1;
      if (m == 4)
        g[n2][k1] = mul_mi(g[n2][k1]);
      else
        g[n2][k1] = cmul(g[n2][k1], kW16C[m], -kW16S[m]);
    }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    cx<T> X0, X1, X2, X3;
    dft4<T>(g[0][k1], g[1][k1], g[2][k1], g[3][k1], X0, X1, X2, X3);
    pw[k1] = vfma(X0.r, X0.r, vmul(X0.i, X0.i));
    pw[k1 + 4] = vfma(X1.r, X1.r, vmul(X1.i, X1.i));
    pw[k1 + 8] = vfma(X2.r, X2.r, vmul(X2.i, X2.i));
    pw[k1 + 12] = vfma(X3.r, X3.r, vmul(X3.i, X3.i));
  }
}

// ---- layouts shared by the kernel, the host harness and the table builders ------------------------------
// z buffer (stage 1 -> stage 2): plane-major, [plane][n1][frame]; plane 0 = Re y[k2=0], then (Re, Im) of k2 = 1..12.
constexpr int kZPlanes = 25;
WFE_DEV int z_plane(int k2, int im) { return k2 == 0 ? 0 : 1 + 2 * (k2 - 1) + im; }
WFE_DEV int z_index(int plane, int n1, int frame) { return (plane * 16 + n1) * kTileF + frame; }
// power buffer (stage 2 -> mel): BIN-major, 201 rows of 32 frames: conflict-free [row][lane] stores in stage 2 and
// conflict-free 64-bit loads of (frame 2p, frame 2p+1) pairs in the mel stage.
constexpr int kPStride = 32;
// bin computed by stage 2 for residue k2 (0..12) and output index k1 (0..15): 25*k1 + k2, folded by conjugate symmetry
inline int stage2_bin(int k2, int k1) {
  const int k = 25 * k1 + k2;
  return k <= 200 ? k : 400 - k;
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

// Per-warp constant block for stage 1 (warp w owns n1 = 2w, 2w+1): 25 float4
//   [0..12]   window pairs: float4 j = (w[n1+16*(2j)], w[n1+1+16*(2j)], w[n1+16*(2j+1)], w[n1+1+16*(2j+1)])
//   [13..24]  W400 twiddles for k2 = 1..12: (cos_n1, cos_n1+1, -sin_n1, -sin_n1+1)
constexpr int kS1ConstVec = 25;
inline void fill_stage1_consts(float* out /* 8 * 25 * 4 */) {
  float win[kNFft];
  float2 tw[16 * 12];
  fill_tables(win, tw);
  for (int w = 0; w < 8; ++w) {
    float* blk = out + w * kS1ConstVec * 4;
    const int n1 = 2 * w;
    for (int j = 0; j < 13; ++j)
      for (int h = 0; h < 2; ++h) {
        const int n2 = 2 * j + h;
        blk[j * 4 + 2 * h + 0] = n2 < 25 ? win[n1 + 16 * n2] : 0.f;
        blk[j * 4 + 2 * h + 1] = n2 < 25 ? win[n1 + 1 + 16 * n2] : 0.f;
      }
    for (int k2 = 1; k2 <= 12; ++k2) {
      float* t = blk + (13 + k2 - 1) * 4;
      t[0] = tw[n1 * 12 + (k2 - 1)].x;
      t[1] = tw[(n1 + 1) * 12 + (k2 - 1)].x;
      t[2] = tw[n1 * 12 + (k2 - 1)].y;
      t[3] = tw[(n1 + 1) * 12 + (k2 - 1)].y;
    }
  }
}

// ---- stage 1 for one (frame, n1 pair): window, real DFT-25, W400 twiddle, scatter to the z planes ----
// sig_frame: this frame's 400 samples in the skewed tile layout (sample n at n + 2*(n/160)), 8-byte aligned;
// cst: the warp's constant block (25 float4, read as broadcast 128-bit loads); zcol = z + frame.
WFE_DEV void stage1_pair(const float* __restrict__ sig_frame, const float4* __restrict__ cst, int n1,
                         float* __restrict__ zcol) {
  f2 x[25];
#pragma unroll
  for (int j = 0; j < 13; ++j) {
    const float4 c = cst[j];
    {
      const int n2 = 2 * j;
      const float2 s = *reinterpret_cast<const float2*>(sig_frame + n1 + 16 * n2 + 2 * (n2 / 10));
      x[n2] = vmul(f2{s}, mk2(c.x, c.y));
    }
    if (2 * j + 1 < 25) {
      const int n2 = 2 * j + 1;
      const float2 s = *reinterpret_cast<const float2*>(sig_frame + n1 + 16 * n2 + 2 * (n2 / 10));
      x[n2] = vmul(f2{s}, mk2(c.z, c.w));
    }
  }
  cx<f2> y[13];
  dft25_real<f2>(x, y);
  zcol[z_index(0, n1, 0)] = y[0].r.v.x;
  zcol[z_index(0, n1 + 1, 0)] = y[0].r.v.y;
#pragma unroll
  for (int k2 = 1; k2 < 13; ++k2) {
    const float4 t = cst[13 + k2 - 1];
    const cx<f2> zz = cmul(y[k2], mk2(t.x, t.y), mk2(t.z, t.w));
    zcol[z_index(z_plane(k2, 0), n1, 0)] = zz.r.v.x;
    zcol[z_index(z_plane(k2, 0), n1 + 1, 0)] = zz.r.v.y;
    zcol[z_index(z_plane(k2, 1), n1, 0)] = zz.i.v.x;
    zcol[z_index(z_plane(k2, 1), n1 + 1, 0)] = zz.i.v.y;
  }
}

// ---- stage 2 for one (frame, k2 pair (a, a+1)), a odd in 1..11: DFT-16 over n1, power -------------------
// split in two so that the power can be written over the z buffer once every warp has finished reading it
WFE_DEV void stage2_pair_compute(const float* __restrict__ zcol, int a, f2 (&pw)[16]) {
  cx<f2> z[16];
  const float* ra = zcol + z_index(z_plane(a, 0), 0, 0);
  constexpr int kPlane = 16 * kTileF;  // floats per plane
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    z[n].r = mk2(ra[n * kTileF], ra[n * kTileF + 2 * kPlane]);
    z[n].i = mk2(ra[n * kTileF + kPlane], ra[n * kTileF + 3 * kPlane]);
  }
  dft16_power<f2>(z, pw);
}
// bin-major store: bins 25*k1 + a (k1 = 0..7) and 400 - (25*k1 + a) (k1 = 8..15); same with a+1.  pcol = P + frame
WFE_DEV void stage2_pair_store(const f2 (&pw)[16], int a, float* __restrict__ pcol) {
  float* lo = pcol + a * kPStride;
  float* hi = pcol + (400 - a) * kPStride;
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    lo[(25 * k1) * kPStride] = pw[k1].v.x;
    lo[(25 * k1 + 1) * kPStride] = pw[k1].v.y;
  }
#pragma unroll
  for (int k1 = 8; k1 < 16; ++k1) {
    hi[-(25 * k1) * kPStride] = pw[k1].v.x;
    hi[-(25 * k1 + 1) * kPStride] = pw[k1].v.y;
  }
}

// k2 = 0: real input, bins 0, 25, ..., 200 (kept in the .x halves of pw[0..8])
WFE_DEV void stage2_k0_compute(const float* __restrict__ zcol, f2 (&pw)[16]) {
  cx<float> z[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) z[n] = {zcol[z_index(0, n, 0)], 0.f};
  float q[16];
  dft16_power<float>(z, q);
#pragma unroll
  for (int k1 = 0; k1 < 9; ++k1) pw[k1].v.x = q[k1];
}
WFE_DEV void stage2_k0_store(const f2 (&pw)[16], float* __restrict__ pcol) {
#pragma unroll
  for (int k1 = 0; k1 < 9; ++k1) pcol[(25 * k1) * kPStride] = pw[k1].v.x;
}

}  // namespace wfe
