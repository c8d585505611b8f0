// Per-pixel colour arithmetic of the graph builder (reference: graph_builder.py:142-154).
//
// The reference evaluates rgb2lab / rgb2hsv in float64 on img/255 and then casts to
// float32; region sums are taken over those float32 values.  To reproduce the float32
// values (and therefore the region means that decide the non-local kNN edges) the same
// quantities are produced here as follows:
//   * sRGB linearisation: 256-entry float64 table (the input is uint8);
//   * XYZ and the cube roots in float64 (B200 runs FP64 at half the FP32 rate);
//   * HSV in float32 from the integer channels.  v = max/255, s = delta/max and
//     h = P/(6 delta) are rationals with denominators <= 1530; a float32 division of the
//     exact integers is the correctly rounded value, and the reference's float64 result
//     (relative error < 2^-43) can never sit closer than 2^-35 to a float32 rounding
//     boundary, so both round identically.  Verified exhaustively over all 2^24 colours
//     (tests/test_pixel_math_host.py for the host build, tests/test_gpu_parity.py on GPU).
#pragma once

#include <math.h>
#include <stdint.h>

#ifndef GG_HD
#define GG_HD __host__ __device__ __forceinline__
#endif

#ifdef __CUDA_ARCH__
#define GG_FDIV(a, b) __fdiv_rn((a), (b))
#define GG_FMUL(a, b) __fmul_rn((a), (b))
#define GG_FADD(a, b) __fadd_rn((a), (b))
#define GG_FSUB(a, b) __fsub_rn((a), (b))
#define GG_FSQRT(a) __fsqrt_rn((a))
#define GG_DFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define GG_FDIV(a, b) ((float)(a) / (float)(b))
#define GG_FMUL(a, b) ((float)(a) * (float)(b))
#define GG_FADD(a, b) ((float)(a) + (float)(b))
#define GG_FSUB(a, b) ((float)(a) - (float)(b))
#define GG_FSQRT(a) sqrtf((a))
#define GG_DFMA(a, b, c) fma((a), (b), (c))
#endif

namespace gg {

// cv2.cvtColor(BGR2GRAY) on uint8: 15-bit fixed point, round half up.
GG_HD int gray_u8(int b, int g, int r) { return (9798 * r + 19235 * g + 3735 * b + 16384) >> 15; }

// skimage xyz_from_rgb rows divided by the D65/2deg white point (0.95047, 1, 1.08883).
struct LabMatrix {
  double m[9];
};

static inline LabMatrix make_lab_matrix() {
  const double M[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160,
                       0.072169, 0.019334, 0.119193, 0.950227};
  const double white[3] = {0.95047, 1.0, 1.08883};
  LabMatrix o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.m[3 * i + j] = M[3 * i + j] / white[i];
  return o;
}

// sRGB -> linear, the table entry for channel value v (host side, float64).
static inline double srgb_linear(int v) {
  double a = (double)v / 255.0;
  return a > 0.04045 ? pow((a + 0.055) / 1.055, 2.4) : a / 12.92;
}

GG_HD double lab_f(double t) {
  return t > 0.008856 ? cbrt(t) : GG_DFMA(7.787, t, 16.0 / 116.0);
}

// lin: 256-entry linearisation table, mat: LabMatrix::m
GG_HD void bgr_to_lab(const double* __restrict__ lin, const double* __restrict__ mat, int b, int g,
                      int r, float& L, float& A, float& B) {
  const double lr = lin[r], lg = lin[g], lb = lin[b];
  const double x = GG_DFMA(mat[2], lb, GG_DFMA(mat[1], lg, mat[0] * lr));
  const double y = GG_DFMA(mat[5], lb, GG_DFMA(mat[4], lg, mat[3] * lr));
  const double z = GG_DFMA(mat[8], lb, GG_DFMA(mat[7], lg, mat[6] * lr));
  const double fx = lab_f(x), fy = lab_f(y), fz = lab_f(z);
  L = (float)(116.0 * fy - 16.0);
  A = (float)(500.0 * (fx - fy));
  B = (float)(200.0 * (fy - fz));
}

GG_HD void bgr_to_hsv(int b, int g, int r, float& h, float& s, float& v) {
  const int mx = max(r, max(g, b)), mn = min(r, min(g, b));
  const int d = mx - mn;
  v = GG_FDIV((float)mx, 255.0f);
  if (d == 0) {
    h = 0.0f;
    s = 0.0f;
    return;
  }
  s = GG_FDIV((float)d, (float)mx);
  int p;
  if (b == mx) p = 4 * d + (r - g);        // "blue is max" overwrites (skimage order R, G, B)
  else if (g == mx) p = 2 * d + (b - r);
  else { p = g - b; if (p < 0) p += 6 * d; }  // python-style (h/6) % 1
  h = GG_FDIV((float)p, (float)(6 * d));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Device fast paths.  They return the same correctly rounded float32 values as the IEEE
// intrinsics on the domains they are used on, with a third of the instructions and no
// slow-path branches; gg_selftest_math checks them against __fsqrt_rn / __fdiv_rn on the GPU
// (exhaustively where the domain is finite) and the all-colours parity test covers Lab / HSV.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// num / den for small non-negative integers held in floats (den >= 1): approximate reciprocal,
// one exact residual (FMA) and one correction.  The quotients are rationals with denominators
// <= 1530, never closer than 2^-35 (relative) to a float32 rounding boundary, and the corrected
// value is within 2^-45 of the true quotient, so the final rounding is the correct one.
__device__ __forceinline__ float fdiv_small(float num, float den) {
  const float r = rcp_approx(den);
  const float q = __fmul_rn(num, r);
  const float e = __fmaf_rn(-q, den, num);
  return __fmaf_rn(e, r, q);
}

// num / den with rden = RN(1/den) precomputed (Markstein's sequence); num >= 0, den > 0 normal.
__device__ __forceinline__ float fdiv_rcp(float num, float den, float rden) {
  const float q = __fmul_rn(num, rden);
  const float e = __fmaf_rn(-q, den, num);
  return __fmaf_rn(e, rden, q);
}

// sqrt of a non-negative integer < 2^24 held in a float (Sobel magnitude): rsqrt seed and one
// Newton correction with an exact residual.  sqrt(0) = 0 through the clamped seed.
__device__ __forceinline__ float fsqrt_int(float a) {
  const float r = rsqrt_approx(fmaxf(a, 1.0f));
  const float s = __fmul_rn(a, r);
  const float h = __fmul_rn(0.5f, r);
  const float e = __fmaf_rn(-s, s, a);
  return __fmaf_rn(e, h, s);
}

// cbrt(t) for t in (0.008, 2): float32 seed of t^(-1/3) (lg2 / ex2), one third-order correction
// in float64 (relative error ~e^3 with e ~ 1e-6), then t * r * r: ~1 ulp of float64, i.e. the
// same float32 after rounding as the reference's libm cbrt except with probability ~1e-8.
__device__ __forceinline__ double cbrt_unit(double t) {
  const float tf = (float)t;
  const double r0 = (double)ex2_approx(__fmul_rn(lg2_approx(tf), -0.33333334f));
  const double e = __fma_rn(-__dmul_rn(t, r0), __dmul_rn(r0, r0), 1.0);       // 1 - t r^3
  const double c = __fma_rn(e, 2.0 / 9.0, 1.0 / 3.0);
  const double r = __fma_rn(__dmul_rn(r0, e), c, r0);                        // r (1 + e/3 + 2e^2/9)
  return __dmul_rn(__dmul_rn(t, r), r);
}

// bgr_to_lab with one branch for "some channel above the linear segment" instead of three;
// float64 L, a, b before the final rounding to float32.
__device__ __forceinline__ void bgr_to_lab_fast_f64(const double* __restrict__ lin, const double* __restrict__ mat,
                                                    int b, int g, int r, double& L, double& A, double& B) {
  const double lr = lin[r], lg = lin[g], lb = lin[b];
  const double x = GG_DFMA(mat[2], lb, GG_DFMA(mat[1], lg, mat[0] * lr));
  const double y = GG_DFMA(mat[5], lb, GG_DFMA(mat[4], lg, mat[3] * lr));
  const double z = GG_DFMA(mat[8], lb, GG_DFMA(mat[7], lg, mat[6] * lr));
  double fx = GG_DFMA(7.787, x, 16.0 / 116.0);
  double fy = GG_DFMA(7.787, y, 16.0 / 116.0);
  double fz = GG_DFMA(7.787, z, 16.0 / 116.0);
  if ((x > 0.008856) | (y > 0.008856) | (z > 0.008856)) {
    if (x > 0.008856) fx = cbrt_unit(x);
    if (y > 0.008856) fy = cbrt_unit(y);
    if (z > 0.008856) fz = cbrt_unit(z);
  }
  L = 116.0 * fy - 16.0;
  A = 500.0 * (fx - fy);
  B = 200.0 * (fy - fz);
}

__device__ __forceinline__ void bgr_to_lab_fast(const double* __restrict__ lin, const double* __restrict__ mat,
                                                int b, int g, int r, float& L, float& A, float& B) {
  double Ld, Ad, Bd;
  bgr_to_lab_fast_f64(lin, mat, b, g, r, Ld, Ad, Bd);
  L = (float)Ld;
  A = (float)Ad;
  B = (float)Bd;
}

// hue and saturation of bgr_to_hsv, branch-free (d == 0 gives 0 / 1 = 0 for both).
__device__ __forceinline__ void hsv_hs_fast(int b, int g, int r, int mx, int mn, float& h, float& s) {
  const int d = mx - mn;
  int pr_ = g - b;
  pr_ += pr_ < 0 ? 6 * d : 0;                   // python-style (h/6) % 1 on the red sector
  const int pg_ = 2 * d + (b - r);
  const int pb_ = 4 * d + (r - g);
  const int p = (b == mx) ? pb_ : ((g == mx) ? pg_ : pr_);   // "blue is max" overwrites (skimage order)
  s = fdiv_small((float)d, (float)max(mx, 1));
  h = fdiv_small((float)p, (float)max(6 * d, 1));
}
#endif  // __CUDACC__

// BORDER_REFLECT_101 index (cv2 default for Sobel / blur): -1 -> 1, n -> n-2.
GG_HD int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
  return i;
}

}  // namespace gg
