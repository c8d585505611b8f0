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

// BORDER_REFLECT_101 index (cv2 default for Sobel / blur): -1 -> 1, n -> n-2.
GG_HD int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
  return i;
}

}  // namespace gg
