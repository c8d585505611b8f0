// SLIC superpixels on the device -- the input producer in front of the trimap path
// (GraphBuilder._compute_superpixels, graph_builder.py:177-188: skimage.segmentation.slic on the
// CIELAB image, compactness 10, sigma 1, 10 iterations, connectivity enforced).
//
// The algorithm follows scikit-image's (oracle/slic_port.py restates it, with citations):
//   k_slic_lab_minmax    float32 CIELAB of the image, global min / max over all three channels
//   k_slic_features      rescale to [0,1], separable Gaussian (scipy.ndimage "reflect" border), the
//                        second rgb2lab that skimage applies to any 3-channel input, x 1/compactness
//   k_slic_init          centres on skimage's regular grid (+ the per-centre record the assignment reads)
//   k_slic_assign        per pixel: the centres of the 3 x 3 (GG_SLIC_NEIGH=2: 5 x 5) grid cells around it,
//                        distance |dc|^2 + |dxy|^2 / step^2.  Column-walk form (default): a thread walks 16 / 8
//                        consecutive rows of one column, its 9 candidates in registers, feature rows staged by
//                        cp.async; tiles with a window that does not cover them, very small grid steps and
//                        NEIGH = 2 take the general form with skimage's window test, lowest index wins ties.
//                        The block accumulates its pixels into per-centre sums (integers: count, y, x,
//                        fixed-point colour -> the result does not depend on the order of the atomics)
//   k_slic_update        centres = means (+ the per-centre record)
//   k_slic_cc_*          connectivity: union-find components of equal label (4-neighbours),
//                        components smaller than half a nominal superpixel join the component of
//                        the pixel above (or left of) their first pixel, consecutive relabelling in
//                        raster order of the components' first pixels -> labels 0..N-1, all used.
// Differences from scikit-image (parity with it is unpinned anyway, no build of it can be run here):
// a pixel looks at the centres that STARTED in the 3 x 3 (GG_SLIC_NEIGH=2: 5 x 5) cells around it
// instead of at every centre whose two-step window reaches it; float32 arithmetic with fused multiply-adds; the merge rule of the
// connectivity pass is order-free instead of the sequential flood fill's "last labelled
// neighbour".  The gate is segmentation quality against the restatement (tests).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "pixel_math.cuh"
#include "slic.cuh"
#include "tc_ptx.cuh"

namespace gg {

GG_D int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
GG_D float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct LabF {
  float m[9];
};
// x^(1/3) for x in (0.008, 2): two special-function instructions (the superpixel features carry no
// bit-level contract; k-means only compares distances)
GG_D float cbrt_fast(float x) { return __powf(x, 0.33333334f); }

// float32 CIELAB of a uint8 pixel (same formulas as pixel_math.cuh, single precision)
GG_D void bgr_to_lab_f32(const float* __restrict__ lin, const LabF& M, int b, int g, int r, float& L, float& A, float& B) {
  const float lr = lin[r], lg = lin[g], lb = lin[b];
  const float x = fmaf(M.m[2], lb, fmaf(M.m[1], lg, M.m[0] * lr));
  const float y = fmaf(M.m[5], lb, fmaf(M.m[4], lg, M.m[3] * lr));
  const float z = fmaf(M.m[8], lb, fmaf(M.m[7], lg, M.m[6] * lr));
  const float fx = x > 0.008856f ? cbrt_fast(x) : fmaf(7.787f, x, 16.0f / 116.0f);
  const float fy = y > 0.008856f ? cbrt_fast(y) : fmaf(7.787f, y, 16.0f / 116.0f);
  const float fz = z > 0.008856f ? cbrt_fast(z) : fmaf(7.787f, z, 16.0f / 116.0f);
  L = fmaf(116.0f, fy, -16.0f);
  A = 500.0f * (fx - fy);
  B = 200.0f * (fy - fz);
}

// rgb2lab of float channels in [0,1] (the second conversion skimage's slic applies)
GG_D void rgbf_to_lab_f32(const LabF& M, float r, float g, float b, float& L, float& A, float& B) {
  auto lin = [](float v) { return v > 0.04045f ? __powf((v + 0.055f) * (1.0f / 1.055f), 2.4f) : v * (1.0f / 12.92f); };
  const float lr = lin(r), lg = lin(g), lb = lin(b);
  const float x = fmaf(M.m[2], lb, fmaf(M.m[1], lg, M.m[0] * lr));
  const float y = fmaf(M.m[5], lb, fmaf(M.m[4], lg, M.m[3] * lr));
  const float z = fmaf(M.m[8], lb, fmaf(M.m[7], lg, M.m[6] * lr));
  const float fx = x > 0.008856f ? cbrt_fast(x) : fmaf(7.787f, x, 16.0f / 116.0f);
  const float fy = y > 0.008856f ? cbrt_fast(y) : fmaf(7.787f, y, 16.0f / 116.0f);
  const float fz = z > 0.008856f ? cbrt_fast(z) : fmaf(7.787f, z, 16.0f / 116.0f);
  L = fmaf(116.0f, fy, -16.0f);
  A = 500.0f * (fx - fy);
  B = 200.0f * (fy - fz);
}

// ---------------------------------------------------------------------------- Lab range
__global__ void k_slic_minmax_init(int* __restrict__ minmax, int B) {      // {+inf, -inf} as ordered ints
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { minmax[2 * b] = 0x7f800000; minmax[2 * b + 1] = (int)(0xff800000u ^ 0x7fffffffu); }
}
__global__ void __launch_bounds__(256)
k_slic_lab_minmax(const uint8_t* __restrict__ bgr, const double* __restrict__ lin_lut, LabF M, int HW,
                  int* __restrict__ minmax /*[B][2] ordered ints*/) {
  __shared__ float s_lin[256];
  __shared__ float sred[32];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lin[i] = (float)lin_lut[i];
  __syncthreads();
  const int b = blockIdx.y;
  const uint8_t* img = bgr + (size_t)b * HW * 3;
  const float INF = __int_as_float(0x7f800000);
  float mn = INF, mx = -INF;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const uint8_t* p = img + (size_t)i * 3;
    float L, A, Bv;
    bgr_to_lab_f32(s_lin, M, p[0], p[1], p[2], L, A, Bv);
    mn = fminf(mn, fminf(L, fminf(A, Bv)));
    mx = fmaxf(mx, fmaxf(L, fmaxf(A, Bv)));
  }
  mn = block_reduce<float>(mn, INF, OpMinF(), sred);
  mx = block_reduce<float>(mx, -INF, OpMaxF(), sred);
  if (threadIdx.x == 0) {
    atomicMin(&minmax[2 * b], f2ord(mn));
    atomicMax(&minmax[2 * b + 1], f2ord(mx));
  }
}

// ---------------------------------------------------------------------------- feature image
constexpr int SF_T = 32;            // output tile
constexpr int SF_MAXR = 8;          // Gaussian radius limit (sigma <= 2)

struct GaussW {
  float w[2 * SF_MAXR + 1];
  int r;
};

GG_D int reflect_sym(int i, int n) {          // scipy.ndimage mode="reflect": d c b a | a b c d | d c b a
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

// RT > 0: Gaussian radius known at compile time (sigma = 1 -> 4): the tile width is a constant, so the
// index decompositions are multiply-shifts instead of runtime integer divisions (they were half of the
// kernel's 1143 warp-instructions per 32 pixels) and the tap loops unroll; RT == 0: runtime radius.
template <int RT>
__global__ void __launch_bounds__(256)
k_slic_features(const uint8_t* __restrict__ bgr, const double* __restrict__ lin_lut, LabF M, GaussW gw,
                const int* __restrict__ minmax, int H, int W, float inv_comp, float4* __restrict__ feat) {
  constexpr int TP = SF_T + 2 * SF_MAXR;
  __shared__ float s_lin[256];
  __shared__ float s_in[3][TP][TP + 1];
  __shared__ float s_v[3][SF_T][TP + 1];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lin[i] = (float)lin_lut[i];
  __syncthreads();
  const int b = blockIdx.z, y0 = blockIdx.y * SF_T, x0 = blockIdx.x * SF_T, r = RT > 0 ? RT : gw.r;
  const uint8_t* img = bgr + (size_t)b * H * W * 3;
  const float mn = ord2f(minmax[2 * b]), rng = ord2f(minmax[2 * b + 1]) - mn;
  const float inv = rng != 0.0f ? 1.0f / rng : 1.0f;
  const int tw = SF_T + 2 * r;
  for (int i = threadIdx.x; i < tw * tw; i += blockDim.x) {
    const int ty = i / tw, tx = i - ty * tw;
    const int y = reflect_sym(y0 + ty - r, H), x = reflect_sym(x0 + tx - r, W);
    const uint8_t* p = img + ((size_t)y * W + x) * 3;
    float L, A, Bv;
    bgr_to_lab_f32(s_lin, M, p[0], p[1], p[2], L, A, Bv);
    s_in[0][ty][tx] = (L - mn) * inv; s_in[1][ty][tx] = (A - mn) * inv; s_in[2][ty][tx] = (Bv - mn) * inv;
  }
  __syncthreads();
  // vertical pass (axis 0 first, like scipy), then horizontal
  for (int i = threadIdx.x; i < 3 * SF_T * tw; i += blockDim.x) {
    const int c = i / (SF_T * tw), rem = i - c * SF_T * tw, ty = rem / tw, tx = rem - ty * tw;
    float s = 0.0f;
    if (RT > 0) {
#pragma unroll
      for (int d = 0; d <= 2 * RT; ++d) s = fmaf(gw.w[d], s_in[c][ty + d][tx], s);
    } else {
      for (int d = -r; d <= r; ++d) s = fmaf(gw.w[d + r], s_in[c][ty + r + d][tx], s);
    }
    s_v[c][ty][tx] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SF_T * SF_T; i += blockDim.x) {
    const int ty = i / SF_T, tx = i - ty * SF_T;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = 0.0f;
      if (RT > 0) {
#pragma unroll
        for (int d = 0; d <= 2 * RT; ++d) s = fmaf(gw.w[d], s_v[c][ty][tx + d], s);
      } else {
        for (int d = -r; d <= r; ++d) s = fmaf(gw.w[d + r], s_v[c][ty][tx + r + d], s);
      }
      v[c] = s;
    }
    float L, A, Bv;
    rgbf_to_lab_f32(M, v[0], v[1], v[2], L, A, Bv);
    feat[((size_t)b * H + y) * W + x] = make_float4(L * inv_comp, A * inv_comp, Bv * inv_comp, 0.0f);
  }
}

// ---------------------------------------------------------------------------- k-means
struct SlicGrid {
  int gy, gx, sy, ty, sx, tx, step;      // centres at (sy + i ty, sx + j tx), i < gy, j < gx
  int general;                           // test switch: the column-walk kernels take their general path for every tile
};
constexpr float SLIC_FIX = 4096.0f;      // fixed-point scale of the colour sums

// Per-centre record the column-walk assignment copies straight into shared memory (two float4):
//   (-2 cy/S, -2 cx/S, -2 L, -2 A), (-2 B, |c|^2, covers, 0)       positions in units of the grid step S
// |f - c|^2 - |f|^2 = |c|^2 - 2 f.c: five FMAs per candidate.  `covers` = 1 when skimage's window of the
// centre (+- 2 steps around its CURRENT position) contains every pixel that can consider it (the home cells
// within one cell of the centre's own cell) -- then the window test can be skipped.
GG_D void slic_centre_record(const SlicGrid& g, int H, int W, int k, float cy, float cx, float cl, float ca, float cb,
                             float4* __restrict__ rec) {
  const float inv_step = 1.0f / (float)g.step;
  const float ys = cy * inv_step, xs = cx * inv_step;
  const int ci = k / g.gx, cj = k - ci * g.gx;
  const int wy0 = (int)fmaxf(cy - 2.0f * g.ty, 0.0f), wy1 = (int)fminf(cy + 2.0f * g.ty + 1.0f, (float)H);
  const int wx0 = (int)fmaxf(cx - 2.0f * g.tx, 0.0f), wx1 = (int)fminf(cx + 2.0f * g.tx + 1.0f, (float)W);
  const int ry0 = ci - 1 <= 0 ? 0 : g.sy - g.ty / 2 + (ci - 1) * g.ty;
  const int ry1 = ci + 1 >= g.gy - 1 ? H : g.sy - g.ty / 2 + (ci + 2) * g.ty;
  const int rx0 = cj - 1 <= 0 ? 0 : g.sx - g.tx / 2 + (cj - 1) * g.tx;
  const int rx1 = cj + 1 >= g.gx - 1 ? W : g.sx - g.tx / 2 + (cj + 2) * g.tx;
  const bool covers = wy0 <= ry0 && wy1 >= ry1 && wx0 <= rx0 && wx1 >= rx1;
  rec[0] = make_float4(-2.0f * ys, -2.0f * xs, -2.0f * cl, -2.0f * ca);
  rec[1] = make_float4(-2.0f * cb, fmaf(ys, ys, fmaf(xs, xs, fmaf(cl, cl, fmaf(ca, ca, cb * cb)))), covers ? 1.0f : 0.0f, 0.0f);
}

__global__ void k_slic_init(const float4* __restrict__ feat, SlicGrid g, int H, int W, float* __restrict__ cen /*[B][K][5]*/,
                            int* __restrict__ sums /*[B][K][6]*/, float4* __restrict__ cq /*[B][K][2]*/) {
  const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = g.gy * g.gx;
  if (k >= K) return;
  const int i = k / g.gx, j = k - i * g.gx;
  const int y = g.sy + i * g.ty, x = g.sx + j * g.tx;
  const float4 f = feat[((size_t)b * H + y) * W + x];
  float* c = cen + ((size_t)b * K + k) * 5;
  c[0] = (float)y; c[1] = (float)x; c[2] = f.x; c[3] = f.y; c[4] = f.z;
  slic_centre_record(g, H, W, k, (float)y, (float)x, f.x, f.y, f.z, cq + ((size_t)b * K + k) * 2);
  int* s = sums + ((size_t)b * K + k) * 6;
#pragma unroll
  for (int q = 0; q < 6; ++q) s[q] = 0;
}

constexpr int SA_TY = 32, SA_TX = 64;    // pixel tile of k_slic_assign (256 threads x 8 pixels)
constexpr int SA_MAXC = 16;              // cells per axis that a tile can touch (tile / step + 5)
constexpr int SA_WALK_SLOTS = 96;        // cells of a tile in the column-walk kernels (small tables: three blocks per SM)

GG_D int slic_cell(int p, int start, int step, int n) {       // home cell of a coordinate (boundaries midway)
  const int num = p - start + step / 2;
  return num < 0 ? 0 : min(num / step, n - 1);
}

// NEIGH = 2: the centres of the 5 x 5 cells around the pixel's home cell (skimage's windows reach two
// steps); NEIGH = 1: 3 x 3 cells (the original SLIC search region).
// TY: rows of the pixel tile (32 or 64).  WALK: the tile's feature rows are staged in (dynamic) shared memory
// by asynchronous copies issued before the centre records are fetched, and the tiles whose windows all cover
// them (practically every tile) take the column-walk path below.
template <int NEIGH, int TY, bool WALK>
__global__ void __launch_bounds__(256, WALK ? 3 : 4)
k_slic_assign(const float4* __restrict__ feat, const float* __restrict__ cen, const float4* __restrict__ cq, SlicGrid g,
              int H, int W, int32_t* __restrict__ labels, int* __restrict__ sums) {
  extern __shared__ __align__(128) float4 s_feat[];      // WALK: [TY][SA_TX] features of the tile
  constexpr int SLOTS = WALK ? SA_WALK_SLOTS : SA_MAXC * SA_MAXC;
  __shared__ __align__(128) float4 s_q[WALK ? SLOTS * 2 : 2];   // WALK: the centre records of the tile's cells
  __shared__ float2 s_yx[SLOTS];       // centre position
  __shared__ float4 s_col[SLOTS];      // centre colour
  __shared__ int4 s_win[SLOTS];        // skimage's window of the centre: y_min, y_max, x_min, x_max
  // fast path: |f - c|^2 - |f|^2 = |c|^2 - 2 f.c  ->  (-2cy, -2cx, -2L, -2A), (-2B, |c|^2): five FMAs per centre
  __shared__ float4 s_q4[SLOTS];
  __shared__ float2 s_q2[SLOTS];
  __shared__ int s_sum[SLOTS][6];
  __shared__ int s_rowcell[TY], s_colcell[SA_TX];
  __shared__ int s_cid[SLOTS];         // global centre index of a slot (the label a pixel gets)
  const int b = blockIdx.z, y0 = blockIdx.y * TY, x0 = blockIdx.x * SA_TX;
  const int K = g.gy * g.gx;
  const int ci0 = max(slic_cell(y0, g.sy, g.ty, g.gy) - NEIGH, 0);
  const int ci1 = min(slic_cell(min(y0 + TY, H) - 1, g.sy, g.ty, g.gy) + NEIGH, g.gy - 1);
  const int cj0 = max(slic_cell(x0, g.sx, g.tx, g.gx) - NEIGH, 0);
  const int cj1 = min(slic_cell(min(x0 + SA_TX, W) - 1, g.sx, g.tx, g.gx) + NEIGH, g.gx - 1);
  const int nci = ci1 - ci0 + 1, ncj = cj1 - cj0 + 1;          // <= SA_MAXC by the launch check
  const int tx = threadIdx.x & 63, ty4 = threadIdx.x >> 6;      // 64 columns x 4 row groups
  const int x = x0 + tx;
  if (WALK && x < W) {
    // every thread fetches the feature rows it is going to walk itself (its column, TY / 4 consecutive rows)
    // with 16-byte asynchronous copies into shared memory: all of them are in flight while the centre records
    // are fetched, no register is held for them, and a thread only ever reads what it copied -- no barrier
    const float4* src = feat + ((size_t)b * H + y0 + ty4 * (TY / 4)) * W + x;
    const uint32_t dst = smem_u32(s_feat + (ty4 * (TY / 4)) * SA_TX + tx);
    const int rows = min(TY / 4, H - (y0 + ty4 * (TY / 4)));
#pragma unroll 4
    for (int r = 0; r < rows; ++r)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)r * SA_TX * 16u), "l"(src + (size_t)r * W) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const float* cb = cen + (size_t)b * K * 5;
  const int y_hi = min(y0 + TY, H), x_hi = min(x0 + SA_TX, W);
  const float inv_step = 1.0f / (float)g.step;
  // tables of the general path (and of the straight-line search of the row-interleaved form): centre
  // positions / colours / windows per slot; returns whether every window covers all the pixels of the tile
  // that consider its centre
  auto slot_tables = [&]() {
    bool full = true;
    for (int i = threadIdx.x; i < nci * ncj; i += blockDim.x) {
      const int k = (ci0 + i / ncj) * g.gx + cj0 + i % ncj;
      const float cy = cb[k * 5], cx = cb[k * 5 + 1];
      s_cid[i] = k;
      s_yx[i] = NEIGH == 1 ? make_float2(cy * inv_step, cx * inv_step) : make_float2(cy, cx);
      s_col[i] = make_float4(cb[k * 5 + 2], cb[k * 5 + 3], cb[k * 5 + 4], 0.0f);
      if (NEIGH == 1) {
        const float ys = cy * inv_step, xs = cx * inv_step, cl = cb[k * 5 + 2], ca = cb[k * 5 + 3], cbb = cb[k * 5 + 4];
        s_q4[i] = make_float4(-2.0f * ys, -2.0f * xs, -2.0f * cl, -2.0f * ca);
        s_q2[i] = make_float2(-2.0f * cbb, fmaf(ys, ys, fmaf(xs, xs, fmaf(cl, cl, fmaf(ca, ca, cbb * cbb)))));
      }
      const int4 w = make_int4((int)fmaxf(cy - 2.0f * g.ty, 0.0f), (int)fminf(cy + 2.0f * g.ty + 1.0f, (float)H),
                               (int)fmaxf(cx - 2.0f * g.tx, 0.0f), (int)fminf(cx + 2.0f * g.tx + 1.0f, (float)W));
      s_win[i] = w;
      // the pixels of this tile that can see centre k are those whose home cell is within one cell of
      // k's cell: rows [begin(i-1), end(i+1)) x columns [begin(j-1), end(j+1)), cut to the tile
      const int ci = ci0 + i / ncj, cj = cj0 + i % ncj;
      const int ry0 = max(ci - 1 <= 0 ? 0 : g.sy - g.ty / 2 + (ci - 1) * g.ty, y0);
      const int ry1 = min(ci + 1 >= g.gy - 1 ? H : g.sy - g.ty / 2 + (ci + 2) * g.ty, y_hi);
      const int rx0 = max(cj - 1 <= 0 ? 0 : g.sx - g.tx / 2 + (cj - 1) * g.tx, x0);
      const int rx1 = min(cj + 1 >= g.gx - 1 ? W : g.sx - g.tx / 2 + (cj + 2) * g.tx, x_hi);
      if (ry0 < ry1 && rx0 < rx1) full = full && w.x <= ry0 && w.y >= ry1 && w.z <= rx0 && w.w >= rx1;
    }
    return full;
  };
  for (int i = threadIdx.x; i < nci * ncj * 6; i += blockDim.x) (&s_sum[0][0])[i] = 0;
  if (threadIdx.x >= 128 && threadIdx.x < 128 + TY) s_rowcell[threadIdx.x - 128] = slic_cell(y0 + threadIdx.x - 128, g.sy, g.ty, g.gy);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + SA_TX) s_colcell[threadIdx.x - 64] = slic_cell(x0 + threadIdx.x - 64, g.sx, g.tx, g.gx);
  bool all_full;
  if (!WALK) {
    all_full = __syncthreads_and(slot_tables()) != 0;
  } else {
    // nothing to compute per centre: k_slic_init / k_slic_update left a ready-made record
    bool full = true;
    for (int i = threadIdx.x; i < nci * ncj; i += blockDim.x) {
      const int k = (ci0 + i / ncj) * g.gx + cj0 + i % ncj;
      const float4* rec = cq + ((size_t)b * K + k) * 2;
      const float4 r0 = rec[0], r1 = rec[1];
      s_cid[i] = k;
      s_q[2 * i] = r0;
      s_q[2 * i + 1] = r1;
      full = full && r1.z != 0.0f && g.general == 0;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    all_full = __syncthreads_and(full) != 0;
    if (!all_full) {                                     // rare: a centre drifted by more than half a cell
      slot_tables();
      __syncthreads();
      all_full = false;
    }
  }
  const float w_sp = 1.0f / (float)(g.step * g.step);
  const int lane = threadIdx.x & 31;
  const int hj = s_colcell[tx];
  const int j_lo = max(hj - NEIGH, cj0), j_hi = min(hj + NEIGH, cj1);
  // NEIGH == 1: the three candidate columns, clamped to the grid (a clamped duplicate is evaluated
  // twice with the same distance and never wins a second time)
  const int jc0 = max(hj - 1, 0) - cj0, jc1 = hj - cj0, jc2 = min(hj + 1, g.gx - 1) - cj0;
  const float fxs = (float)x * inv_step;
  if (WALK && NEIGH == 1 && all_full) {
    // Column walk (the common case: every window covers the tile).  A thread owns 8 CONSECUTIVE rows of
    // one column.  Its three candidate columns are fixed and the candidate rows change only when the home
    // row cell does (warp-uniform, at most twice in 8 rows), so the coefficients of the 9 candidate centres
    // live in registers -- with the column term folded into the constant, a candidate costs 4 FMAs + a
    // compare and two selects, and no shared-memory access.  The per-centre sums are run-length
    // aggregated down the column in registers (a thread flushes when its winner changes) and the runs
    // still open at the end are combined across the warp before they reach shared memory.
    constexpr int RPT = TY / 4;                             // rows per thread
    const int rg = threadIdx.x >> 6;
    const int yb = y0 + rg * RPT;
    const int jc[3] = {jc0, jc1, jc2};
    float qy[9], qL[9], qA[9], qB[9], qc[9];
    int rb0 = 0, rb1 = 0, rb2 = 0;
#pragma unroll
    for (int s = 0; s < 9; ++s) qy[s] = qL[s] = qA[s] = qB[s] = qc[s] = 0.0f;
    int hi_prev = -1;
    int cur = -1, n = 0, sy = 0, sL = 0, sA = 0, sB = 0;
    if (x < W) {                                              // (no warp-collective inside: columns right of a cut tile idle)
      const int rows = min(RPT, H - yb);
      const float4* fp = s_feat + (rg * RPT) * SA_TX + tx;    // running pointers instead of per-row index arithmetic
      const int* rc = s_rowcell + rg * RPT;
      int32_t* lp = labels ? labels + ((size_t)b * H + yb) * W + x : nullptr;
#pragma unroll 1
      for (int r = 0; r < rows; ++r, fp += SA_TX, ++rc) {
        const int y = yb + r;
        const float4 f = *fp;
        const int hi = *rc;
        if (hi != hi_prev) {                                  // warp-uniform
          hi_prev = hi;
          rb0 = (max(hi - 1, 0) - ci0) * ncj; rb1 = (hi - ci0) * ncj; rb2 = (min(hi + 1, g.gy - 1) - ci0) * ncj;
          const int rb[3] = {rb0, rb1, rb2};
#pragma unroll
          for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int slot = rb[a] + jc[c];
              const float4 q4 = s_q[2 * slot];
              const float4 q2 = s_q[2 * slot + 1];
              qy[a * 3 + c] = q4.x;
              qc[a * 3 + c] = fmaf(q4.y, fxs, q2.y);
              qL[a * 3 + c] = q4.z;
              qA[a * 3 + c] = q4.w;
              qB[a * 3 + c] = q2.x;
            }
          }
        }
        const float fys = (float)y * inv_step;
        // The candidate number rides in the four low mantissa bits of its distance, so the minimum over the
        // candidates is a tree of 8 float min instructions that carries the winner along (no compare / select
        // pairs, no slot registers); the distances are compared on their upper 28 bits (2^-19 relative: a
        // hundredth of a pixel on the boundary position).
        float d[9];
#pragma unroll
        for (int s = 0; s < 9; ++s) {
          const float v = fmaf(qy[s], fys, fmaf(qL[s], f.x, fmaf(qA[s], f.y, fmaf(qB[s], f.z, qc[s]))));
          d[s] = __uint_as_float((__float_as_uint(v) & 0xfffffff0u) | (unsigned)s);
        }
        const float m01 = fminf(d[0], d[1]), m23 = fminf(d[2], d[3]), m45 = fminf(d[4], d[5]), m67 = fminf(d[6], d[7]);
        const float mw = fminf(fminf(fminf(m01, m23), fminf(m45, m67)), d[8]);
        const int bi = (int)(__float_as_uint(mw) & 15u);
        const int ba = (bi * 11) >> 5, bc = bi - 3 * ba;        // candidate = 3 * row + column
        const int best_slot = (ba == 0 ? rb0 : ba == 1 ? rb1 : rb2) + (bc == 0 ? jc0 : bc == 1 ? jc1 : jc2);
        if (lp) { *lp = s_cid[best_slot]; lp += W; }
        // a thread whose winner changed hands its finished run over with plain shared-memory atomics (combining
        // the lanes that give up the same centre with match.any / redux first was measured: 4.0 -> 5.5 ms)
        if (best_slot != cur) {
          if (n > 0) {
            atomicAdd(&s_sum[cur][0], n);
            atomicAdd(&s_sum[cur][1], sy);
            atomicAdd(&s_sum[cur][2], x * n);
            atomicAdd(&s_sum[cur][3], sL);
            atomicAdd(&s_sum[cur][4], sA);
            atomicAdd(&s_sum[cur][5], sB);
          }
          cur = best_slot;
          n = sy = sL = sA = sB = 0;
        }
        // round-to-nearest fixed point without the conversion unit: 1.5 * 2^23 + v has the integer v in its
        // low mantissa bits (|v| < 2^22), the same value as __float2int_rn(f * SLIC_FIX)
        n += 1;
        sy += y;
        sL += __float_as_int(fmaf(f.x, SLIC_FIX, 12582912.0f)) - 0x4B400000;
        sA += __float_as_int(fmaf(f.y, SLIC_FIX, 12582912.0f)) - 0x4B400000;
        sB += __float_as_int(fmaf(f.z, SLIC_FIX, 12582912.0f)) - 0x4B400000;
      }
    }
    const unsigned act = __ballot_sync(0xffffffffu, n > 0);
    if (n > 0) {
      const int first = __shfl_sync(act, cur, __ffs(act) - 1);
      const unsigned grp = __all_sync(act, cur == first) ? act : __match_any_sync(act, cur);
      const int v0 = __reduce_add_sync(grp, n);
      const int v1 = __reduce_add_sync(grp, sy);
      const int v2 = __reduce_add_sync(grp, x * n);
      const int v3 = __reduce_add_sync(grp, sL);
      const int v4 = __reduce_add_sync(grp, sA);
      const int v5 = __reduce_add_sync(grp, sB);
      if (lane == __ffs(grp) - 1) {
        atomicAdd(&s_sum[cur][0], v0);
        atomicAdd(&s_sum[cur][1], v1);
        atomicAdd(&s_sum[cur][2], v2);
        atomicAdd(&s_sum[cur][3], v3);
        atomicAdd(&s_sum[cur][4], v4);
        atomicAdd(&s_sum[cur][5], v5);
      }
    }
  } else {
  float4 f_next = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y0 + ty4 < H && x < W) f_next = feat[((size_t)b * H + y0 + ty4) * W + x];
  for (int rr = 0; rr < TY / 4; ++rr) {
    const int y = y0 + ty4 + 4 * rr;
    const bool in = y < H && x < W;
    int best_slot = -1;
    const float4 f = f_next;                                  // loaded one iteration ago
    if (rr + 1 < TY / 4 && y + 4 < H && x < W) f_next = feat[((size_t)b * H + y + 4) * W + x];
    if (in) {
      const int hi = s_rowcell[ty4 + 4 * rr];
      float best = __int_as_float(0x7f7fffff);
      if (NEIGH == 1 && all_full) {
        // straight-line search over the 3 x 3 cells in increasing centre index; ties keep the lower
        // index (strict <), exactly the rule of the general path below.  Positions are in units of the
        // grid step, so the 5-d distance needs no weights.
        const float fys = (float)y * inv_step;
        const int rb[3] = {(max(hi - 1, 0) - ci0) * ncj, (hi - ci0) * ncj, (min(hi + 1, g.gy - 1) - ci0) * ncj};
        const int jc[3] = {jc0, jc1, jc2};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int slot = rb[a] + jc[c];
            const float4 q4 = s_q4[slot];
            const float2 q2 = s_q2[slot];
            const float d = fmaf(q4.x, fys, fmaf(q4.y, fxs, fmaf(q4.z, f.x, fmaf(q4.w, f.y, fmaf(q2.x, f.z, q2.y)))));
            if (d < best) { best = d; best_slot = slot; }
          }
        }
      } else {
      const float ps = NEIGH == 1 ? (float)g.step : 1.0f;         // centre positions are stored in step units for NEIGH == 1
      const float fy = (float)y, fx = (float)x;
      // distance with the spatial term first -- it alone rules out most centres once a near one has
      // been seen -- then skimage's window test, then the colour term; slots are visited in
      // increasing centre index after the home cell, equal distances go to the lower index
      auto consider = [&](int slot, bool home) {
        const float2 c = s_yx[slot];
        const float dy = c.x * ps - fy, dx = c.y * ps - fx;
        float d = (dy * dy + dx * dx) * w_sp;
        if (d > best) return;
        const int4 w = s_win[slot];
        if (y < w.x || y >= w.y || x < w.z || x >= w.w) return;
        const float4 cc = s_col[slot];
        const float d0 = f.x - cc.x, d1 = f.y - cc.y, d2 = f.z - cc.z;
        d += d0 * d0 + d1 * d1 + d2 * d2;
        if (d < best || (d == best && !home && slot < best_slot)) { best = d; best_slot = slot; }
      };
      const int home_slot = (hi - ci0) * ncj + (hj - cj0);
      consider(home_slot, true);                          // nearest first: a small `best` prunes the rest
      for (int i = max(hi - NEIGH, ci0); i <= min(hi + NEIGH, ci1); ++i) {
        const int row = (i - ci0) * ncj - cj0;
        for (int j = j_lo; j <= j_hi; ++j)
          if (row + j != home_slot) consider(row + j, false);
      }
      if (best_slot < 0) best_slot = home_slot;           // no window reaches the pixel (cannot happen on a regular grid)
      }
      if (labels) labels[((size_t)b * H + y) * W + x] = s_cid[best_slot];
    }
    // per-centre sums: the lanes of a warp that chose the same centre are reduced first (one
    // shared-memory atomic per field and distinct centre instead of one per pixel)
    const unsigned act = __ballot_sync(0xffffffffu, in);
    if (in) {
      // a warp covers 32 consecutive pixels of ONE row: mostly one or two centres; the row sum is y * count
      const int first = __shfl_sync(act, best_slot, __ffs(act) - 1);
      const unsigned grp = __all_sync(act, best_slot == first) ? act : __match_any_sync(act, best_slot);
      const int v2 = __reduce_add_sync(grp, x);
      const int v3 = __reduce_add_sync(grp, __float2int_rn(f.x * SLIC_FIX));
      const int v4 = __reduce_add_sync(grp, __float2int_rn(f.y * SLIC_FIX));
      const int v5 = __reduce_add_sync(grp, __float2int_rn(f.z * SLIC_FIX));
      if (lane == __ffs(grp) - 1) {
        const int n = __popc(grp);
        atomicAdd(&s_sum[best_slot][0], n);
        atomicAdd(&s_sum[best_slot][1], y * n);
        atomicAdd(&s_sum[best_slot][2], v2);
        atomicAdd(&s_sum[best_slot][3], v3);
        atomicAdd(&s_sum[best_slot][4], v4);
        atomicAdd(&s_sum[best_slot][5], v5);
      }
    }
  }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nci * ncj; i += blockDim.x) {
    if (s_sum[i][0] == 0) continue;
    const int k = (ci0 + i / ncj) * g.gx + cj0 + i % ncj;
    int* s = sums + ((size_t)b * K + k) * 6;
#pragma unroll
    for (int q = 0; q < 6; ++q) atomicAdd(&s[q], s_sum[i][q]);
  }
}

__global__ void k_slic_update(SlicGrid g, int H, int W, float* __restrict__ cen, int* __restrict__ sums, float4* __restrict__ cq) {
  const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = g.gy * g.gx;
  if (k >= K) return;
  int* s = sums + ((size_t)b * K + k) * 6;
  float* c = cen + ((size_t)b * K + k) * 5;
  const int n = s[0];
  if (n > 0) {                               // an empty cluster keeps its centre (skimage: NaN, never wins again)
    const float inv = 1.0f / (float)n;
    c[0] = (float)s[1] * inv; c[1] = (float)s[2] * inv;
    c[2] = (float)s[3] * inv * (1.0f / SLIC_FIX); c[3] = (float)s[4] * inv * (1.0f / SLIC_FIX);
    c[4] = (float)s[5] * inv * (1.0f / SLIC_FIX);
    slic_centre_record(g, H, W, k, c[0], c[1], c[2], c[3], c[4], cq + ((size_t)b * K + k) * 2);
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) s[q] = 0;
}

// ---------------------------------------------------------------------------- connectivity
GG_D int cc_find(int* L, int i) {
  int r = i;
  while (true) {
    const int pr = reinterpret_cast<volatile int*>(L)[r];
    if (pr == r) break;
    r = pr;
  }
  return r;
}
// find with path halving: every other node on the way is re-pointed at its grandparent.  Only nodes that
// were seen with a parent are written, and only with one of their ancestors, so concurrent unions
// (atomicMin on roots) stay correct; the trees the later finds walk are about half as deep.
GG_D int cc_find_halve(int* L, int i) {
  volatile int* V = L;
  int r = i, p = V[r];
  while (p != r) {
    const int gp = V[p];
    if (gp == p) return p;
    V[r] = gp;
    r = gp;
    p = V[r];
  }
  return r;
}
template <bool HALVE>
GG_D void cc_union(int* L, int a, int b) {
  while (true) {
    a = HALVE ? cc_find_halve(L, a) : cc_find(L, a);
    b = HALVE ? cc_find_halve(L, b) : cc_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// Every pixel first points at the start of its horizontal run inside its 32-pixel segment (one
// ballot, no atomics); a union with the pixel above is issued only by the first pixel of a stretch
// that lies under one and the same upper run, and runs that continue across a segment boundary are
// joined by the segment's first lane -- an order of magnitude fewer atomics than one union per
// pixel and neighbour, and shallow trees for the flatten pass.
__global__ void __launch_bounds__(256)
k_slic_cc_init(const int32_t* __restrict__ labels, int H, int W, int* __restrict__ L, int* __restrict__ size) {
  const int b = blockIdx.y;
  const int HW = H * W;
  const int seg_per_row = (W + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp_id >= (long long)H * seg_per_row) return;
  const int y = (int)(warp_id / seg_per_row), x = (int)(warp_id % seg_per_row) * 32 + lane;
  const int32_t* lab = labels + (size_t)b * HW;
  const bool in = x < W;
  const int i = y * W + min(x, W - 1);
  const int v = in ? lab[i] : -1;
  const int left = __shfl_up_sync(0xffffffffu, v, 1);
  const bool start = lane == 0 || left != v;
  const unsigned starts = __ballot_sync(0xffffffffu, start);
  if (!in) return;
  const int run_lane = 31 - __clz(starts & (0xffffffffu >> (31 - lane)));     // last start at or before this lane
  L[(size_t)b * HW + i] = i - (lane - run_lane);
  size[(size_t)b * HW + i] = 0;
}
// grid (ceil(W / 256), H, B): a block owns 256 consecutive pixels of ONE row (no division per pixel); the left
// and upper-left neighbours come from the lane to the left (lane 0 loads them)
template <bool HALVE>
__global__ void __launch_bounds__(256)
k_slic_cc_merge(const int32_t* __restrict__ labels, int H, int W, int* __restrict__ L) {
  const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W, lane = threadIdx.x & 31;
  const bool in = x < W;
  const int32_t* lab = labels + (size_t)b * HW;
  int* Lb = L + (size_t)b * HW;
  const int i = y * W + min(x, W - 1);
  const int v = lab[i];
  const int up = y > 0 ? lab[i - W] : -1;
  int left = __shfl_up_sync(0xffffffffu, v, 1), upleft = __shfl_up_sync(0xffffffffu, up, 1);
  if (lane == 0) {
    left = x > 0 ? lab[i - 1] : -1;
    upleft = (x > 0 && y > 0) ? lab[i - W - 1] : -1;
  }
  if (!in) return;
  const bool left_same = left == v;                                    // labels are >= 0: -1 never matches
  if (left_same && (x & 31) == 0) cc_union<HALVE>(Lb, i, i - 1);        // the run continues across the segment boundary
  if (up == v) {
    // the pixel to the left already joins this run to the same upper run
    const bool covered = left_same && upleft == v;
    if (!covered) cc_union<HALVE>(Lb, i, i - W);
  }
}
__global__ void __launch_bounds__(256)
k_slic_cc_flatten(int HW, int* __restrict__ L, int* __restrict__ size) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned act = __ballot_sync(0xffffffffu, i < HW);
  if (i >= HW) return;
  int* Lb = L + (size_t)b * HW;
  const int r = cc_find(Lb, i);
  Lb[i] = r;
  // neighbouring pixels mostly share their root: one atomic per distinct root of the warp (a few hundred
  // roots per image collect 150k increments -- per-pixel atomics serialise on them)
  const unsigned grp = __match_any_sync(act, r);
  if ((threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&size[(size_t)b * HW + r], __popc(grp));
}
// roots only: a small component points at the component of the pixel above / left of its first pixel
__global__ void __launch_bounds__(256)
k_slic_cc_target(int H, int W, const int* __restrict__ L, const int* __restrict__ size, int min_size,
                 int* __restrict__ target) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (i >= HW) return;
  const int* Lb = L + (size_t)b * HW;
  int t = -1;                                             // not a root
  if (Lb[i] == i) {
    t = i;                                                // kept
    if (size[(size_t)b * HW + i] < min_size && i > 0) {
      const int y = i / W;
      t = y > 0 ? Lb[i - W] : Lb[i - 1];                  // an earlier component: chains end at a kept one
    }
  }
  target[(size_t)b * HW + i] = t;
}
// kept[i] = 1 for the roots that survive; block counts for the scan
__global__ void __launch_bounds__(1024)
k_slic_cc_count(int HW, const int* __restrict__ target, int* __restrict__ block_cnt, int n_blocks) {
  __shared__ int s_w[32];
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool kept = i < HW && target[(size_t)b * HW + i] == i;
  const unsigned m = __ballot_sync(0xffffffffu, kept);          // kept roots are sparse: one ballot per warp
  if (lane == 0) s_w[wid] = __popc(m);
  __syncthreads();
  if (wid == 0) {
    int v = s_w[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) block_cnt[(size_t)b * n_blocks + blockIdx.x] = v;
  }
}
// ---- four pixels per thread (H W % 4 == 0): the same passes with 128-bit accesses and a quarter of the
// threads; a block of 256 threads covers the same 1024 pixels as a block of the scalar kernels, so the
// per-block counts and the scan over them are shared.
// target of the roots (see k_slic_cc_target) + number of kept roots per block (k_slic_cc_count) in one pass
__global__ void __launch_bounds__(256)
k_slic_cc_target_count_v4(int H, int W, const int* __restrict__ L, const int* __restrict__ size, int min_size,
                          int* __restrict__ target, int* __restrict__ block_cnt, int n_blocks) {
  __shared__ int s_w[8];
  const int b = blockIdx.y, i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int HW = H * W;
  const int* Lb = L + (size_t)b * HW;
  int kept = 0;
  if (i0 < HW) {
    const int4 l4 = *reinterpret_cast<const int4*>(Lb + i0);
    const int lv[4] = {l4.x, l4.y, l4.z, l4.w};
    int tv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + q;
      int t = -1;                                           // not a root
      if (lv[q] == i) {
        t = i;                                              // kept
        if (size[(size_t)b * HW + i] < min_size && i > 0) {
          const int y = i / W;
          t = y > 0 ? Lb[i - W] : Lb[i - 1];                // an earlier component: chains end at a kept one
        }
      }
      tv[q] = t;
      kept += (t == i);
    }
    *reinterpret_cast<int4*>(target + (size_t)b * HW + i0) = make_int4(tv[0], tv[1], tv[2], tv[3]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wsum = __reduce_add_sync(0xffffffffu, kept);
  if (lane == 0) s_w[wid] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int v = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += s_w[q];
    block_cnt[(size_t)b * n_blocks + blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(256)
k_slic_cc_newid_v4(int HW, const int* __restrict__ target, const int* __restrict__ block_cnt, int n_blocks,
                   int* __restrict__ newid) {
  __shared__ int s_w[8];
  const int b = blockIdx.y, i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int4 t4 = make_int4(-1, -1, -1, -1);
  if (i0 < HW) t4 = *reinterpret_cast<const int4*>(target + (size_t)b * HW + i0);
  const int k0 = t4.x == i0, k1 = t4.y == i0 + 1, k2 = t4.z == i0 + 2, k3 = t4.w == i0 + 3;
  const int c = k0 + k1 + k2 + k3;
  const unsigned any = __ballot_sync(0xffffffffu, c != 0);      // kept roots are sparse: most warps hold none
  int inc = c;
  if (any) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
  }
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  if (c != 0) {
    int base = block_cnt[(size_t)b * n_blocks + blockIdx.x] + inc - c;
    for (int q = 0; q < wid; ++q) base += s_w[q];
    int* out = newid + (size_t)b * HW + i0;
    if (k0) out[0] = base;
    base += k0;
    if (k1) out[1] = base;
    base += k1;
    if (k2) out[2] = base;
    base += k2;
    if (k3) out[3] = base;
  }
}
__global__ void __launch_bounds__(256)
k_slic_cc_relabel_v4(int HW, const int* __restrict__ L, const int* __restrict__ target, const int* __restrict__ newid,
                     int32_t* __restrict__ labels) {
  const int b = blockIdx.y, i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= HW) return;
  const size_t o = (size_t)b * HW;
  const int4 l4 = *reinterpret_cast<const int4*>(L + o + i0);
  int r[4] = {l4.x, l4.y, l4.z, l4.w};
  int out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q > 0 && r[q] == l4.x) { out[q] = out[0]; continue; }   // the four pixels mostly share their component
    int rr = r[q];
    for (int guard = 0; guard < HW; ++guard) {            // follow the merge chain to a kept root
      const int t = target[o + rr];
      if (t == rr) break;
      rr = t;
    }
    out[q] = newid[o + rr];
  }
  *reinterpret_cast<int4*>(labels + o + i0) = make_int4(out[0], out[1], out[2], out[3]);
}
__global__ void __launch_bounds__(1024)
k_slic_cc_scan_blocks(int* __restrict__ block_cnt, int n_blocks, int32_t* __restrict__ n_labels) {
  __shared__ int scratch[40];
  const int b = blockIdx.x;
  int* c = block_cnt + (size_t)b * n_blocks;
  int carry = 0;
  for (int base = 0; base < n_blocks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? c[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, scratch, &total);
    if (i < n_blocks) c[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && n_labels) n_labels[b] = carry;
}
__global__ void __launch_bounds__(1024)
k_slic_cc_newid(int HW, const int* __restrict__ target, const int* __restrict__ block_cnt, int n_blocks,
                int* __restrict__ newid) {
  __shared__ int s_w[32];
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool kept = i < HW && target[(size_t)b * HW + i] == i;
  const unsigned m = __ballot_sync(0xffffffffu, kept);
  if (lane == 0) s_w[wid] = __popc(m);
  __syncthreads();
  if (wid == 0) {                                               // exclusive prefix of the 32 warp counts
    const int v = s_w[lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    s_w[lane] = inc - v;
  }
  __syncthreads();
  if (kept)
    newid[(size_t)b * HW + i] = block_cnt[(size_t)b * n_blocks + blockIdx.x] + s_w[wid] + __popc(m & ((1u << lane) - 1u));
}
__global__ void __launch_bounds__(256)
k_slic_cc_relabel(int HW, const int* __restrict__ L, const int* __restrict__ target, const int* __restrict__ newid,
                  int32_t* __restrict__ labels) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const size_t o = (size_t)b * HW;
  int r = L[o + i];
  for (int guard = 0; guard < HW; ++guard) {            // follow the merge chain to a kept root
    const int t = target[o + r];
    if (t == r) break;
    r = t;
  }
  labels[o + i] = newid[o + r];
}

// ---------------------------------------------------------------------------- host side
// skimage.util.regular_grid for the (1, H, W) volume (see oracle/slic_port.py)
static SlicGrid make_grid(int H, int W, int n_segments) {
  double dims[3] = {1.0, (double)H, (double)W};
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return dims[a] < dims[b] || (dims[a] == dims[b] && a < b); });
  double sd[3] = {dims[order[0]], dims[order[1]], dims[order[2]]};
  double space = sd[0] * sd[1] * sd[2];
  double steps[3];
  for (double& s : steps) s = pow(space / n_segments, 1.0 / 3.0);
  if (sd[0] < steps[0] || sd[1] < steps[1] || sd[2] < steps[2]) {
    for (int dim = 0; dim < 3; ++dim) {
      steps[dim] = sd[dim];
      space = 1.0;
      for (int q = dim + 1; q < 3; ++q) space *= sd[q];
      if (dim < 2) {
        const double s = pow(space / n_segments, 1.0 / (3 - dim - 1));
        for (int q = dim + 1; q < 3; ++q) steps[q] = s;
      }
      if (sd[0] >= steps[0] && sd[1] >= steps[1] && sd[2] >= steps[2]) break;
    }
  }
  int start[3], stepi[3];
  for (int q = 0; q < 3; ++q) {
    start[order[q]] = (int)floor(steps[q] / 2.0);
    stepi[order[q]] = std::max(1, (int)nearbyint(steps[q]));      // numpy round: half to even
  }
  SlicGrid g;
  g.sy = start[1]; g.ty = stepi[1]; g.sx = start[2]; g.tx = stepi[2];
  g.gy = (H - g.sy + g.ty - 1) / g.ty; g.gx = (W - g.sx + g.tx - 1) / g.tx;
  g.step = std::max(g.ty, g.tx);
  static const int general = getenv("GG_SLIC_GENERAL") ? atoi(getenv("GG_SLIC_GENERAL")) : 0;
  g.general = general;
  return g;
}

size_t slic_workspace_bytes(int B, int H, int W, int n_segments) {
  const SlicGrid g = make_grid(H, W, n_segments);
  const size_t HW = (size_t)H * W, K = (size_t)g.gy * g.gx;
  const size_t nb = (HW + 1023) / 1024;
  return Arena::padded((size_t)B * HW, 16) + Arena::padded((size_t)B * 2, 4) + Arena::padded((size_t)B * K * 5, 4) +
         Arena::padded((size_t)B * K * 6, 4) + Arena::padded((size_t)B * K * 2, 16) + Arena::padded((size_t)B * HW, 4) * 4 + Arena::padded((size_t)B * nb, 4) + 4096;
}

int slic_nominal_segments(int H, int W, int n_segments) {
  const SlicGrid g = make_grid(H, W, n_segments);
  return g.gy * g.gx;
}

int slic_labels(gg_context* ctx, Arena& ar, const uint8_t* bgr, int B, int H, int W, int n_segments,
                double compactness, double sigma, int max_iter, int32_t* labels, int32_t* n_labels, cudaStream_t st) {
  GG_REQUIRE(B > 0 && H >= 2 && W >= 2 && n_segments >= 1, "slic: bad sizes");
  GG_REQUIRE(compactness > 0 && sigma >= 0 && sigma <= 2.0, "slic: compactness must be > 0, sigma in [0, 2]");
  GG_REQUIRE((long long)H * W < (1ll << 31), "slic: image too large");
  const SlicGrid g = make_grid(H, W, n_segments);
  GG_REQUIRE(g.gy >= 1 && g.gx >= 1, "slic: empty grid");
  GG_REQUIRE(SA_TY / g.ty + 6 <= SA_MAXC && SA_TX / g.tx + 6 <= SA_MAXC,
             "slic: n_segments too large for this image (grid step %d x %d pixels; need >= 11 x 11)", g.ty, g.tx);
  const int HW = H * W, K = g.gy * g.gx;
  const int nb = ceil_div(HW, 1024);
  // search region of a pixel: 1 = the centres of the 3 x 3 cells around it (the 2S x 2S region of the SLIC
  // paper; default: 8.0 ms per 256 images of 320x480 for the 10 rounds), 2 = the 5 x 5 cells that skimage's
  // two-step windows can reach (11.8 ms; the same segmentation quality on the test images)
  static const int neigh = getenv("GG_SLIC_NEIGH") ? atoi(getenv("GG_SLIC_NEIGH")) : 1;
  // 1 (default): column walk (register-resident candidates, tile staged by bulk copies; 64-row tiles when they
  // fit), 32: the same with 32-row tiles only, 0: the row-interleaved form
  static const int column_walk = getenv("GG_SLIC_WALK") ? atoi(getenv("GG_SLIC_WALK")) : 1;
  float4* feat = ar.take<float4>((size_t)B * HW);
  int* minmax = ar.take<int>((size_t)B * 2);
  float* cen = ar.take<float>((size_t)B * K * 5);
  int* sums = ar.take<int>((size_t)B * K * 6);
  float4* cq = ar.take<float4>((size_t)B * K * 2);
  int* L = ar.take<int>((size_t)B * HW);
  int* size = ar.take<int>((size_t)B * HW);
  int* target = ar.take<int>((size_t)B * HW);
  int* newid = ar.take<int>((size_t)B * HW);
  int* block_cnt = ar.take<int>((size_t)B * nb);

  LabF M;
  {
    const LabMatrix Md = make_lab_matrix();
    for (int i = 0; i < 9; ++i) M.m[i] = (float)Md.m[i];
  }
  GaussW gw{};
  gw.r = sigma > 0 ? (int)(4.0 * sigma + 0.5) : 0;
  {
    double w[2 * SF_MAXR + 1], tot = 0.0;
    for (int d = -gw.r; d <= gw.r; ++d) { w[d + gw.r] = sigma > 0 ? exp(-0.5 * d * d / (sigma * sigma)) : 1.0; tot += w[d + gw.r]; }
    for (int d = 0; d <= 2 * gw.r; ++d) gw.w[d] = (float)(w[d] / tot);
  }
  GG_LAUNCH(ctx, k_slic_minmax_init, ceil_div(B, 256), 256, 0, st, minmax, B);
  {
    dim3 grid(std::min(ceil_div(HW, 256), 64), B);
    GG_LAUNCH(ctx, k_slic_lab_minmax, grid, 256, 0, st, bgr, ctx->d_lin, M, HW, minmax);
  }
  {
    dim3 grid(ceil_div(W, SF_T), ceil_div(H, SF_T), B);
    if (gw.r == 4) GG_LAUNCH(ctx, k_slic_features<4>, grid, 256, 0, st, bgr, ctx->d_lin, M, gw, minmax, H, W, (float)(1.0 / compactness), feat);
    else GG_LAUNCH(ctx, k_slic_features<0>, grid, 256, 0, st, bgr, ctx->d_lin, M, gw, minmax, H, W, (float)(1.0 / compactness), feat);
  }
  {
    dim3 grid(ceil_div(K, 256), B);
    GG_LAUNCH(ctx, k_slic_init, grid, 256, 0, st, feat, g, H, W, cen, sums, cq);
    dim3 ga(ceil_div(W, SA_TX), ceil_div(H, SA_TY), B), ga_tall(ceil_div(W, SA_TX), ceil_div(H, 2 * SA_TY), B);
    // 64-row tiles (16 rows per thread: the centre tables and the final flush are amortised over twice the
    // pixels) when the cell table still fits and the image is not left with a mostly empty last tile row
    auto walk_slots = [&](int ty) { return ((ty - 1) / g.ty + 4) * ((SA_TX - 1) / g.tx + 4); };   // cells a tile can touch
    const bool tall = column_walk >= 1 && column_walk != 32 && walk_slots(2 * SA_TY) <= SA_WALK_SLOTS &&
                      (H % (2 * SA_TY) == 0 || H % (2 * SA_TY) > SA_TY || H > 8 * SA_TY);
    const bool walk = column_walk >= 1 && walk_slots(SA_TY) <= SA_WALK_SLOTS;     // else: very small grid steps
    if (column_walk) {
      static bool attr_set[64] = {};
      if (!attr_set[ctx->device & 63]) {       // static + dynamic shared memory exceed 48 KB
        GG_CUDA_OK(cudaFuncSetAttribute(k_slic_assign<1, 2 * SA_TY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(2 * SA_TY * SA_TX * sizeof(float4))));
        GG_CUDA_OK(cudaFuncSetAttribute(k_slic_assign<1, SA_TY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(SA_TY * SA_TX * sizeof(float4))));
        attr_set[ctx->device & 63] = true;
      }
    }
    for (int it = 0; it < max_iter; ++it) {
      // only the labels of the last iteration are read (the centre sums are accumulated inside the kernel)
      int32_t* lab_out = it == max_iter - 1 ? labels : nullptr;
      if (neigh >= 2) {
        GG_LAUNCH(ctx, (k_slic_assign<2, SA_TY, false>), ga, 256, 0, st, feat, cen, cq, g, H, W, lab_out, sums);
      } else if (!walk) {
        GG_LAUNCH(ctx, (k_slic_assign<1, SA_TY, false>), ga, 256, 0, st, feat, cen, cq, g, H, W, lab_out, sums);
      } else if (tall) {
        GG_LAUNCH(ctx, (k_slic_assign<1, 2 * SA_TY, true>), ga_tall, 256, 2 * SA_TY * SA_TX * sizeof(float4), st, feat, cen, cq, g, H, W,
                  lab_out, sums);
      } else {
        GG_LAUNCH(ctx, (k_slic_assign<1, SA_TY, true>), ga, 256, SA_TY * SA_TX * sizeof(float4), st, feat, cen, cq, g, H, W, lab_out, sums);
      }
      GG_LAUNCH(ctx, k_slic_update, grid, 256, 0, st, g, H, W, cen, sums, cq);
    }
  }
  // connectivity (skimage: min_size = int(0.5 * H W / K))
  {
    const int min_size = (int)(0.5 * ((double)HW / (double)K));
    dim3 grid(ceil_div(HW, 256), B), grid1k(nb, B);
    {
      dim3 gi(ceil_div((long long)H * ((W + 31) / 32) * 32, 256), B);
      GG_LAUNCH(ctx, k_slic_cc_init, gi, 256, 0, st, labels, H, W, L, size);
    }
    static const bool cc_halve = getenv("GG_SLIC_CC_HALVE") ? atoi(getenv("GG_SLIC_CC_HALVE")) != 0 : true;
    GG_REQUIRE(H <= 65535 && B <= 65535, "slic: image height / batch too large for the connectivity grid");
    dim3 grid_rows(ceil_div(W, 256), H, B);
    if (cc_halve) GG_LAUNCH(ctx, k_slic_cc_merge<true>, grid_rows, 256, 0, st, labels, H, W, L);
    else GG_LAUNCH(ctx, k_slic_cc_merge<false>, grid_rows, 256, 0, st, labels, H, W, L);
    GG_LAUNCH(ctx, k_slic_cc_flatten, grid, 256, 0, st, HW, L, size);
    static const bool cc_scalar = getenv("GG_SLIC_CC_SCALAR") != nullptr;
    if (HW % 4 == 0 && !cc_scalar) {
      GG_LAUNCH(ctx, k_slic_cc_target_count_v4, grid1k, 256, 0, st, H, W, L, size, min_size, target, block_cnt, nb);
      GG_LAUNCH(ctx, k_slic_cc_scan_blocks, B, 1024, 0, st, block_cnt, nb, n_labels);
      GG_LAUNCH(ctx, k_slic_cc_newid_v4, grid1k, 256, 0, st, HW, target, block_cnt, nb, newid);
      GG_LAUNCH(ctx, k_slic_cc_relabel_v4, grid1k, 256, 0, st, HW, L, target, newid, labels);
    } else {
      GG_LAUNCH(ctx, k_slic_cc_target, grid, 256, 0, st, H, W, L, size, min_size, target);
      GG_LAUNCH(ctx, k_slic_cc_count, grid1k, 1024, 0, st, HW, target, block_cnt, nb);
      GG_LAUNCH(ctx, k_slic_cc_scan_blocks, B, 1024, 0, st, block_cnt, nb, n_labels);
      GG_LAUNCH(ctx, k_slic_cc_newid, grid1k, 1024, 0, st, HW, target, block_cnt, nb, newid);
      GG_LAUNCH(ctx, k_slic_cc_relabel, grid, 256, 0, st, HW, L, target, newid, labels);
    }
  }
  return GG_OK;
}

}  // namespace gg
