// Graph construction: label map + BGR image -> attributed region graph (batched).
//
// Replaces graph_builder.py:142-154, 190-350, 357-454 of the reference.  Stages:
//   k_gray_gradmax(_v4) grey plane (uint8) + per-image max of the squared Sobel magnitude
//   k_coord_tables      y/H, x/W in the two precisions the reference uses (+ prefix sums), cached
//   k_region_stats      ONE pass over pixels: per-region sums (fp64) + adjacency transitions
//   k_finalize_regions  region sums -> means / std / centroids ...
//   k_adj_sort          adjacency hash table -> (lo,hi)-sorted pair list + shared lengths
//   k_knn_sel(_smem) / k_knn / k_nl_pairs   non-local colour edges
//   k_offsets           ragged offsets (prefix sums over the batch)
//   k_node_features, k_prior_contrast, k_prior_finish, k_edge_attrs, k_csr
//   k_mask_counts, k_region_labels          training labels (dataset.py), k_selftest_math
//
// Float32 epilogues use explicit round-to-nearest intrinsics (never contracted to FMA) so
// that, given identical region sums, every feature is bit-identical to numpy's.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "pixel_math.cuh"
#include "graph_build.cuh"

namespace gg {

constexpr int CSR_LOCAL_ROW = 64;   // longest row that k_csr / k_nl_pairs sort in a thread-local array


// ============================================================================ K0
// Tile of TY x TX pixels (+1 halo): grey values to shared memory, interior written to the
// grey plane, Sobel 3x3 (BORDER_REFLECT_101) squared magnitude max-reduced per image.
constexpr int K0_TY = 16, K0_TX = 128;

__global__ void __launch_bounds__(256)
k_gray_gradmax(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray,
               int* __restrict__ gradmax_sq, int H, int W) {
  __shared__ uint8_t sg[K0_TY + 2][K0_TX + 2];
  __shared__ int sred[32];
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * K0_TY, x0 = blockIdx.x * K0_TX;
  const uint8_t* img = bgr + (size_t)b * H * W * 3;
  for (int i = threadIdx.x; i < (K0_TY + 2) * (K0_TX + 2); i += blockDim.x) {
    const int ty = i / (K0_TX + 2), tx = i - ty * (K0_TX + 2);
    const int y = reflect101(y0 + ty - 1, H), x = reflect101(x0 + tx - 1, W);
    const uint8_t* p = img + ((size_t)y * W + x) * 3;
    sg[ty][tx] = (uint8_t)gray_u8(p[0], p[1], p[2]);
  }
  __syncthreads();
  int m = 0;
  for (int i = threadIdx.x; i < K0_TY * K0_TX; i += blockDim.x) {
    const int ty = i / K0_TX, tx = i - ty * K0_TX;
    const int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) {
      const int a00 = sg[ty][tx], a01 = sg[ty][tx + 1], a02 = sg[ty][tx + 2];
      const int a10 = sg[ty + 1][tx], a11 = sg[ty + 1][tx + 1], a12 = sg[ty + 1][tx + 2];
      const int a20 = sg[ty + 2][tx], a21 = sg[ty + 2][tx + 1], a22 = sg[ty + 2][tx + 2];
      const int gx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
      const int gy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
      m = max(m, gx * gx + gy * gy);
      gray[((size_t)b * H + y) * W + x] = (uint8_t)a11;
    }
  }
  m = block_reduce<int>(m, 0, OpMaxI(), sred);
  if (threadIdx.x == 0 && m > 0) atomicMax(&gradmax_sq[b], m);
}

// Vector form for W % 4 == 0 (every benchmark shape): a thread converts 4 pixels from three
// aligned 32-bit loads, writes the 4 grey bytes with one store, and evaluates the Sobel
// magnitudes of 4 pixels from six 32-bit shared-memory loads (column sums V = a0 + 2 a1 + a2 and
// differences D = a2 - a0 shared between neighbouring pixels).  Tile: 32 x 128 pixels (+1 halo).
constexpr int K0V_TY = 32, K0V_TX = 128, K0V_PITCH = 136;   // tile row: [0] = x0-1, [1..128], [129] = x0+128

__global__ void __launch_bounds__(256)
k_gray_gradmax_v4(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray,
                  int* __restrict__ gradmax_sq, int H, int W) {
  __shared__ __align__(16) uint8_t sg[(K0V_TY + 2) * K0V_PITCH];
  __shared__ int sred[32];
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * K0V_TY, x0 = blockIdx.x * K0V_TX;
  const uint8_t* img = bgr + (size_t)b * H * W * 3;
  uint8_t* gimg = gray + (size_t)b * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- phase 1: grey values of rows y0-1 .. y0+TY (reflected), columns x0-1 .. x0+TX
  for (int ty = warp; ty < K0V_TY + 2; ty += 8) {
    const int yy = y0 + ty - 1;
    const int y = reflect101(yy, H);
    const uint8_t* row = img + (size_t)y * W * 3;
    uint8_t* srow = sg + ty * K0V_PITCH;
    const int x = x0 + 4 * lane;
    if (x < W) {                                         // W % 4 == 0: the group is inside the row
      const uint32_t* p32 = reinterpret_cast<const uint32_t*>(row + (size_t)x * 3);
      const uint32_t w0 = p32[0], w1 = p32[1], w2 = p32[2];   // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
      const int g0 = gray_u8(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255);
      const int g1 = gray_u8(w0 >> 24, w1 & 255, (w1 >> 8) & 255);
      const int g2 = gray_u8((w1 >> 16) & 255, w1 >> 24, w2 & 255);
      const int g3 = gray_u8((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24);
      const uint32_t packed = (uint32_t)g0 | ((uint32_t)g1 << 8) | ((uint32_t)g2 << 16) | ((uint32_t)g3 << 24);
      // tile column of pixel x is 1 + (x - x0): unaligned by one byte -> byte stores are avoided
      // by keeping the tile shifted: bytes [1 + 4 lane, 5 + 4 lane)
      srow[1 + 4 * lane] = (uint8_t)g0; srow[2 + 4 * lane] = (uint8_t)g1;
      srow[3 + 4 * lane] = (uint8_t)g2; srow[4 + 4 * lane] = (uint8_t)g3;
      if (yy >= y0 && yy < min(H, y0 + K0V_TY))
        *reinterpret_cast<uint32_t*>(gimg + (size_t)yy * W + x) = packed;
    } else if (x == W) {                                 // right image border inside the tile: REFLECT_101
      const uint8_t* px = row + (size_t)(W - 2) * 3;
      srow[1 + 4 * lane] = (uint8_t)gray_u8(px[0], px[1], px[2]);
    }
    if (lane < 2) {                                      // the two halo columns
      const int xx = reflect101(lane == 0 ? x0 - 1 : x0 + K0V_TX, W);
      const uint8_t* px = row + (size_t)xx * 3;
      srow[lane == 0 ? 0 : K0V_TX + 1] = (uint8_t)gray_u8(px[0], px[1], px[2]);
    }
  }
  __syncthreads();
  // ---- phase 2: Sobel of 4 pixels per thread; a tile row past the image edge holds the
  // reflected values, columns past W are masked
  int m = 0;
  for (int ty = warp; ty < K0V_TY; ty += 8) {
    const int y = y0 + ty, x = x0 + 4 * lane;
    if (y < H && x < W) {
      int V[6], Dv[6];
      // bytes 4 lane .. 4 lane + 5 of the three tile rows ty, ty+1, ty+2
      const uint8_t* r0 = sg + ty * K0V_PITCH + 4 * lane;
      const uint2 a = make_uint2(*reinterpret_cast<const uint32_t*>(r0), *reinterpret_cast<const uint32_t*>(r0 + 4));
      const uint2 c = make_uint2(*reinterpret_cast<const uint32_t*>(r0 + K0V_PITCH),
                                 *reinterpret_cast<const uint32_t*>(r0 + K0V_PITCH + 4));
      const uint2 e = make_uint2(*reinterpret_cast<const uint32_t*>(r0 + 2 * K0V_PITCH),
                                 *reinterpret_cast<const uint32_t*>(r0 + 2 * K0V_PITCH + 4));
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int a0 = (k < 4 ? a.x >> (8 * k) : a.y >> (8 * (k - 4))) & 255;
        const int a1 = (k < 4 ? c.x >> (8 * k) : c.y >> (8 * (k - 4))) & 255;
        const int a2 = (k < 4 ? e.x >> (8 * k) : e.y >> (8 * (k - 4))) & 255;
        V[k] = a0 + 2 * a1 + a2;
        Dv[k] = a2 - a0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int gx = V[k + 2] - V[k];
        const int gy = Dv[k] + 2 * Dv[k + 1] + Dv[k + 2];
        m = max(m, gx * gx + gy * gy);
      }
    }
  }
  m = block_reduce<int>(m, 0, OpMaxI(), sred);
  if (threadIdx.x == 0 && m > 0) atomicMax(&gradmax_sq[b], m);
}

// ============================================================================ coordinate tables
// tab[0..H)      double(float(y)/float(H))   graph_builder.py:207  (float32 coordinates)
// tab[H..2H)     double(y)/double(H)         graph_builder.py:401  (float64 coordinates)
// tab[2H..2H+W)  double(float(x)/float(W));  tab[2H+W..2H+2W)  double(x)/double(W)
// tab[2H+2W ..)  exclusive prefix sums of the two y tables, H+1 entries each: the sum of y/H over
//                a vertical run [y0, y1) is P[y1] - P[y0]
__global__ void k_coord_tables(double* __restrict__ tab, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H) {
    tab[i] = (double)__fdiv_rn((float)i, (float)H);
    tab[H + i] = (double)i / (double)H;
  }
  if (i < W) {
    tab[2 * H + i] = (double)__fdiv_rn((float)i, (float)W);
    tab[2 * H + W + i] = (double)i / (double)W;
  }
  if (i < 2) {     // two threads, one sequential prefix each (H <= 4096 additions)
    double* P = tab + 2 * (H + W) + i * (H + 1);
    double s = 0.0;
    for (int y = 0; y < H; ++y) {
      P[y] = s;
      s += i == 0 ? (double)__fdiv_rn((float)y, (float)H) : (double)y / (double)H;
    }
    P[H] = s;
  }
}

// ============================================================================ K1
// One warp walks down a strip of 32 columns x `rows` rows, lane = column.  Each lane keeps fp64
// register accumulators for the vertical run of its current label and hands them to a per-warp
// shared-memory table when the label changes (plain read-modify-write, no atomics: 64-bit
// shared atomics are CAS loops on sm_100).  Table slots go to the global per-region
// accumulators with one RED.F64 per field when evicted / at the end of the strip.
//
// Per-region fields (RS_NF doubles): 0-2 sum Lab, 3-5 sum Lab^2, 6-8 sum HSV, 9 sum y/H (f32
// coords), 10 sum x/W (f32 coords), 11 sum |grad|, 12 sum |grad|/(max+1e-6), 13 sum y/H
// (f64 coords), 14 sum x/W (f64 coords), 15 pixel count, 16 boundary pixels, 17 frame pixels
// (the three counters are integers, exact in float64).  Fields that are functions of the run
// geometry alone (9, 10, 13, 14, 15, 17) are produced when a run is handed over, not per pixel.
//
// Everything the pixel loop needs from the neighbouring columns comes from warp shuffles; the two
// edge lanes fetch their outer neighbour one row ahead with one predicated load.  Out-of-image
// neighbours are replaced by the pixel's own label, so that "differs from a neighbour" needs no
// bounds tests.  Labels are only range-checked where they are used as an index (run hand-over,
// pair emission).
constexpr int RS_SLOTS = 16;
constexpr int RS_NF = 18;
constexpr int RS_STAGE_LD = 19;  // doubles per lane in the staging area (odd: conflict-free)
// the direct hand-over needs the two look-up tables only; the table variant adds per-warp slots + staging
constexpr size_t rs_smem_bytes(int warps, bool direct) {
  return (size_t)(256 + 256 + (direct ? 0 : warps * RS_SLOTS * RS_NF + warps * 32 * RS_STAGE_LD)) * sizeof(double);
}

// x rounded to float32 precision (24 significant bits), computed in the FP64 pipe: Veltkamp's
// split with 2^29 + 1.  Same value as (double)(float)x -- round to nearest; an exact tie (the 29
// dropped bits are 1000...0: probability 2^-29 per value) may resolve differently from the
// conversion instruction -- without the two quarter-rate F2F conversions on the XU pipe.
GG_D double round_to_f32(double x) {
  const double t = __dmul_rn(x, 536870913.0);
  return __dadd_rn(t, -__dadd_rn(t, -x));
}

struct RegionStatsParams {
  const uint8_t* bgr;
  const uint8_t* gray;
  const int32_t* labels;
  const int* gradmax_sq;   // [B]
  const double* coord;     // k_coord_tables
  const double* lin_lut;   // [256]
  double* acc;             // [B][node_cap][RS_NF]
  int* label_max;          // [B]
  unsigned long long* pair_keys;  // [B][table_cap]
  int* pair_cnts;                 // [B][table_cap]
  int* status;
  int B, H, W, node_cap, table_cap, connectivity;
  int n_sx, n_sy, rows;
  LabMatrix lab;
};

GG_D uint32_t pair_hash(uint32_t lo, uint32_t hi) {
  uint32_t h = lo * 0x9E3779B1u ^ (hi + 0x7F4A7C15u) * 0x85EBCA6Bu;
  h ^= h >> 15;
  return h;
}

// Insert / increment an undirected pair (lo < hi) in the per-image open-addressing table.
__device__ __noinline__ void pair_emit(unsigned long long* keys, int* cnts, int cap, int lo_, int hi_,
                                       int count, int* status) {
  const uint32_t lo = (uint32_t)lo_, hi = (uint32_t)hi_;
  const unsigned long long key = ((unsigned long long)lo << 32) | hi;
  const unsigned long long EMPTY = ~0ull;
  uint32_t h = pair_hash(lo, hi) & (uint32_t)(cap - 1);
  for (int probe = 0; probe < cap; ++probe) {
    unsigned long long cur = keys[h];
    if (cur == EMPTY) cur = atomicCAS(&keys[h], EMPTY, key);
    if (cur == EMPTY || cur == key) {
      atomicAdd(&cnts[h], count);
      return;
    }
    h = (h + 1) & (uint32_t)(cap - 1);
  }
  atomicOr(status, ST_PAIR_TABLE);
}

// DIRECT: finished runs go straight to the global accumulators (one RED.F64 per field and run)
// instead of through the per-warp shared-memory table: ~14x more L2 atomics, ~80 fewer
// instructions per row.
// PAIRS: the kernel also collects the adjacency transitions (label pairs + shared-boundary counts); with
// PAIRS == false that is k_adjacency_pairs' job (a labels-only pass) and this kernel carries neither the
// pair cache nor the run-aggregation registers (measured by ablation: 0.127 of the 0.71 ms).
template <int RS_WARPS, int MINB, bool DIRECT, bool VELT, bool PAIRS>
__global__ void __launch_bounds__(RS_WARPS * 32, MINB)
k_region_stats(const RegionStatsParams p) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  double* s_lin = reinterpret_cast<double*>(rs_smem);                       // [256] sRGB -> linear
  double* s_vd = s_lin + 256;                                               // [256] (double)float(v/255)
  double* s_vals_all = s_vd + 256;                                          // [W][SLOTS][NF]
  double* s_stage_all = s_vals_all + RS_WARPS * RS_SLOTS * RS_NF;           // [W][32*LD]

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* vals = s_vals_all + (DIRECT ? 0 : wid * RS_SLOTS * RS_NF);
  double* stage = s_stage_all + (DIRECT ? 0 : wid * 32 * RS_STAGE_LD);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_lin[i] = p.lin_lut[i];
    s_vd[i] = (double)__fdiv_rn((float)i, 255.0f);          // HSV value channel: max/255
  }
  if (!DIRECT)
    for (int i = lane; i < RS_SLOTS * RS_NF; i += 32) vals[i] = 0.0;
  __syncthreads();

  const long long task = (long long)blockIdx.x * RS_WARPS + wid;
  const long long n_tasks = (long long)p.B * p.n_sy * p.n_sx;
  if (task >= n_tasks) return;
  const int sx = (int)(task % p.n_sx);
  const int sy = (int)((task / p.n_sx) % p.n_sy);
  const int b = (int)(task / ((long long)p.n_sx * p.n_sy));
  const int H = p.H, W = p.W;
  const int x = sx * 32 + lane;
  const bool valid = x < W;
  const int xc = valid ? x : W - 1;
  const int y_begin = sy * p.rows, y_end = min(H, y_begin + p.rows);

  const uint8_t* img = p.bgr + (size_t)b * H * W * 3;
  const uint8_t* gry = p.gray + (size_t)b * H * W;
  const int32_t* lab = p.labels + (size_t)b * H * W;
  double* acc_g = p.acc + (size_t)b * p.node_cap * RS_NF;
  unsigned long long* keys = p.pair_keys + (size_t)b * p.table_cap;
  int* cnts = p.pair_cnts + (size_t)b * p.table_cap;
  const unsigned node_cap = (unsigned)p.node_cap;

  // loop-invariant lane roles
  const bool is_l0 = lane == 0, is_l31 = lane == 31;
  const bool own_r = x >= W - 1;                       // no pixel to the right
  const bool edge_lane = (is_l0 || is_l31) && valid;
  // outer neighbour of the edge lanes: labels -> own pixel when outside, grey -> REFLECT_101
  const int ldelta = is_l0 ? (x > 0 ? -1 : 0) : (x + 1 < W ? 1 : 0);
  const int gdelta = is_l0 ? (x > 0 ? -1 : (W > 1 ? 1 : 0)) : (x + 1 < W ? 1 : (W > 1 ? -1 : 0));
  const double xtf = p.coord[2 * H + xc], xtd = p.coord[2 * H + W + xc];
  const double* PF = p.coord + 2 * (H + W);            // prefix of y/H, float32 coordinates
  const double* PD = PF + (H + 1);                     // prefix of y/H, float64 coordinates
  const int gm = p.gradmax_sq[b];
  const float gden = __fadd_rn(__fsqrt_rn((float)gm), 1e-6f);  // grad.max() + 1e-6 (float32)
  const float grcp = __frcp_rn(gden);
  const int frame_x = (x == 0) + (x == W - 1);

  // ---- software pipeline.  Iteration y hands over / accumulates row y with the per-pixel values
  // computed one iteration earlier, computes the values of row y+1 (independent instruction
  // chains, interleaved by the scheduler) and issues the loads two rows ahead:
  //   BGR(y+2), labels(y+2), grey(y+3) -- so no load is consumed in the iteration that issues it.
  // VELT: the Lab values stay in float64 registers, rounded to float32 precision inside the FP64
  // pipe (round_to_f32) instead of a D2F + F2D round trip through the quarter-rate XU pipe; their
  // squares (float32 products in the reference) are formed and rounded the same way.
  struct PixelVals { float L, A, Bv, hh, ss, g, gs; int mx; double Ld, Ad, Bd; };
  auto pixel_math = [&](int b_, int g_, int r_, int hs_a, int hd_a, int hd_b, int hs_c_, int hd_c_) {
    PixelVals v;
    if (VELT) {
      bgr_to_lab_fast_f64(s_lin, p.lab.m, b_, g_, r_, v.Ld, v.Ad, v.Bd);
      v.Ld = round_to_f32(v.Ld); v.Ad = round_to_f32(v.Ad); v.Bd = round_to_f32(v.Bd);
      v.L = v.A = v.Bv = 0.0f;
    } else {
      bgr_to_lab_fast(s_lin, p.lab.m, b_, g_, r_, v.L, v.A, v.Bv);
      v.Ld = v.Ad = v.Bd = 0.0;
    }
    const int mx = max(r_, max(g_, b_)), mn = min(r_, min(g_, b_));
    hsv_hs_fast(b_, g_, r_, mx, mn, v.hh, v.ss);
    v.mx = mx;
    const int gx = hd_a + 2 * hd_b + hd_c_;          // rows r-1, r, r+1
    const int gy = hs_c_ - hs_a;
    v.g = fsqrt_int((float)(gx * gx + gy * gy));
    v.gs = fdiv_rcp(v.g, gden, grcp);
    return v;
  };
  // horizontal Sobel partials of a grey row from its centre values and the edge lanes' outer
  // neighbours: hs = g[x-1] + 2 g[x] + g[x+1], hd = g[x+1] - g[x-1]
  auto partials = [&](int gc, int ge, int& hs, int& hd) {
    int gl = __shfl_up_sync(0xffffffffu, gc, 1);
    int gr = __shfl_down_sync(0xffffffffu, gc, 1);
    if (is_l0) gl = ge;
    if (is_l31) gr = ge;
    if (own_r) gr = gl;                                // REFLECT_101 at the last column
    if (W == 1) gl = gr = gc;
    hs = gl + 2 * gc + gr;
    hd = gr - gl;
  };
  auto grey_row_off = [&](int r) -> uint32_t {         // BORDER_REFLECT_101 row, clamped
    r = r < 0 ? -r : r;
    r = r >= H ? 2 * H - 2 - r : r;
    return (uint32_t)(min(max(r, 0), H - 1) * W + xc);
  };
  auto load_grey = [&](int r, int& gc, int& ge) {
    const uint8_t* pgr = gry + grey_row_off(r);
    gc = pgr[0];
    ge = 0;
    if (edge_lane) ge = pgr[gdelta];
  };
  const uint32_t off0 = (uint32_t)(y_begin * W + xc);  // pixel offset of (y_begin, xc) in the image
  int hs_m, hd_m, hs_c, hd_c, g2c, g2e;
  PixelVals cv;
  int pb, pg, pr;
  {
    int gc, ge, hs_a, hd_a, hs_p0, hd_p0;
    load_grey(y_begin - 1, gc, ge); partials(gc, ge, hs_a, hd_a);
    load_grey(y_begin, gc, ge);     partials(gc, ge, hs_m, hd_m);
    load_grey(y_begin + 1, gc, ge); partials(gc, ge, hs_p0, hd_p0);
    load_grey(y_begin + 2, g2c, g2e);
    const uint8_t* px0 = img + (size_t)3 * off0;
    cv = pixel_math(px0[0], px0[1], px0[2], hs_a, hd_a, hd_m, hs_p0, hd_p0);
    hs_c = hs_p0; hd_c = hd_p0;
    const uint8_t* px1 = img + (size_t)3 * (size_t)(min(y_begin + 1, H - 1) * W + xc);
    pb = px1[0]; pg = px1[1]; pr = px1[2];
  }
  int lab_c = valid ? lab[off0] : -1;
  int lab_up = (valid && y_begin > 0) ? lab[off0 - W] : lab_c;
  int lab_dn = (valid && y_begin + 1 < H) ? lab[off0 + W] : lab_c;
  int e_lab = edge_lane ? lab[off0 + ldelta] : 0;       // outer label neighbour of the edge lanes
  int e_lab_n = (edge_lane && y_begin + 1 < H) ? lab[off0 + W + ldelta] : 0;

  // per-lane run accumulators
  int cur = -1, cnt = 0, bnd = 0, ys = y_begin;
  double aL = 0, aA = 0, aB = 0, aL2 = 0, aA2 = 0, aB2 = 0, aH = 0, aS = 0, aV = 0, aG = 0, aGs = 0;
  // run of identical right-neighbour transitions (a | b) down the column
  int rp_a = -1, rp_b = -1, rp_cnt = 0;
  // region table tags (lane s < RS_SLOTS owns the tag of slot s) and the label-pair cache
  // (lane = slot): registers, looked up with shuffles / ballots
  int mytag = -1;
  int pk_lo = -1, pk_hi = -1, pk_n = 0, victim = 0;
  int lmax = -1;

  auto evict_slot = [&](int slot, int tag) {  // warp-uniform arguments
    if (tag >= 0 && lane < RS_NF) {
      const double v = vals[slot * RS_NF + lane];
      if (v != 0.0) atomicAdd(acc_g + (size_t)tag * RS_NF + lane, v);
      vals[slot * RS_NF + lane] = 0.0;
    }
  };

  // one undirected pair (a, b, n), warp-uniform: bump the cached slot or take a slot over
  auto pair_add = [&](int a, int b_, int n) {
    if (!PAIRS) return;
    const int lo = min(a, b_), hi = max(a, b_);
    if (lo == hi || (unsigned)lo >= node_cap || (unsigned)hi >= node_cap) return;
    const unsigned hit = __ballot_sync(0xffffffffu, pk_lo == lo && pk_hi == hi);
    if (hit) {
      if (lane == __ffs(hit) - 1) pk_n += n;
      return;
    }
    const unsigned empty = __ballot_sync(0xffffffffu, pk_lo < 0);
    const int slot = empty ? __ffs(empty) - 1 : victim;
    if (!empty) victim = (victim + 1) & 31;
    if (lane == slot) {
      if (pk_lo >= 0) pair_emit(keys, cnts, p.table_cap, pk_lo, pk_hi, pk_n, p.status);
      pk_lo = lo; pk_hi = hi; pk_n = n;
    }
  };

  // hand the finished runs of the lanes in `mask` (warp-uniform) to the warp table; y_now = first
  // row after the runs; nxt = the label that follows each run (its down-neighbour pair)
  auto flush_lanes = [&](unsigned mask, int y_now, int nxt) {
    const bool mine = mask >> lane & 1u;
    if (DIRECT) {
      if (mine) {
        lmax = max(lmax, cur);
        if ((unsigned)cur < node_cap) {
          double* dst = acc_g + (size_t)cur * RS_NF;
          const double dc = (double)cnt;
          atomicAdd(dst + 0, aL); atomicAdd(dst + 1, aA); atomicAdd(dst + 2, aB);
          atomicAdd(dst + 3, aL2); atomicAdd(dst + 4, aA2); atomicAdd(dst + 5, aB2);
          atomicAdd(dst + 6, aH); atomicAdd(dst + 7, aS); atomicAdd(dst + 8, aV);
          atomicAdd(dst + 9, PF[y_now] - PF[ys]); atomicAdd(dst + 10, dc * xtf);
          atomicAdd(dst + 11, aG); atomicAdd(dst + 12, aGs);
          atomicAdd(dst + 13, PD[y_now] - PD[ys]); atomicAdd(dst + 14, dc * xtd);
          atomicAdd(dst + 15, dc);
          if (bnd) atomicAdd(dst + 16, (double)bnd);
          const int brd = cnt * frame_x + (ys == 0) + (y_now == H);   // corners count twice
          if (brd) atomicAdd(dst + 17, (double)brd);
        } else {
          atomicOr(p.status, ST_LABEL_RANGE);
        }
        cnt = bnd = 0;
        aL = aA = aB = aL2 = aA2 = aB2 = aH = aS = aV = aG = aGs = 0.0;
      }
      if (!PAIRS) return;
      // down-neighbour pairs of the finished runs, one insertion per distinct pair
      const unsigned long long key =
          ((unsigned long long)(uint32_t)min(cur, nxt) << 32) | (uint32_t)max(cur, nxt);
      unsigned todo = __ballot_sync(0xffffffffu, mine && cur != nxt);
      while (todo) {
        const int src = __ffs(todo) - 1;
        const unsigned long long k0 = __shfl_sync(0xffffffffu, key, src);
        const unsigned same = __ballot_sync(0xffffffffu, (todo >> lane & 1u) && key == k0);
        todo &= ~same;
        pair_add((int)(k0 >> 32), (int)(k0 & 0xffffffffu), __popc(same));
      }
      return;
    }
    if (mine) {
      double* st = stage + lane * RS_STAGE_LD;
      const double dc = (double)cnt;
      st[0] = aL; st[1] = aA; st[2] = aB; st[3] = aL2; st[4] = aA2; st[5] = aB2;
      st[6] = aH; st[7] = aS; st[8] = aV;
      st[9] = PF[y_now] - PF[ys]; st[10] = dc * xtf;
      st[11] = aG; st[12] = aGs;
      st[13] = PD[y_now] - PD[ys]; st[14] = dc * xtd;
      st[15] = dc; st[16] = (double)bnd;
      st[17] = (double)(cnt * frame_x + (ys == 0) + (y_now == H));   // corners count twice
      cnt = bnd = 0;
      aL = aA = aB = aL2 = aA2 = aB2 = aH = aS = aV = aG = aGs = 0.0;
      lmax = max(lmax, cur);
    }
    __syncwarp();
    unsigned m = mask;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int L = __shfl_sync(0xffffffffu, cur, src);
      const int Ln = __shfl_sync(0xffffffffu, nxt, src);
      if ((unsigned)L < node_cap) {
        const int slot = (int)(((uint32_t)L * 0x9E3779B1u) >> 28);
        const int tag = __shfl_sync(0xffffffffu, mytag, slot);
        if (tag != L) {
          evict_slot(slot, tag);
          if (lane == slot) mytag = L;
        }
        // field f is always handled by lane f: no cross-lane hazard between iterations
        if (lane < RS_NF) vals[slot * RS_NF + lane] += stage[src * RS_STAGE_LD + lane];
      } else {
        if (lane == 0) atomicOr(p.status, ST_LABEL_RANGE);
      }
      pair_add(L, Ln, 1);      // vertical transition between the run and the pixel below it
    }
    __syncwarp();
  };

  for (int y = y_begin; y < y_end; ++y) {
    // ---- loads two rows ahead (row indices are warp-uniform)
    int g3c, g3e;
    load_grey(y + 3, g3c, g3e);
    // one address per pixel / label (pointer + constant offsets: 32-bit index arithmetic such as
    // 3 * ob + 1 may wrap and would cost an address computation per load)
    const int ob = min(y + 2, H - 1) * W + xc;
    const uint8_t* px2 = img + (size_t)3 * (size_t)ob;
    const int nb_ = px2[0], ng_ = px2[1], nr_ = px2[2];
    int lab_dn2 = lab_dn, e_lab_n2 = 0;                 // below the image: the own label
    if (y + 2 < H) {
      const int32_t* pl2 = lab + ob;
      if (valid) lab_dn2 = pl2[0];
      if (edge_lane) e_lab_n2 = pl2[ldelta];
    }
    const bool has_dn = y + 1 < H;

    // ---- horizontal neighbours: labels of this row, grey partials of row y+2
    int lab_l = __shfl_up_sync(0xffffffffu, lab_c, 1);
    int lab_r = __shfl_down_sync(0xffffffffu, lab_c, 1);
    if (is_l0) lab_l = e_lab;
    if (is_l31) lab_r = e_lab;
    if (own_r) lab_r = lab_c;
    int hs_p, hd_p;
    partials(g2c, g2e, hs_p, hd_p);

    // ---- run bookkeeping: hand finished runs to the warp table
    const unsigned fm = __ballot_sync(0xffffffffu, cnt > 0 && lab_c != cur);
    if (fm) flush_lanes(fm, y, lab_c);
    if (cnt == 0) ys = y;
    cur = lab_c;

    // ---- values of row y+1 (BGR(y+1), grey rows y, y+1, y+2) ...
    const PixelVals nv = pixel_math(pb, pg, pr, hs_m, hd_m, hd_c, hs_p, hd_p);
    // ---- ... while row y is accumulated (all lanes; lanes beyond the image feed a dead run)
    {
      if (VELT) {
        aL += cv.Ld; aA += cv.Ad; aB += cv.Bd;
        // the float32 product of two float32 values = their exact float64 product, rounded once
        aL2 += round_to_f32(__dmul_rn(cv.Ld, cv.Ld)); aA2 += round_to_f32(__dmul_rn(cv.Ad, cv.Ad));
        aB2 += round_to_f32(__dmul_rn(cv.Bd, cv.Bd));
      } else {
        aL += (double)cv.L; aA += (double)cv.A; aB += (double)cv.Bv;
        aL2 += (double)__fmul_rn(cv.L, cv.L); aA2 += (double)__fmul_rn(cv.A, cv.A);
        aB2 += (double)__fmul_rn(cv.Bv, cv.Bv);
      }
      aH += (double)cv.hh; aS += (double)cv.ss; aV += s_vd[cv.mx];
      aG += (double)cv.g; aGs += (double)cv.gs;
      cnt += 1;
      // find_boundaries(mode="inner"): differs from an in-bounds 4-neighbour and label != 0
      const bool diff = (lab_up != lab_c) | (lab_dn != lab_c) | (lab_l != lab_c) | (lab_r != lab_c);
      bnd += (diff && lab_c != 0) ? 1 : 0;
    }

    // ---- adjacency transitions (graph_builder.py:267-281)
    // right neighbour: identical transitions are run-aggregated down the column in registers;
    // down neighbour: emitted with the run hand-over above.
    if (PAIRS) {
      const int cb = (lab_r != lab_c) ? lab_r : -1;
      const bool changed = (lab_c != rp_a) | (cb != rp_b);
      unsigned em = __ballot_sync(0xffffffffu, changed && rp_b >= 0 && valid);
      while (em) {
        const int src = __ffs(em) - 1;
        em &= em - 1;
        pair_add(__shfl_sync(0xffffffffu, rp_a, src), __shfl_sync(0xffffffffu, rp_b, src),
                 __shfl_sync(0xffffffffu, rp_cnt, src));
      }
      if (changed) { rp_a = lab_c; rp_b = cb; rp_cnt = 0; }
      rp_cnt += 1;
    }
    if (PAIRS && p.connectivity == 8) {
      // the two diagonals below: (y+1, x+1) and (y+1, x-1); outside the image -> own label
      int dn_r = __shfl_down_sync(0xffffffffu, lab_dn, 1);
      int dn_l = __shfl_up_sync(0xffffffffu, lab_dn, 1);
      if (is_l31) dn_r = e_lab_n;
      if (is_l0) dn_l = e_lab_n;
      if (own_r || !has_dn) dn_r = lab_c;
      if (x == 0 || !has_dn) dn_l = lab_c;
#pragma unroll
      for (int sdir = 0; sdir < 2; ++sdir) {
        const int o = sdir ? dn_l : dn_r;
        unsigned tm = __ballot_sync(0xffffffffu, valid && o != lab_c);
        while (tm) {
          const int src = __ffs(tm) - 1;
          tm &= tm - 1;
          pair_add(__shfl_sync(0xffffffffu, lab_c, src), __shfl_sync(0xffffffffu, o, src), 1);
        }
      }
    }

    // ---- roll
    cv = nv;
    hs_m = hs_c; hd_m = hd_c; hs_c = hs_p; hd_c = hd_p;
    g2c = g3c; g2e = g3e;
    lab_up = lab_c; lab_c = lab_dn; lab_dn = lab_dn2;
    e_lab = e_lab_n; e_lab_n = e_lab_n2;
    pb = nb_; pg = ng_; pr = nr_;
  }

  // ---- end of strip: flush runs (lab_c now holds the label below the strip, or the own label
  //      at the bottom of the image), pending right-neighbour runs, the pair cache, the table
  const unsigned fm = __ballot_sync(0xffffffffu, cnt > 0 && valid);
  if (fm) flush_lanes(fm, y_end, lab_c);
  if (PAIRS) {
    unsigned em = __ballot_sync(0xffffffffu, rp_b >= 0 && valid);
    while (em) {
      const int src = __ffs(em) - 1;
      em &= em - 1;
      pair_add(__shfl_sync(0xffffffffu, rp_a, src), __shfl_sync(0xffffffffu, rp_b, src),
               __shfl_sync(0xffffffffu, rp_cnt, src));
    }
  }
  if (PAIRS && pk_lo >= 0) pair_emit(keys, cnts, p.table_cap, pk_lo, pk_hi, pk_n, p.status);
  for (int s_ = 0; s_ < RS_SLOTS; ++s_) evict_slot(s_, __shfl_sync(0xffffffffu, mytag, s_));
  lmax = warp_max_i(lmax);
  if (lane == 0 && lmax >= 0) atomicMax(&p.label_max[b], lmax);
}

// ============================================================================ S0b
// Adjacency transitions (graph_builder.py:265-286) as a labels-only pass: one warp per 32-column x
// `rows`-row strip, lane = column, walking down the rows.  Right-neighbour transitions are run-aggregated
// down the column in registers (a vertical boundary repeats the same pair row after row), down-neighbour
// (and diagonal) transitions are aggregated across the lanes by key; distinct pairs go through a 32-entry
// register cache (lane = slot) to the per-image open-addressing hash table -- the bookkeeping that used to
// ride inside k_region_stats, without its 119 registers and float64 chains around it.
struct AdjParams {
  const int32_t* labels;
  unsigned long long* pair_keys;
  int* pair_cnts;
  int* status;
  int B, H, W, node_cap, table_cap, connectivity, n_sx, n_sy, rows;
};

constexpr int ADJ_QUEUE = 128;      // pending (lo, hi, count) records per warp

__global__ void __launch_bounds__(256)
k_adjacency_pairs(const AdjParams p) {
  __shared__ int s_q[8][ADJ_QUEUE][3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_tasks = (long long)p.B * p.n_sy * p.n_sx;
  if (task >= n_tasks) return;
  const int sx = (int)(task % p.n_sx);
  const int sy = (int)((task / p.n_sx) % p.n_sy);
  const int b = (int)(task / ((long long)p.n_sx * p.n_sy));
  const int H = p.H, W = p.W;
  const int x = sx * 32 + lane;
  const bool valid = x < W;
  const int xc = valid ? x : W - 1;
  const int y_begin = sy * p.rows, y_end = min(H, y_begin + p.rows);
  const int32_t* lab = p.labels + (size_t)b * H * W;
  unsigned long long* keys = p.pair_keys + (size_t)b * p.table_cap;
  int* cnts = p.pair_cnts + (size_t)b * p.table_cap;
  const unsigned node_cap = (unsigned)p.node_cap;
  const bool is_l0 = lane == 0, is_l31 = lane == 31;
  const bool own_r = x >= W - 1;                       // no pixel to the right
  const bool edge_lane = (is_l0 || is_l31) && valid;
  const int ldelta = is_l0 ? (x > 0 ? -1 : 0) : (x + 1 < W ? 1 : 0);
  const unsigned lt_mask = (1u << lane) - 1u;

  // Finished records are queued lane-parallel (slot = ballot prefix, the fill level is warp-uniform) and
  // the queue is drained lane-parallel into the hash table: no warp-uniform loop over the lanes that
  // have something to report (such loops were 60 % of this kernel's instructions).
  int (*q)[3] = s_q[wid];
  int qn = 0;
  auto drain = [&]() {
    __syncwarp();
    for (int i = lane; i < qn; i += 32) {
      const int a = q[i][0], c = q[i][1], n = q[i][2];
      const int lo = min(a, c), hi = max(a, c);
      if (lo != hi && (unsigned)lo < node_cap && (unsigned)hi < node_cap)
        pair_emit(keys, cnts, p.table_cap, lo, hi, n, p.status);
    }
    __syncwarp();
    qn = 0;
  };
  auto push = [&](bool want, int a, int c, int n) {    // called by the whole warp
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const int k = __popc(m);
    if (qn + k > ADJ_QUEUE) drain();
    if (want) {
      const int pos = qn + __popc(m & lt_mask);
      q[pos][0] = a; q[pos][1] = c; q[pos][2] = n;
    }
    qn += k;
  };
  // transitions towards `other` (per lane): equal neighbouring transitions along the row are merged --
  // the first lane of a stretch reports the stretch length
  auto push_stretches = [&](bool trans, int mine, int other) {
    const unsigned tm = __ballot_sync(0xffffffffu, trans);
    if (!tm) return;
    const int lm = __shfl_up_sync(0xffffffffu, mine, 1), lo_ = __shfl_up_sync(0xffffffffu, other, 1);
    const bool head = trans && (is_l0 || !(tm >> (lane - 1) & 1u) || lm != mine || lo_ != other);
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    const unsigned stop = (hm | ~tm) & ~lt_mask & ~(1u << lane);     // next head or non-transition lane above me
    const int len = (stop ? __ffs(stop) - 1 : 32) - lane;
    push(head, mine, other, len);
  };

  const uint32_t off0 = (uint32_t)(y_begin * W + xc);
  int lab_c = valid ? lab[off0] : -1;
  int e_lab = edge_lane ? lab[off0 + ldelta] : 0;      // outer label neighbour of the edge lanes
  int rp_a = -1, rp_b = -1, rp_cnt = 0;                // run of identical right-neighbour transitions
  // the rows below are loaded two iterations ahead: the register hand-down at the end of an iteration
  // (row y+2 becomes row y+1) then never waits for a load issued in the same iteration
  auto load_row = [&](int yy, int fallback, int& c, int& e) {   // labels of row yy (below the image: `fallback`)
    c = fallback; e = 0;
    if (yy < H) {
      const int32_t* pl = lab + (size_t)yy * W + xc;
      if (valid) c = pl[0];
      if (edge_lane) e = pl[ldelta];
    }
  };
  int lab_dn, e_lab_n, lab_dn2, e_lab_n2;
  load_row(y_begin + 1, lab_c, lab_dn, e_lab_n);
  load_row(y_begin + 2, lab_dn, lab_dn2, e_lab_n2);
  for (int y = y_begin; y < y_end; ++y) {
    const bool has_dn = y + 1 < H;
    int lab_dn3, e_lab_n3;                             // row y+3
    load_row(y + 3, lab_dn2, lab_dn3, e_lab_n3);
    int lab_r = __shfl_down_sync(0xffffffffu, lab_c, 1);
    if (is_l31) lab_r = e_lab;
    if (own_r) lab_r = lab_c;
    // right neighbour, run-aggregated down the column: a finished run is one record
    {
      const int cb = (lab_r != lab_c) ? lab_r : -1;
      const bool changed = (lab_c != rp_a) | (cb != rp_b);
      push(changed && rp_b >= 0 && valid, rp_a, rp_b, rp_cnt);
      if (changed) { rp_a = lab_c; rp_b = cb; rp_cnt = 0; }
      rp_cnt += 1;
    }
    // down neighbour
    push_stretches(valid && lab_dn != lab_c, lab_c, lab_dn);
    if (p.connectivity == 8) {
      // the two diagonals below: (y+1, x+1) and (y+1, x-1); outside the image -> own label
      int dn_r = __shfl_down_sync(0xffffffffu, lab_dn, 1);
      int dn_l = __shfl_up_sync(0xffffffffu, lab_dn, 1);
      if (is_l31) dn_r = e_lab_n;
      if (is_l0) dn_l = e_lab_n;
      if (own_r || !has_dn) dn_r = lab_c;
      if (x == 0 || !has_dn) dn_l = lab_c;
      push_stretches(valid && dn_r != lab_c, lab_c, dn_r);
      push_stretches(valid && dn_l != lab_c, lab_c, dn_l);
    }
    lab_c = lab_dn; lab_dn = lab_dn2; lab_dn2 = lab_dn3;
    e_lab = e_lab_n; e_lab_n = e_lab_n2; e_lab_n2 = e_lab_n3;
  }
  push(rp_b >= 0 && valid, rp_a, rp_b, rp_cnt);
  drain();
}

// ============================================================================ S1
// Region sums -> per-region statistics (graph_builder.py:194-226 and :391-403).
__global__ void k_finalize_regions(const double* __restrict__ acc, const int* __restrict__ label_max,
                                   float* __restrict__ st, int B, int H, int W, int node_cap) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = label_max[b] + 1;
  if (i >= n || i >= node_cap) return;
  const double* a = acc + ((size_t)b * node_cap + i) * RS_NF;
  const float counts = (float)a[15];
  const float bnd = (float)a[16];
  const float brd = (float)a[17];
  const float safe = fmaxf(counts, 1.0f);
  float* s = st + (size_t)b * ST_FIELDS * node_cap + i;
  auto put = [&](int f, float v) { s[(size_t)f * node_cap] = v; };
  put(ST_COUNT, counts);
  put(ST_SAFE, safe);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float mean = __fdiv_rn((float)a[c], safe);
    const float sq = __fdiv_rn((float)a[3 + c], safe);
    const float var = fmaxf(__fsub_rn(sq, __fmul_rn(mean, mean)), 0.0f);
    put(ST_MEAN_L + c, mean);
    put(ST_STD_L + c, __fsqrt_rn(var));
    put(ST_MEAN_H + c, __fdiv_rn((float)a[6 + c], safe));
  }
  put(ST_CY, __fdiv_rn((float)a[9], safe));
  put(ST_CX, __fdiv_rn((float)a[10], safe));
  put(ST_BND, bnd);
  put(ST_AREA, __fdiv_rn(counts, (float)((double)H * (double)W)));
  put(ST_MGRAD, __fdiv_rn((float)a[11], safe));
  put(ST_MGRADN, __fdiv_rn((float)a[12], safe));
  // compute_auto_prior: float64 sum / float32 safe -> float64 -> float32 (…:401-403)
  put(ST_PCY, (float)(a[13] / (double)safe));
  put(ST_PCX, (float)(a[14] / (double)safe));
  put(ST_BORDER, brd);
}

// ============================================================================ S2
// Per image: hash table -> pairs sorted by (lo, hi) via counting sort on lo + per-lo
// insertion sort on hi.  start[] (node_cap+1 ints) is kept for adjacency lookups.
__global__ void __launch_bounds__(512)
k_adj_sort(const unsigned long long* __restrict__ keys_all, const int* __restrict__ cnts_all,
           const int* __restrict__ label_max, int2* __restrict__ pairs_all,
           int* __restrict__ shared_all, int* __restrict__ start_all, int* __restrict__ cursor_all,
           int* __restrict__ n_adj, int* __restrict__ max_shared, int* __restrict__ status,
           int node_cap, int table_cap, int pair_cap) {
  __shared__ int scratch[40];
  const int b = blockIdx.x;
  const int n = min(label_max[b] + 1, node_cap);
  const unsigned long long* keys = keys_all + (size_t)b * table_cap;
  const int* cnts = cnts_all + (size_t)b * table_cap;
  int2* pairs = pairs_all + (size_t)b * pair_cap;
  int* shared = shared_all + (size_t)b * pair_cap;
  int* start = start_all + (size_t)b * (node_cap + 1);
  int* cursor = cursor_all + (size_t)b * node_cap;
  const unsigned long long EMPTY = ~0ull;

  for (int i = threadIdx.x; i < n; i += blockDim.x) { start[i] = 0; cursor[i] = 0; }
  __syncthreads();
  for (int s = threadIdx.x; s < table_cap; s += blockDim.x) {
    const unsigned long long k = keys[s];
    if (k != EMPTY) atomicAdd(&start[(int)(k >> 32)], 1);
  }
  __syncthreads();
  int carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? start[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, scratch, &total);
    if (i < n) start[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  const int n_pairs = carry;
  if (threadIdx.x == 0) {
    start[n] = n_pairs;
    n_adj[b] = n_pairs > pair_cap ? 0 : n_pairs;
    if (n_pairs > pair_cap) { atomicOr(status, ST_EDGE_CAP); max_shared[b] = 0; }
  }
  __syncthreads();
  if (n_pairs > pair_cap) return;
  int mx = 0;
  for (int s = threadIdx.x; s < table_cap; s += blockDim.x) {
    const unsigned long long k = keys[s];
    if (k != EMPTY) {
      const int lo = (int)(k >> 32), hi = (int)(k & 0xFFFFFFFFull);
      const int pos = start[lo] + atomicAdd(&cursor[lo], 1);
      pairs[pos] = make_int2(lo, hi);
      shared[pos] = cnts[s];
      mx = max(mx, cnts[s]);
    }
  }
  mx = block_reduce<int>(mx, 0, OpMaxI(), scratch);
  if (threadIdx.x == 0) max_shared[b] = mx;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int s0 = start[i], s1 = start[i + 1];
    for (int a = s0 + 1; a < s1; ++a) {
      const int2 pv = pairs[a];
      const int cv = shared[a];
      int j = a - 1;
      while (j >= s0 && pairs[j].y > pv.y) {
        pairs[j + 1] = pairs[j];
        shared[j + 1] = shared[j];
        --j;
      }
      pairs[j + 1] = pv;
      shared[j + 1] = cv;
    }
  }
}

// ============================================================================ S3
// k nearest neighbours in mean-Lab space, self and spatially adjacent regions excluded
// (graph_builder.py:333-347).  One warp per region; ties on distance -> lower index.
GG_D bool knn_less(float d, int j, float bd, int bj) { return d < bd || (d == bd && j < bj); }

// Non-local neighbours for graphs of any size.  k_knn (below) keeps ONE sorted candidate list per warp and
// tests adjacency for the winners only; this routine is its exact fall-back for a row whose nearest
// regions are mostly adjacent ones.  It is the original form of the kernel -- per-lane sorted lists,
// adjacency tested before every insertion -- whose insertions run in divergent code, one lane at a
// time (ncu at config E: 83 % issue slots, 17 of 32 threads active, ~220 instructions per 32
// candidates; 3.96 ms for 8 x 10^4 regions).
template <int K>
__device__ void knn_row_exact(const float* __restrict__ mL, const float* __restrict__ mA, const float* __restrict__ mB,
                              const int2* __restrict__ pairs, const int* __restrict__ start, int n, int i, int k,
                              int lane, int* __restrict__ out) {
  const float INF = __int_as_float(0x7f800000);
  float bd[K];
  int bj[K];
#pragma unroll
  for (int t = 0; t < K; ++t) { bd[t] = INF; bj[t] = 0x7fffffff; }
  const float li = mL[i], ai = mA[i], bi = mB[i];
  for (int j = lane; j < n; j += 32) {
    if (j == i) continue;
    const float dx = __fsub_rn(li, mL[j]), dy = __fsub_rn(ai, mA[j]), dz = __fsub_rn(bi, mB[j]);
    const float d = __fsqrt_rn(
        __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    if (!(d < INF)) continue;   // inf / nan never selected (np.isfinite filter, …:346)
    // only the k best of a lane can make the warp's k best
    bool cand = false;
#pragma unroll
    for (int t = 0; t < K; ++t)
      if (t == k - 1) cand = knn_less(d, j, bd[t], bj[t]);
    if (!cand) continue;
    const int lo = min(i, j), hi = max(i, j);
    bool adjacent = false;
    for (int q = start[lo]; q < start[lo + 1]; ++q)
      if (pairs[q].y == hi) { adjacent = true; break; }
    if (adjacent) continue;
    float cd = d;
    int cj = j;
#pragma unroll
    for (int t = 0; t < K; ++t) {
      if (knn_less(cd, cj, bd[t], bj[t])) {
        const float td = bd[t]; const int tj = bj[t];
        bd[t] = cd; bj[t] = cj;
        cd = td; cj = tj;
      }
    }
  }
  // merge the 32 sorted lists: k rounds of warp arg-min on (d, j)
  for (int r = 0; r < k; ++r) {
    float d = bd[0];
    int j = bj[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, o);
      const int oj = __shfl_xor_sync(0xffffffffu, j, o);
      if (knn_less(od, oj, d, j)) { d = od; j = oj; }
    }
    if (lane == 0) out[r] = (d < INF) ? j : -1;
    if (bj[0] == j && bd[0] == d && d < INF) {   // the winner pops its head
#pragma unroll
      for (int t = 0; t + 1 < K; ++t) { bd[t] = bd[t + 1]; bj[t] = bj[t + 1]; }
      bd[K - 1] = INF; bj[K - 1] = 0x7fffffff;
    }
  }
}

template <int K>
__global__ void __launch_bounds__(256)
k_knn(const float* __restrict__ st_all, const int* __restrict__ label_max,
      const int2* __restrict__ pairs_all, const int* __restrict__ start_all,
      int* __restrict__ picks_all, int node_cap, int pair_cap, int k) {
  const int b = blockIdx.y;
  const int n = min(label_max[b] + 1, node_cap);
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* st = st_all + (size_t)b * ST_FIELDS * node_cap;
  const float* mL = st + (size_t)ST_MEAN_L * node_cap;
  const float* mA = mL + node_cap;
  const float* mB = mA + node_cap;
  const int2* pairs = pairs_all + (size_t)b * pair_cap;
  const int* start = start_all + (size_t)b * (node_cap + 1);
  const float INF = __int_as_float(0x7f800000);
  int* out = picks_all + ((size_t)b * node_cap + i) * k;

  // ONE sorted list per warp, rank t in lane t: the kl = min(32, k + 8) nearest regions of region i,
  // adjacent or not.  A sweep step looks at 32 candidates; almost always none beats the list's last
  // entry (one compare + one ballot).  A candidate that does is inserted by the whole warp (rank by
  // ballot, shift by shuffle): ~kl ln(n / kl) insertions per row, none of them in divergent code.
  const int kl = min(32, k + 8);
  float my_d = INF;                            // this lane's list entry (rank = lane)
  int my_j = 0x7fffffff;
  float thr_d = INF;                           // entry of rank kl - 1
  int thr_j = 0x7fffffff;
  const float li = mL[i], ai = mA[i], bi = mB[i];
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    float d = INF;
    if (j < n && j != i) {
      const float dx = __fsub_rn(li, mL[j]), dy = __fsub_rn(ai, mA[j]), dz = __fsub_rn(bi, mB[j]);
      d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      if (!(d < INF)) d = INF;                 // inf / nan never selected (np.isfinite filter, …:346)
    }
    unsigned todo = __ballot_sync(0xffffffffu, d < INF && knn_less(d, j, thr_d, thr_j));
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const float cd = __shfl_sync(0xffffffffu, d, src);
      const int cj = __shfl_sync(0xffffffffu, j, src);
      if (!knn_less(cd, cj, thr_d, thr_j)) continue;          // the list moved on since the ballot
      const int pos = __popc(__ballot_sync(0xffffffffu, knn_less(my_d, my_j, cd, cj)));   // entries before it
      const float up_d = __shfl_up_sync(0xffffffffu, my_d, 1);
      const int up_j = __shfl_up_sync(0xffffffffu, my_j, 1);
      if (lane == pos) { my_d = cd; my_j = cj; }
      else if (lane > pos) { my_d = up_d; my_j = up_j; }
      thr_d = __shfl_sync(0xffffffffu, my_d, kl - 1);
      thr_j = __shfl_sync(0xffffffffu, my_j, kl - 1);
    }
  }
  // the first k entries that are not spatially adjacent (warp-parallel scan of the sorted pair list);
  // the list is the exact, ordered top kl, so the result is exact iff k of them are found in it --
  // otherwise (more than kl - k adjacent regions among the nearest kl: never on SLIC-like maps) the row
  // is redone by the per-candidate routine
  int r = 0, t = 0;
  bool complete = false;
  for (; t < kl; ++t) {
    const float d = __shfl_sync(0xffffffffu, my_d, t);
    const int j = __shfl_sync(0xffffffffu, my_j, t);
    if (!(d < INF)) { complete = true; break; }               // fewer than kl finite candidates in total: all seen
    const int lo = min(i, j), hi = max(i, j);
    bool found = false;
    for (int q = start[lo] + lane; q < start[lo + 1]; q += 32) found |= pairs[q].y == hi;
    if (__any_sync(0xffffffffu, found)) continue;
    if (lane == 0) out[r] = j;
    if (++r == k) { complete = true; break; }
  }
  if (!complete) {
    knn_row_exact<K>(mL, mA, mB, pairs, start, n, i, k, lane, out);
    return;
  }
  if (lane == 0)
    for (; r < k; ++r) out[r] = -1;
}

// Selection form of the same search for images with at most 32*T regions: each lane keeps the
// distances to regions lane, lane+32, ... in registers; every round takes the warp-wide minimum
// of (distance, index), retires it, and only then asks whether that region is adjacent (a
// warp-parallel scan of the sorted pair list) -- adjacency is tested for the handful of nearest
// regions instead of for every candidate.  Same result as k_knn (ties -> lower index).
template <int T>
__global__ void __launch_bounds__(256)
k_knn_sel(const float* __restrict__ st_all, const int* __restrict__ label_max,
          const int2* __restrict__ pairs_all, const int* __restrict__ start_all,
          int* __restrict__ picks_all, int node_cap, int pair_cap, int k) {
  const int b = blockIdx.y;
  const int n = min(label_max[b] + 1, node_cap);
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* st = st_all + (size_t)b * ST_FIELDS * node_cap;
  const float* mL = st + (size_t)ST_MEAN_L * node_cap;
  const float* mA = mL + node_cap;
  const float* mB = mA + node_cap;
  const int2* pairs = pairs_all + (size_t)b * pair_cap;
  const int* start = start_all + (size_t)b * (node_cap + 1);
  const float INF = __int_as_float(0x7f800000);
  const float li = mL[i], ai = mA[i], bi = mB[i];
  float dv[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int j = lane + 32 * t;
    float d = INF;
    if (j < n && j != i) {
      const float dx = __fsub_rn(li, mL[j]), dy = __fsub_rn(ai, mA[j]), dz = __fsub_rn(bi, mB[j]);
      d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      if (!(d < INF)) d = INF;      // inf / nan never selected (np.isfinite filter, …:346)
    }
    dv[t] = d;
  }
  int* out = picks_all + ((size_t)b * node_cap + i) * k;
  int r = 0;
  while (r < k) {
    // lane-local minimum; the strict '<' keeps the lowest index of equal distances
    float md = INF;
    int mt = 0;
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (dv[t] < md) { md = dv[t]; mt = t; }
    float d = md;
    int j = md < INF ? lane + 32 * mt : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, o);
      const int oj = __shfl_xor_sync(0xffffffffu, j, o);
      if (knn_less(od, oj, d, j)) { d = od; j = oj; }
    }
    if (!(d < INF)) break;          // fewer than k finite candidates
    if (lane == (j & 31)) {
      const int wt = j >> 5;
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (t == wt) dv[t] = INF;
    }
    const int lo = min(i, j), hi = max(i, j);
    bool found = false;
    for (int q = start[lo] + lane; q < start[lo + 1]; q += 32) found |= pairs[q].y == hi;
    if (__any_sync(0xffffffffu, found)) continue;        // spatially adjacent: excluded
    if (lane == 0) out[r] = j;
    ++r;
  }
  if (lane == 0)
    for (; r < k; ++r) out[r] = -1;
}

// The same selection for images with up to 2048 regions (T = 32 or 64 distances per lane): the
// distances live in shared memory (one padded row per lane), every lane keeps only its current
// minimum in registers, and after a round only the WINNER's minimum is recomputed -- by the whole
// warp, from the winner's row -- instead of every lane rescanning its T registers.  Rounds cost
// ~60 instructions instead of ~5 T.  Same result as k_knn_sel / k_knn.
template <int T>
__global__ void __launch_bounds__(256)
k_knn_sel_smem(const float* __restrict__ st_all, const int* __restrict__ label_max,
               const int2* __restrict__ pairs_all, const int* __restrict__ start_all,
               int* __restrict__ picks_all, int node_cap, int pair_cap, int k) {
  static_assert(T % 32 == 0, "T is a multiple of the warp size");
  extern __shared__ float knn_smem[];
  constexpr int LD = T + 1;                               // odd row length: conflict-free both ways
  const int b = blockIdx.y;
  const int n = min(label_max[b] + 1, node_cap);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + wid;
  if (i >= n) return;
  float* rows = knn_smem + (size_t)wid * 32 * LD;         // rows[lane][t]: distance to region lane + 32 t
  float* mine = rows + lane * LD;
  const float* st = st_all + (size_t)b * ST_FIELDS * node_cap;
  const float* mL = st + (size_t)ST_MEAN_L * node_cap;
  const float* mA = mL + node_cap;
  const float* mB = mA + node_cap;
  const int2* pairs = pairs_all + (size_t)b * pair_cap;
  const int* start = start_all + (size_t)b * (node_cap + 1);
  const float INF = __int_as_float(0x7f800000);
  const float li = mL[i], ai = mA[i], bi = mB[i];
  float md = INF;
  int mt = 0;
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    const int j = lane + 32 * t;
    float d = INF;
    if (j < n && j != i) {
      const float dx = __fsub_rn(li, mL[j]), dy = __fsub_rn(ai, mA[j]), dz = __fsub_rn(bi, mB[j]);
      d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      if (!(d < INF)) d = INF;      // inf / nan never selected (np.isfinite filter, …:346)
    }
    mine[t] = d;
    if (d < md) { md = d; mt = t; }   // strict '<': lowest index among equal distances
  }
  __syncwarp();
  int* out = picks_all + ((size_t)b * node_cap + i) * k;
  int r = 0;
  while (r < k) {
    float d = md;
    int j = md < INF ? lane + 32 * mt : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, o);
      const int oj = __shfl_xor_sync(0xffffffffu, j, o);
      if (knn_less(od, oj, d, j)) { d = od; j = oj; }
    }
    if (!(d < INF)) break;          // fewer than k finite candidates
    const int wl = j & 31;
    if (lane == wl) mine[j >> 5] = INF;                  // retire the winner
    __syncwarp();
    {
      // the winner's new minimum, computed by the whole warp from the winner's row
      const float* wrow = rows + wl * LD;
      float cd = INF;
      int ct = 0x7fffffff;
#pragma unroll
      for (int u = 0; u < T / 32; ++u) {
        const float v = wrow[lane + 32 * u];
        if (v < cd) { cd = v; ct = lane + 32 * u; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, cd, o);
        const int ot = __shfl_xor_sync(0xffffffffu, ct, o);
        if (knn_less(od, ot, cd, ct)) { cd = od; ct = ot; }
      }
      if (lane == wl) { md = cd; mt = cd < INF ? ct : 0; }
    }
    const int lo = min(i, j), hi = max(i, j);
    bool found = false;
    for (int q = start[lo] + lane; q < start[lo + 1]; q += 32) found |= pairs[q].y == hi;
    if (__any_sync(0xffffffffu, found)) continue;        // spatially adjacent: excluded
    if (lane == 0) out[r] = j;
    ++r;
  }
  if (lane == 0)
    for (; r < k; ++r) out[r] = -1;
}

// ============================================================================ S4
// Symmetrise the picks into sorted unique (lo,hi) pairs (graph_builder.py:348-350).
__global__ void __launch_bounds__(512)
k_nl_pairs(const int* __restrict__ picks_all, const int* __restrict__ label_max,
           const int* __restrict__ n_adj, int2* __restrict__ pairs_all,
           int* __restrict__ start_all, int* __restrict__ cursor_all, int* __restrict__ n_nl,
           int* __restrict__ status, int node_cap, int pair_cap, int k) {
  __shared__ int scratch[40];
  const int b = blockIdx.x;
  const int n = min(label_max[b] + 1, node_cap);
  const int* picks = picks_all + (size_t)b * node_cap * k;
  const int base_adj = n_adj[b];
  int2* out = pairs_all + (size_t)b * pair_cap + base_adj;   // appended after the adjacency pairs
  int* start = start_all + (size_t)b * (node_cap + 1);
  int* cursor = cursor_all + (size_t)b * node_cap;

  if (k <= 0 || n <= k + 1) {                                   // graph_builder.py:291
    if (threadIdx.x == 0) n_nl[b] = 0;
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) { start[i] = 0; cursor[i] = 0; }
  __syncthreads();
  // keep (i -> j) if i is the smaller id, or if the smaller id did not pick i itself
  auto keeps = [&](int i, int j) -> bool {
    if (j < 0) return false;
    if (i < j) return true;
    for (int t = 0; t < k; ++t)
      if (picks[(size_t)j * k + t] == i) return false;
    return true;
  };
  for (int e = threadIdx.x; e < n * k; e += blockDim.x) {
    const int i = e / k, j = picks[e];
    if (keeps(i, j)) atomicAdd(&start[min(i, j)], 1);
  }
  __syncthreads();
  int carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? start[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, scratch, &total);
    if (i < n) start[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  const int n_pairs = carry;
  if (threadIdx.x == 0) {
    start[n] = n_pairs;
    const bool over = base_adj + n_pairs > pair_cap;
    n_nl[b] = over ? 0 : n_pairs;
    if (over) atomicOr(status, ST_EDGE_CAP);
  }
  __syncthreads();
  if (base_adj + n_pairs > pair_cap) return;
  for (int e = threadIdx.x; e < n * k; e += blockDim.x) {
    const int i = e / k, j = picks[e];
    if (keeps(i, j)) {
      const int lo = min(i, j), hi = max(i, j);
      out[start[lo] + atomicAdd(&cursor[lo], 1)] = make_int2(lo, hi);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int s0 = start[i], s1 = start[i + 1];
    if (s1 - s0 <= CSR_LOCAL_ROW) {                    // all pairs of the row share lo = i: sort the partners locally
      int hi[CSR_LOCAL_ROW];
      const int len = s1 - s0;
      for (int a = 0; a < len; ++a) hi[a] = out[s0 + a].y;
      for (int a = 1; a < len; ++a) {
        const int hv = hi[a];
        int j = a - 1;
        while (j >= 0 && hi[j] > hv) { hi[j + 1] = hi[j]; --j; }
        hi[j + 1] = hv;
      }
      for (int a = 0; a < len; ++a) out[s0 + a] = make_int2(i, hi[a]);
      continue;
    }
    for (int a = s0 + 1; a < s1; ++a) {
      const int2 pv = out[a];
      int j = a - 1;
      while (j >= s0 && out[j].y > pv.y) { out[j + 1] = out[j]; --j; }
      out[j + 1] = pv;
    }
  }
}

// ============================================================================ S5
// Ragged offsets over the batch (single block).
__global__ void __launch_bounds__(1024)
k_offsets(const int* __restrict__ label_max, const int* __restrict__ n_adj,
          const int* __restrict__ n_nl, int B, int node_cap, int32_t* __restrict__ n_nodes,
          int32_t* __restrict__ n_edges, int64_t* __restrict__ node_off,
          int64_t* __restrict__ edge_off, int32_t* __restrict__ o_n_adj,
          int32_t* __restrict__ o_n_nl) {
  __shared__ int scratch[40];
  int carry_n = 0, carry_e = 0;
  for (int base = 0; base < B; base += blockDim.x) {
    const int b = base + threadIdx.x;
    const int nn = (b < B) ? min(label_max[b] + 1, node_cap) : 0;
    const int ne = (b < B) ? 2 * (n_adj[b] + n_nl[b]) : 0;
    int tn, te;
    const int en = block_exclusive_scan(nn, scratch, &tn);
    const int ee = block_exclusive_scan(ne, scratch, &te);
    if (b < B) {
      n_nodes[b] = nn;
      n_edges[b] = ne;
      node_off[b] = carry_n + en;
      edge_off[b] = carry_e + ee;
      if (o_n_adj) o_n_adj[b] = n_adj[b];
      if (o_n_nl) o_n_nl[b] = n_nl[b];
    }
    carry_n += tn;
    carry_e += te;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    node_off[B] = carry_n;
    edge_off[B] = carry_e;
  }
}

// ============================================================================ S6
// The 16 image-derived node attributes (graph_builder.py:228-255).
GG_D float nan_to_num(float v) {
  if (isnan(v)) return 0.0f;
  if (isinf(v)) return v > 0 ? 1.0f : 0.0f;
  return v;
}

__global__ void __launch_bounds__(256)
k_node_features(const float* __restrict__ st_all, const int* __restrict__ label_max,
                const int64_t* __restrict__ node_off, float* __restrict__ x,
                float* __restrict__ centroids, float* __restrict__ areas, int node_cap) {
  __shared__ float sred[32];
  __shared__ float s_mn[6], s_mx[6];
  const int b = blockIdx.x;
  const int n = min(label_max[b] + 1, node_cap);
  const float* st = st_all + (size_t)b * ST_FIELDS * node_cap;
  auto S = [&](int f, int i) { return st[(size_t)f * node_cap + i]; };
  const int64_t no = node_off[b];
  const float INF = __int_as_float(0x7f800000);

  float mn[6], mx[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) { mn[c] = INF; mx[c] = -INF; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float m = S(ST_MEAN_L + c, i), sd = S(ST_STD_L + c, i);
      mn[c] = fminf(mn[c], m); mx[c] = fmaxf(mx[c], m);
      mn[3 + c] = fminf(mn[3 + c], sd); mx[3 + c] = fmaxf(mx[3 + c], sd);
    }
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const float a = block_reduce<float>(mn[c], INF, OpMinF(), sred);
    const float z = block_reduce<float>(mx[c], -INF, OpMaxF(), sred);
    if (threadIdx.x == 0) { s_mn[c] = a; s_mx[c] = z; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float* f = x + (size_t)(no + i) * GG_N_NODE_FEATS;
    const float counts = S(ST_COUNT, i), safe = S(ST_SAFE, i), bnd = S(ST_BND, i);
    const float cy = S(ST_CY, i), cx = S(ST_CX, i);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float den0 = __fadd_rn(__fsub_rn(s_mx[c], s_mn[c]), 1e-6f);
      const float den1 = __fadd_rn(__fsub_rn(s_mx[3 + c], s_mn[3 + c]), 1e-6f);
      f[c] = nan_to_num(__fdiv_rn(__fsub_rn(S(ST_MEAN_L + c, i), s_mn[c]), den0));
      f[3 + c] = nan_to_num(__fdiv_rn(__fsub_rn(S(ST_STD_L + c, i), s_mn[3 + c]), den1));
      f[6 + c] = nan_to_num(S(ST_MEAN_H + c, i));
    }
    f[9] = nan_to_num(cy);
    f[10] = nan_to_num(cx);
    f[11] = nan_to_num(S(ST_AREA, i));
    const float perim = fmaxf(bnd, 1.0f);
    const float iso = __fdiv_rn(__fmul_rn(12.566370614359172f, counts), __fmul_rn(perim, perim));
    f[12] = nan_to_num(fminf(fmaxf(iso, 0.0f), 1.0f));
    f[13] = nan_to_num(__fdiv_rn(S(ST_MGRAD, i), 255.0f));
    f[14] = nan_to_num(__fdiv_rn(bnd, safe));
    const float dy = __fsub_rn(cy, 0.5f), dx = __fsub_rn(cx, 0.5f);
    f[15] = nan_to_num(__fdiv_rn(
        __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx))), 0.707f));
    if (centroids) { centroids[(size_t)(no + i) * 2] = cy; centroids[(size_t)(no + i) * 2 + 1] = cx; }
    if (areas) areas[no + i] = S(ST_AREA, i);
  }
}

// ============================================================================ S7
// Global colour contrast (graph_builder.py:405-412): one warp per region.
// PC_ROWS = regions per warp: 2 for small graphs (more warps in flight), 4 for large ones (fewer loads per pair)
template <int PC_ROWS>
__global__ void __launch_bounds__(256)
k_prior_contrast(const float* __restrict__ st_all, const int* __restrict__ label_max,
                 float* __restrict__ contrast_all, int node_cap, float two_sig2 /* float32(2*contrast_sigma**2) */) {
  const int b = blockIdx.y;
  const int n = min(label_max[b] + 1, node_cap);
  const int lane = threadIdx.x & 31;
  // a warp owns PC_ROWS consecutive regions: the values of region j are loaded once for all of them
  // and the independent rows hide each other's latency (same per-row summation order as one row per warp)
  const int i0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PC_ROWS;
  if (i0 >= n) return;
  const float* st = st_all + (size_t)b * ST_FIELDS * node_cap;
  const float* mL = st + (size_t)ST_MEAN_L * node_cap;
  const float* mA = mL + node_cap;
  const float* mB = mA + node_cap;
  const float* pcy = st + (size_t)ST_PCY * node_cap;
  const float* pcx = st + (size_t)ST_PCX * node_cap;
  // area_w = counts / counts.sum() (graph_builder.py:409): every pixel carries a label, so
  // counts.sum() is H*W (exact in float32) and area_w is the area ratio k_finalize_regions stored
  const float* area = st + (size_t)ST_AREA * node_cap;
  float li[PC_ROWS], ai[PC_ROWS], bi[PC_ROWS], yi[PC_ROWS], xi[PC_ROWS];
  double s[PC_ROWS];
#pragma unroll
  for (int k = 0; k < PC_ROWS; ++k) {
    const int i = min(i0 + k, n - 1);
    li[k] = mL[i]; ai[k] = mA[i]; bi[k] = mB[i]; yi[k] = pcy[i]; xi[k] = pcx[i];
    s[k] = 0.0;
  }
  for (int j = lane; j < n; j += 32) {
    const float lj = mL[j], aj = mA[j], bj = mB[j], yj = pcy[j], xj = pcx[j], wj = area[j];
#pragma unroll
    for (int k = 0; k < PC_ROWS; ++k) {
      const float dl = __fsub_rn(li[k], lj), da = __fsub_rn(ai[k], aj), db = __fsub_rn(bi[k], bj);
      const float cd = __fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db)));
      const float dy = __fsub_rn(yi[k], yj), dx = __fsub_rn(xi[k], xj);
      const float sd = __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
      const float sw = expf(__fdiv_rn(-__fmul_rn(sd, sd), two_sig2));
      s[k] += (double)__fmul_rn(__fmul_rn(cd, sw), wj);
    }
  }
#pragma unroll
  for (int k = 0; k < PC_ROWS; ++k) {
    const double t = warp_sum(s[k]);
    if (lane == 0 && i0 + k < n) contrast_all[(size_t)b * node_cap + i0 + k] = (float)t;
  }
}

// Block-wide unit-norm helpers: min and max of v over [0,n) strided by the block.
template <typename F>
GG_D void block_minmax(int n, F value, float* sred, float& mn, float& mx) {
  const float INF = __int_as_float(0x7f800000);
  float a = INF, z = -INF;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = value(i);
    a = fminf(a, v);
    z = fmaxf(z, v);
  }
  mn = block_reduce<float>(a, INF, OpMinF(), sred);
  mx = block_reduce<float>(z, -INF, OpMaxF(), sred);
}
// _unit_norm (graph_builder.py:447-454): (v - mn) / float32(mx - mn), zeros if flat.
GG_D float unit_norm_apply(float v, float mn, float mx) {
  if ((double)mx - (double)mn < 1e-8) return 0.0f;
  return __fdiv_rn(__fsub_rn(v, mn), (float)((double)mx - (double)mn));
}

// Centre prior, border colour model, ambiguity (graph_builder.py:414-444).
__global__ void __launch_bounds__(256)
k_prior_finish(const float* __restrict__ st_all, const int* __restrict__ label_max,
               const float* __restrict__ contrast_all, float* __restrict__ tmp_all,
               const int64_t* __restrict__ node_off, float* __restrict__ x, int node_cap,
               float two_c2 /* float32(2*centre_sigma**2) */, int out_stride, int out_col) {
  __shared__ float sred[32];
  __shared__ double sredd[32];
  const int b = blockIdx.x;
  const int n = min(label_max[b] + 1, node_cap);
  const float* st = st_all + (size_t)b * ST_FIELDS * node_cap;
  auto S = [&](int f, int i) { return st[(size_t)f * node_cap + i]; };
  const float* contrast = contrast_all + (size_t)b * node_cap;
  float* fgv = tmp_all + (size_t)b * 2 * node_cap;   // fg-ness then bg-ness scratch
  float* bgv = fgv + node_cap;
  const int64_t no = node_off ? node_off[b] : (int64_t)b * node_cap;     // dense rows when there are no offsets

  float cmn, cmx;
  block_minmax(n, [&](int i) { return contrast[i]; }, sred, cmn, cmx);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float cn = unit_norm_apply(contrast[i], cmn, cmx);
    const float dy = __fsub_rn(S(ST_PCY, i), 0.5f), dx = __fsub_rn(S(ST_PCX, i), 0.5f);
    const float cd = __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
    const float cw = expf(__fdiv_rn(-__fmul_rn(cd, cd), two_c2));
    fgv[i] = __fmul_rn(cn, cw);
  }
  __syncthreads();
  float fmn, fmx;
  block_minmax(n, [&](int i) { return fgv[i]; }, sred, fmn, fmx);

  // border colour model
  double bs = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bs += (double)S(ST_BORDER, i);
  bs = block_reduce<double>(bs, 0.0, OpAdd(), sredd);
  const float bsum = (float)bs;   // exact: integer < 2^24
  float mu[3] = {0.f, 0.f, 0.f};
  float den = 1.0f;
  if (bsum > 0.0f) {
    double m[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float w = __fdiv_rn(S(ST_BORDER, i), bsum);
#pragma unroll
      for (int c = 0; c < 3; ++c) m[c] += (double)__fmul_rn(S(ST_MEAN_L + c, i), w);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) mu[c] = (float)block_reduce<double>(m[c], 0.0, OpAdd(), sredd);
    double v[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float w = __fdiv_rn(S(ST_BORDER, i), bsum);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = __fsub_rn(S(ST_MEAN_L + c, i), mu[c]);
        v[c] += (double)__fmul_rn(__fmul_rn(d, d), w);
      }
    }
    float var = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      var = __fadd_rn(var, (float)block_reduce<double>(v[c], 0.0, OpAdd(), sredd));
    // sigma_bg = float(np.sqrt(max(var_bg, 1e-6)))
    const double sigma = ((double)var > 1e-6) ? (double)__fsqrt_rn(var) : sqrt(1e-6);
    den = (float)(2.0 * (sigma + 1e-6) * (sigma + 1e-6));
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float bg = 0.0f;
    if (bsum > 0.0f) {
      const float dl = __fsub_rn(S(ST_MEAN_L, i), mu[0]), da = __fsub_rn(S(ST_MEAN_L + 1, i), mu[1]),
                  db = __fsub_rn(S(ST_MEAN_L + 2, i), mu[2]);
      const float d = __fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db)));
      bg = expf(__fdiv_rn(-__fmul_rn(d, d), den));
    }
    const float ratio = __fdiv_rn(S(ST_BORDER, i), S(ST_SAFE, i));
    const float touch = fminf(fmaxf(__fmul_rn(ratio, 4.0f), 0.0f), 1.0f);
    bgv[i] = fmaxf(bg, touch);
  }
  __syncthreads();
  float bmn, bmx;
  block_minmax(n, [&](int i) { return bgv[i]; }, sred, bmn, bmx);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float fg = unit_norm_apply(fgv[i], fmn, fmx);
    const float bg = unit_norm_apply(bgv[i], bmn, bmx);
    float* f = x + (size_t)(no + i) * out_stride + out_col;
    f[0] = nan_to_num(fg);
    f[1] = nan_to_num(bg);
    f[2] = nan_to_num(__fsub_rn(1.0f, fabsf(__fsub_rn(fg, bg))));
  }
}

// ============================================================================ S8
// Edge attributes + COO emission in the reference order (graph_builder.py:288-322).
__global__ void __launch_bounds__(256)
k_edge_attrs(const float* __restrict__ st_all, const int2* __restrict__ pairs_all,
             const int* __restrict__ shared_all, const int* __restrict__ n_adj,
             const int* __restrict__ n_nl, const int* __restrict__ max_shared,
             const int64_t* __restrict__ edge_off, int64_t* __restrict__ edge_index,
             int64_t edge_index_stride, float* __restrict__ edge_attr, int node_cap, int pair_cap) {
  __shared__ float sred[32];
  const int b = blockIdx.x;
  const float* st = st_all + (size_t)b * ST_FIELDS * node_cap;
  auto S = [&](int f, int i) { return st[(size_t)f * node_cap + i]; };
  const int2* pairs = pairs_all + (size_t)b * pair_cap;
  const int* shared = shared_all + (size_t)b * pair_cap;
  const int na = n_adj[b], nn = n_nl[b], P = na + nn;
  const int64_t eo = edge_off[b];

  // pass 1: raw colour / centroid distances, parked in the forward rows; per-set maxima
  float mde[2] = {0.f, 0.f}, mdx[2] = {0.f, 0.f};
  for (int q = threadIdx.x; q < P; q += blockDim.x) {
    const int i = pairs[q].x, j = pairs[q].y;
    const float dl = __fsub_rn(S(ST_MEAN_L, i), S(ST_MEAN_L, j));
    const float da = __fsub_rn(S(ST_MEAN_L + 1, i), S(ST_MEAN_L + 1, j));
    const float db = __fsub_rn(S(ST_MEAN_L + 2, i), S(ST_MEAN_L + 2, j));
    const float de = __fsqrt_rn(
        __fadd_rn(__fadd_rn(__fmul_rn(dl, dl), __fmul_rn(da, da)), __fmul_rn(db, db)));
    const float dy = __fsub_rn(S(ST_CY, i), S(ST_CY, j)), dx = __fsub_rn(S(ST_CX, i), S(ST_CX, j));
    const float dxy = __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
    float* r = edge_attr + (size_t)(eo + q) * GG_N_EDGE_FEATS;
    r[0] = de;
    r[1] = dxy;
    const int set = q < na ? 0 : 1;
    mde[set] = fmaxf(mde[set], de);
    mdx[set] = fmaxf(mdx[set], dxy);
  }
  float den_de[2], den_dx[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    den_de[s] = __fadd_rn(block_reduce<float>(mde[s], 0.f, OpMaxF(), sred), 1e-6f);
    den_dx[s] = __fadd_rn(block_reduce<float>(mdx[s], 0.f, OpMaxF(), sred), 1e-6f);
  }
  __syncthreads();
  const double den_sh = (double)max_shared[b] + 1e-6;   // float64 division (…:286)
  for (int q = threadIdx.x; q < P; q += blockDim.x) {
    const int i = pairs[q].x, j = pairs[q].y;
    const int set = q < na ? 0 : 1;
    float* r = edge_attr + (size_t)(eo + q) * GG_N_EDGE_FEATS;
    float* rr = edge_attr + (size_t)(eo + P + q) * GG_N_EDGE_FEATS;
    const float a0 = __fdiv_rn(r[0], den_de[set]);
    const float a1 = __fdiv_rn(r[1], den_dx[set]);
    const float a2 = set == 0 ? (float)((double)(float)shared[q] / den_sh) : 0.0f;
    const float a3 = fabsf(__fsub_rn(S(ST_MGRADN, i), S(ST_MGRADN, j)));
    const float a4 = set == 0 ? 0.0f : 1.0f;
    r[0] = a0; r[1] = a1; r[2] = a2; r[3] = a3; r[4] = a4;
    rr[0] = a0; rr[1] = a1; rr[2] = a2; rr[3] = a3; rr[4] = a4;
    edge_index[eo + q] = i;
    edge_index[edge_index_stride + eo + q] = j;
    edge_index[eo + P + q] = j;
    edge_index[edge_index_stride + eo + P + q] = i;
  }
}

// ============================================================================ S9
// Destination-sorted CSR over the directed edges, global ids (consumed by the network).
__global__ void __launch_bounds__(512)
k_csr(const int2* __restrict__ pairs_all, const int* __restrict__ n_adj,
      const int* __restrict__ n_nl, const int* __restrict__ label_max,
      const int64_t* __restrict__ node_off, const int64_t* __restrict__ edge_off,
      int* __restrict__ cursor_all, int32_t* __restrict__ rowptr, int32_t* __restrict__ csr_src,
      int32_t* __restrict__ csr_eid, int node_cap, int pair_cap, int B) {
  __shared__ int scratch[40];
  const int b = blockIdx.x;
  const int n = min(label_max[b] + 1, node_cap);
  const int2* pairs = pairs_all + (size_t)b * pair_cap;
  const int P = n_adj[b] + n_nl[b];
  const int64_t no = node_off[b], eo = edge_off[b];
  int* cursor = cursor_all + (size_t)b * node_cap;
  int32_t* rp = rowptr + no;

  for (int i = threadIdx.x; i < n; i += blockDim.x) { rp[i] = 0; cursor[i] = 0; }
  __syncthreads();
  for (int q = threadIdx.x; q < P; q += blockDim.x) {
    atomicAdd(&rp[pairs[q].x], 1);
    atomicAdd(&rp[pairs[q].y], 1);
  }
  __syncthreads();
  int carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? rp[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, scratch, &total);
    __syncthreads();
    if (i < n) rp[i] = (int)eo + carry + ex;
    carry += total;
    __syncthreads();
  }
  if (b == B - 1 && threadIdx.x == 0) rp[n] = (int)eo + carry;   // rowptr[total nodes]
  __syncthreads();
  for (int q = threadIdx.x; q < P; q += blockDim.x) {
    const int lo = pairs[q].x, hi = pairs[q].y;
    int pos = rp[hi] + atomicAdd(&cursor[hi], 1);      // edge lo -> hi  (forward row q)
    csr_src[pos] = (int)no + lo;
    csr_eid[pos] = (int)eo + q;
    pos = rp[lo] + atomicAdd(&cursor[lo], 1);          // edge hi -> lo  (reversed row P+q)
    csr_src[pos] = (int)no + hi;
    csr_eid[pos] = (int)eo + P + q;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int s0 = rp[i], s1 = s0 + cursor[i];
    if (s1 - s0 <= CSR_LOCAL_ROW) {
      // the row fits a thread-local array: the insertion sort's dependent compare / move chain runs at
      // L1 latency instead of one L2 round trip per step (config C: 26 entries per row on average)
      unsigned long long key[CSR_LOCAL_ROW];            // (source id << 32) | edge id: sources are unique in a row
      const int len = s1 - s0;
      for (int a = 0; a < len; ++a) key[a] = ((unsigned long long)(uint32_t)csr_src[s0 + a] << 32) | (uint32_t)csr_eid[s0 + a];
      for (int a = 1; a < len; ++a) {
        const unsigned long long kv = key[a];
        int j = a - 1;
        while (j >= 0 && (key[j] >> 32) > (kv >> 32)) { key[j + 1] = key[j]; --j; }
        key[j + 1] = kv;
      }
      for (int a = 0; a < len; ++a) { csr_src[s0 + a] = (int)(key[a] >> 32); csr_eid[s0 + a] = (int)(key[a] & 0xffffffffu); }
      continue;
    }
    for (int a = s0 + 1; a < s1; ++a) {
      const int sv = csr_src[a], ev = csr_eid[a];
      int j = a - 1;
      while (j >= s0 && csr_src[j] > sv) {
        csr_src[j + 1] = csr_src[j];
        csr_eid[j + 1] = csr_eid[j];
        --j;
      }
      csr_src[j + 1] = sv;
      csr_eid[j + 1] = ev;
    }
  }
}

// The same CSR built in shared memory (one block per image, when 2 * pair_cap entries fit): row
// counts, scan, fill and the per-row insertion sort touch shared memory only -- the sort's dependent
// load / compare / store chains run at shared-memory instead of L2 latency -- and the finished arrays
// leave with coalesced stores.
__global__ void __launch_bounds__(512)
k_csr_smem(const int2* __restrict__ pairs_all, const int* __restrict__ n_adj,
           const int* __restrict__ n_nl, const int* __restrict__ label_max,
           const int64_t* __restrict__ node_off, const int64_t* __restrict__ edge_off,
           int32_t* __restrict__ rowptr, int32_t* __restrict__ csr_src,
           int32_t* __restrict__ csr_eid, int node_cap, int pair_cap, int B) {
  extern __shared__ int cs_smem[];
  __shared__ int scratch[40];
  int* s_rp = cs_smem;                       // [node_cap + 1]
  int* s_cur = s_rp + node_cap + 1;          // [node_cap]
  int* s_src = s_cur + node_cap;             // [2 * pair_cap]
  int* s_eid = s_src + 2 * pair_cap;         // [2 * pair_cap]
  const int b = blockIdx.x;
  const int n = min(label_max[b] + 1, node_cap);
  const int2* pairs = pairs_all + (size_t)b * pair_cap;
  const int P = min(n_adj[b] + n_nl[b], pair_cap);     // (an overflow is reported through the status word upstream)
  const int64_t no = node_off[b], eo = edge_off[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x) { s_rp[i] = 0; s_cur[i] = 0; }
  __syncthreads();
  for (int q = threadIdx.x; q < P; q += blockDim.x) {
    const int2 pr = pairs[q];
    atomicAdd(&s_rp[pr.x], 1);
    atomicAdd(&s_rp[pr.y], 1);
  }
  __syncthreads();
  int carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < n) ? s_rp[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, scratch, &total);
    __syncthreads();
    if (i < n) s_rp[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) s_rp[n] = carry;
  __syncthreads();
  for (int q = threadIdx.x; q < P; q += blockDim.x) {
    const int2 pr = pairs[q];
    int pos = s_rp[pr.y] + atomicAdd(&s_cur[pr.y], 1);     // edge lo -> hi  (forward row q)
    s_src[pos] = (int)no + pr.x;
    s_eid[pos] = (int)eo + q;
    pos = s_rp[pr.x] + atomicAdd(&s_cur[pr.x], 1);         // edge hi -> lo  (reversed row P+q)
    s_src[pos] = (int)no + pr.y;
    s_eid[pos] = (int)eo + P + q;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int s0 = s_rp[i], s1 = s_rp[i + 1];
    for (int a = s0 + 1; a < s1; ++a) {
      const int sv = s_src[a], ev = s_eid[a];
      int j = a - 1;
      while (j >= s0 && s_src[j] > sv) {
        s_src[j + 1] = s_src[j];
        s_eid[j + 1] = s_eid[j];
        --j;
      }
      s_src[j + 1] = sv;
      s_eid[j + 1] = ev;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) rowptr[no + i] = (int)eo + s_rp[i];
  if (b == B - 1 && threadIdx.x == 0) rowptr[no + n] = (int)eo + carry;   // rowptr[total nodes]
  for (int k = threadIdx.x; k < 2 * P; k += blockDim.x) {
    csr_src[eo + k] = s_src[k];
    csr_eid[eo + k] = s_eid[k];
  }
}

// ============================================================================ debug planes
// GraphBuilder.__init__ planes (_lab, _hsv, _gray, _grad) for parity tests.
__global__ void k_pixel_planes(const uint8_t* __restrict__ bgr, const double* __restrict__ lin_lut,
                               LabMatrix lab, int H, int W, float* __restrict__ o_lab,
                               float* __restrict__ o_hsv, float* __restrict__ o_gray,
                               float* __restrict__ o_grad) {
  __shared__ double s_lin[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lin[i] = lin_lut[i];
  __syncthreads();
  const int b = blockIdx.y;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)H * W) return;
  const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
  const uint8_t* img = bgr + (size_t)b * H * W * 3;
  const uint8_t* px = img + i * 3;
  const size_t o = (size_t)b * H * W + i;
  float L, A, Bv, hh, ss, vv;
  // the same fast paths as k_region_stats, so that the all-colours parity test covers them
  bgr_to_lab_fast(s_lin, lab.m, px[0], px[1], px[2], L, A, Bv);
  {
    const int mx = max(px[2], max(px[1], px[0])), mn = min(px[2], min(px[1], px[0]));
    hsv_hs_fast(px[0], px[1], px[2], mx, mn, hh, ss);
    vv = __fdiv_rn((float)mx, 255.0f);
  }
  if (o_lab) { o_lab[o * 3] = L; o_lab[o * 3 + 1] = A; o_lab[o * 3 + 2] = Bv; }
  if (o_hsv) { o_hsv[o * 3] = hh; o_hsv[o * 3 + 1] = ss; o_hsv[o * 3 + 2] = vv; }
  auto G = [&](int yy, int xx) {
    const uint8_t* q = img + ((size_t)reflect101(yy, H) * W + reflect101(xx, W)) * 3;
    return gray_u8(q[0], q[1], q[2]);
  };
  if (o_gray) o_gray[o] = (float)G(y, x);
  if (o_grad) {
    const int gx = (G(y - 1, x + 1) + 2 * G(y, x + 1) + G(y + 1, x + 1)) -
                   (G(y - 1, x - 1) + 2 * G(y, x - 1) + G(y + 1, x - 1));
    const int gy = (G(y + 1, x - 1) + 2 * G(y + 1, x) + G(y + 1, x + 1)) -
                   (G(y - 1, x - 1) + 2 * G(y - 1, x) + G(y - 1, x + 1));
    o_grad[o] = fsqrt_int((float)(gx * gx + gy * gy));
  }
}

// ============================================================================ training labels
// derive_trimap_labels and the fg_ratio of prepare_sample (dataset.py:175-205, 239-249):
// per region, pixels and ground-truth-foreground pixels are counted exactly (integers; one
// MATCH.ANY per 32 pixels aggregates equal labels before the atomics), the ratio is formed in
// float64 like numpy does and thresholded.
__global__ void __launch_bounds__(256)
k_mask_counts(const int32_t* __restrict__ labels, const uint8_t* __restrict__ mask,
              const int64_t* __restrict__ node_off, int HW, int* __restrict__ cnt,
              int* __restrict__ fg, int* __restrict__ status) {
  const int b = blockIdx.y;
  const int64_t n0 = node_off[b];
  const int n = (int)(node_off[b + 1] - n0);
  const int lane = threadIdx.x & 31;
  const int32_t* l = labels + (size_t)b * HW;
  const uint8_t* m = mask + (size_t)b * HW;
  for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < HW; i0 += gridDim.x * blockDim.x) {
    const int i = i0 + lane;
    const bool in = i < HW;
    const int lab = in ? l[i] : -1;
    const bool ok = in && lab >= 0 && lab < n;
    if (in && !ok) atomicOr(status, ST_LABEL_RANGE);
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    const unsigned fgm = __ballot_sync(0xffffffffu, ok && m[i] > 0);
    if (ok) {
      const unsigned grp = __match_any_sync(act, lab);
      if (lane == __ffs(grp) - 1) {
        atomicAdd(&cnt[n0 + lab], __popc(grp));
        const int f = __popc(grp & fgm);
        if (f) atomicAdd(&fg[n0 + lab], f);
      }
    }
  }
}

__global__ void k_region_labels(const int* __restrict__ cnt, const int* __restrict__ fg,
                                const int64_t* __restrict__ node_off, int n_graphs, double fg_thr,
                                double bg_cut, float* __restrict__ fg_ratio, long long* __restrict__ y) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= node_off[n_graphs]) return;
  const double c = (double)cnt[v];
  const double ratio = (double)fg[v] / fmax(c, 1.0);
  long long lab = 1;                       // CLASS_UNK
  if (ratio >= fg_thr) lab = 2;            // CLASS_FG
  if (ratio <= bg_cut) lab = 0;            // CLASS_BG (assigned second in the reference: wins)
  if (cnt[v] == 0) lab = 1;
  if (fg_ratio) fg_ratio[v] = (float)ratio;
  if (y) y[v] = lab;
}

int region_labels(gg_context* ctx, Arena& ar, const int32_t* labels, const uint8_t* mask,
                  const int64_t* node_off, int B, int H, int W, long long node_cap_total,
                  double fg_thr, double bg_thr, float* fg_ratio, long long* y, cudaStream_t st) {
  const int HW = H * W;
  int* cnt = ar.take<int>((size_t)node_cap_total);
  int* fg = ar.take<int>((size_t)node_cap_total);
  GG_CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)node_cap_total * sizeof(int), st));
  GG_CUDA_OK(cudaMemsetAsync(fg, 0, (size_t)node_cap_total * sizeof(int), st));
  GG_TRY(status_epoch(ctx, st));
  dim3 grid(std::min(ceil_div(HW, 256), 148), B);
  GG_LAUNCH(ctx, k_mask_counts, grid, 256, 0, st, labels, mask, node_off, HW, cnt, fg, ctx->status_word);
  GG_LAUNCH(ctx, k_region_labels, ceil_div(node_cap_total, 256), 256, 0, st, cnt, fg, node_off, B, fg_thr,
            1.0 - bg_thr, fg_ratio, y);
  return GG_OK;
}

// ============================================================================ self test
// The float32 fast paths of pixel_math.cuh against the IEEE intrinsics: bad[0] fsqrt_int over all
// integers < 2^24; bad[1] saturation quotients d/max; bad[2] hue quotients p/(6d) (both exhaustive);
// bad[3] normalised gradient g/(gmax+1e-6) for every Sobel magnitude and 64 image maxima each.
__global__ void k_selftest_math(unsigned long long* __restrict__ bad) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;      // grid covers [0, 2^24)
  unsigned long long b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  {
    const float a = (float)idx;
    b0 += fsqrt_int(a) != __fsqrt_rn(a);
  }
  if (idx < 256u * 256u) {
    const int d = idx >> 8, mx = idx & 255;
    if (mx >= 1 && d <= mx) b1 += fdiv_small((float)d, (float)mx) != __fdiv_rn((float)d, (float)mx);
  }
  if (idx < 256u * 1536u) {
    const int d = idx / 1536, pn = idx - d * 1536;
    if (d >= 1 && pn < 6 * d) b2 += fdiv_small((float)pn, (float)(6 * d)) != __fdiv_rn((float)pn, (float)(6 * d));
  }
  if (idx <= 2080800u) {                                             // 2 * (4 * 255)^2
    const float g = __fsqrt_rn((float)idx);
    for (unsigned j = 0; j < 64; ++j) {
      const unsigned gm = min(idx + j * j * 517u + j, 2080800u);
      const float gden = __fadd_rn(__fsqrt_rn((float)gm), 1e-6f);
      if (gden > 0.0f) b3 += fdiv_rcp(g, gden, __frcp_rn(gden)) != __fdiv_rn(g, gden);
    }
  }
  b0 = (unsigned long long)warp_sum((double)b0);
  b1 = (unsigned long long)warp_sum((double)b1);
  b2 = (unsigned long long)warp_sum((double)b2);
  b3 = (unsigned long long)warp_sum((double)b3);
  if ((threadIdx.x & 31) == 0) {
    if (b0) atomicAdd(&bad[0], b0);
    if (b1) atomicAdd(&bad[1], b1);
    if (b2) atomicAdd(&bad[2], b2);
    if (b3) atomicAdd(&bad[3], b3);
  }
}

int selftest_math(gg_context* ctx, Arena& ar, long long* mismatches, cudaStream_t st) {
  unsigned long long* bad = ar.take<unsigned long long>(4);
  GG_CUDA_OK(cudaMemsetAsync(bad, 0, 4 * sizeof(unsigned long long), st));
  GG_LAUNCH(ctx, k_selftest_math, (1 << 24) / 256, 256, 0, st, bad);
  unsigned long long hb[4];
  GG_CUDA_OK(cudaMemcpyAsync(hb, bad, sizeof(hb), cudaMemcpyDeviceToHost, st));
  GG_CUDA_OK(cudaStreamSynchronize(st));
  for (int i = 0; i < 4; ++i) mismatches[i] = (long long)hb[i];
  return GG_OK;
}

// ============================================================================ host driver
// Slots of the per-image adjacency hash table.  Only ADJACENCY pairs live in it (the non-local
// pairs are produced later, sorted, without a table): a 4-connected label map is a planar graph
// (< 3 N pairs), the diagonals of 8-connectivity add at most one crossing pair per pixel corner,
// far below 8 N in practice -- so the table is sized for the adjacency share of pair_cap, not for
// a multiple of the full pair capacity (which grows with n_nonlocal).  An overflow is reported (ST_PAIR_TABLE).
static int next_pow2(int v);
static int adjacency_table_slots(int node_cap, int pair_cap, int k_nonlocal) {
  // what the caller's pair capacity leaves for adjacency once every node has its k non-local
  // pairs, but at least 8 N (a caller with unusual label maps raises pair_cap)
  const long long adj_cap = std::min<long long>(
      pair_cap, std::max<long long>(8ll * node_cap, (long long)pair_cap - (long long)node_cap * k_nonlocal));
  return next_pow2((int)std::max<long long>(2 * adj_cap, 64));
}

// strip height of k_region_stats: the image is cut into equal strips of about 107 rows (every
// strip boundary costs one hand-over per column; shorter strips give more warps)
static int rs_rows(int H) {
  static int target = getenv("GG_RS_ROWS") ? atoi(getenv("GG_RS_ROWS")) : 107;
  const int n = ceil_div(H, target > 0 ? target : 107);
  return ceil_div(H, n);
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t graph_workspace_bytes(int B, int H, int W, const gg_graph_config& cfg) {
  const int nc = cfg.node_cap, pc = cfg.pair_cap, tc = adjacency_table_slots(nc, pc, cfg.n_nonlocal);
  const int k = cfg.n_nonlocal > 0 ? cfg.n_nonlocal : 1;
  size_t s = 0;
  s += Arena::padded((size_t)B * H * W, 1);                 // gray
  s += Arena::padded((size_t)B * nc * RS_NF, 8);            // acc
  s += Arena::padded((size_t)B * tc, 8);                    // pair keys
  s += Arena::padded((size_t)B * tc, 4);                    // pair counts
  s += Arena::padded((size_t)B * 4, 4) * 6;                 // per-image ints
  s += Arena::padded((size_t)B * ST_FIELDS * nc, 4);        // stats
  s += Arena::padded((size_t)B * pc, 8);                    // pairs
  s += Arena::padded((size_t)B * pc, 4);                    // shared counts
  s += Arena::padded((size_t)B * (nc + 1), 4) * 2;          // start arrays (adj, nl)
  s += Arena::padded((size_t)B * nc, 4);                    // cursor
  s += Arena::padded((size_t)B * nc * k, 4);                // picks
  s += Arena::padded((size_t)B * nc, 4) * 3;                // contrast + 2 tmp
  return s + 4096;
}

template <int K>
static int launch_knn(gg_context* ctx, cudaStream_t st, dim3 grid, const float* stats,
                      const int* label_max, const int2* pairs, const int* start, int* picks,
                      int node_cap, int pair_cap, int k) {
  GG_LAUNCH(ctx, k_knn<K>, grid, 256, 0, st, stats, label_max, pairs, start, picks, node_cap,
            pair_cap, k);
  return GG_OK;
}

int build_graphs(gg_context* ctx, Arena& ar, const uint8_t* bgr, const int32_t* labels, int B,
                 int H, int W, const gg_graph_config& cfg_in, const gg_graph_out& out,
                 cudaStream_t st, const uint8_t** gray_out) {
  gg_graph_config cfg = cfg_in;
  if (cfg.pair_cap <= 0) cfg.pair_cap = 8 * cfg.node_cap;
  GG_REQUIRE(B > 0 && H >= 2 && W >= 2, "build_graphs: need B>0, H,W>=2 (got %d,%d,%d)", B, H, W);
  GG_REQUIRE((long long)H * W < (1ll << 24), "build_graphs: H*W must be < 2^24");
  GG_REQUIRE(cfg.connectivity == 4 || cfg.connectivity == 8, "connectivity must be 4 or 8");
  GG_REQUIRE(cfg.n_nonlocal >= 0 && cfg.n_nonlocal <= 32, "n_nonlocal must be in [0,32]");
  GG_REQUIRE(cfg.node_cap > 0, "node_cap must be > 0");
  GG_REQUIRE(out.n_nodes && out.n_edges && out.node_off && out.edge_off && out.x &&
                 out.edge_index && out.edge_attr && out.csr_rowptr && out.csr_src && out.csr_eid,
             "build_graphs: a mandatory output pointer is NULL");
  GG_REQUIRE((long long)B * 2 * cfg.pair_cap < (1ll << 31) && (long long)B * cfg.node_cap < (1ll << 31),
             "build_graphs: batch too large for 32-bit CSR ids");
  const int nc = cfg.node_cap, pc = cfg.pair_cap, tc = adjacency_table_slots(nc, pc, cfg.n_nonlocal);
  const int k = cfg.n_nonlocal;

  // coordinate tables: a function of (H, W) alone, kept in the handle between calls.  A change of
  // shape drains the device first (concurrent sub-batches on other streams may read the old one).
  if (ctx->coord_H != H || ctx->coord_W != W) {
    GG_CUDA_OK(cudaDeviceSynchronize());
    if (ctx->d_coord) GG_CUDA_OK(cudaFree(ctx->d_coord));
    ctx->d_coord = nullptr;
    ctx->coord_H = ctx->coord_W = 0;
    GG_CUDA_OK(cudaMalloc((void**)&ctx->d_coord, (2 * (size_t)(H + W) + 2 * (size_t)(H + 1)) * sizeof(double)));
    GG_LAUNCH(ctx, k_coord_tables, ceil_div(H > W ? H : W, 256), 256, 0, st, ctx->d_coord, H, W);
    GG_CUDA_OK(cudaStreamSynchronize(st));
    ctx->coord_H = H;
    ctx->coord_W = W;
  }
  const double* coord = ctx->d_coord;
  uint8_t* gray = ar.take<uint8_t>((size_t)B * H * W);
  const double* lin = ctx->d_lin;
  double* acc = ar.take<double>((size_t)B * nc * RS_NF);
  unsigned long long* pkeys = ar.take<unsigned long long>((size_t)B * tc);
  int* pcnts = ar.take<int>((size_t)B * tc);
  int* gradmax = ar.take<int>((size_t)B * 4);
  int* label_max = ar.take<int>((size_t)B * 4);
  int* n_adj = ar.take<int>((size_t)B * 4);
  int* n_nl = ar.take<int>((size_t)B * 4);
  int* max_shared = ar.take<int>((size_t)B * 4);
  (void)ar.take<int>((size_t)B * 4);
  float* stats = ar.take<float>((size_t)B * ST_FIELDS * nc);
  int2* pairs = ar.take<int2>((size_t)B * pc);
  int* shared = ar.take<int>((size_t)B * pc);
  int* start_adj = ar.take<int>((size_t)B * (nc + 1));
  int* start_nl = ar.take<int>((size_t)B * (nc + 1));
  int* cursor = ar.take<int>((size_t)B * nc);
  int* picks = ar.take<int>((size_t)B * nc * (k > 0 ? k : 1));
  float* contrast = ar.take<float>((size_t)B * nc);
  float* tmp2 = ar.take<float>((size_t)B * nc * 2);
  if (gray_out) *gray_out = gray;

  GG_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)B * nc * RS_NF * sizeof(double), st));
  GG_CUDA_OK(cudaMemsetAsync(pkeys, 0xFF, (size_t)B * tc * sizeof(unsigned long long), st));
  GG_CUDA_OK(cudaMemsetAsync(pcnts, 0, (size_t)B * tc * sizeof(int), st));
  GG_CUDA_OK(cudaMemsetAsync(gradmax, 0, (size_t)B * 4 * sizeof(int), st));
  GG_CUDA_OK(cudaMemsetAsync(label_max, 0xFF, (size_t)B * 4 * sizeof(int), st));
  GG_TRY(status_epoch(ctx, st));

  {
    if (W % 4 == 0 && ((uintptr_t)bgr & 3) == 0 && ((uintptr_t)gray & 3) == 0) {
      dim3 grid(ceil_div(W, K0V_TX), ceil_div(H, K0V_TY), B);
      GG_LAUNCH(ctx, k_gray_gradmax_v4, grid, 256, 0, st, bgr, gray, gradmax, H, W);
    } else {
      dim3 grid(ceil_div(W, K0_TX), ceil_div(H, K0_TY), B);
      GG_LAUNCH(ctx, k_gray_gradmax, grid, 256, 0, st, bgr, gray, gradmax, H, W);
    }
  }
  {
    RegionStatsParams p;
    p.bgr = bgr; p.gray = gray; p.labels = labels; p.gradmax_sq = gradmax; p.coord = coord;
    p.lin_lut = lin; p.acc = acc; p.label_max = label_max; p.pair_keys = pkeys;
    p.pair_cnts = pcnts; p.status = ctx->status_word; p.B = B; p.H = H; p.W = W; p.node_cap = nc;
    p.table_cap = tc; p.connectivity = cfg.connectivity;
    p.rows = rs_rows(H);
    p.n_sx = ceil_div(W, 32); p.n_sy = ceil_div(H, p.rows);
    p.lab = make_lab_matrix();
    const long long tasks = (long long)B * p.n_sx * p.n_sy;
    // Direct RED.F64 hand-over is the faster kernel (0.74 vs 0.88 ms per 256 images), but its
    // ~14x more L2 atomics slow down a concurrent host-to-device copy: the streaming host path,
    // which is bound by that copy, asks for the table variant (ctx->rs_direct = 0).
    static const int forced = getenv("GG_RS_DIRECT") ? atoi(getenv("GG_RS_DIRECT")) : -1;
    // launch shape of the direct variant: 0 = 8 warps x 2 blocks/SM (128 registers), 1 = 4 warps x 5
    // blocks/SM (102 registers), 2 = 4 warps x 4 blocks/SM; +4 = keep the F2F conversions (no Veltkamp)
    static const int variant = getenv("GG_RS_VARIANT") ? atoi(getenv("GG_RS_VARIANT")) : 6;
    const int direct = forced >= 0 ? forced : ctx->rs_direct;
    // adjacency transitions: inside k_region_stats (GG_RS_PAIRS=1) or by the labels-only kernel (default)
    static const int pairs_inside = getenv("GG_RS_PAIRS") ? atoi(getenv("GG_RS_PAIRS")) : 0;
#define GG_RS_LAUNCH(BIT, W_, MINB_, DIRECT_, VELT_)                                                              \
    do {                                                                                                          \
      GG_SMEM_ATTR_ONCE(ctx, BIT, (k_region_stats<W_, MINB_, DIRECT_, VELT_, true>), rs_smem_bytes(W_, DIRECT_));  \
      GG_LAUNCH(ctx, (k_region_stats<W_, MINB_, DIRECT_, VELT_, true>), ceil_div(tasks, W_), W_ * 32,              \
                rs_smem_bytes(W_, DIRECT_), st, p);                                                               \
    } while (0)
#define GG_RS_LAUNCH_NP(BIT, W_, MINB_, DIRECT_, VELT_)                                                           \
    do {                                                                                                          \
      GG_SMEM_ATTR_ONCE(ctx, BIT, (k_region_stats<W_, MINB_, DIRECT_, VELT_, false>), rs_smem_bytes(W_, DIRECT_)); \
      GG_LAUNCH(ctx, (k_region_stats<W_, MINB_, DIRECT_, VELT_, false>), ceil_div(tasks, W_), W_ * 32,             \
                rs_smem_bytes(W_, DIRECT_), st, p);                                                               \
    } while (0)
    if (!pairs_inside) {
      AdjParams ap;
      ap.labels = labels; ap.pair_keys = pkeys; ap.pair_cnts = pcnts; ap.status = ctx->status_word;
      ap.B = B; ap.H = H; ap.W = W; ap.node_cap = nc; ap.table_cap = tc; ap.connectivity = cfg.connectivity;
      static const int adj_rows = getenv("GG_ADJ_ROWS") ? atoi(getenv("GG_ADJ_ROWS")) : 32;
      ap.rows = adj_rows > 0 ? adj_rows : 32; ap.n_sx = ceil_div(W, 32); ap.n_sy = ceil_div(H, ap.rows);
      const long long atasks = (long long)B * ap.n_sx * ap.n_sy;
      GG_LAUNCH(ctx, k_adjacency_pairs, ceil_div(atasks, 8), 256, 0, st, ap);
      if (!direct) GG_RS_LAUNCH_NP(52, 8, 2, false, true);
      else {
        // blocks per SM of the direct variant without the pair bookkeeping (register cap 128 / 102 / 85)
        static const int np_minb = getenv("GG_RS_NP_MINB") ? atoi(getenv("GG_RS_NP_MINB")) : 4;
        if (np_minb == 5) GG_RS_LAUNCH_NP(57, 4, 5, true, false);
        else if (np_minb == 6) GG_RS_LAUNCH_NP(58, 4, 6, true, false);
        else GG_RS_LAUNCH_NP(56, 4, 4, true, false);
      }
    } else
    if (!direct) GG_RS_LAUNCH(30, 8, 2, false, true);
    else if (variant == 0) GG_RS_LAUNCH(0, 8, 2, true, true);
    else if (variant == 1) GG_RS_LAUNCH(42, 4, 5, true, true);
    else if (variant == 2) GG_RS_LAUNCH(43, 4, 4, true, true);
    else if (variant == 4) GG_RS_LAUNCH(44, 8, 2, true, false);
    else if (variant == 5) GG_RS_LAUNCH(45, 4, 5, true, false);
    else GG_RS_LAUNCH(46, 4, 4, true, false);
#undef GG_RS_LAUNCH
#undef GG_RS_LAUNCH_NP
  }
  {
    dim3 grid(ceil_div(nc, 256), B);
    GG_LAUNCH(ctx, k_finalize_regions, grid, 256, 0, st, acc, label_max, stats, B, H, W, nc);
  }
  GG_LAUNCH(ctx, k_adj_sort, B, 512, 0, st, pkeys, pcnts, label_max, pairs, shared, start_adj,
            cursor, n_adj, max_shared, ctx->status_word, nc, tc, pc);
  if (k > 0) {
    dim3 grid(ceil_div(nc, 8), B);
    static const bool knn_env = getenv("GG_KNN_LEGACY") != nullptr;
    const bool knn_legacy = knn_env || ctx->knn_legacy;
    if (!knn_legacy && nc <= 2048) {
      if (nc <= 320) GG_LAUNCH(ctx, k_knn_sel<10>, grid, 256, 0, st, stats, label_max, pairs, start_adj, picks, nc, pc, k);
      else if (nc <= 512) GG_LAUNCH(ctx, k_knn_sel<16>, grid, 256, 0, st, stats, label_max, pairs, start_adj, picks, nc, pc, k);
      else if (nc <= 1024) {
        const size_t smem = (size_t)8 * 32 * (32 + 1) * sizeof(float);
        GG_LAUNCH(ctx, k_knn_sel_smem<32>, grid, 256, smem, st, stats, label_max, pairs, start_adj, picks, nc, pc, k);
      } else {
        const size_t smem = (size_t)8 * 32 * (64 + 1) * sizeof(float);
        GG_SMEM_ATTR_ONCE(ctx, 41, k_knn_sel_smem<64>, smem);
        GG_LAUNCH(ctx, k_knn_sel_smem<64>, grid, 256, smem, st, stats, label_max, pairs, start_adj, picks, nc, pc, k);
      }
    } else if (k <= 4) GG_TRY(launch_knn<4>(ctx, st, grid, stats, label_max, pairs, start_adj, picks, nc, pc, k));
    else if (k <= 8) GG_TRY(launch_knn<8>(ctx, st, grid, stats, label_max, pairs, start_adj, picks, nc, pc, k));
    else if (k <= 16) GG_TRY(launch_knn<16>(ctx, st, grid, stats, label_max, pairs, start_adj, picks, nc, pc, k));
    else GG_TRY(launch_knn<32>(ctx, st, grid, stats, label_max, pairs, start_adj, picks, nc, pc, k));
  }
  GG_LAUNCH(ctx, k_nl_pairs, B, 512, 0, st, picks, label_max, n_adj, pairs, start_nl, cursor, n_nl,
            ctx->status_word, nc, pc, k);
  GG_LAUNCH(ctx, k_offsets, 1, 1024, 0, st, label_max, n_adj, n_nl, B, nc, out.n_nodes,
            out.n_edges, out.node_off, out.edge_off, out.n_adj_pairs, out.n_nl_pairs);
  GG_LAUNCH(ctx, k_node_features, B, 256, 0, st, stats, label_max, out.node_off, out.x,
            out.centroids, out.areas, nc);
  {
    if (nc >= 1024) {
      dim3 grid(ceil_div(nc, 8 * 4), B);
      GG_LAUNCH(ctx, k_prior_contrast<4>, grid, 256, 0, st, stats, label_max, contrast, nc, (float)(2 * 0.40 * 0.40));
    } else {
      dim3 grid(ceil_div(nc, 8 * 2), B);
      GG_LAUNCH(ctx, k_prior_contrast<2>, grid, 256, 0, st, stats, label_max, contrast, nc, (float)(2 * 0.40 * 0.40));
    }
  }
  GG_LAUNCH(ctx, k_prior_finish, B, 256, 0, st, stats, label_max, contrast, tmp2, out.node_off,
            out.x, nc, (float)(2 * 0.45 * 0.45), GG_N_NODE_FEATS, GG_N_IMAGE_FEATS);
  GG_LAUNCH(ctx, k_edge_attrs, B, 256, 0, st, stats, pairs, shared, n_adj, n_nl, max_shared,
            out.edge_off, out.edge_index, (int64_t)2 * B * pc, out.edge_attr, nc, pc);
  {
    const size_t csr_smem = ((size_t)2 * nc + 1 + (size_t)4 * pc) * sizeof(int);
    if (csr_smem <= 160 * 1024) {
      GG_SMEM_ATTR_ONCE(ctx, 54, k_csr_smem, 160 * 1024);
      GG_LAUNCH(ctx, k_csr_smem, B, 512, csr_smem, st, pairs, n_adj, n_nl, label_max, out.node_off, out.edge_off,
                out.csr_rowptr, out.csr_src, out.csr_eid, nc, pc, B);
    } else {
      GG_LAUNCH(ctx, k_csr, B, 512, 0, st, pairs, n_adj, n_nl, label_max, out.node_off, out.edge_off,
                cursor, out.csr_rowptr, out.csr_src, out.csr_eid, nc, pc, B);
    }
  }
  if (out.shared_cnt)
    GG_CUDA_OK(cudaMemcpyAsync(out.shared_cnt, shared, (size_t)B * pc * sizeof(int),
                               cudaMemcpyDeviceToDevice, st));
  return GG_OK;
}

// ============================================================================ standalone prior
// compute_auto_prior(segments, lab, centre_sigma, contrast_sigma) (graph_builder.py:357-444) with a
// caller-supplied float32 Lab plane: the region sums it needs (count, Lab, float64 centroids, frame
// pixels) in the accumulator layout of k_region_stats, equal labels of a warp aggregated first.
__global__ void __launch_bounds__(256)
k_lab_plane_sums(const int32_t* __restrict__ labels, const float* __restrict__ lab, int H, int W, int node_cap,
                 double* __restrict__ acc_all, int* __restrict__ label_max, int* __restrict__ status) {
  const int b = blockIdx.y;
  const int HW = H * W;
  const int lane = threadIdx.x & 31;
  const int32_t* l = labels + (size_t)b * HW;
  const float* p = lab + (size_t)b * HW * 3;
  double* acc = acc_all + (size_t)b * node_cap * RS_NF;
  int lmax = -1;
  for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < HW; i0 += gridDim.x * blockDim.x) {
    const int i = i0 + lane;
    const bool in = i < HW;
    const int v = in ? l[i] : -1;
    const bool ok = in && v >= 0 && v < node_cap;
    if (in && !ok) atomicOr(status, ST_LABEL_RANGE);
    lmax = max(lmax, ok ? v : -1);
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (!ok) continue;
    const int y = i / W, x = i - y * W;
    double f[8] = {(double)p[3 * (size_t)i], (double)p[3 * (size_t)i + 1], (double)p[3 * (size_t)i + 2],
                   (double)y / (double)H, (double)x / (double)W, 1.0,
                   (double)((y == 0) + (y == H - 1) + (x == 0) + (x == W - 1)), 0.0};
    const unsigned grp = __match_any_sync(act, v);
    const int leader = __ffs(grp) - 1;
    // sum the group's values into its leader (sequential over the group's lanes: the groups are short)
    double s[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) s[k] = 0.0;
    for (unsigned m = grp; m; m &= m - 1) {
      const int src = __ffs(m) - 1;
#pragma unroll
      for (int k = 0; k < 7; ++k) s[k] += __shfl_sync(grp, f[k], src);
    }
    if (lane == leader) {
      double* a = acc + (size_t)v * RS_NF;
      atomicAdd(a + 0, s[0]); atomicAdd(a + 1, s[1]); atomicAdd(a + 2, s[2]);
      atomicAdd(a + 13, s[3]); atomicAdd(a + 14, s[4]); atomicAdd(a + 15, s[5]);
      if (s[6] != 0.0) atomicAdd(a + 17, s[6]);
    }
  }
  lmax = warp_max_i(lmax);
  if (lane == 0 && lmax >= 0) atomicMax(&label_max[b], lmax);
}

size_t auto_prior_workspace_bytes(int B, int node_cap) {
  return Arena::padded((size_t)B * node_cap * RS_NF, 8) + Arena::padded((size_t)B * 4, 4) +
         Arena::padded((size_t)B * ST_FIELDS * node_cap, 4) + Arena::padded((size_t)B * node_cap, 4) * 3 + 2048;
}

int auto_prior(gg_context* ctx, Arena& ar, const int32_t* labels, const float* lab, int B, int H, int W, int node_cap,
               double centre_sigma, double contrast_sigma, float* prior, int32_t* n_nodes, cudaStream_t st) {
  GG_REQUIRE(B > 0 && H >= 1 && W >= 1 && node_cap > 0, "auto_prior: bad sizes");
  double* acc = ar.take<double>((size_t)B * node_cap * RS_NF);
  int* label_max = ar.take<int>((size_t)B * 4);
  float* stats = ar.take<float>((size_t)B * ST_FIELDS * node_cap);
  float* contrast = ar.take<float>((size_t)B * node_cap);
  float* tmp2 = ar.take<float>((size_t)B * node_cap * 2);
  GG_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)B * node_cap * RS_NF * sizeof(double), st));
  GG_CUDA_OK(cudaMemsetAsync(label_max, 0xFF, (size_t)B * 4 * sizeof(int), st));
  GG_TRY(status_epoch(ctx, st));
  {
    dim3 grid(std::min(ceil_div((long long)H * W, 256), 296), B);
    GG_LAUNCH(ctx, k_lab_plane_sums, grid, 256, 0, st, labels, lab, H, W, node_cap, acc, label_max, ctx->status_word);
  }
  {
    dim3 grid(ceil_div(node_cap, 256), B);
    GG_LAUNCH(ctx, k_finalize_regions, grid, 256, 0, st, acc, label_max, stats, B, H, W, node_cap);
  }
  {
    if (node_cap >= 1024) {
      dim3 grid(ceil_div(node_cap, 8 * 4), B);
      GG_LAUNCH(ctx, k_prior_contrast<4>, grid, 256, 0, st, stats, label_max, contrast, node_cap,
                (float)(2 * contrast_sigma * contrast_sigma));
    } else {
      dim3 grid(ceil_div(node_cap, 8 * 2), B);
      GG_LAUNCH(ctx, k_prior_contrast<2>, grid, 256, 0, st, stats, label_max, contrast, node_cap,
                (float)(2 * contrast_sigma * contrast_sigma));
    }
  }
  GG_LAUNCH(ctx, k_prior_finish, B, 256, 0, st, stats, label_max, contrast, tmp2, (const int64_t*)nullptr, prior,
            node_cap, (float)(2 * centre_sigma * centre_sigma), 3, 0);
  if (n_nodes) {
    // n_nodes[b] = label_max[b] + 1 (label_max is strided by 1 here)
    GG_CUDA_OK(cudaMemcpyAsync(n_nodes, label_max, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  }
  return GG_OK;
}

int pixel_planes(gg_context* ctx, Arena& ar, const uint8_t* bgr, int B, int H, int W, float* lab,
                 float* hsv, float* gray, float* grad, cudaStream_t st) {
  (void)ar;
  const double* lin = ctx->d_lin;
  dim3 grid(ceil_div((long long)H * W, 256), B);
  GG_LAUNCH(ctx, k_pixel_planes, grid, 256, 0, st, bgr, lin, make_lab_matrix(), H, W, lab, hsv,
            gray, grad);
  return GG_OK;
}

}  // namespace gg
