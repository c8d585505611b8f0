// Region -> pixel projection: guided filter + thresholds -> OpenCV trimap.
//
// Replaces pipeline.py:71-146 (guided_filter, refine_trimap) and model.py:623-678
// (probs_to_node_trimap, project_to_pixels, _probs_to_trimap) of the reference.
//
// cv2.blur on float32 planes = float64 window sums over BORDER_REFLECT_101, times 1/k^2,
// cast to float32; the element-wise guided-filter algebra is float32 with
// separate multiply / add roundings.  Both are reproduced: window sums are float64
// sliding sums (vertical pass, then horizontal pass, through shared memory), the algebra
// uses round-to-nearest intrinsics that are never contracted into FMAs.
//
// Two column-walker kernels: (1) means of {g, g^2, s_c, g s_c} -> a_c, b_c planes;
//                            (2) means of {a_c, b_c} -> q_c = mean(a_c) g + mean(b_c) -> clip -> labels.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "pixel_math.cuh"
#include "trimap.cuh"

namespace gg {

constexpr int GF_MAX_STRIP = 512;   // output rows per block (strip height), runtime <= this
constexpr int GF_MAX_RADIUS = 24;
constexpr int GF_PROB_TABLE = 2048;  // most labels whose (p_bg, p_fg) are staged in (dynamic) shared memory
constexpr int GF_PROB_DEFAULT = 1024;  // ... when the caller gives no bound on the labels of an image

struct GuidedParams {
  // trimap mode
  const uint8_t* gray;      // [B,H,W]
  const int32_t* labels;    // [B,H,W]
  const float* probs;       // [SN,3]
  const int64_t* node_off;  // [B+1]
  // plane mode
  const float* guide;       // [B,H,W]
  const float* src;         // [B,H,W]
  // stage-1 outputs / stage-2 inputs, interleaved per pixel: [B,H,W][a_0, b_0(, a_1, b_1)]
  // float32 -- one 16-byte (trimap mode) / 8-byte (plane mode) access per pixel
  float* ab;
  size_t plane_stride;      // B*H*W (pixels)
  // stage-2 outputs
  uint8_t* trimap;
  float* q0;                // optional filtered planes (p_bg / out)
  float* q1;                // optional (p_fg)
  int H, W, radius, strip;
  int prob_cap;             // entries of the shared (p_bg, p_fg) table (dynamic shared memory of stage 1)
  float eps, thr_fg, thr_bg;
};

template <bool kTrimap>
struct GfTraits {
  static constexpr int NSRC = kTrimap ? 2 : 1;
  static constexpr int NPLANE1 = 2 + 2 * NSRC;   // g, g^2, s_c, g*s_c
  static constexpr int NPLANE2 = 2 * NSRC;       // a_c, b_c
};

// Column walker.  Thread t owns halo column x0 - r + t of a strip and walks down its rows with
// float64 vertical running sums of all NP planes in registers (add row y+r, subtract row
// y-r-1; old rows are re-read through L1/L2; the reflected row offsets of the strip come from a
// small shared table).  Per row the sums go to a shared row buffer and (plane, segment) tasks
// slide the horizontal window over it; segment lengths are odd so that the 64-bit reads of a
// warp hit distinct banks.  Means = float32(sum * 1/k^2), exactly cv2.blur's float32 path.
// RT > 0: radius known at compile time (loops unrolled); RT == 0: runtime radius.
// NT = threads per block = halo columns per block (128 or 256, whichever wastes fewer columns).
template <int NP, int NT>
struct HorizontalPlan {
  int seglen, nseg, tx;
  int off, len;          // this thread's task: elements [off, off+len) of plane-major row buffers
  __device__ HorizontalPlan(int radius) {
    tx = NT - 2 * radius;
    const int per = NT / NP;
    seglen = ((tx + per - 1) / per) | 1;
    nseg = (tx + seglen - 1) / seglen;
    const int t = threadIdx.x;
    const int pl = t / nseg, seg = t - pl * nseg;
    const int o0 = seg * seglen;
    off = pl * NT + o0;
    len = (t < NP * nseg) ? min(seglen, tx - o0) : 0;
  }
};

template <int NP, int NT, int RT>
GG_D void horizontal_means(const double* sV, float* sM, const HorizontalPlan<NP, NT>& hp, int radius,
                           double scale) {
  // sV / sM: [NP][NT]; one task = one (plane, segment)
  if (hp.len > 0) {
    const int len = hp.len;
    const double* in = sV + hp.off;
    float* out = sM + hp.off;
    if (RT > 0) {
      constexpr int K = 2 * RT + 1;
      constexpr int PER = NT / NP;
      constexpr int SEG = (((NT - 2 * RT) + PER - 1) / PER) | 1;      // == hp.seglen
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < K; ++d) s += in[d];
      out[0] = (float)(s * scale);
#pragma unroll
      for (int i = 1; i < SEG; ++i) {
        if (i < len) {
          s += in[i + K - 1];
          s -= in[i - 1];
          out[i] = (float)(s * scale);
        }
      }
    } else {
      const int k = 2 * radius + 1;
      double s = 0.0;
      for (int d = 0; d < k; ++d) s += in[d];
      out[0] = (float)(s * scale);
      for (int i = 1; i < len; ++i) {
        s += in[i + k - 1];
        s -= in[i - 1];
        out[i] = (float)(s * scale);
      }
    }
  }
}

// row index of BORDER_REFLECT_101 for |overshoot| < n (one fold), else the general loop
GG_D int reflect_row(int y, int n) {
  if (y < 0) y = -y;
  if (y >= n) y = 2 * n - 2 - y;
  return (y < 0 || y >= n) ? reflect101(y, n) : y;
}

// element offsets (row * W) of rows y_begin - r - 1 .. y_end + r - 1 of the strip, reflected
GG_D void fill_row_table(int* sRow, int y_begin, int y_end, int r, int H, int W) {
  const int n = (y_end - y_begin) + 2 * r + 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sRow[i] = reflect_row(y_begin - r - 1 + i, H) * W;
}

// float32 division whose zero numerators (very common: cov == 0 wherever the posterior is constant
// over the window) do not take the slow path of __fdiv_rn; the sign of a zero is kept.
GG_D float fdiv_zero_guard(float num, float den) {
  const float q = __fdiv_rn(num == 0.0f ? den : num, den);
  return num == 0.0f ? num : q;
}

template <bool kTrimap, int RT, int NT>
__global__ void __launch_bounds__(NT)
k_guided_ab(const GuidedParams p) {
  using T = GfTraits<kTrimap>;
  constexpr int NP = T::NPLANE1;
  constexpr int NS = T::NSRC;
  __shared__ double sV[NP * NT];
  __shared__ float sM[NP * NT];
  __shared__ int sRow[GF_MAX_STRIP + 2 * GF_MAX_RADIUS + 2];
  __shared__ float sLut[kTrimap ? 256 : 1];
  extern __shared__ float2 sProb[];                         // [p.prob_cap] (p_bg, p_fg) per label, if it fits
  const int r = RT > 0 ? RT : p.radius, H = p.H, W = p.W, b = blockIdx.z, t = threadIdx.x;
  const HorizontalPlan<NP, NT> hp(r);
  const int x0 = blockIdx.x * hp.tx;
  const int y_begin = blockIdx.y * p.strip, y_end = min(H, y_begin + p.strip);
  const int xs = reflect101(x0 - r + t, W);
  const double scale = 1.0 / ((double)(2 * r + 1) * (double)(2 * r + 1));
  const size_t img_off = (size_t)b * H * W;
  // column base pointers (element (y, xs) lives at col[y * W])
  const uint8_t* gcol = kTrimap ? p.gray + img_off + xs : nullptr;
  const int32_t* lcol = kTrimap ? p.labels + img_off + xs : nullptr;
  const float* ucol = kTrimap ? nullptr : p.guide + img_off + xs;
  const float* scol = kTrimap ? nullptr : p.src + img_off + xs;
  const float* prob0 = nullptr;
  int nn = 0;
  fill_row_table(sRow, y_begin, y_end, r, H, W);
  if (kTrimap) {
    const int64_t no = p.node_off[b];
    nn = (int)(p.node_off[b + 1] - no);
    prob0 = p.probs + (size_t)no * 3;
    for (int i = t; i < 256; i += NT) sLut[i] = __fdiv_rn((float)i, 255.0f);   // guide = gray/255
    if (nn <= p.prob_cap)
      for (int i = t; i < nn; i += NT) sProb[i] = make_float2(prob0[(size_t)i * 3], prob0[(size_t)i * 3 + 2]);
  }
  __syncthreads();
  const bool prob_in_smem = kTrimap && nn <= p.prob_cap;
  const int* rowtab = sRow + (r + 1) - y_begin;     // rowtab[yy] for yy in [y_begin-r-1, y_end+r)

  // base planes of one pixel of this thread's column: g and the NSRC source planes.  Two steps, so
  // that a prefetch only ISSUES the global loads (raw grey byte + label, or the two floats) and the
  // look-ups that depend on them run one iteration later, when the data has arrived.
  struct Px { float g, sv[NS]; };
  struct Raw { int gr, l; float u, s; };
  auto fetch_raw = [&](int yy) {
    const int ro = rowtab[yy];
    Raw rw;
    if (kTrimap) {
      rw.gr = gcol[ro];
      rw.l = lcol[ro];
      rw.u = rw.s = 0.0f;
    } else {
      rw.gr = rw.l = 0;
      rw.u = ucol[ro];
      rw.s = scol[ro];
    }
    return rw;
  };
  auto resolve = [&](const Raw& rw, Px& px) {
    if (kTrimap) {
      px.g = sLut[rw.gr];
      px.sv[0] = 0.0f;
      px.sv[NS - 1] = 0.0f;                            // project_to_pixels zero padding
      if (rw.l >= 0 && rw.l < nn) {
        if (prob_in_smem) {
          const float2 pr = sProb[rw.l];
          px.sv[0] = pr.x;
          px.sv[NS - 1] = pr.y;
        } else {
          const float* row = prob0 + (size_t)rw.l * 3;
          px.sv[0] = row[0];
          px.sv[NS - 1] = row[2];
        }
      }
    } else {
      px.g = rw.u;
      px.sv[0] = rw.s;
    }
  };
  double vs[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) vs[q] = 0.0;
  auto add = [&](const Px& px) {
    vs[0] += (double)px.g;
    vs[1] += (double)__fmul_rn(px.g, px.g);
#pragma unroll
    for (int c = 0; c < NS; ++c) {
      vs[2 + 2 * c] += (double)px.sv[c];
      vs[3 + 2 * c] += (double)__fmul_rn(px.g, px.sv[c]);
    }
  };
  auto sub = [&](const Px& px) {
    vs[0] -= (double)px.g;
    vs[1] -= (double)__fmul_rn(px.g, px.g);
#pragma unroll
    for (int c = 0; c < NS; ++c) {
      vs[2 + 2 * c] -= (double)px.sv[c];
      vs[3 + 2 * c] -= (double)__fmul_rn(px.g, px.sv[c]);
    }
  };
  for (int yy = y_begin - r; yy < y_begin + r; yy += 4) {       // window warm-up: loads of four rows in flight
    Raw rw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) rw[k] = fetch_raw(min(yy + k, y_begin + r - 1));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (yy + k < y_begin + r) {
        Px px;
        resolve(rw[k], px);
        add(px);
      }
    }
  }
  // software pipeline: the rows entering / leaving the window of the NEXT row are in flight
  // while the current row is reduced
  Raw rn = fetch_raw(y_begin + r), ro_ = rn;
  for (int y = y_begin; y < y_end; ++y) {
    Px cn, co;
    resolve(rn, cn);
    if (y > y_begin) resolve(ro_, co);
    if (y + 1 < y_end) {
      rn = fetch_raw(y + 1 + r);
      ro_ = fetch_raw(y - r);
    }
    add(cn);
    if (y > y_begin) sub(co);
#pragma unroll
    for (int q = 0; q < NP; ++q) sV[q * NT + t] = vs[q];
    __syncthreads();
    horizontal_means<NP, NT, RT>(sV, sM, hp, r, scale);
    __syncthreads();
    const int o = t - r, x = x0 + o;
    if (o >= 0 && o < hp.tx && x < W) {
      const float* m = sM + o;
      const float mg = m[0], mgg = m[NT];
      const float var = __fsub_rn(mgg, __fmul_rn(mg, mg));
      const float den = __fadd_rn(var, p.eps);
      const size_t op = img_off + (size_t)y * W + x;
      float abv[2 * NS];
#pragma unroll
      for (int c = 0; c < NS; ++c) {
        const float ms = m[(2 + 2 * c) * NT], mgs = m[(3 + 2 * c) * NT];
        const float cov = __fsub_rn(mgs, __fmul_rn(mg, ms));
        const float a = fdiv_zero_guard(cov, den);
        abv[2 * c] = a;
        abv[2 * c + 1] = __fsub_rn(ms, __fmul_rn(a, mg));
      }
      if (NS == 2) *reinterpret_cast<float4*>(p.ab + op * 4) = make_float4(abv[0], abv[1], abv[2 * NS - 2], abv[2 * NS - 1]);
      else *reinterpret_cast<float2*>(p.ab + op * 2) = make_float2(abv[0], abv[1]);
    }
  }
}

template <bool kTrimap, int RT, int NT>
__global__ void __launch_bounds__(NT)
k_guided_out(const GuidedParams p) {
  using T = GfTraits<kTrimap>;
  constexpr int NP = T::NPLANE2;
  __shared__ double sV[NP * NT];
  __shared__ float sM[NP * NT];
  __shared__ int sRow[GF_MAX_STRIP + 2 * GF_MAX_RADIUS + 2];
  __shared__ float sLut[kTrimap ? 256 : 1];
  const int r = RT > 0 ? RT : p.radius, H = p.H, W = p.W, b = blockIdx.z, t = threadIdx.x;
  const HorizontalPlan<NP, NT> hp(r);
  const int x0 = blockIdx.x * hp.tx;
  const int y_begin = blockIdx.y * p.strip, y_end = min(H, y_begin + p.strip);
  const int xs = reflect101(x0 - r + t, W);
  const size_t img_off = (size_t)b * H * W;
  const double scale = 1.0 / ((double)(2 * r + 1) * (double)(2 * r + 1));
  const float* col = p.ab + (img_off + xs) * NP;          // NP interleaved values per pixel
  fill_row_table(sRow, y_begin, y_end, r, H, W);
  if (kTrimap)
    for (int i = t; i < 256; i += NT) sLut[i] = __fdiv_rn((float)i, 255.0f);
  __syncthreads();
  const int* rowtab = sRow + (r + 1) - y_begin;
  double vs[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) vs[q] = 0.0;
  auto load_row = [&](int yy, float (&v)[NP]) {
    const float* rp = col + (size_t)rowtab[yy] * NP;
    if (NP == 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(rp));
      v[0] = t4.x; v[1] = t4.y; v[NP - 2] = t4.z; v[NP - 1] = t4.w;
    } else {
      const float2 t2 = __ldg(reinterpret_cast<const float2*>(rp));
      v[0] = t2.x; v[1] = t2.y;
    }
  };
  for (int yy = y_begin - r; yy < y_begin + r; yy += 4) {       // window warm-up: loads of four rows in flight
    float v[4][NP];
#pragma unroll
    for (int k = 0; k < 4; ++k) load_row(min(yy + k, y_begin + r - 1), v[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (yy + k < y_begin + r) {
#pragma unroll
        for (int q = 0; q < NP; ++q) vs[q] += (double)v[k][q];
      }
    }
  }
  float cn[NP], co[NP], nn_[NP], no_[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) co[q] = nn_[q] = no_[q] = 0.f;
  load_row(y_begin + r, cn);
  const int o = t - r, x = x0 + o;
  const bool writer = o >= 0 && o < hp.tx && x < W;
  for (int y = y_begin; y < y_end; ++y) {
    if (y + 1 < y_end) {
      load_row(y + 1 + r, nn_);
      load_row(y - r, no_);
    }
    // guide value of the output pixel: issued here, consumed after the two barriers
    const size_t op = img_off + (size_t)y * W + (writer ? x : 0);
    int g_raw = 0;
    float g_val = 0.0f;
    if (writer) { if (kTrimap) g_raw = p.gray[op]; else g_val = p.guide[op]; }
#pragma unroll
    for (int q = 0; q < NP; ++q) vs[q] += (double)cn[q];
    if (y > y_begin) {
#pragma unroll
      for (int q = 0; q < NP; ++q) vs[q] -= (double)co[q];
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) sV[q * NT + t] = vs[q];
    __syncthreads();
    horizontal_means<NP, NT, RT>(sV, sM, hp, r, scale);
    __syncthreads();
    if (writer) {
      const float* m = sM + o;
      const float g = kTrimap ? sLut[g_raw] : g_val;
      float q[T::NSRC];
#pragma unroll
      for (int c = 0; c < T::NSRC; ++c)
        q[c] = __fadd_rn(__fmul_rn(m[(2 * c) * NT], g), m[(2 * c + 1) * NT]);
      if (kTrimap) {
        const float pbg = fminf(fmaxf(q[0], 0.0f), 1.0f);          // np.clip(., 0, 1)
        const float pfg = fminf(fmaxf(q[T::NSRC - 1], 0.0f), 1.0f);
        uint8_t tv = (pfg > pbg) ? 3 : 2;                          // pipeline.py:143-145
        if (pbg >= p.thr_bg) tv = 0;
        if (pfg >= p.thr_fg) tv = 1;
        p.trimap[op] = tv;
        if (p.q0) p.q0[op] = pbg;
        if (p.q1) p.q1[op] = pfg;
      } else {
        p.q0[op] = q[0];
      }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) { cn[q] = nn_[q]; co[q] = no_[q]; }
  }
}

// Block width: 128 or 256 halo columns, whichever covers W with fewer wasted columns; strip
// height: the image is cut into equal strips of at most GF_MAX_STRIP rows.
static int gf_block_width(int W, int radius) {
  const long long w128 = (long long)ceil_div(W, 128 - 2 * radius) * 128;
  const long long w256 = (long long)ceil_div(W, 256 - 2 * radius) * 256;
  return w256 <= w128 ? 256 : 128;
}
static int gf_strip(int H) {
  static int forced = getenv("GG_GF_STRIP") ? atoi(getenv("GG_GF_STRIP")) : 0;
  const int target = forced > 0 ? std::min(forced, GF_MAX_STRIP) : 64;
  const int n = ceil_div(H, target);
  return std::min(GF_MAX_STRIP, ceil_div(H, n));
}

template <bool kTrimap, int NT>
static int launch_guided_ab(gg_context* ctx, GuidedParams p, int B, cudaStream_t st) {
  p.strip = gf_strip(p.H);
  const size_t dyn = kTrimap ? (size_t)p.prob_cap * sizeof(float2) : 0;
  dim3 grid(ceil_div(p.W, NT - 2 * p.radius), ceil_div(p.H, p.strip), B);
  if (p.radius == 8) GG_LAUNCH(ctx, (k_guided_ab<kTrimap, 8, NT>), grid, NT, dyn, st, p);
  else if (p.radius == 4) GG_LAUNCH(ctx, (k_guided_ab<kTrimap, 4, NT>), grid, NT, dyn, st, p);
  else GG_LAUNCH(ctx, (k_guided_ab<kTrimap, 0, NT>), grid, NT, dyn, st, p);
  return GG_OK;
}

template <bool kTrimap, int NT>
static int launch_guided_out(gg_context* ctx, GuidedParams p, int B, cudaStream_t st) {
  p.strip = gf_strip(p.H);
  dim3 grid(ceil_div(p.W, NT - 2 * p.radius), ceil_div(p.H, p.strip), B);
  if (p.radius == 8) GG_LAUNCH(ctx, (k_guided_out<kTrimap, 8, NT>), grid, NT, 0, st, p);
  else if (p.radius == 4) GG_LAUNCH(ctx, (k_guided_out<kTrimap, 4, NT>), grid, NT, 0, st, p);
  else GG_LAUNCH(ctx, (k_guided_out<kTrimap, 0, NT>), grid, NT, 0, st, p);
  return GG_OK;
}

// Measured on B200 (320x480, r=8): stage 1 (6 planes, more registers and shared memory per
// thread) is fastest with 128-column blocks, stage 2 (4 planes) with 256-column blocks.
template <bool kTrimap>
static int launch_guided(gg_context* ctx, const GuidedParams& p, int B, cudaStream_t st) {
  static int f_ab = getenv("GG_GF_NT_AB") ? atoi(getenv("GG_GF_NT_AB")) : 0;
  static int f_out = getenv("GG_GF_NT_OUT") ? atoi(getenv("GG_GF_NT_OUT")) : 0;
  const int nt_ab = f_ab == 256 ? 256 : 128;
  const int nt_out = (f_out == 128 || f_out == 256) ? f_out : gf_block_width(p.W, p.radius);
  if (nt_ab == 256) GG_TRY((launch_guided_ab<kTrimap, 256>(ctx, p, B, st)));
  else GG_TRY((launch_guided_ab<kTrimap, 128>(ctx, p, B, st)));
  if (nt_out == 256) GG_TRY((launch_guided_out<kTrimap, 256>(ctx, p, B, st)));
  else GG_TRY((launch_guided_out<kTrimap, 128>(ctx, p, B, st)));
  return GG_OK;
}

// predict_trimap / _probs_to_trimap (model.py:623-678): node rule gathered through the map.
__global__ void k_project_trimap(const int32_t* __restrict__ labels, const float* __restrict__ probs,
                                 const int64_t* __restrict__ node_off, int HW, float thr_fg,
                                 float thr_bg, uint8_t* __restrict__ trimap) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const int64_t no = node_off[b];
  const int nn = (int)(node_off[b + 1] - no);
  const size_t o = (size_t)b * HW + i;
  const int l = labels[o];
  uint8_t t = 2;                                                   // GC_PR_BGD padding
  if (l >= 0 && l < nn) {
    const float* row = probs + (size_t)(no + l) * 3;
    const float bg = row[0], fg = row[2];
    t = (fg > bg) ? 3 : 2;
    if (bg >= thr_bg) t = 0;
    if (fg >= thr_fg) t = 1;
  }
  trimap[o] = t;
}

__global__ void k_gray_only(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = bgr + i * 3;
  gray[i] = (uint8_t)gray_u8(p[0], p[1], p[2]);
}

// ----------------------------------------------------------------------------- seeding
// _seed_from_prior (pipeline.py:149-186): a trimap without any foreground label (1, 3) or
// without any background label (0, 2) gets the ceil-ish 10 % most confident regions of the
// automatic prior promoted to the missing side (FG_PROBABLE = 3 / BG_PROBABLE = 2).
// flags[b]: bit0 = some pixel is FG_DEFINITE / FG_PROBABLE, bit1 = some pixel is BG_*.
__global__ void __launch_bounds__(256)
k_trimap_sides(const uint8_t* __restrict__ trimap, int HW, int* __restrict__ flags) {
  const int b = blockIdx.y;
  const uint8_t* t = trimap + (size_t)b * HW;
  int bits = 0;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 16; i < HW; i += gridDim.x * blockDim.x * 16) {
    if (i + 16 <= HW && (((uintptr_t)(t + i)) & 15) == 0) {
      const uint4 v = *reinterpret_cast<const uint4*>(t + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t odd = w[k] & 0x01010101u;          // labels 1 and 3 are odd, 0 and 2 even
        bits |= odd ? 1 : 0;
        bits |= (odd != 0x01010101u) ? 2 : 0;
      }
    } else {
      for (int k = i; k < min(HW, i + 16); ++k) bits |= (t[k] & 1) ? 1 : 2;
    }
  }
  bits = __reduce_or_sync(0xffffffffu, bits);
  if ((threadIdx.x & 31) == 0 && bits) atomicOr(&flags[b], bits);
}

// Block per image; does nothing when both sides are present.  sel[v] (node-major, the image's
// slice of a [SN] array): 1 = promote to FG_PROBABLE, 2 = promote to BG_PROBABLE, 3 = both (the
// background assignment comes second in the reference and wins), 0 = untouched.  Rank by the
// prior column, descending; equal values: the larger region index first (a stable ascending
// argsort, reversed).
__global__ void __launch_bounds__(256)
k_seed_select(const float* __restrict__ x, const int64_t* __restrict__ node_off,
              const int* __restrict__ flags, double seed_frac, uint8_t* __restrict__ sel) {
  const int b = blockIdx.x;
  const int f = flags[b];
  const int64_t n0 = node_off[b];
  const int n = (int)(node_off[b + 1] - n0);
  if (n <= 0) return;
  uint8_t* s = sel + n0;
  if ((f & 3) == 3) return;                     // k_seed_apply never reads sel for this image
  const int n_seed = max(1, (int)rint(seed_frac * (double)n));
  const float* pr = x + (size_t)n0 * GG_N_NODE_FEATS + GG_N_IMAGE_FEATS;     // prior columns 16..18
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int out = 0;
#pragma unroll
    for (int side = 0; side < 2; ++side) {      // 0: foreground prior (column 0), 1: background (column 1)
      if (f & (1 << side)) continue;
      const float pi = pr[(size_t)i * GG_N_NODE_FEATS + side];
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float pj = pr[(size_t)j * GG_N_NODE_FEATS + side];
        rank += (pj > pi) || (pj == pi && j > i);
      }
      if (rank < n_seed) out |= 1 << side;
    }
    s[i] = (uint8_t)out;
  }
}

__global__ void __launch_bounds__(256)
k_seed_apply(const int32_t* __restrict__ labels, const int64_t* __restrict__ node_off,
             const int* __restrict__ flags, const uint8_t* __restrict__ sel, int HW,
             uint8_t* __restrict__ trimap) {
  const int b = blockIdx.y;
  if ((flags[b] & 3) == 3) return;
  const int64_t n0 = node_off[b];
  const int n = (int)(node_off[b + 1] - n0);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const size_t o = (size_t)b * HW + i;
  const int l = labels[o];
  if (l < 0 || l >= n) return;
  const int sv = sel[n0 + l];
  if (sv & 2) trimap[o] = 2;                    // BG_PROBABLE (applied last in the reference)
  else if (sv & 1) trimap[o] = 3;               // FG_PROBABLE
}

// ----------------------------------------------------------------------------- GrabCut hand-off guards
// GrabCut.run_with_trimap (grabcut.py:127-140): cv2.grabCut needs at least one definite foreground
// and one definite background pixel.  An image without GC_FGD gets its GC_PR_FGD pixels promoted,
// likewise GC_PR_BGD -> GC_BGD; an image that still lacks one side afterwards is degenerate (the
// reference then returns the trimap's own labelling instead of calling cv2.grabCut).
// present[b]: bit l set = label l occurs in image b.
__global__ void __launch_bounds__(256)
k_trimap_labels_present(const uint8_t* __restrict__ trimap, int HW, int* __restrict__ present) {
  const int b = blockIdx.y;
  const uint8_t* t = trimap + (size_t)b * HW;
  int bits = 0;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 16; i < HW; i += gridDim.x * blockDim.x * 16) {
    if (i + 16 <= HW && (((uintptr_t)(t + i)) & 15) == 0) {
      const uint4 v = *reinterpret_cast<const uint4*>(t + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) bits |= 1 << ((w[k] >> (8 * j)) & 3);
    } else {
      for (int k = i; k < min(HW, i + 16); ++k) bits |= 1 << (t[k] & 3);
    }
  }
  bits = __reduce_or_sync(0xffffffffu, bits);
  if ((threadIdx.x & 31) == 0 && bits) atomicOr(&present[b], bits);
}

__global__ void __launch_bounds__(256)
k_trimap_promote(uint8_t* __restrict__ trimap, int HW, const int* __restrict__ present,
                 int32_t* __restrict__ degenerate) {
  const int b = blockIdx.y;
  const int pr = present[b];
  const bool no_fgd = !(pr & 2), no_bgd = !(pr & 1);
  if (blockIdx.x == 0 && threadIdx.x == 0 && degenerate)
    degenerate[b] = (!(pr & (2 | 8)) || !(pr & (1 | 4))) ? 1 : 0;     // a side is missing even after promotion
  if (!no_fgd && !no_bgd) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  uint8_t* t = trimap + (size_t)b * HW;
  const uint8_t v = t[i];
  if (no_fgd && v == 3) t[i] = 1;
  else if (no_bgd && v == 2) t[i] = 0;
}

int grabcut_guards(gg_context* ctx, Arena& ar, uint8_t* trimap, int B, int H, int W, int32_t* degenerate,
                   cudaStream_t st) {
  const int HW = H * W;
  int* present = ar.take<int>((size_t)B);
  GG_CUDA_OK(cudaMemsetAsync(present, 0, (size_t)B * sizeof(int), st));
  dim3 g1(std::min(ceil_div(HW, 256 * 16), 64), B);
  GG_LAUNCH(ctx, k_trimap_labels_present, g1, 256, 0, st, trimap, HW, present);
  dim3 g2(ceil_div(HW, 256), B);
  GG_LAUNCH(ctx, k_trimap_promote, g2, 256, 0, st, trimap, HW, present, degenerate);
  return GG_OK;
}

// ----------------------------------------------------------------------------- mask clean-up
// clean_mask (pipeline.py:189-227): 8-connected components of a binary mask, drop the components
// smaller than min_area_ratio * H * W (keep the largest one if nothing survives), or keep only the
// largest.  Label-equivalence union-find on the pixel grid: a component's root is its smallest
// pixel index = its first pixel in raster order, so "the first of the largest components" is the
// one cv2.connectedComponentsWithStats + argmax picks.
GG_D int ccl_find(int* L, int i) {
  int r = i;
  while (true) {
    const int pr = reinterpret_cast<volatile int*>(L)[r];
    if (pr == r) break;
    r = pr;
  }
  return r;
}
GG_D void ccl_union(int* L, int a, int b) {
  while (true) {
    a = ccl_find(L, a);
    b = ccl_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }          // hook the larger root under the smaller
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

__global__ void __launch_bounds__(256)
k_ccl_init(const uint8_t* __restrict__ mask, int HW, int* __restrict__ L, int* __restrict__ area) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const size_t o = (size_t)b * HW + i;
  L[o] = mask[o] ? i : -1;
  area[o] = 0;
}

__global__ void __launch_bounds__(256)
k_ccl_merge(const uint8_t* __restrict__ mask, int H, int W, int* __restrict__ L) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (i >= HW) return;
  const uint8_t* m = mask + (size_t)b * HW;
  if (!m[i]) return;
  int* Lb = L + (size_t)b * HW;
  const int y = i / W, x = i - y * W;
  if (x > 0 && m[i - 1]) ccl_union(Lb, i, i - 1);
  if (y > 0) {
    if (m[i - W]) ccl_union(Lb, i, i - W);
    if (x > 0 && m[i - W - 1]) ccl_union(Lb, i, i - W - 1);
    if (x + 1 < W && m[i - W + 1]) ccl_union(Lb, i, i - W + 1);
  }
}

__global__ void __launch_bounds__(256)
k_ccl_flatten(int HW, int* __restrict__ L, int* __restrict__ area) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  int* Lb = L + (size_t)b * HW;
  if (Lb[i] < 0) return;
  const int r = ccl_find(Lb, i);
  Lb[i] = r;
  atomicAdd(&area[(size_t)b * HW + r], 1);
}

// best[b]: (area << 32) | (0xffffffff - root) of the largest component (smallest root among equals);
// any_big[b]: some component reaches min_area
__global__ void __launch_bounds__(256)
k_ccl_select(int HW, const int* __restrict__ L, const int* __restrict__ area, double min_area,
             unsigned long long* __restrict__ best, int* __restrict__ any_big) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const size_t o = (size_t)b * HW + i;
  if (L[o] != i) return;                                     // roots only
  const int a = area[o];
  atomicMax(&best[b], ((unsigned long long)(unsigned)a << 32) | (unsigned long long)(0xffffffffu - (unsigned)i));
  if ((double)a >= min_area) any_big[b] = 1;
}

__global__ void __launch_bounds__(256)
k_ccl_apply(int HW, const int* __restrict__ L, const int* __restrict__ area, double min_area, int keep_largest,
            const unsigned long long* __restrict__ best, const int* __restrict__ any_big,
            uint8_t* __restrict__ out) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const size_t o = (size_t)b * HW + i;
  const int r = L[o];
  uint8_t keep = 0;
  if (r >= 0) {
    const int largest = (int)(0xffffffffu - (unsigned)(best[b] & 0xffffffffull));
    if (keep_largest || !any_big[b]) keep = r == largest;
    else keep = (double)area[(size_t)b * HW + r] >= min_area;
  }
  out[o] = keep;
}

size_t clean_workspace_bytes(int B, int H, int W) {
  return 2 * Arena::padded((size_t)B * H * W, 4) + Arena::padded((size_t)B, 8) + Arena::padded((size_t)B, 4) + 1024;
}

int clean_masks(gg_context* ctx, Arena& ar, const uint8_t* mask, uint8_t* out, int B, int H, int W,
                double min_area_ratio, int keep_largest, cudaStream_t st) {
  const int HW = H * W;
  int* L = ar.take<int>((size_t)B * HW);
  int* area = ar.take<int>((size_t)B * HW);
  unsigned long long* best = ar.take<unsigned long long>((size_t)B);
  int* any_big = ar.take<int>((size_t)B);
  GG_CUDA_OK(cudaMemsetAsync(best, 0, (size_t)B * sizeof(unsigned long long), st));
  GG_CUDA_OK(cudaMemsetAsync(any_big, 0, (size_t)B * sizeof(int), st));
  const double min_area = min_area_ratio * (double)HW;       // min_area_ratio * mask.size (float64)
  dim3 grid(ceil_div(HW, 256), B);
  GG_LAUNCH(ctx, k_ccl_init, grid, 256, 0, st, mask, HW, L, area);
  GG_LAUNCH(ctx, k_ccl_merge, grid, 256, 0, st, mask, H, W, L);
  GG_LAUNCH(ctx, k_ccl_flatten, grid, 256, 0, st, HW, L, area);
  GG_LAUNCH(ctx, k_ccl_select, grid, 256, 0, st, HW, L, area, min_area, best, any_big);
  GG_LAUNCH(ctx, k_ccl_apply, grid, 256, 0, st, HW, L, area, min_area, keep_largest, best, any_big, out);
  return GG_OK;
}

// ----------------------------------------------------------------------------- host side
size_t trimap_workspace_bytes(int B, int H, int W, bool need_gray) {
  size_t s = Arena::padded((size_t)B * H * W * 4, 4);      // a/b planes
  if (need_gray) s += Arena::padded((size_t)B * H * W, 1);
  return s + 1024;
}

int refine_trimap(gg_context* ctx, Arena& ar, const uint8_t* bgr, const uint8_t* gray_in,
                  const int32_t* labels, const float* probs, const int64_t* node_off, int B, int H,
                  int W, int radius, float eps, float thr_fg, float thr_bg, uint8_t* trimap,
                  float* p_bg, float* p_fg, cudaStream_t st, int node_cap_hint) {
  GG_REQUIRE(radius >= 0 && radius <= GF_MAX_RADIUS, "refine_trimap: radius must be in [0,%d]",
             GF_MAX_RADIUS);
  GG_REQUIRE(B > 0 && H >= 2 && W >= 2, "refine_trimap: bad shape");
  const size_t npx = (size_t)B * H * W;
  const uint8_t* gray = gray_in;
  if (!gray) {
    uint8_t* g = ar.take<uint8_t>(npx);
    GG_LAUNCH(ctx, k_gray_only, ceil_div((long long)npx, 256), 256, 0, st, bgr, g, npx);
    gray = g;
  }
  float* ab = ar.take<float>(npx * 4);
  GuidedParams p{};
  p.gray = gray; p.labels = labels; p.probs = probs; p.node_off = node_off;
  p.ab = ab; p.plane_stride = npx; p.trimap = trimap; p.q0 = p_bg; p.q1 = p_fg;
  p.H = H; p.W = W; p.radius = radius; p.eps = eps; p.thr_fg = thr_fg; p.thr_bg = thr_bg;
  p.prob_cap = node_cap_hint > 0 ? std::min(node_cap_hint, GF_PROB_TABLE) : GF_PROB_DEFAULT;
  GG_TRY(launch_guided<true>(ctx, p, B, st));
  return GG_OK;
}

int guided_filter_plane(gg_context* ctx, Arena& ar, const float* guide, const float* src, int H,
                        int W, int radius, float eps, float* out, cudaStream_t st) {
  GG_REQUIRE(radius >= 0 && radius <= GF_MAX_RADIUS, "guided_filter: radius must be in [0,%d]",
             GF_MAX_RADIUS);
  GG_REQUIRE(H >= 2 && W >= 2, "guided_filter: bad shape");
  const size_t npx = (size_t)H * W;
  float* ab = ar.take<float>(npx * 2);
  GuidedParams p{};
  p.guide = guide; p.src = src; p.ab = ab; p.plane_stride = npx; p.q0 = out;
  p.H = H; p.W = W; p.radius = radius; p.eps = eps;
  GG_TRY(launch_guided<false>(ctx, p, 1, st));
  return GG_OK;
}

size_t seed_workspace_bytes(int B, long long node_cap_total) {
  return Arena::padded((size_t)B, 4) + Arena::padded((size_t)node_cap_total, 1) + 512;
}

int seed_from_prior(gg_context* ctx, Arena& ar, uint8_t* trimap, const int32_t* labels, const float* x,
                    const int64_t* node_off, int B, int H, int W, long long node_cap_total,
                    double seed_frac, cudaStream_t st) {
  GG_REQUIRE(seed_frac > 0.0 && seed_frac <= 1.0, "seed_from_prior: seed_frac must be in (0, 1]");
  const int HW = H * W;
  int* flags = ar.take<int>((size_t)B);
  uint8_t* sel = ar.take<uint8_t>((size_t)node_cap_total);
  GG_CUDA_OK(cudaMemsetAsync(flags, 0, (size_t)B * sizeof(int), st));
  {
    dim3 grid(std::min(ceil_div(HW, 256 * 16), 64), B);
    GG_LAUNCH(ctx, k_trimap_sides, grid, 256, 0, st, trimap, HW, flags);
  }
  GG_LAUNCH(ctx, k_seed_select, B, 256, 0, st, x, node_off, flags, seed_frac, sel);
  {
    dim3 grid(ceil_div(HW, 256), B);
    GG_LAUNCH(ctx, k_seed_apply, grid, 256, 0, st, labels, node_off, flags, sel, HW, trimap);
  }
  return GG_OK;
}

int project_trimap(gg_context* ctx, const int32_t* labels, const float* probs,
                   const int64_t* node_off, int B, int H, int W, float thr_fg, float thr_bg,
                   uint8_t* trimap, cudaStream_t st) {
  dim3 grid(ceil_div((long long)H * W, 256), B);
  GG_LAUNCH(ctx, k_project_trimap, grid, 256, 0, st, labels, probs, node_off, H * W, thr_fg, thr_bg,
            trimap);
  return GG_OK;
}

}  // namespace gg
