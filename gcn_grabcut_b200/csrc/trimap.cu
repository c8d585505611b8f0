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
// Two tiled kernels: (1) means of {g, g^2, s_c, g s_c} -> a_c, b_c planes;
//                    (2) means of {a_c, b_c} -> q_c = mean(a_c) g + mean(b_c) -> clip -> labels.
#include "common.cuh"
#include "pixel_math.cuh"
#include "trimap.cuh"

namespace gg {

constexpr int GF_TY = 16, GF_TX = 64, GF_SEGX = 16, GF_THREADS = 256;
constexpr int GF_MAX_RADIUS = 24;

struct GuidedParams {
  // trimap mode
  const uint8_t* gray;      // [B,H,W]
  const int32_t* labels;    // [B,H,W]
  const float* probs;       // [SN,3]
  const int64_t* node_off;  // [B+1]
  // plane mode
  const float* guide;       // [B,H,W]
  const float* src;         // [B,H,W]
  // stage-1 outputs / stage-2 inputs: NSRC * 2 planes of [B,H,W] float32 (a_0, b_0, a_1, b_1)
  float* ab;
  size_t plane_stride;      // B*H*W
  // stage-2 outputs
  uint8_t* trimap;
  float* q0;                // optional filtered planes (p_bg / out)
  float* q1;                // optional (p_fg)
  int H, W, radius;
  float eps, thr_fg, thr_bg;
};

template <bool kTrimap>
struct GfTraits {
  static constexpr int NSRC = kTrimap ? 2 : 1;
  static constexpr int NPLANE1 = 2 + 2 * NSRC;   // g, g^2, s_c, g*s_c
  static constexpr int NPLANE2 = 2 * NSRC;       // a_c, b_c
};

size_t guided_smem_bytes(int radius, bool trimap) {
  const int tyh = GF_TY + 2 * radius, txh = GF_TX + 2 * radius;
  const int ld = txh | 1;
  const int nsrc = trimap ? 2 : 1;
  const int np1 = 2 + 2 * nsrc, np2 = 2 * nsrc;
  const size_t in1 = (size_t)(1 + nsrc) * tyh * txh * 4, in2 = (size_t)np2 * tyh * txh * 4;
  const size_t means1 = (size_t)np1 * GF_TY * GF_TX * 4, means2 = (size_t)np2 * GF_TY * GF_TX * 4;
  const size_t a1 = in1 > means1 ? in1 : means1, a2 = in2 > means2 ? in2 : means2;
  const size_t v1 = (size_t)np1 * GF_TY * ld * 8, v2 = (size_t)np2 * GF_TY * ld * 8;
  const size_t s1 = ((a1 + 15) & ~size_t(15)) + v1, s2 = ((a2 + 15) & ~size_t(15)) + v2;
  return (s1 > s2 ? s1 : s2) + 16;
}

// Vertical sliding window sums: for plane p and halo column c, V[p][y][c] =
// sum_{dy=0..2r} in_p[y+dy][c], y = 0..TY-1 (float64).  `value(p, row, col)` yields the
// float32 plane value.  Then horizontal sliding sums and scaling by 1/k^2.
template <int NP, typename ValueFn>
GG_D void box_means(ValueFn value, double* sV, float* sM, int radius, int txh, int ld) {
  const int k = 2 * radius + 1;
  const double scale = 1.0 / ((double)k * (double)k);
  // ---- vertical pass
  for (int t = threadIdx.x; t < NP * txh; t += blockDim.x) {
    const int p = t / txh, c = t - p * txh;
    double s = 0.0;
    for (int dy = 0; dy < k; ++dy) s += (double)value(p, dy, c);
    double* out = sV + (size_t)p * GF_TY * ld + c;
    out[0] = s;
    for (int y = 1; y < GF_TY; ++y) {
      s += (double)value(p, y + k - 1, c);
      s -= (double)value(p, y - 1, c);
      out[(size_t)y * ld] = s;
    }
  }
  __syncthreads();
  // ---- horizontal pass (lane -> row fastest: conflict-free 64-bit reads, ld odd)
  constexpr int NSEG = GF_TX / GF_SEGX;
  for (int t = threadIdx.x; t < NP * GF_TY * NSEG; t += blockDim.x) {
    const int y = t % GF_TY, seg = (t / GF_TY) % NSEG, p = t / (GF_TY * NSEG);
    const double* in = sV + (size_t)p * GF_TY * ld + (size_t)y * ld + seg * GF_SEGX;
    float* out = sM + ((size_t)p * GF_TY + y) * GF_TX + seg * GF_SEGX;
    double s = 0.0;
    for (int dx = 0; dx < k; ++dx) s += in[dx];
    out[0] = (float)(s * scale);
    for (int x = 1; x < GF_SEGX; ++x) {
      s += in[x + k - 1];
      s -= in[x - 1];
      out[x] = (float)(s * scale);
    }
  }
  __syncthreads();
}

template <bool kTrimap>
__global__ void __launch_bounds__(GF_THREADS)
k_guided_ab(const GuidedParams p) {
  using T = GfTraits<kTrimap>;
  extern __shared__ __align__(16) unsigned char gf_smem[];
  const int r = p.radius, tyh = GF_TY + 2 * r, txh = GF_TX + 2 * r, ld = txh | 1;
  const int H = p.H, W = p.W, b = blockIdx.z;
  const int y0 = blockIdx.y * GF_TY, x0 = blockIdx.x * GF_TX;
  float* sIn = reinterpret_cast<float*>(gf_smem);                     // [(1+NSRC)][tyh][txh]
  float* sM = sIn;                                                    // aliases (inputs dead by then)
  const size_t in_bytes = (size_t)(1 + T::NSRC) * tyh * txh * 4;
  const size_t m_bytes = (size_t)T::NPLANE1 * GF_TY * GF_TX * 4;
  const size_t a_bytes = ((in_bytes > m_bytes ? in_bytes : m_bytes) + 15) & ~size_t(15);
  double* sV = reinterpret_cast<double*>(gf_smem + a_bytes);
  const int plane = tyh * txh;

  // ---- load tile + halo (BORDER_REFLECT_101)
  const size_t img_off = (size_t)b * H * W;
  int64_t no = 0;
  int nn = 0;
  if (kTrimap) { no = p.node_off[b]; nn = (int)(p.node_off[b + 1] - no); }
  for (int i = threadIdx.x; i < plane; i += blockDim.x) {
    const int ty = i / txh, tx = i - ty * txh;
    const int y = reflect101(y0 + ty - r, H), x = reflect101(x0 + tx - r, W);
    const size_t o = img_off + (size_t)y * W + x;
    if (kTrimap) {
      sIn[i] = __fdiv_rn((float)p.gray[o], 255.0f);
      const int l = p.labels[o];
      float pb = 0.0f, pf = 0.0f;                      // project_to_pixels zero padding
      if (l >= 0 && l < nn) { const float* row = p.probs + (size_t)(no + l) * 3; pb = row[0]; pf = row[2]; }
      sIn[plane + i] = pb;
      sIn[2 * plane + i] = pf;
    } else {
      sIn[i] = p.guide[o];
      sIn[plane + i] = p.src[o];
    }
  }
  __syncthreads();
  auto value = [&](int pl, int row, int col) -> float {
    const int i = row * txh + col;
    const float g = sIn[i];
    if (pl == 0) return g;
    if (pl == 1) return __fmul_rn(g, g);
    const int c = (pl - 2) >> 1;
    const float s = sIn[(1 + c) * plane + i];
    return ((pl - 2) & 1) ? __fmul_rn(g, s) : s;
  };
  // means are written over the input area: every thread must be done reading inputs, which
  // box_means guarantees (its first barrier separates the vertical pass from the writes)
  box_means<T::NPLANE1>(value, sV, sM, r, txh, ld);

  // ---- a_c = cov/(var+eps), b_c = mean_s - a_c mean_g
  for (int i = threadIdx.x; i < GF_TY * GF_TX; i += blockDim.x) {
    const int ty = i / GF_TX, tx = i - ty * GF_TX;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const float mg = sM[i], mgg = sM[GF_TY * GF_TX + i];
    const float var = __fsub_rn(mgg, __fmul_rn(mg, mg));
    const float den = __fadd_rn(var, p.eps);
    const size_t o = img_off + (size_t)y * W + x;
#pragma unroll
    for (int c = 0; c < T::NSRC; ++c) {
      const float ms = sM[(2 + 2 * c) * GF_TY * GF_TX + i], mgs = sM[(3 + 2 * c) * GF_TY * GF_TX + i];
      const float cov = __fsub_rn(mgs, __fmul_rn(mg, ms));
      const float a = __fdiv_rn(cov, den);
      const float bb = __fsub_rn(ms, __fmul_rn(a, mg));
      p.ab[(size_t)(2 * c) * p.plane_stride + o] = a;
      p.ab[(size_t)(2 * c + 1) * p.plane_stride + o] = bb;
    }
  }
}

template <bool kTrimap>
__global__ void __launch_bounds__(GF_THREADS)
k_guided_out(const GuidedParams p) {
  using T = GfTraits<kTrimap>;
  extern __shared__ __align__(16) unsigned char gf_smem[];
  const int r = p.radius, tyh = GF_TY + 2 * r, txh = GF_TX + 2 * r, ld = txh | 1;
  const int H = p.H, W = p.W, b = blockIdx.z;
  const int y0 = blockIdx.y * GF_TY, x0 = blockIdx.x * GF_TX;
  float* sIn = reinterpret_cast<float*>(gf_smem);                     // [NPLANE2][tyh][txh]
  float* sM = sIn;
  const size_t in_bytes = (size_t)T::NPLANE2 * tyh * txh * 4;
  const size_t m_bytes = (size_t)T::NPLANE2 * GF_TY * GF_TX * 4;
  const size_t a_bytes = ((in_bytes > m_bytes ? in_bytes : m_bytes) + 15) & ~size_t(15);
  double* sV = reinterpret_cast<double*>(gf_smem + a_bytes);
  const int plane = tyh * txh;
  const size_t img_off = (size_t)b * H * W;

  for (int i = threadIdx.x; i < plane; i += blockDim.x) {
    const int ty = i / txh, tx = i - ty * txh;
    const int y = reflect101(y0 + ty - r, H), x = reflect101(x0 + tx - r, W);
    const size_t o = img_off + (size_t)y * W + x;
#pragma unroll
    for (int q = 0; q < T::NPLANE2; ++q) sIn[q * plane + i] = p.ab[(size_t)q * p.plane_stride + o];
  }
  __syncthreads();
  auto value = [&](int pl, int row, int col) -> float { return sIn[pl * plane + row * txh + col]; };
  box_means<T::NPLANE2>(value, sV, sM, r, txh, ld);

  for (int i = threadIdx.x; i < GF_TY * GF_TX; i += blockDim.x) {
    const int ty = i / GF_TX, tx = i - ty * GF_TX;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const size_t o = img_off + (size_t)y * W + x;
    const float g = kTrimap ? __fdiv_rn((float)p.gray[o], 255.0f) : p.guide[o];
    float q[T::NSRC];
#pragma unroll
    for (int c = 0; c < T::NSRC; ++c)
      q[c] = __fadd_rn(__fmul_rn(sM[(2 * c) * GF_TY * GF_TX + i], g), sM[(2 * c + 1) * GF_TY * GF_TX + i]);
    if (kTrimap) {
      const float pbg = fminf(fmaxf(q[0], 0.0f), 1.0f);          // np.clip(., 0, 1)
      const float pfg = fminf(fmaxf(q[T::NSRC - 1], 0.0f), 1.0f);
      uint8_t t = (pfg > pbg) ? 3 : 2;                           // pipeline.py:143-145
      if (pbg >= p.thr_bg) t = 0;
      if (pfg >= p.thr_fg) t = 1;
      p.trimap[o] = t;
      if (p.q0) p.q0[o] = pbg;
      if (p.q1) p.q1[o] = pfg;
    } else {
      p.q0[o] = q[0];
    }
  }
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

// ----------------------------------------------------------------------------- host side
size_t trimap_workspace_bytes(int B, int H, int W, bool need_gray) {
  size_t s = Arena::padded((size_t)B * H * W * 4, 4);      // a/b planes
  if (need_gray) s += Arena::padded((size_t)B * H * W, 1);
  return s + 1024;
}

int refine_trimap(gg_context* ctx, Arena& ar, const uint8_t* bgr, const uint8_t* gray_in,
                  const int32_t* labels, const float* probs, const int64_t* node_off, int B, int H,
                  int W, int radius, float eps, float thr_fg, float thr_bg, uint8_t* trimap,
                  float* p_bg, float* p_fg, cudaStream_t st) {
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
  const size_t smem = guided_smem_bytes(radius, true);
  GG_CUDA_OK(cudaFuncSetAttribute(k_guided_ab<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GG_CUDA_OK(cudaFuncSetAttribute(k_guided_out<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(W, GF_TX), ceil_div(H, GF_TY), B);
  GG_LAUNCH(ctx, k_guided_ab<true>, grid, GF_THREADS, smem, st, p);
  GG_LAUNCH(ctx, k_guided_out<true>, grid, GF_THREADS, smem, st, p);
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
  const size_t smem = guided_smem_bytes(radius, false);
  GG_CUDA_OK(cudaFuncSetAttribute(k_guided_ab<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GG_CUDA_OK(cudaFuncSetAttribute(k_guided_out<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(W, GF_TX), ceil_div(H, GF_TY), 1);
  GG_LAUNCH(ctx, k_guided_ab<false>, grid, GF_THREADS, smem, st, p);
  GG_LAUNCH(ctx, k_guided_out<false>, grid, GF_THREADS, smem, st, p);
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
