// Baseline and attention variants of the trimap network: GCNTrimapNet and GATTrimapNet
// (reference model.py:142-162 EdgeInjectionLayer, :216-233 ResGCNBlock, :239-316 GCNTrimapNet,
// :323-414 GATTrimapNet, PyG GCNConv / GATv2Conv), eval mode, fp32.
//
// ResGCNNet is the path's network (resgcn.cu, gcn_fused.cu, gemm_tc.cu); these two are the
// build_model("gcn" | "gat") alternatives of SURVEY 8(f)4.  They share one design:
//   * dense per-node transforms: k_lin (64x64x16 SIMT tiles; optional input / output affine =
//     eval-mode BatchNorm, bias, activation; strided rows so the jumping-knowledge concat of
//     GCNTrimapNet is written in place and read by the head without a copy);
//   * EdgeInjectionLayer: the per-edge MLP 5 -> D -> D (E x D x D multiply-adds per layer, >90 % of the
//     variants' FLOPs) runs on the tensor cores for D = 64 / 128: k_tc_gemm (gemm_tc.cu) with the first
//     layer as its A-operand producer (relu(W1 a_e + b1), 5 -> D per row, never stored), bf16x3 split
//     operands, fp32 accumulators in TMEM, sigmoid epilogue; the weight images are packed at load time.
//     Other widths (and gemm_impl = 0) use the SIMT tile kernel with the same on-the-fly A operand.
//     Rows are taken in CSR order (eid gather), so the scatter-mean over incoming edges is a mean over
//     CONTIGUOUS rows inside the node kernel;
//   * message passing: one warp per destination node over the dst-sorted CSR -- GCN aggregation +
//     BatchNorm + ReLU + residual + gate (k_gcn_block), or GATv2 attention with an online softmax
//     per head + LayerNorm + GELU + gate (k_gat_layer).  No atomics, fixed summation order.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "resgcn.cuh"
#include "variants.cuh"

namespace gg {

GG_D float gelu_erf_v(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
GG_D float sigmoid_v(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------ dense transform
struct LinArgs {
  const float* A; int lda;                         // rows [M, K]
  const float* in_scale; const float* in_shift;    // optional per-column affine of A
  const float* W; const float* bias;               // [O, K], [O] (optional)
  const float* out_scale; const float* out_shift;  // optional per-output affine after the bias
  float* C; int ldc;
  int M, O, K, act;                                // act: 0 none, 1 GELU(erf), 2 sigmoid, 3 ReLU
  // EDGE: A[j, k] = relu(w1[k, :] . attr[eid[j], :] + b1[k])  (row j = CSR position)
  const float* attr; const int32_t* eid; const float* w1; const float* b1;
};

template <bool EDGE>
__global__ void __launch_bounds__(256)
k_lin(const LinArgs a) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  __shared__ float sAt[EDGE ? 64 : 1][5];
  const int t = threadIdx.x, m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int ty = t >> 4, tx = t & 15;
  if (EDGE) {
    for (int i = t; i < 64 * 5; i += 256) {
      const int m = i / 5, c = i - 5 * m, row = m0 + m;
      sAt[m][c] = row < a.M ? a.attr[(size_t)a.eid[row] * 5 + c] : 0.0f;
    }
    __syncthreads();
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < a.K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (EDGE) {
        const int m = t & 63, k = (t >> 6) + 4 * i, kk = k0 + k;
        float v = 0.0f;
        if (kk < a.K) {
          const float* w = a.w1 + (size_t)kk * 5;
          v = a.b1[kk];
#pragma unroll
          for (int c = 0; c < 5; ++c) v = fmaf(w[c], sAt[m][c], v);
          v = fmaxf(v, 0.0f);
        }
        sA[k][m] = v;
      } else {
        const int idx = t + 256 * i, m = idx >> 4, k = idx & 15, row = m0 + m, kk = k0 + k;
        float v = 0.0f;
        if (row < a.M && kk < a.K) {
          v = a.A[(size_t)row * a.lda + kk];
          if (a.in_scale) v = fmaf(v, a.in_scale[kk], a.in_shift[kk]);
        }
        sA[k][m] = v;
      }
      const int idx = t + 256 * i, n = idx >> 4, k = idx & 15, col = n0 + n, kk = k0 + k;
      sB[k][n] = (col < a.O && kk < a.K) ? a.W[(size_t)col * a.K + kk] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= a.O) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[col];
      if (a.out_scale) v = fmaf(v, a.out_scale[col], a.out_shift[col]);
      if (a.act == 1) v = gelu_erf_v(v);
      else if (a.act == 2) v = sigmoid_v(v);
      else if (a.act == 3) v = fmaxf(v, 0.0f);
      a.C[(size_t)row * a.ldc + col] = v;
    }
  }
}

static int launch_lin(gg_context* ctx, cudaStream_t st, const LinArgs& a, bool edge) {
  if (a.M <= 0) return GG_OK;
  dim3 grid(ceil_div(a.M, 64), ceil_div(a.O, 64));
  if (edge) GG_LAUNCH(ctx, k_lin<true>, grid, 256, 0, st, a);
  else GG_LAUNCH(ctx, k_lin<false>, grid, 256, 0, st, a);
  return GG_OK;
}

// ------------------------------------------------------------------ row kernels
// h = GELU(LayerNorm(t)) in place, warp per row, channel lane + 32 j   (GATTrimapNet.input_proj)
template <int CPL>
__global__ void __launch_bounds__(256)
k_ln_gelu(float* __restrict__ t, const float* __restrict__ g, const float* __restrict__ b, int N) {
  constexpr int D = CPL * 32;
  const int lane = threadIdx.x & 31, v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= N) return;
  float x[CPL], s = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { x[j] = t[(size_t)v * D + lane + 32 * j]; s += x[j]; }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { const float d = x[j] - mean; q = fmaf(d, d, q); }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    t[(size_t)v * D + c] = gelu_erf_v((x[j] - mean) * rstd * g[c] + b[c]);
  }
}

// GCN normalisation: deg = 1 + #incoming edges that are not self loops (PyG gcn_norm removes the
// loops and adds one per node); dinv = deg^-1/2
__global__ void k_var_dinv(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, int N,
                           float* __restrict__ dinv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  int deg = 1;
  for (int e = rowptr[v]; e < rowptr[v + 1]; ++e) deg += src[e] != v;
  dinv[v] = 1.0f / sqrtf((float)deg);
}

// ResGCNBlock after the node transform (model.py:225-232): out = (relu(BN(A_hat xp + bias)) + h) * gate,
// gate = mean over the node's incoming edges of the EdgeInjectionLayer rows G (CSR order).
template <int CPL>
__global__ void __launch_bounds__(256)
k_gcn_block(const float* __restrict__ xp, const float* __restrict__ hprev, int ldh, const float* __restrict__ G,
            const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, const float* __restrict__ dinv,
            const float* __restrict__ bias, const float* __restrict__ bn_s, const float* __restrict__ bn_t,
            float* __restrict__ out, int ldo, int N) {
  constexpr int D = CPL * 32;
  const int lane = threadIdx.x & 31, v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= N) return;
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  const float dv = dinv[v];
  float acc[CPL], gs[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { acc[j] = 0.0f; gs[j] = 0.0f; }
  for (int e = e0; e < e1; ++e) {
    const int u = src[e];
    const float w = u != v ? dinv[u] * dv : 0.0f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
      gs[j] += G[(size_t)e * D + c];
      acc[j] = fmaf(w, xp[(size_t)u * D + c], acc[j]);
    }
  }
  const float inv = 1.0f / (float)max(e1 - e0, 1);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    float val = fmaf(dv * dv, xp[(size_t)v * D + c], acc[j]) + bias[c];
    val = fmaxf(fmaf(val, bn_s[c], bn_t[c]), 0.0f) + hprev[(size_t)v * ldh + c];
    out[(size_t)v * ldo + c] = val * (gs[j] * inv);
  }
}

// One GATTrimapNet layer after the two node transforms (model.py:394-400): GATv2 attention over the
// incoming edges + the mean-attribute self loop, bias, LayerNorm, GELU, edge gate.
// A lane owns the CPL consecutive channels lane*CPL..; a head spans lph = 32/H consecutive lanes.
template <int CPL>
__global__ void __launch_bounds__(256)
k_gat_layer(const float* __restrict__ xl, const float* __restrict__ xr, const float* __restrict__ attr,
            const float* __restrict__ G, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
            const int32_t* __restrict__ eid, const float* __restrict__ We, const float* __restrict__ att,
            const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
            float* __restrict__ out, int N, int lph) {
  constexpr int D = CPL * 32;
  const int lane = threadIdx.x & 31, v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= N) return;
  const int c0 = lane * CPL;
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  float we[CPL][5], at[CPL], xrv[CPL], gs[CPL], acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
#pragma unroll
    for (int k = 0; k < 5; ++k) we[j][k] = We[(size_t)(c0 + j) * 5 + k];
    at[j] = att[c0 + j];
    xrv[j] = xr[(size_t)v * D + c0 + j];
    gs[j] = 0.0f;
    acc[j] = 0.0f;
  }
  float la[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  int cnt = 0;
  for (int e = e0; e < e1; ++e) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) gs[j] += G[(size_t)e * D + c0 + j];
    if (src[e] != v) {
      const float* ap = attr + (size_t)eid[e] * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) la[k] += ap[k];
      ++cnt;
    }
  }
  const float cinv = 1.0f / (float)max(cnt, 1);
#pragma unroll
  for (int k = 0; k < 5; ++k) la[k] *= cinv;
  float mrun = -INFINITY, den = 0.0f;
  auto edge = [&](int u, const float (&a)[5]) {
    float xlu[CPL], part = 0.0f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      xlu[j] = xl[(size_t)u * D + c0 + j];
      float m = xlu[j] + xrv[j];
      float ea = 0.0f;
#pragma unroll
      for (int k = 0; k < 5; ++k) ea = fmaf(we[j][k], a[k], ea);
      m += ea;
      m = m > 0.0f ? m : 0.2f * m;
      part = fmaf(at[j], m, part);
    }
    for (int o = 1; o < lph; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float nm = fmaxf(mrun, part);
    const float sc = expf(mrun - nm), p = expf(part - nm);
    den = fmaf(den, sc, p);
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = fmaf(acc[j], sc, p * xlu[j]);
    mrun = nm;
  };
  for (int e = e0; e < e1; ++e) {
    const int u = src[e];
    if (u == v) continue;                       // warp-uniform: every lane sees the same edge
    const float* ap = attr + (size_t)eid[e] * 5;
    const float a[5] = {ap[0], ap[1], ap[2], ap[3], ap[4]};
    edge(u, a);
  }
  edge(v, la);
  float o[CPL], s = 0.0f;
  const float dinv = 1.0f / (den + 1e-16f);
#pragma unroll
  for (int j = 0; j < CPL; ++j) { o[j] = fmaf(acc[j], dinv, bias[c0 + j]); s += o[j]; }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { const float d = o[j] - mean; q = fmaf(d, d, q); }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
  const float ginv = 1.0f / (float)max(e1 - e0, 1);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = c0 + j;
    out[(size_t)v * D + c] = gelu_erf_v((o[j] - mean) * rstd * ln_g[c] + ln_b[c]) * (gs[j] * ginv);
  }
}

// h += skip, then GlobalContextModule (model.py:165-188) per graph, in place: block per graph.
__global__ void __launch_bounds__(256)
k_var_context(float* __restrict__ h, const float* __restrict__ skip, const int64_t* __restrict__ graph_off, int D,
              int n_graphs, const float* __restrict__ attn_w, const float* __restrict__ attn_b,
              const float* __restrict__ cmp_w, const float* __restrict__ cmp_b, const float* __restrict__ exp_w,
              const float* __restrict__ exp_b, float* __restrict__ score) {
  __shared__ float s_g[256], s_c[128], s_gate[256], sred[32];
  const int g = blockIdx.x, Dh = D / 2;
  const int v0 = (int)graph_off[g], v1 = (int)graph_off[g + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int v = v0 + wid; v < v1; v += nw) {
    float s = 0.0f;
    for (int c = lane; c < D; c += 32) {
      const float z = h[(size_t)v * D + c] + skip[(size_t)v * D + c];
      h[(size_t)v * D + c] = z;
      s = fmaf(attn_w[c], z, s);
    }
    s = warp_sum(s) + attn_b[0];
    if (lane == 0) score[v] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce<float>(mx, -INFINITY, OpMaxF(), sred);
  __syncthreads();
  float tot = 0.0f;
  for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    const float e = expf(score[v] - mx);
    score[v] = e;
    tot += e;
  }
  tot = block_reduce<float>(tot, 0.0f, OpAdd(), sred);
  const float inv = 1.0f / (n_graphs > 1 ? tot + 1e-12f : tot);     // _graph_softmax: +1e-12 in the batched branch
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {               // fixed order over the graph's nodes
    float a = 0.0f;
    for (int v = v0; v < v1; ++v) a = fmaf(score[v] * inv, h[(size_t)v * D + c], a);
    s_g[c] = a;
  }
  __syncthreads();
  for (int u = wid; u < Dh; u += nw) {
    float s = 0.0f;
    for (int c = lane; c < D; c += 32) s = fmaf(cmp_w[(size_t)u * D + c], s_g[c], s);
    s = warp_sum(s);
    if (lane == 0) s_c[u] = fmaxf(s + cmp_b[u], 0.0f);
  }
  __syncthreads();
  for (int c = wid; c < D; c += nw) {
    float s = 0.0f;
    for (int u = lane; u < Dh; u += 32) s = fmaf(exp_w[(size_t)c * Dh + u], s_c[u], s);
    s = warp_sum(s);
    if (lane == 0) s_gate[c] = sigmoid_v(s + exp_b[c]);
  }
  __syncthreads();
  for (size_t i = threadIdx.x; i < (size_t)(v1 - v0) * D; i += blockDim.x) {
    const int c = (int)(i % D);
    h[(size_t)v0 * D + i] *= s_gate[c];
  }
}

__global__ void k_var_softmax3(const float* __restrict__ logits, float* __restrict__ probs, int N) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const float s0 = logits[(size_t)v * 3], s1 = logits[(size_t)v * 3 + 1], s2 = logits[(size_t)v * 3 + 2];
  const float m = fmaxf(s0, fmaxf(s1, s2));
  const float e0 = expf(s0 - m), e1 = expf(s1 - m), e2 = expf(s2 - m);
  const float inv = 1.0f / (e0 + e1 + e2);
  probs[(size_t)v * 3] = e0 * inv;
  probs[(size_t)v * 3 + 1] = e1 * inv;
  probs[(size_t)v * 3 + 2] = e2 * inv;
}

__global__ void k_var_set_int(int* p, int v) { *p = v; }

// ------------------------------------------------------------------ weights
// Tensor order of gg_variant_weights (see the header).  BatchNorm groups (weight, bias, mean, var) are
// folded at load time into (scale, shift, -, -) in double precision.
namespace {
struct Layout {
  std::vector<long long> numel;
  std::vector<int> bn_groups;     // index of the first tensor of every BatchNorm group
};

Layout variant_layout(int kind, int D, int n, int H) {
  Layout L;
  auto add = [&](long long k) { L.numel.push_back(k); };
  auto bn = [&](int c) { L.bn_groups.push_back((int)L.numel.size()); for (int i = 0; i < 4; ++i) add(c); };
  bn(19);
  add(19LL * D); add(D);
  if (kind == GG_VARIANT_GCN) {
    bn(D);
    for (int i = 0; i < n; ++i) {
      add(D); add((long long)D * D); bn(D);
      add(5LL * D); add(D); add((long long)D * D); add(D);
    }
    add((long long)D * D * (n + 1)); add(D); bn(D);
    add((long long)(D / 2) * D); add(D / 2);
    add(3LL * (D / 2)); add(3);
  } else {
    add(D); add(D);
    for (int i = 0; i < n; ++i) {
      add(D); add(D);
      add((long long)D * D); add(D); add((long long)D * D); add(D);
      add(5LL * D);
      add(D); add(D);
      add(5LL * D); add(D); add((long long)D * D); add(D);
    }
    add((long long)D * D);
    add(D); add(1);
    add((long long)(D / 2) * D); add(D / 2);
    add((long long)D * (D / 2)); add(D);
    add((long long)D * D); add(D);
    add(3LL * D); add(3);
  }
  (void)H;
  return L;
}
}  // namespace

int variant_load_weights(gg_context* ctx, const gg_variant_weights* w) {
  GG_REQUIRE(w && w->tensors && w->numel, "gg_variant_load_weights: null argument");
  const int kind = w->variant, D = w->hidden, n = w->n_layers, H = w->n_heads;
  GG_REQUIRE(kind == GG_VARIANT_GCN || kind == GG_VARIANT_GAT, "gg_variant_load_weights: variant must be 1 (gcn) or 2 (gat)");
  GG_REQUIRE(D >= 32 && D <= 256 && D % 32 == 0, "hidden must be a multiple of 32 in [32, 256], got %d", D);
  GG_REQUIRE(n >= 1 && n <= 32, "n_layers must be in [1, 32], got %d", n);
  if (kind == GG_VARIANT_GAT)
    GG_REQUIRE(H >= 1 && H <= 32 && (32 % H) == 0, "n_heads must divide 32 (1, 2, 4, 8, 16, 32), got %d", H);
  const Layout L = variant_layout(kind, D, n, H);
  GG_REQUIRE(w->n_tensors == (int)L.numel.size(), "expected %d tensors for this variant, got %d", (int)L.numel.size(),
             w->n_tensors);
  std::vector<float> blob;
  std::vector<size_t> off(L.numel.size());
  for (size_t i = 0; i < L.numel.size(); ++i) {
    GG_REQUIRE(w->numel[i] == L.numel[i], "tensor %d: expected %lld elements, got %lld", (int)i, L.numel[i],
               (long long)w->numel[i]);
    GG_REQUIRE(w->tensors[i] != nullptr, "tensor %d is NULL", (int)i);
    off[i] = blob.size();
    blob.insert(blob.end(), w->tensors[i], w->tensors[i] + L.numel[i]);
    while (blob.size() % 4) blob.push_back(0.0f);        // 16-byte aligned tensors
  }
  for (int g0 : L.bn_groups) {                           // (weight, bias, mean, var) -> (scale, shift)
    const long long c = L.numel[g0];
    float *wt = &blob[off[g0]], *bs = &blob[off[g0 + 1]], *mu = &blob[off[g0 + 2]], *var = &blob[off[g0 + 3]];
    for (long long i = 0; i < c; ++i) {
      const double sc = (double)wt[i] / sqrt((double)var[i] + 1e-5);
      const double sh = (double)bs[i] - (double)mu[i] * sc;
      wt[i] = (float)sc;
      bs[i] = (float)sh;
    }
  }
  VariantWeights& vw = ctx->variant;
  if (vw.blob) { cudaFree(vw.blob); vw.blob = nullptr; }
  vw.loaded = false;
  GG_CUDA_OK(cudaMalloc(&vw.blob, blob.size() * sizeof(float)));
  GG_CUDA_OK(cudaMemcpy(vw.blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  vw.blob_floats = blob.size();
  vw.off = off;
  vw.kind = kind; vw.D = D; vw.n_layers = n; vw.heads = H;
  // tensor-core operand images of the edge-gate second layers (the bulk of the FLOPs: E x D x D per layer)
  if (vw.tc_blob) { cudaFree(vw.tc_blob); vw.tc_blob = nullptr; }
  if (D == 64 || D == 128) {
    vw.tc_stride = tc_image_bytes_padded(D, D);
    std::vector<unsigned char> img((size_t)n * vw.tc_stride, 0);
    for (int i = 0; i < n; ++i) {
      const int t2 = kind == GG_VARIANT_GCN ? 10 + 10 * i + 8 : 8 + 13 * i + 11;      // proj.2.weight
      tc_pack_weight(&blob[off[t2]], D, D, D, img.data() + (size_t)i * vw.tc_stride);
    }
    GG_CUDA_OK(cudaMalloc(&vw.tc_blob, img.size()));
    GG_CUDA_OK(cudaMemcpy(vw.tc_blob, img.data(), img.size(), cudaMemcpyHostToDevice));
  }
  if (!vw.d_rows) GG_CUDA_OK(cudaMalloc(&vw.d_rows, 2 * sizeof(int)));
  vw.loaded = true;
  return GG_OK;
}

size_t variant_workspace_bytes(const VariantWeights& vw, long long N, long long E) {
  const size_t D = vw.D;
  size_t b = 0;
  const size_t rows = (size_t)std::max<long long>(N, 1), erows = (size_t)std::max<long long>(E, 1);
  b += Arena::padded(erows * D, 4);                                         // G
  b += Arena::padded(rows * 3, 4) + Arena::padded(rows, 4) * 2;             // logits, dinv, score
  if (vw.kind == GG_VARIANT_GCN) b += Arena::padded(rows * D * (vw.n_layers + 1), 4) + 3 * Arena::padded(rows * D, 4);
  else b += 6 * Arena::padded(rows * D, 4);
  return b + 4096;
}

template <typename F>
static int dispatch_cpl(int D, F&& f) {
  switch (D / 32) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 5: return f(std::integral_constant<int, 5>());
    case 6: return f(std::integral_constant<int, 6>());
    case 7: return f(std::integral_constant<int, 7>());
    case 8: return f(std::integral_constant<int, 8>());
  }
  set_error("hidden width %d not supported", D);
  return GG_ERR_INVALID;
}

int variant_forward(gg_context* ctx, Arena& ar, int kind, const float* x, const int32_t* rowptr, const int32_t* src,
                    const int32_t* eid, const float* edge_attr, const int64_t* graph_off, int n_graphs, long long N,
                    long long E, float* logits, float* probs, cudaStream_t st) {
  const VariantWeights& vw = ctx->variant;
  const int D = vw.D, n = vw.n_layers;
  const float* wb = vw.blob;
  auto T = [&](int i) { return wb + vw.off[i]; };
  const int Ni = (int)N, Ei = (int)E;
  const int node_blocks = ceil_div((long long)Ni * 32, 256);
  float* G = ar.take<float>((size_t)std::max<long long>(E, 1) * D);
  float* lg = ar.take<float>((size_t)std::max<long long>(N, 1) * 3);
  float* dinv = ar.take<float>((size_t)std::max<long long>(N, 1));
  float* score = ar.take<float>((size_t)std::max<long long>(N, 1));
  const size_t nd = (size_t)std::max<long long>(N, 1) * D;
  if (N <= 0) return GG_OK;
  auto lin = [&](const float* A, int lda, int K, const float* W, const float* bias, float* C, int ldc, int O, int act,
                 const float* is = nullptr, const float* ish = nullptr, const float* os = nullptr,
                 const float* osh = nullptr) {
    LinArgs a{};
    a.A = A; a.lda = lda; a.in_scale = is; a.in_shift = ish; a.W = W; a.bias = bias; a.out_scale = os; a.out_shift = osh;
    a.C = C; a.ldc = ldc; a.M = Ni; a.O = O; a.K = K; a.act = act;
    return launch_lin(ctx, st, a, false);
  };
  const bool edge_tc = vw.tc_blob != nullptr && ctx->gemm_impl == 1 && Ei > 0;
  if (edge_tc) GG_LAUNCH(ctx, k_var_set_int, 1, 1, 0, st, vw.d_rows, Ei);
  auto edge_gate = [&](const float* w1, const float* b1, const float* w2, const float* b2, int layer) {
    if (edge_tc) {
      // G = sigmoid(relu(W1 a_e + b1) W2^T + b2) on the tensor cores: the first layer is the A-operand
      // producer (5 -> D per row, never stored), rows in CSR order through the eid gather
      TcPrologue pro{};
      pro.mode = 2; pro.w0 = w1; pro.b0 = b1; pro.row_index = eid; pro.relu = 1;
      return gemm_tc_image(ctx, st, vw.tc_blob + (size_t)layer * vw.tc_stride, edge_attr, b2, G, vw.d_rows, Ei, D, D, 5, D,
                           2, 0, pro);
    }
    LinArgs a{};
    a.W = w2; a.bias = b2; a.C = G; a.ldc = D; a.M = Ei; a.O = D; a.K = D; a.act = 2;
    a.attr = edge_attr; a.eid = eid; a.w1 = w1; a.b1 = b1;
    return launch_lin(ctx, st, a, true);
  };
  if (kind == GG_VARIANT_GCN) {
    const int ld = D * (n + 1);
    float* allh = ar.take<float>((size_t)Ni * ld);
    float* xp = ar.take<float>(nd);
    float* t1 = ar.take<float>(nd);
    float* t2 = ar.take<float>(nd);
    if (ar.overflowed) { set_error("variant_forward: workspace overflow"); return GG_ERR_INVALID; }
    GG_LAUNCH(ctx, k_var_dinv, ceil_div(Ni, 256), 256, 0, st, rowptr, src, Ni, dinv);
    // h0 = relu(BN(Linear(in_norm(x))))                                   model.py:266-270, 293
    GG_TRY(lin(x, GG_N_NODE_FEATS, GG_N_NODE_FEATS, T(4), T(5), allh, ld, D, 3, T(0), T(1), T(6), T(7)));
    for (int i = 0; i < n; ++i) {
      const int b = 10 + 10 * i;
      const float* hprev = allh + (size_t)i * D;
      GG_TRY(lin(hprev, ld, D, T(b + 1), nullptr, xp, D, D, 0));            // GCNConv.lin (no bias)
      GG_TRY(edge_gate(T(b + 6), T(b + 7), T(b + 8), T(b + 9), i));         // EdgeInjectionLayer.proj per edge
      GG_TRY(dispatch_cpl(D, [&](auto cpl) -> int {
        constexpr int CPL = decltype(cpl)::value;
        GG_LAUNCH(ctx, k_gcn_block<CPL>, node_blocks, 256, 0, st, xp, hprev, ld, G, rowptr, src, dinv, T(b), T(b + 2),
                  T(b + 3), allh + (size_t)(i + 1) * D, ld, Ni);
        return GG_OK;
      }));
    }
    const int hb = 10 + 10 * n;
    GG_TRY(lin(allh, ld, ld, T(hb), T(hb + 1), t1, D, D, 3, nullptr, nullptr, T(hb + 2), T(hb + 3)));   // head.0-2
    GG_TRY(lin(t1, D, D, T(hb + 6), T(hb + 7), t2, D / 2, D / 2, 3));                                    // head.4-5
    GG_TRY(lin(t2, D / 2, D / 2, T(hb + 8), T(hb + 9), logits ? logits : lg, 3, 3, 0));                  // head.6
  } else {
    float* h = ar.take<float>(nd);
    float* h2 = ar.take<float>(nd);
    float* skip = ar.take<float>(nd);
    float* xl = ar.take<float>(nd);
    float* xr = ar.take<float>(nd);
    float* t1 = ar.take<float>(nd);
    if (ar.overflowed) { set_error("variant_forward: workspace overflow"); return GG_ERR_INVALID; }
    const int lph = 32 / vw.heads;
    // h = gelu(LN(Linear(in_norm(x))))                                    model.py:347-351, 391
    GG_TRY(lin(x, GG_N_NODE_FEATS, GG_N_NODE_FEATS, T(4), T(5), h, D, D, 0, T(0), T(1)));
    GG_TRY(dispatch_cpl(D, [&](auto cpl) -> int {
      constexpr int CPL = decltype(cpl)::value;
      GG_LAUNCH(ctx, k_ln_gelu<CPL>, node_blocks, 256, 0, st, h, T(6), T(7), Ni);
      return GG_OK;
    }));
    const int tb = 8 + 13 * n;
    GG_TRY(lin(h, D, D, T(tb), nullptr, skip, D, D, 0));                    // skip_proj
    float* cur = h;
    float* nxt = h2;
    for (int i = 0; i < n; ++i) {
      const int b = 8 + 13 * i;
      GG_TRY(lin(cur, D, D, T(b + 2), T(b + 3), xl, D, D, 0));
      GG_TRY(lin(cur, D, D, T(b + 4), T(b + 5), xr, D, D, 0));
      GG_TRY(edge_gate(T(b + 9), T(b + 10), T(b + 11), T(b + 12), i));
      GG_TRY(dispatch_cpl(D, [&](auto cpl) -> int {
        constexpr int CPL = decltype(cpl)::value;
        GG_LAUNCH(ctx, k_gat_layer<CPL>, node_blocks, 256, 0, st, xl, xr, edge_attr, G, rowptr, src, eid, T(b + 6), T(b),
                  T(b + 1), T(b + 7), T(b + 8), nxt, Ni, lph);
        return GG_OK;
      }));
      std::swap(cur, nxt);
    }
    GG_LAUNCH(ctx, k_var_context, n_graphs, 256, 0, st, cur, skip, graph_off, D, n_graphs, T(tb + 1), T(tb + 2), T(tb + 3),
              T(tb + 4), T(tb + 5), T(tb + 6), score);
    GG_TRY(lin(cur, D, D, T(tb + 7), T(tb + 8), t1, D, D, 1));              // head.0 + GELU
    GG_TRY(lin(t1, D, D, T(tb + 9), T(tb + 10), logits ? logits : lg, 3, 3, 0));
  }
  if (probs) GG_LAUNCH(ctx, k_var_softmax3, ceil_div(Ni, 256), 256, 0, st, logits ? logits : lg, probs, Ni);
  return GG_OK;
}

}  // namespace gg
