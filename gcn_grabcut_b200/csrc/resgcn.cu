// ResGCNNet forward (eval mode, fp32 semantics) over a batch of region graphs.
//
// Replaces model.py:449-546 of the reference (ResGCNNet.forward / predict_probs) together
// with the PyG GCNConv / SAGEConv layers it calls and the helpers _scatter_mean,
// _graph_softmax, EdgeContext, GlobalContextModule, InputNorm (model.py:69-213).
//
// Layout: all node tensors are row-major [nodes, D] fp32 over the whole batch (ragged,
// graph g owns rows [graph_off[g], graph_off[g+1])); message passing walks the
// destination-sorted CSR (no atomics, deterministic summation order); the dense per-node
// transforms go through gemm() -- tcgen05 tensor cores (bf16x3 split, fp32 accumulate in
// TMEM; gemm_tc.cu) or the SIMT fp32 kernel below (validation path).
//
// Row counts are only known on the device: every kernel reads sizes[0] = total nodes,
// sizes[1] = total directed edges and is launched for the capacity.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "resgcn.cuh"

namespace gg {

GG_D float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// The same exact-erf GELU x Phi(x) with Phi from the complementary error function in Abramowitz-Stegun
// form 7.1.26 (|error of erf| <= 1.5e-7, no cancellation on the negative side): 16 instructions instead
// of erff's 30.  Used by the per-node / per-edge kernels (as by the fused GCN blocks and the tensor-core
// epilogues); the SIMT validation GEMM keeps erff.
GG_D float gelu_fast(float x) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float pl = fmaf(t, 1.061405429f, -1.453152027f);
  pl = fmaf(t, pl, 1.421413741f);
  pl = fmaf(t, pl, -0.284496736f);
  pl = fmaf(t, pl, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  const float half_erfc = 0.5f * t * pl * e;   // Phi(-|x|)
  return x * (x >= 0.0f ? 1.0f - half_erfc : half_erfc);
}
// 1 / (1 + 2^(-x log2 e)) with the approximate ex2 / rcp units (2 + 1 ulp): 4 instructions instead of ~20
GG_D float sigmoidf(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

__global__ void k_sizes(const int64_t* __restrict__ graph_off, int n_graphs,
                        const int32_t* __restrict__ rowptr, int* __restrict__ sizes,
                        long long node_cap, long long edge_cap, int* status) {
  long long nt = graph_off[n_graphs];
  if (nt > node_cap) { nt = node_cap; atomicOr(status, ST_EDGE_CAP); }
  long long et = nt > 0 ? rowptr[nt] : 0;
  if (et > edge_cap) { et = edge_cap; atomicOr(status, ST_EDGE_CAP); }
  sizes[0] = (int)nt;
  sizes[1] = (int)et;
}

// node -> graph id, and GCN normalisation dinv[v] = (1 + #in-edges with src != v)^-1/2
__global__ void k_node_meta(const int64_t* __restrict__ graph_off, int n_graphs,
                            const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                            const int* __restrict__ sizes, int* __restrict__ node_graph,
                            float* __restrict__ dinv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= sizes[0]) return;
  int lo = 0, hi = n_graphs;            // last g with graph_off[g] <= v
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (graph_off[mid] <= v) lo = mid; else hi = mid;
  }
  node_graph[v] = lo;
  int deg = 1;
  for (int e = rowptr[v]; e < rowptr[v + 1]; ++e) deg += (src[e] != v);
  dinv[v] = 1.0f / sqrtf((float)deg);
}

// ------------------------------------------------------------------ input stage
// h = GELU(LN(W_in BN(x) + b)) * (1 + sigmoid(W2 GELU(W1 prior + b1) + b2)); z = w0 h
// (model.py:516-518, 191-213, 465-476).  One warp per group of R nodes, lane owns channels lane+32j:
// every weight read from shared memory feeds R nodes (the kernel is bound by the shared-memory /
// shuffle pipe: R = 1 needs 5 such instructions per 4 FMAs, R = 4 needs 8 per 16); the arithmetic of
// a node does not depend on R.
template <int CPL, int R>   // channels per lane = D/32, nodes per warp iteration
__global__ void __launch_bounds__(256)
k_input_stage(const float* __restrict__ x, const float* __restrict__ wb, NetOffsets o,
              const int* __restrict__ sizes, float* __restrict__ h, float* __restrict__ z,
              float2* __restrict__ row_stats) {
  extern __shared__ float sw[];
  const int D = CPL * 32, q = o.q;
  float* s_win = sw;                    // [D][19]
  float* s_pb0 = s_win + D * 19;        // [q][3]
  float* s_pb2 = s_pb0 + q * 3;         // [q][D] (transposed: conflict-free across lanes)
  for (int i = threadIdx.x; i < D * 19; i += blockDim.x) s_win[i] = wb[o.w_in + i];
  for (int i = threadIdx.x; i < q * 3; i += blockDim.x) s_pb0[i] = wb[o.pb0_w + i];
  // transposed copy with LINEAR shared-memory stores (the strided side is the L2-cached global read)
  for (int i = threadIdx.x; i < D * q; i += blockDim.x) s_pb2[i] = wb[o.pb2_w + (size_t)(i % D) * q + i / D];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nt = sizes[0];
  const float jk0 = wb[o.jk_w];
  const float bn_s = lane < 19 ? wb[o.bn_scale + lane] : 0.0f, bn_b = lane < 19 ? wb[o.bn_shift + lane] : 0.0f;
  for (int v0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R; v0 < nt; v0 += warps * R) {
    float xv[R], xraw[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      xraw[r] = 0.0f;
      if (lane < 19) xraw[r] = x[(size_t)min(v0 + r, nt - 1) * 19 + lane];
      xv[r] = xraw[r] * bn_s + bn_b;
    }
    float t[R][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const float bj = wb[o.b_in + lane + 32 * j];
#pragma unroll
      for (int r = 0; r < R; ++r) t[r][j] = bj;
    }
    for (int kk = 0; kk < 19; ++kk) {
      float xk[R];
#pragma unroll
      for (int r = 0; r < R; ++r) xk[r] = __shfl_sync(0xffffffffu, xv[r], kk);
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const float w = s_win[(lane + 32 * j) * 19 + kk];
#pragma unroll
        for (int r = 0; r < R; ++r) t[r][j] = fmaf(w, xk[r], t[r][j]);
      }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) sum += t[r][j];
      mean[r] = warp_sum(sum) / (float)D;
      float sq = 0.0f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) { const float d = t[r][j] - mean[r]; sq += d * d; }
      rstd[r] = 1.0f / sqrtf(warp_sum(sq) / (float)D + 1e-5f);
    }
    // prior booster: q hidden units, lane owns units lane, lane+32, ...
    float p0[R], p1[R], p2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      p0[r] = __shfl_sync(0xffffffffu, xraw[r], 16);
      p1[r] = __shfl_sync(0xffffffffu, xraw[r], 17);
      p2[r] = __shfl_sync(0xffffffffu, xraw[r], 18);
    }
    float boost[R][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const float bj = wb[o.pb2_b + lane + 32 * j];
#pragma unroll
      for (int r = 0; r < R; ++r) boost[r][j] = bj;
    }
    for (int u0 = 0; u0 < q; u0 += 32) {
      const int u = u0 + lane;
      float hu[R];
#pragma unroll
      for (int r = 0; r < R; ++r) hu[r] = 0.0f;
      if (u < q) {
        const float w0 = s_pb0[u * 3], w1 = s_pb0[u * 3 + 1], w2 = s_pb0[u * 3 + 2], bu = wb[o.pb0_b + u];
#pragma unroll
        for (int r = 0; r < R; ++r) hu[r] = gelu_fast(fmaf(w2, p2[r], fmaf(w1, p1[r], fmaf(w0, p0[r], bu))));
      }
      const int lim = min(32, q - u0);
      for (int s = 0; s < lim; ++s) {
        float hs[R];
#pragma unroll
        for (int r = 0; r < R; ++r) hs[r] = __shfl_sync(0xffffffffu, hu[r], s);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const float w = s_pb2[(u0 + s) * D + lane + 32 * j];
#pragma unroll
          for (int r = 0; r < R; ++r) boost[r][j] = fmaf(w, hs[r], boost[r][j]);
        }
      }
    }
    float lg[CPL], lb[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) { lg[j] = wb[o.ln_in_g + lane + 32 * j]; lb[j] = wb[o.ln_in_b + lane + 32 * j]; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int v = v0 + r;
      if (v >= nt) break;                                    // warp-uniform
      float hsum = 0.0f, hvv[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        const float ln = (t[r][j] - mean[r]) * rstd[r] * lg[j] + lb[j];
        const float hv = gelu_fast(ln) * (1.0f + sigmoidf(boost[r][j]));
        h[(size_t)v * D + c] = hv;
        z[(size_t)v * D + c] = jk0 * hv;
        hvv[j] = hv;
        hsum += hv;
      }
      // mean / rstd of the new row for the next layer's LayerNorm (consumed by the GEMM producers)
      const float hmean = warp_sum(hsum) / (float)D;
      float hsq = 0.0f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) { const float d = hvv[j] - hmean; hsq += d * d; }
      const float hrstd = 1.0f / sqrtf(warp_sum(hsq) / (float)D + 1e-5f);
      if (lane == 0) row_stats[v] = make_float2(hmean, hrstd);
    }
  }
}

// ------------------------------------------------------------------ edge context
// e1 = GELU(W0 attr + b0)  [E, c]   (model.py:124-128, first layer; K = 5)
__global__ void k_edge_enc1(const float* __restrict__ attr, const float* __restrict__ wb,
                            NetOffsets o, const int* __restrict__ sizes, float* __restrict__ e1) {
  const int c = o.c;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)sizes[1] * c;
  if (i >= total) return;
  const int e = (int)(i / c), u = (int)(i - (long long)e * c);
  const float* a = attr + (size_t)e * 5;
  const float* w = wb + o.ee0_w + (size_t)u * 5;
  float s = wb[o.ee0_b + u];
#pragma unroll
  for (int kk = 0; kk < 5; ++kk) s = fmaf(w[kk], a[kk], s);
  e1[i] = gelu_fast(s);
}

// The second encoder layer is linear, so the scatter-mean commutes with it:
//     mean_e (W2 g_e + b2) = W2 (mean_e g_e) + b2,     g_e = GELU(W0 a_e + b0)
// (a node without incoming edges keeps ctx = 0, _scatter_mean's empty sum).  k_edge_gelu_mean
// produces m_v = mean_{e: dst = v} g_e straight from the 5-d edge attributes -- the [E, c]
// encodings are never materialised and W2 is applied to N node rows instead of E edge rows.
// Warp per node; lane owns hidden units lane, lane+32, ... (c <= 256).
template <int UPL>   // units per lane = ceil(c / 32)
__global__ void __launch_bounds__(256)
k_edge_gelu_mean(const float* __restrict__ attr, const int32_t* __restrict__ rowptr,
                 const int32_t* __restrict__ eid, const float* __restrict__ wb, NetOffsets o,
                 const int* __restrict__ sizes, float* __restrict__ m) {
  const int c = o.c, lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  float w[UPL][5], b[UPL], acc[UPL];
#pragma unroll
  for (int j = 0; j < UPL; ++j) {
    const int u = min(lane + 32 * j, c - 1);
    b[j] = wb[o.ee0_b + u];
    acc[j] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 5; ++kk) w[j][kk] = wb[o.ee0_w + (size_t)u * 5 + kk];
  }
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  for (int e = e0; e < e1; e += 2) {
    const bool two = e + 1 < e1;
    const float* a0 = attr + (size_t)eid[e] * 5;
    const float* a1 = attr + (size_t)eid[two ? e + 1 : e] * 5;
    float x0[5], x1[5];
#pragma unroll
    for (int kk = 0; kk < 5; ++kk) { x0[kk] = __ldg(a0 + kk); x1[kk] = __ldg(a1 + kk); }
#pragma unroll
    for (int j = 0; j < UPL; ++j) {
      float s0 = b[j], s1 = b[j];
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) { s0 = fmaf(w[j][kk], x0[kk], s0); s1 = fmaf(w[j][kk], x1[kk], s1); }
      acc[j] += gelu_fast(s0);
      if (two) acc[j] += gelu_fast(s1);
    }
  }
  const float inv = 1.0f / (float)max(e1 - e0, 1);
#pragma unroll
  for (int j = 0; j < UPL; ++j) {
    const int u = lane + 32 * j;
    if (u < c) m[(size_t)v * c + u] = acc[j] * inv;
  }
}

// ctx_v = LN_c(W2 m_v + b2) for nodes with incoming edges, LN_c(0) otherwise; t = W2 m (no bias)
__global__ void __launch_bounds__(256)
k_edge_ctx_ln(const float* __restrict__ t, const int32_t* __restrict__ rowptr,
              const float* __restrict__ wb, NetOffsets o, const int* __restrict__ sizes,
              float* __restrict__ ctx) {
  const int c = o.c, lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  const bool has = rowptr[v + 1] > rowptr[v];
  float acc[8];
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = lane + 32 * j;
    acc[j] = (has && ch < c) ? t[(size_t)v * c + ch] + wb[o.ee2_b + ch] : 0.0f;
    sum += acc[j];
  }
  const float mean = warp_sum(sum) / (float)c;
  float sq = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) if (lane + 32 * j < c) { const float d = acc[j] - mean; sq += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)c + 1e-5f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = lane + 32 * j;
    if (ch < c) ctx[(size_t)v * c + ch] = (acc[j] - mean) * rstd * wb[o.eg_ln_g + ch] + wb[o.eg_ln_b + ch];
  }
}

// ctx_v = LN_c(mean_{e: dst = v} enc_e)   (model.py:69-74, 129-139); warp per node
__global__ void __launch_bounds__(256)
k_edge_ctx(const float* __restrict__ enc, const int32_t* __restrict__ rowptr,
           const int32_t* __restrict__ eid, const float* __restrict__ wb, NetOffsets o,
           const int* __restrict__ sizes, float* __restrict__ ctx) {
  const int c = o.c, lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  const float inv = 1.0f / (float)max(e1 - e0, 1);
  // c <= 256: up to 8 channels per lane
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  for (int e = e0; e < e1; ++e) {
    const float* row = enc + (size_t)eid[e] * c;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int ch = lane + 32 * j; if (ch < c) acc[j] += row[ch]; }
  }
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] *= inv; if (lane + 32 * j < c) sum += acc[j]; }
  const float mean = warp_sum(sum) / (float)c;
  float sq = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) if (lane + 32 * j < c) { const float d = acc[j] - mean; sq += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)c + 1e-5f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = lane + 32 * j;
    if (ch < c) ctx[(size_t)v * c + ch] = (acc[j] - mean) * rstd * wb[o.eg_ln_g + ch] + wb[o.eg_ln_b + ch];
  }
}

// ------------------------------------------------------------------ row LayerNorm
// out = LN(in [* scale_rows[node_graph]]) with affine (g, b); warp per row.
template <int CPL>
__global__ void __launch_bounds__(256)
k_layernorm(const float* __restrict__ in, const float* __restrict__ g, const float* __restrict__ bta,
            const float* __restrict__ gvec, const int* __restrict__ node_graph,
            const int* __restrict__ sizes, float* __restrict__ out) {
  constexpr int D = CPL * 32;
  const int lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  float t[CPL];
  const float* gv = gvec ? gvec + (size_t)node_graph[v] * D : nullptr;
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    t[j] = in[(size_t)v * D + lane + 32 * j];
    if (gv) t[j] *= gv[lane + 32 * j];
    sum += t[j];
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { const float d = t[j] - mean; sq += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + 1e-5f);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    out[(size_t)v * D + c] = (t[j] - mean) * rstd * g[c] + bta[c];
  }
}

// ------------------------------------------------------------------ GCN aggregation
// u_v = sum_{j in N(v), j != v} dinv_v dinv_j xw_j + dinv_v^2 xw_v + bias
// h_v += GELU(u_v * gate_v);  z_v += w_l h_v          (model.py:523-528; GCNConv defaults)
// Warp per node; a lane owns CPL consecutive channels (one 128-bit load per neighbour row for
// D = 128); neighbour rows are fetched four at a time so that four gathers are in flight.
template <int CPL>
struct ChanVec {
  float v[CPL];
  GG_D void load(const float* row, int lane) {
    if (CPL == 4) {
      const float4 t = *reinterpret_cast<const float4*>(row + 4 * lane);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] = row[CPL * lane + j];
    }
  }
  GG_D void store(float* row, int lane) const {
    if (CPL == 4) {
      *reinterpret_cast<float4*>(row + 4 * lane) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < CPL; ++j) row[CPL * lane + j] = v[j];
    }
  }
};

template <int CPL>
__global__ void __launch_bounds__(1024)
k_gcn_aggregate(const float* __restrict__ xw, const int32_t* __restrict__ rowptr,
                const int32_t* __restrict__ src, const float* __restrict__ dinv,
                const float* __restrict__ bias, const float* __restrict__ gate, float jkw,
                const int* __restrict__ sizes, float* __restrict__ h, float* __restrict__ z,
                float2* __restrict__ row_stats) {
  constexpr int D = CPL * 32;
  const int lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  const float dv = dinv[v];
  ChanVec<CPL> acc, self;
  self.load(xw + (size_t)v * D, lane);
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc.v[j] = dv * dv * self.v[j];
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  for (int e = e0; e < e1; e += 4) {
    int s4[4];
    float w4[4];
    ChanVec<CPL> r4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = e + u < e1;
      s4[u] = ok ? src[e + u] : v;
      w4[u] = (ok && s4[u] != v) ? dv * dinv[s4[u]] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) r4[u].load(xw + (size_t)s4[u] * D, lane);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc.v[j] = fmaf(w4[u], r4[u].v[j], acc.v[j]);
  }
  ChanVec<CPL> hv, zv, gv, bv;
  hv.load(h + (size_t)v * D, lane);
  zv.load(z + (size_t)v * D, lane);
  gv.load(gate + (size_t)v * D, lane);
  bv.load(bias, lane);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const float u = acc.v[j] + bv.v[j];
    hv.v[j] += gelu_fast(u * gv.v[j]);
    zv.v[j] += jkw * hv.v[j];
  }
  hv.store(h + (size_t)v * D, lane);
  zv.store(z + (size_t)v * D, lane);
  // mean / rstd of the new row for the next layer's LayerNorm (the GEMM producers then skip their
  // own reductions); the same two-pass formula the producers use
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) sum += hv.v[j];
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { const float d = hv.v[j] - mean; sq += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + 1e-5f);
  if (lane == 0) row_stats[v] = make_float2(mean, rstd);
}

// SAGE mean aggregation: m_v = mean_{j -> v} h_j (self loops kept, empty -> 0)
template <int CPL>
__global__ void __launch_bounds__(256)
k_sage_mean(const float* __restrict__ h, const int32_t* __restrict__ rowptr,
            const int32_t* __restrict__ src, const int* __restrict__ sizes, float* __restrict__ m) {
  constexpr int D = CPL * 32;
  const int lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  ChanVec<CPL> acc;
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc.v[j] = 0.0f;
  for (int e = e0; e < e1; e += 4) {
    ChanVec<CPL> r4[4];
    float w4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = e + u < e1;
      w4[u] = ok ? 1.0f : 0.0f;
      r4[u].load(h + (size_t)(ok ? src[e + u] : v) * D, lane);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc.v[j] = fmaf(w4[u], r4[u].v[j], acc.v[j]);
  }
  const float inv = 1.0f / (float)max(e1 - e0, 1);
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc.v[j] *= inv;
  acc.store(m + (size_t)v * D, lane);
}

// s = GELU(LN(t)); z += w s      (model.py:530-533)
template <int CPL>
__global__ void __launch_bounds__(256)
k_sage_finish(const float* __restrict__ t_in, const float* __restrict__ g, const float* __restrict__ bta,
              float jkw, const int* __restrict__ sizes, float* __restrict__ z) {
  constexpr int D = CPL * 32;
  const int lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  float t[CPL];
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { t[j] = t_in[(size_t)v * D + lane + 32 * j]; sum += t[j]; }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.0f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { const float d = t[j] - mean; sq += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + 1e-5f);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    const float s = gelu_fast((t[j] - mean) * rstd * g[c] + bta[c]);
    z[(size_t)v * D + c] += jkw * s;
  }
}

// ------------------------------------------------------------------ global context
// One block per graph: a = softmax_graph(u.z + c); g = sum a_i z_i;
// gvec = sigmoid(W_e relu(W_c g + b_c) + b_e)          (model.py:165-188, 90-108)
constexpr int GC_WARPS = 16;      // warps of k_graph_context (one block per graph; the node loops are latency-bound)
__global__ void __launch_bounds__(GC_WARPS * 32)
k_graph_context(const float* __restrict__ z, const int64_t* __restrict__ graph_off,
                const float* __restrict__ wb, NetOffsets o, int n_graphs,
                float* __restrict__ score, float* __restrict__ gvec) {
  extern __shared__ float sm[];
  const int D = o.D, Dh = D / 2;
  float* s_g = sm;            // [D]
  float* s_c = sm + D;        // [Dh]
  __shared__ float sred[32];
  const int g = blockIdx.x;
  const int v0 = (int)graph_off[g], v1 = (int)graph_off[g + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float INF = __int_as_float(0x7f800000);
  // scores
  float mx = -INF;
  for (int v = v0 + wid; v < v1; v += nw) {
    float s = 0.0f;
    for (int c = lane; c < D; c += 32) s = fmaf(wb[o.attn_w + c], z[(size_t)v * D + c], s);
    s = warp_sum(s) + wb[o.attn_b];
    if (lane == 0) score[v] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce<float>(mx, -INF, OpMaxF(), sred);
  __syncthreads();
  float tot = 0.0f;
  for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    const float e = expf(score[v] - mx);
    score[v] = e;
    tot += e;
  }
  tot = block_reduce<float>(tot, 0.0f, OpAdd(), sred);
  const float inv = 1.0f / (n_graphs > 1 ? tot + 1e-12f : tot);
  __syncthreads();
  // g = sum_v a_v z_v: the nodes are split over the warps (a lane owns channels lane, lane+32,
  // ...); the per-warp partial sums are combined in a fixed order (deterministic)
  {
    __shared__ float s_part[GC_WARPS][256];      // blockDim.x == GC_WARPS * 32, D <= 256
    float part[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) part[j] = 0.0f;
    for (int v = v0 + wid; v < v1; v += nw) {
      const float a = score[v] * inv;
      const float* zr = z + (size_t)v * D;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        if (c < D) part[j] = fmaf(a, zr[c], part[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      if (c < D) s_part[wid][c] = part[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float a = 0.0f;
      for (int w = 0; w < nw; ++w) a += s_part[w][c];
      s_g[c] = a;
    }
  }
  __syncthreads();
  for (int u = wid; u < Dh; u += nw) {
    float s = 0.0f;
    for (int c = lane; c < D; c += 32) s = fmaf(wb[o.cmp_w + (size_t)u * D + c], s_g[c], s);
    s = warp_sum(s);
    if (lane == 0) s_c[u] = fmaxf(s + wb[o.cmp_b + u], 0.0f);
  }
  __syncthreads();
  for (int c = wid; c < D; c += nw) {
    float s = 0.0f;
    for (int u = lane; u < Dh; u += 32) s = fmaf(wb[o.exp_w + (size_t)c * Dh + u], s_c[u], s);
    s = warp_sum(s);
    if (lane == 0) gvec[(size_t)g * D + c] = sigmoidf(s + wb[o.exp_b + c]);
  }
}

// The same readout for large graphs, split over many blocks (a graph of 10^4 regions on one
// block leaves the GPU idle): scores + per-graph maximum (ordered-int atomicMax), then partial
// softmax sums per (graph, block) written to scratch, then one block per graph adds the partials
// in a fixed order (deterministic) and applies the two small linear layers.
GG_D int float_to_ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
GG_D float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
k_ctx_scores(const float* __restrict__ z, const int* __restrict__ node_graph, const float* __restrict__ wb,
             NetOffsets o, const int* __restrict__ sizes, float* __restrict__ score, int* __restrict__ gmax) {
  const int D = o.D, lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  float s = 0.0f;
  for (int c = lane; c < D; c += 32) s = fmaf(wb[o.attn_w + c], z[(size_t)v * D + c], s);
  s = warp_sum(s) + wb[o.attn_b];
  if (lane == 0) {
    score[v] = s;
    atomicMax(&gmax[node_graph[v]], float_to_ordered(s));
  }
}

constexpr int CTX_PARTS = 32;     // blocks per graph in the partial-sum pass

__global__ void __launch_bounds__(256)
k_ctx_partial(const float* __restrict__ z, const int64_t* __restrict__ graph_off, const float* __restrict__ score,
              const int* __restrict__ gmax, int D, float* __restrict__ part /*[G][PARTS][D+1]*/) {
  __shared__ float s_part[8][257];
  const int g = blockIdx.y, pb = blockIdx.x;
  const int v0 = (int)graph_off[g], v1 = (int)graph_off[g + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float mx = ordered_to_float(gmax[g]);
  float acc[8], tot = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  // node v of the graph belongs to part (v - v0) / span, warp (v - v0) % 8 within the part
  const int span = (v1 - v0 + CTX_PARTS - 1) / CTX_PARTS;
  const int lo = v0 + pb * span, hi = min(v1, lo + span);
  for (int v = lo + wid; v < hi; v += 8) {
    const float e = expf(score[v] - mx);
    tot += e;
    const float* zr = z + (size_t)v * D;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      if (c < D) acc[j] = fmaf(e, zr[c], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    if (c < D) s_part[wid][c] = acc[j];
  }
  if (lane == 0) s_part[wid][256] = tot;
  __syncthreads();
  float* out = part + ((size_t)g * CTX_PARTS + pb) * (D + 1);
  for (int c = threadIdx.x; c <= D; c += blockDim.x) {
    const int col = c < D ? c : 256;
    float a = 0.0f;
    for (int w = 0; w < 8; ++w) a += s_part[w][col];
    out[c] = a;
  }
}

__global__ void __launch_bounds__(256)
k_ctx_finish(const float* __restrict__ part, const float* __restrict__ wb, NetOffsets o, int n_graphs,
             float* __restrict__ gvec) {
  extern __shared__ float sm[];
  const int D = o.D, Dh = D / 2;
  float* s_g = sm;            // [D]
  float* s_c = sm + D;        // [Dh]
  __shared__ float s_tot;
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* pg = part + (size_t)g * CTX_PARTS * (D + 1);
  for (int c = threadIdx.x; c <= D; c += blockDim.x) {
    float a = 0.0f;
    for (int pb = 0; pb < CTX_PARTS; ++pb) a += pg[(size_t)pb * (D + 1) + c];
    if (c < D) s_g[c] = a; else s_tot = a;
  }
  __syncthreads();
  const float inv = 1.0f / (n_graphs > 1 ? s_tot + 1e-12f : s_tot);
  for (int u = wid; u < Dh; u += nw) {
    float s = 0.0f;
    for (int c = lane; c < D; c += 32) s = fmaf(wb[o.cmp_w + (size_t)u * D + c], s_g[c] * inv, s);
    s = warp_sum(s);
    if (lane == 0) s_c[u] = fmaxf(s + wb[o.cmp_b + u], 0.0f);
  }
  __syncthreads();
  for (int c = wid; c < D; c += nw) {
    float s = 0.0f;
    for (int u = lane; u < Dh; u += 32) s = fmaf(wb[o.exp_w + (size_t)c * Dh + u], s_c[u], s);
    s = warp_sum(s);
    if (lane == 0) gvec[(size_t)g * D + c] = sigmoidf(s + wb[o.exp_b + c]);
  }
}

// ------------------------------------------------------------------ head
// logits = W_h f + b_h; probs = softmax(logits)   (model.py:536, 543-546); warp per node
__global__ void __launch_bounds__(256)
k_head(const float* __restrict__ f, const float* __restrict__ wb, NetOffsets o,
       const int* __restrict__ sizes, float* __restrict__ logits, float* __restrict__ probs) {
  const int D = o.D, lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= sizes[0]) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float fv = f[(size_t)v * D + c];
    s0 = fmaf(wb[o.head_w + c], fv, s0);
    s1 = fmaf(wb[o.head_w + D + c], fv, s1);
    s2 = fmaf(wb[o.head_w + 2 * D + c], fv, s2);
  }
  s0 = warp_sum(s0) + wb[o.head_b];
  s1 = warp_sum(s1) + wb[o.head_b + 1];
  s2 = warp_sum(s2) + wb[o.head_b + 2];
  if (lane == 0) {
    if (logits) { logits[(size_t)v * 3] = s0; logits[(size_t)v * 3 + 1] = s1; logits[(size_t)v * 3 + 2] = s2; }
    if (probs) {
      const float m = fmaxf(s0, fmaxf(s1, s2));
      const float e0 = expf(s0 - m), e1 = expf(s1 - m), e2 = expf(s2 - m);
      const float inv = 1.0f / (e0 + e1 + e2);
      probs[(size_t)v * 3] = e0 * inv; probs[(size_t)v * 3 + 1] = e1 * inv; probs[(size_t)v * 3 + 2] = e2 * inv;
    }
  }
}

// ------------------------------------------------------------------ SIMT fp32 GEMM (validation path)
// C[M,N] (+)= act(A[M,K] W[N,K]^T + bias).  64x64 tiles, 256 threads, 4x4 micro-tiles.
template <int ACT>   // 0 none, 1 gelu, 2 sigmoid
__global__ void __launch_bounds__(256)
k_gemm_simt(const float* __restrict__ A, const float* __restrict__ Wt, const float* __restrict__ bias,
            float* __restrict__ C, const int* __restrict__ m_ptr, int N, int K, int accumulate) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int M = *m_ptr;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  if (m0 >= M) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      const int m = m0 + r, n = n0 + r, k = k0 + kk;
      sA[kk][r] = (m < M && k < K) ? A[(size_t)m * K + k] : 0.0f;
      sB[kk][r] = (n < N && k < K) ? Wt[(size_t)n * K + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.0f);
      if (accumulate) v += C[(size_t)m * N + n];
      if (ACT == 1) v = gelu_erf(v);
      if (ACT == 2) v = 1.0f / (1.0f + expf(-v));        // the validation GEMM keeps libm
      C[(size_t)m * N + n] = v;
    }
  }
}

int gemm_simt(gg_context* ctx, cudaStream_t st, const float* A, const float* W, const float* bias,
              float* C, const int* m_ptr, long long m_cap, int N, int K, int act, int accumulate) {
  dim3 grid(ceil_div(m_cap, 64), ceil_div(N, 64));
  if (act == 0) GG_LAUNCH(ctx, k_gemm_simt<0>, grid, 256, 0, st, A, W, bias, C, m_ptr, N, K, accumulate);
  else if (act == 1) GG_LAUNCH(ctx, k_gemm_simt<1>, grid, 256, 0, st, A, W, bias, C, m_ptr, N, K, accumulate);
  else GG_LAUNCH(ctx, k_gemm_simt<2>, grid, 256, 0, st, A, W, bias, C, m_ptr, N, K, accumulate);
  return GG_OK;
}

int gemm(gg_context* ctx, cudaStream_t st, int which, const float* A, const float* W,
         const float* bias, float* C, const int* m_ptr, long long m_cap, int N, int K, int act,
         int accumulate) {
  if (ctx->gemm_impl == 1 && gemm_tc_supported(ctx, which, N, K))
    return gemm_tc(ctx, st, which, A, bias, C, m_ptr, m_cap, N, K, act, accumulate);
  return gemm_simt(ctx, st, A, W, bias, C, m_ptr, m_cap, N, K, act, accumulate);
}

// ------------------------------------------------------------------ weights
static size_t push(std::vector<float>& blob, const float* p, size_t n) {
  size_t off = blob.size();
  off = (off + 3) & ~size_t(3);
  blob.resize(off + n);
  memcpy(blob.data() + off, p, n * sizeof(float));
  return off;
}

int load_weights(gg_context* ctx, const gg_resgcn_weights* w) {
  GG_REQUIRE(w != nullptr, "load_weights: null");
  const int D = w->hidden, n = w->n_layers;
  GG_REQUIRE(D >= 32 && D <= 256 && D % 32 == 0, "load_weights: hidden must be a multiple of 32 in [32,256] (got %d)", D);
  GG_REQUIRE(n >= 0 && n <= 64, "load_weights: n_layers out of range");
  NetWeights& nw = ctx->net;
  nw.loaded = false;
  nw.D = D; nw.n_layers = n; nw.q = D / 4 > 8 ? D / 4 : 8; nw.c = D / 2 > 8 ? D / 2 : 8;
  const int q = nw.q, c = nw.c;
  std::vector<float> blob;
  // softmax(jk_logits)   (model.py:532)
  {
    std::vector<float> jk(n + 2);
    float mx = -1e30f;
    for (int i = 0; i < n + 2; ++i) mx = fmaxf(mx, w->jk_logits[i]);
    float tot = 0.f;
    for (int i = 0; i < n + 2; ++i) { jk[i] = expf(w->jk_logits[i] - mx); tot += jk[i]; }
    for (int i = 0; i < n + 2; ++i) jk[i] /= tot;
    nw.jk_w = push(blob, jk.data(), n + 2);
  }
  // eval BatchNorm folded to scale/shift (model.py:191-213; eps 1e-5)
  {
    float sc[19], sh[19];
    for (int i = 0; i < 19; ++i) {
      sc[i] = w->in_norm_weight[i] / sqrtf(w->in_norm_var[i] + 1e-5f);
      sh[i] = w->in_norm_bias[i] - w->in_norm_mean[i] * sc[i];
    }
    nw.bn_scale = push(blob, sc, 19);
    nw.bn_shift = push(blob, sh, 19);
  }
  nw.w_in = push(blob, w->input_proj_0_weight, (size_t)D * 19);
  nw.b_in = push(blob, w->input_proj_0_bias, D);
  nw.ln_in_g = push(blob, w->input_proj_1_weight, D);
  nw.ln_in_b = push(blob, w->input_proj_1_bias, D);
  nw.pb0_w = push(blob, w->prior_booster_0_weight, (size_t)q * 3);
  nw.pb0_b = push(blob, w->prior_booster_0_bias, q);
  nw.pb2_w = push(blob, w->prior_booster_2_weight, (size_t)D * q);
  nw.pb2_b = push(blob, w->prior_booster_2_bias, D);
  nw.ee0_w = push(blob, w->edge_enc_0_weight, (size_t)c * 5);
  nw.ee0_b = push(blob, w->edge_enc_0_bias, c);
  nw.ee2_w = push(blob, w->edge_enc_2_weight, (size_t)c * c);
  nw.ee2_b = push(blob, w->edge_enc_2_bias, c);
  nw.eg_ln_g = push(blob, w->edge_gate_0_weight, c);
  nw.eg_ln_b = push(blob, w->edge_gate_0_bias, c);
  nw.eg_w = push(blob, w->edge_gate_1_weight, (size_t)D * c);
  nw.eg_b = push(blob, w->edge_gate_1_bias, D);
  nw.gcn_w.resize(n); nw.gcn_b.resize(n); nw.norm_g.resize(n); nw.norm_b.resize(n);
  for (int i = 0; i < n; ++i) {
    nw.gcn_w[i] = push(blob, w->gcn_lin_weight[i], (size_t)D * D);
    nw.gcn_b[i] = push(blob, w->gcn_bias[i], D);
    nw.norm_g[i] = push(blob, w->norm_weight[i], D);
    nw.norm_b[i] = push(blob, w->norm_bias[i], D);
  }
  nw.sage_wl = push(blob, w->sage_lin_l_weight, (size_t)D * D);
  nw.sage_bl = push(blob, w->sage_lin_l_bias, D);
  nw.sage_wr = push(blob, w->sage_lin_r_weight, (size_t)D * D);
  nw.sage_ln_g = push(blob, w->sage_norm_weight, D);
  nw.sage_ln_b = push(blob, w->sage_norm_bias, D);
  nw.attn_w = push(blob, w->ctx_attn_weight, D);
  nw.attn_b = push(blob, w->ctx_attn_bias, 1);
  nw.cmp_w = push(blob, w->ctx_compress_weight, (size_t)(D / 2) * D);
  nw.cmp_b = push(blob, w->ctx_compress_bias, D / 2);
  nw.exp_w = push(blob, w->ctx_expand_weight, (size_t)D * (D / 2));
  nw.exp_b = push(blob, w->ctx_expand_bias, D);
  nw.fuse_ln_g = push(blob, w->fuse_0_weight, D);
  nw.fuse_ln_b = push(blob, w->fuse_0_bias, D);
  nw.fuse_w = push(blob, w->fuse_1_weight, (size_t)D * D);
  nw.fuse_b = push(blob, w->fuse_1_bias, D);
  nw.head_w = push(blob, w->head_weight, (size_t)3 * D);
  nw.head_b = push(blob, w->head_bias, 3);

  GG_CUDA_OK(cudaSetDevice(ctx->device));
  if (nw.blob) {
    GG_CUDA_OK(cudaDeviceSynchronize());     // queued work may still read the old weights
    cudaFree(nw.blob);
    nw.blob = nullptr;
  }
  GG_CUDA_OK(cudaMalloc(&nw.blob, blob.size() * sizeof(float)));
  GG_CUDA_OK(cudaMemcpy(nw.blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  nw.blob_floats = blob.size();
  nw.h_jk.assign(blob.begin() + nw.jk_w, blob.begin() + nw.jk_w + n + 2);
  GG_TRY(gemm_tc_prepare_weights(ctx, blob));
  nw.loaded = true;
  return GG_OK;
}

static NetOffsets make_offsets(const NetWeights& nw) {
  NetOffsets o;
  o.D = nw.D; o.q = nw.q; o.c = nw.c;
  o.jk_w = nw.jk_w; o.bn_scale = nw.bn_scale; o.bn_shift = nw.bn_shift;
  o.w_in = nw.w_in; o.b_in = nw.b_in; o.ln_in_g = nw.ln_in_g; o.ln_in_b = nw.ln_in_b;
  o.pb0_w = nw.pb0_w; o.pb0_b = nw.pb0_b; o.pb2_w = nw.pb2_w; o.pb2_b = nw.pb2_b;
  o.ee0_w = nw.ee0_w; o.ee0_b = nw.ee0_b; o.ee2_b = nw.ee2_b; o.eg_ln_g = nw.eg_ln_g; o.eg_ln_b = nw.eg_ln_b;
  o.attn_w = nw.attn_w; o.attn_b = nw.attn_b; o.cmp_w = nw.cmp_w; o.cmp_b = nw.cmp_b;
  o.exp_w = nw.exp_w; o.exp_b = nw.exp_b; o.head_w = nw.head_w; o.head_b = nw.head_b;
  return o;
}

size_t resgcn_workspace_bytes(const NetWeights& nw, long long node_cap, long long edge_cap, int n_graphs) {
  const int D = nw.D, c = nw.c;
  size_t s = 0;
  s += Arena::padded((size_t)node_cap * D, 4) * 5;      // h, z, gate, t0, t1
  s += Arena::padded((size_t)std::max<long long>(std::max<long long>(edge_cap, node_cap), 1) * c, 4) * 2;   // e1, enc
  s += Arena::padded((size_t)node_cap * c, 4);          // ctx
  s += Arena::padded((size_t)node_cap, 4) * 3;          // node_graph, dinv, score
  s += Arena::padded((size_t)node_cap, 8);              // row_stats
  s += Arena::padded((size_t)n_graphs * D, 4);          // gvec
  s += Arena::padded((size_t)n_graphs, 4) + Arena::padded((size_t)n_graphs * 32 * (D + 1), 4);   // split readout
  s += 4096;
  return s;
}

// dispatch on D/32
#define GG_CPL_SWITCH(D, CALL)                      \
  switch ((D) / 32) {                               \
    case 1: { constexpr int CPL = 1; CALL; } break; \
    case 2: { constexpr int CPL = 2; CALL; } break; \
    case 3: { constexpr int CPL = 3; CALL; } break; \
    case 4: { constexpr int CPL = 4; CALL; } break; \
    case 5: { constexpr int CPL = 5; CALL; } break; \
    case 6: { constexpr int CPL = 6; CALL; } break; \
    case 7: { constexpr int CPL = 7; CALL; } break; \
    case 8: { constexpr int CPL = 8; CALL; } break; \
    default: GG_REQUIRE(false, "unsupported hidden width %d", (D)); \
  }

int resgcn_forward(gg_context* ctx, Arena& ar, const float* x, const int32_t* rowptr,
                   const int32_t* src, const int32_t* eid, const float* edge_attr,
                   const int64_t* graph_off, int n_graphs, long long node_cap, long long edge_cap,
                   float* logits, float* probs, cudaStream_t st, int graph_node_cap, int graph_edge_cap) {
  NetWeights& nw = ctx->net;
  if (!nw.loaded) { set_error("resgcn_forward: no weights loaded (gg_load_weights)"); return GG_ERR_STATE; }
  GG_REQUIRE(n_graphs > 0 && node_cap > 0 && edge_cap >= 0, "resgcn_forward: bad sizes");
  GG_REQUIRE(node_cap < (1ll << 31) && edge_cap < (1ll << 31), "resgcn_forward: batch too large");
  const int D = nw.D, c = nw.c, n = nw.n_layers;
  const NetOffsets o = make_offsets(nw);
  const float* wb = nw.blob;

  float* h = ar.take<float>((size_t)node_cap * D);
  float* z = ar.take<float>((size_t)node_cap * D);
  float* gate = ar.take<float>((size_t)node_cap * D);
  float* t0 = ar.take<float>((size_t)node_cap * D);
  float* t1 = ar.take<float>((size_t)node_cap * D);
  const size_t enc_rows = (size_t)std::max<long long>(std::max<long long>(edge_cap, node_cap), 1);
  float* e1 = ar.take<float>(enc_rows * c);
  float* enc = ar.take<float>(enc_rows * c);
  float* ctxv = ar.take<float>((size_t)node_cap * c);
  int* node_graph = ar.take<int>((size_t)node_cap);
  float* dinv = ar.take<float>((size_t)node_cap);
  float* score = ar.take<float>((size_t)node_cap);
  float* gvec = ar.take<float>((size_t)n_graphs * D);
  float2* row_stats = ar.take<float2>((size_t)node_cap);
  int* ctx_gmax = ar.take<int>((size_t)n_graphs);
  float* ctx_part = ar.take<float>((size_t)n_graphs * CTX_PARTS * (D + 1));
  int* sizes = ar.take<int>(64);
  const int* n_nodes_p = sizes;
  const int* n_edges_p = sizes + 1;

  const int warp_blocks = ceil_div(node_cap, 8);   // warp-per-node kernels, 8 warps per block
  GG_LAUNCH(ctx, k_sizes, 1, 1, 0, st, graph_off, n_graphs, rowptr, sizes, node_cap, edge_cap, ctx->status_word);
  GG_LAUNCH(ctx, k_node_meta, ceil_div(node_cap, 256), 256, 0, st, graph_off, n_graphs, rowptr, src,
            sizes, node_graph, dinv);

  // ---- input stage
  {
    const size_t smem = ((size_t)D * 19 + (size_t)nw.q * 3 + (size_t)D * nw.q) * sizeof(float);
    // few resident blocks per SM, many nodes per warp: the 26 KB weight image is staged once per block
    const int blocks = min(warp_blocks, ctx->sm_count * 3);
    static const int in_r = getenv("GG_IN_R") ? atoi(getenv("GG_IN_R")) : 4;      // nodes per warp iteration (A/B knob)
    GG_CPL_SWITCH(D, {
      if (in_r >= 4 && CPL <= 4) {
        GG_SMEM_ATTR_ONCE(ctx, 1 + CPL, (k_input_stage<CPL, 4>), smem);
        GG_LAUNCH(ctx, (k_input_stage<CPL, 4>), min(blocks, ctx->sm_count * 2), 256, smem, st, x, wb, o, sizes, h, z, row_stats);
      } else if (in_r >= 2) {
        if (smem > 48 * 1024)
          GG_CUDA_OK(cudaFuncSetAttribute(k_input_stage<CPL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GG_LAUNCH(ctx, (k_input_stage<CPL, 2>), blocks, 256, smem, st, x, wb, o, sizes, h, z, row_stats);
      } else {
        if (smem > 48 * 1024)
          GG_CUDA_OK(cudaFuncSetAttribute(k_input_stage<CPL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GG_LAUNCH(ctx, (k_input_stage<CPL, 1>), blocks, 256, smem, st, x, wb, o, sizes, h, z, row_stats);
      }
    });
  }
  // ---- edge context -> gate
  const bool use_tc = ctx->gemm_impl == 1;
  static const bool edge_legacy = getenv("GG_EDGE_LEGACY") != nullptr;
  if (!edge_legacy && c <= 256) {
    // mean of the first-layer features per node, then the (linear) second layer on node rows
    float* m = e1;        // [node_cap, c] fits: e1 / enc hold max(edge_cap, node_cap) rows
    float* tc_ = enc;
    switch ((c + 31) / 32) {
      case 1: GG_LAUNCH(ctx, k_edge_gelu_mean<1>, warp_blocks, 256, 0, st, edge_attr, rowptr, eid, wb, o, sizes, m); break;
      case 2: GG_LAUNCH(ctx, k_edge_gelu_mean<2>, warp_blocks, 256, 0, st, edge_attr, rowptr, eid, wb, o, sizes, m); break;
      case 3: GG_LAUNCH(ctx, k_edge_gelu_mean<3>, warp_blocks, 256, 0, st, edge_attr, rowptr, eid, wb, o, sizes, m); break;
      case 4: GG_LAUNCH(ctx, k_edge_gelu_mean<4>, warp_blocks, 256, 0, st, edge_attr, rowptr, eid, wb, o, sizes, m); break;
      default: GG_LAUNCH(ctx, k_edge_gelu_mean<8>, warp_blocks, 256, 0, st, edge_attr, rowptr, eid, wb, o, sizes, m); break;
    }
    GG_TRY(gemm(ctx, st, GEMM_ENC2, m, wb + nw.ee2_w, nullptr, tc_, n_nodes_p, node_cap, c, c, 0, 0));
    GG_LAUNCH(ctx, k_edge_ctx_ln, warp_blocks, 256, 0, st, tc_, rowptr, wb, o, sizes, ctxv);
  } else {
    if (edge_cap > 0) {
      if (use_tc && gemm_tc_supported(ctx, GEMM_ENC2, c, c) && c == 64) {
        TcPrologue pro;                                   // first encoder layer fused into the A producer
        pro.mode = 2; pro.w0 = wb + nw.ee0_w; pro.b0 = wb + nw.ee0_b;
        GG_TRY(gemm_tc(ctx, st, GEMM_ENC2, edge_attr, wb + nw.ee2_b, enc, n_edges_p, edge_cap, c, c, 0, 0, &pro));
      } else {
        GG_LAUNCH(ctx, k_edge_enc1, ceil_div(edge_cap * c, 256), 256, 0, st, edge_attr, wb, o, sizes, e1);
        GG_TRY(gemm(ctx, st, GEMM_ENC2, e1, wb + nw.ee2_w, wb + nw.ee2_b, enc, n_edges_p, edge_cap, c, c, 0, 0));
      }
    }
    GG_LAUNCH(ctx, k_edge_ctx, warp_blocks, 256, 0, st, enc, rowptr, eid, wb, o, sizes, ctxv);
  }
  GG_TRY(gemm(ctx, st, GEMM_GATE, ctxv, wb + nw.eg_w, wb + nw.eg_b, gate, n_nodes_p, node_cap, D, c, 2, 0));

  // ---- residual GCN blocks
  // the per-graph kernel wins when there are enough graphs to fill the SMs (one CTA per graph); for a handful
  // of graphs the layer-wise kernels, which spread every graph over many SMs, have the shorter latency
  static const int fused_min_graphs = getenv("GG_FUSED_MIN_GRAPHS") ? atoi(getenv("GG_FUSED_MIN_GRAPHS")) : 24;
  const bool fused = graph_node_cap > 0 && n > 0 && (n_graphs >= fused_min_graphs || ctx->gcn_fused == 2) &&
                     gcn_fused_supported(ctx, graph_node_cap, graph_edge_cap);
  if (fused)
    GG_TRY(gcn_layers_fused(ctx, st, h, z, gate, row_stats, dinv, rowptr, src, graph_off, n_graphs, graph_node_cap,
                            graph_edge_cap));
  for (int l = 0; l < n && !fused; ++l) {
    if (use_tc && gemm_tc_supported(ctx, GEMM_GCN0 + l, D, D) && D == 128) {
      TcPrologue pro;                                   // LayerNorm fused into the A producer
      pro.mode = 1; pro.ln_g = wb + nw.norm_g[l]; pro.ln_b = wb + nw.norm_b[l];
      pro.row_stats = row_stats;                        // written by the kernel that produced h
      GG_TRY(gemm_tc(ctx, st, GEMM_GCN0 + l, h, nullptr, t1, n_nodes_p, node_cap, D, D, 0, 0, &pro));
    } else {
      GG_CPL_SWITCH(D, {
        GG_LAUNCH(ctx, k_layernorm<CPL>, warp_blocks, 256, 0, st, h, wb + nw.norm_g[l], wb + nw.norm_b[l],
                  (const float*)nullptr, (const int*)nullptr, sizes, t0);
      });
      GG_TRY(gemm(ctx, st, GEMM_GCN0 + l, t0, wb + nw.gcn_w[l], nullptr, t1, n_nodes_p, node_cap, D, D, 0, 0));
    }
    static const int agg_threads = getenv("GG_AGG_THREADS") ? atoi(getenv("GG_AGG_THREADS")) : 256;
    GG_CPL_SWITCH(D, {
      GG_LAUNCH(ctx, k_gcn_aggregate<CPL>, ceil_div(node_cap, agg_threads / 32), agg_threads, 0, st, t1, rowptr, src,
                dinv, wb + nw.gcn_b[l], gate, nw.h_jk[l + 1], sizes, h, z, row_stats);
    });
  }
  // ---- SAGE branch
  GG_CPL_SWITCH(D, { GG_LAUNCH(ctx, k_sage_mean<CPL>, warp_blocks, 256, 0, st, h, rowptr, src, sizes, t0); });
  GG_TRY(gemm(ctx, st, GEMM_SAGE_L, t0, wb + nw.sage_wl, wb + nw.sage_bl, t1, n_nodes_p, node_cap, D, D, 0, 0));
  GG_TRY(gemm(ctx, st, GEMM_SAGE_R, h, wb + nw.sage_wr, nullptr, t1, n_nodes_p, node_cap, D, D, 0, 1));
  GG_CPL_SWITCH(D, {
    GG_LAUNCH(ctx, k_sage_finish<CPL>, warp_blocks, 256, 0, st, t1, wb + nw.sage_ln_g, wb + nw.sage_ln_b,
              nw.h_jk[n + 1], sizes, z);
  });
  // ---- global context, fuse, head
  {
    const size_t smem = (size_t)(D + D / 2) * sizeof(float);
    if (ceil_div(node_cap, n_graphs) <= 1024) {
      GG_LAUNCH(ctx, k_graph_context, n_graphs, GC_WARPS * 32, smem, st, z, graph_off, wb, o, n_graphs, score, gvec);
    } else {
      // large graphs: the readout is split over CTX_PARTS blocks per graph
      GG_CUDA_OK(cudaMemsetAsync(ctx_gmax, 0x80, (size_t)n_graphs * sizeof(int), st));   // ordered(-huge)
      GG_LAUNCH(ctx, k_ctx_scores, warp_blocks, 256, 0, st, z, node_graph, wb, o, sizes, score, ctx_gmax);
      dim3 grid(CTX_PARTS, n_graphs);
      GG_LAUNCH(ctx, k_ctx_partial, grid, 256, 0, st, z, graph_off, score, ctx_gmax, D, ctx_part);
      GG_LAUNCH(ctx, k_ctx_finish, n_graphs, 256, smem, st, ctx_part, wb, o, n_graphs, gvec);
    }
  }
  if (use_tc && gemm_tc_supported(ctx, GEMM_FUSE, D, D) && D == 128) {
    TcPrologue pro;                                     // z * gvec[graph] -> LayerNorm fused into the A producer
    pro.mode = 1; pro.ln_g = wb + nw.fuse_ln_g; pro.ln_b = wb + nw.fuse_ln_b;
    pro.gvec = gvec; pro.node_graph = node_graph;
    GG_TRY(gemm_tc(ctx, st, GEMM_FUSE, z, wb + nw.fuse_b, t1, n_nodes_p, node_cap, D, D, 1, 0, &pro));
  } else {
    GG_CPL_SWITCH(D, {
      GG_LAUNCH(ctx, k_layernorm<CPL>, warp_blocks, 256, 0, st, z, wb + nw.fuse_ln_g, wb + nw.fuse_ln_b,
                (const float*)gvec, (const int*)node_graph, sizes, t0);
    });
    GG_TRY(gemm(ctx, st, GEMM_FUSE, t0, wb + nw.fuse_w, wb + nw.fuse_b, t1, n_nodes_p, node_cap, D, D, 1, 0));
  }
  GG_LAUNCH(ctx, k_head, warp_blocks, 256, 0, st, t1, wb, o, sizes, logits, probs);
  return GG_OK;
}

// ------------------------------------------------------------------ COO -> CSR
__global__ void k_coo_count(const int64_t* __restrict__ ei, long long E, long long N, int* __restrict__ cnt,
                            int* status) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long d = ei[E + e], s = ei[e];
  if (d < 0 || d >= N || s < 0 || s >= N) { atomicOr(status, ST_LABEL_RANGE); return; }
  atomicAdd(&cnt[d], 1);
}
__global__ void __launch_bounds__(1024)
k_coo_scan(int* __restrict__ cnt, int32_t* __restrict__ rowptr, long long N) {
  __shared__ int scratch[40];
  int carry = 0;
  for (long long base = 0; base < N; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    const int v = (i < N) ? cnt[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, scratch, &total);
    if (i < N) { rowptr[i] = carry + ex; cnt[i] = 0; }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) rowptr[N] = carry;
}
__global__ void k_coo_fill(const int64_t* __restrict__ ei, long long E, long long N,
                           const int32_t* __restrict__ rowptr, int* __restrict__ cursor,
                           int32_t* __restrict__ src, int32_t* __restrict__ eid) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long d = ei[E + e], s = ei[e];
  if (d < 0 || d >= N || s < 0 || s >= N) return;
  const int pos = rowptr[d] + atomicAdd(&cursor[d], 1);
  src[pos] = (int)s;
  eid[pos] = (int)e;
}
__global__ void k_coo_sort_rows(const int32_t* __restrict__ rowptr, long long N, int32_t* __restrict__ src,
                                int32_t* __restrict__ eid) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int s0 = rowptr[v], s1 = rowptr[v + 1];
  for (int a = s0 + 1; a < s1; ++a) {
    const int sv = src[a], ev = eid[a];
    int j = a - 1;
    while (j >= s0 && (src[j] > sv || (src[j] == sv && eid[j] > ev))) {
      src[j + 1] = src[j]; eid[j + 1] = eid[j]; --j;
    }
    src[j + 1] = sv; eid[j + 1] = ev;
  }
}

int coo_to_csr(gg_context* ctx, Arena& ar, const int64_t* ei, long long E, long long N,
               int32_t* rowptr, int32_t* src, int32_t* eid, cudaStream_t st) {
  GG_REQUIRE(N > 0 && E >= 0 && N < (1ll << 31) && E < (1ll << 31), "coo_to_csr: bad sizes");
  int* cnt = ar.take<int>((size_t)N);
  GG_CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)N * sizeof(int), st));
  GG_TRY(status_epoch(ctx, st));
  if (E > 0) GG_LAUNCH(ctx, k_coo_count, ceil_div(E, 256), 256, 0, st, ei, E, N, cnt, ctx->status_word);
  GG_LAUNCH(ctx, k_coo_scan, 1, 1024, 0, st, cnt, rowptr, N);
  if (E > 0) GG_LAUNCH(ctx, k_coo_fill, ceil_div(E, 256), 256, 0, st, ei, E, N, rowptr, cnt, src, eid);
  GG_LAUNCH(ctx, k_coo_sort_rows, ceil_div(N, 256), 256, 0, st, rowptr, N, src, eid);
  return GG_OK;
}

}  // namespace gg
