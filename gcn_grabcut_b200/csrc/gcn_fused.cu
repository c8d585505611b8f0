// The residual GCN blocks of ResGCNNet (model.py:523-528), all n layers in ONE kernel, one CTA per
// graph, for graphs of at most FUSED_MAX_NODES regions and hidden width 128:
//
//     for l in 1..n:   u = GCNConv_l(LN_l(h));   h = h + GELU(u * gate);   z += w_l h
//
// Per layer and graph the CTA
//   1. builds the A operand of the node transform tile by tile: LayerNorm(h) rows (row statistics
//      kept in shared memory from the previous layer), split exactly into three bf16 terms, K-major
//      SWIZZLE_128B images in shared memory (one 64-wide K atom of a 128-row tile at a time);
//   2. multiplies by W_l on the tensor cores (tcgen05.mma, 6 bf16 products = fp32-class accuracy,
//      fp32 accumulators in TMEM: 128 columns per 128-row tile); W_l (96 KB, pre-split and
//      pre-swizzled at gg_load_weights time) arrives by one bulk-TMA copy that was issued while the
//      previous layer was aggregating;
//   3. reads the transformed rows x' back from TMEM a quarter of the channels at a time
//      (tcgen05.ld) into shared memory and runs the degree-normalised neighbour sum there:
//      8 lanes per node, one 128-bit shared-memory read per neighbour and lane, the CSR slice of the
//      graph and the d^-1/2 table in shared memory as well -- x' never touches L2 / HBM and the
//      gather (11 neighbour rows per node and layer) runs at shared-memory bandwidth;
//   4. applies bias, gate, GELU, the residual and the jumping-knowledge accumulation on the row
//      slices of h / z / gate in global memory (coalesced 128-byte pieces, L2-resident) and
//      accumulates the row statistics of the new h for the next layer's LayerNorm.
//
// The layer-wise path (gemm_tc.cu + k_gcn_aggregate) remains for larger graphs and other widths.
#include <stdlib.h>

#include "common.cuh"
#include "resgcn.cuh"
#include "tc_ptx.cuh"

namespace gg {

constexpr int FUSED_THREADS = 1024;
constexpr int FUSED_D = 128;
constexpr uint32_t FUSED_B_BYTES = 3u * 2u * 128u * 128u;   // W image: 3 splits x 2 K-atoms x 128 rows x 128 B
constexpr uint32_t FUSED_A_BYTES = 3u * 128u * 128u;        // A chunk: 3 splits x 128 rows x 128 B (one K atom)
constexpr uint32_t FUSED_A_SPLIT = 128u * 128u;
constexpr uint32_t FUSED_B_SPLIT = 2u * 128u * 128u;
constexpr uint32_t FUSED_B_ATOM = 128u * 128u;

struct FusedParams {
  float* h;                   // [SN,128] residual stream, updated in place
  float* z;                   // [SN,128] jumping-knowledge accumulator, updated in place
  const float* gate;          // [SN,128]
  const float2* row_stats;    // [SN] (mean, rstd) of the rows of h on entry
  const float* dinv;          // [SN] (1 + in-degree)^-1/2
  const int32_t* rowptr;      // [SN+1] global CSR
  const int32_t* src;         // [SE]   global source ids
  const int64_t* graph_off;   // [G+1]
  const uint8_t* tc_blob;
  const float* wb;
  unsigned long long w_off[FUSED_MAX_LAYERS];
  unsigned ln_g[FUSED_MAX_LAYERS], ln_b[FUSED_MAX_LAYERS], bias[FUSED_MAX_LAYERS];
  float jkw[FUSED_MAX_LAYERS];
  int n_layers, node_cap, edge_cap;
  int* status;
  volatile int* dbg;          // optional progress words in mapped host memory (GG_DEBUG_PTR), debugging only
};

#define GG_FUSED_DBG(v)                                              \
  do {                                                               \
    if (p.dbg && lane == 0) { p.dbg[warp] = (v); __threadfence_system(); } \
  } while (0)

GG_D float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// swizzled position of the 16-byte chunk c (4 channels) of row r in the x' quarter buffer
// (rows of 128 B): conflict-free both for "8 lanes read one row" and "32 lanes write 32 rows"
GG_D uint32_t xq_offset(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(FUSED_THREADS, 1)
k_gcn_layers_fused(const FusedParams p) {
  extern __shared__ unsigned char fused_smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.x;
  const int v0 = (int)p.graph_off[g];
  const int N = (int)p.graph_off[g + 1] - v0;
  if (N <= 0) return;
  const int NC = p.node_cap;
  const int n_tiles = (N + 127) >> 7;
  // ---- shared memory carve-up
  const uint32_t raw = smem_u32(fused_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = fused_smem_raw + (base - raw);
  const uint32_t sB = base, sA = base + FUSED_B_BYTES;
  unsigned char* pA = smem + FUSED_B_BYTES;
  unsigned char* pX = pA + FUSED_A_BYTES;                              // [NC8][128 B]
  const int NC8 = (NC + 7) & ~7;
  double2* sStat = reinterpret_cast<double2*>(pX + (size_t)NC8 * 128);  // [NC8] sum, sum of squares
  float2* sMean = reinterpret_cast<float2*>(sStat + NC8);               // [NC8] mean, rstd
  float* sDinv = reinterpret_cast<float*>(sMean + NC8);                 // [NC8]
  int* sRow = reinterpret_cast<int*>(sDinv + NC8);                      // [NC8 + 8]
  float* sLnG = reinterpret_cast<float*>(sRow + NC8 + 8);               // [128]
  float* sLnB = sLnG + 128;
  float* sBias = sLnB + 128;
  unsigned char* ctrl_p = reinterpret_cast<unsigned char*>(sBias + 128);   // 64 B: barriers + TMEM pointer
  const uint32_t ctrl = base + (uint32_t)(ctrl_p - smem);
  const uint32_t bar_w = ctrl, bar_mma = ctrl + 8, tmem_slot = ctrl + 16;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(ctrl_p + 16);
  uint16_t* sSrc = reinterpret_cast<uint16_t*>(ctrl_p + 64);            // [edge_cap]

  const int e_base = p.rowptr[v0];
  const int E = p.rowptr[v0 + N] - e_base;
  if (N > NC || E > p.edge_cap) {                  // cannot happen for graphs built with these capacities
    if (tid == 0) atomicOr(p.status, ST_EDGE_CAP);
    return;
  }

  GG_FUSED_DBG(1);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = n_tiles <= 1 ? 128u : (n_tiles == 2 ? 256u : 512u);   // 128 columns per row tile
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  GG_FUSED_DBG(2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  GG_FUSED_DBG(3);
  const uint32_t tmem_base = *tmem_slot_p;
  if (tid == 0) {
    mbar_expect_tx(bar_w, FUSED_B_BYTES);
    bulk_g2s(sB, p.tc_blob + p.w_off[0], FUSED_B_BYTES, bar_w);
  }
  // ---- graph slice: CSR (local ids), d^-1/2, the row statistics of the incoming h
  for (int i = tid; i <= N; i += FUSED_THREADS) sRow[i] = p.rowptr[v0 + i] - e_base;
  for (int i = tid; i < N; i += FUSED_THREADS) {
    sDinv[i] = p.dinv[v0 + i];
    sMean[i] = p.row_stats[v0 + i];
    sStat[i] = make_double2(0.0, 0.0);
  }
  for (int e = tid; e < E; e += FUSED_THREADS) sSrc[e] = (uint16_t)(p.src[e_base + e] - v0);
  __syncthreads();
  GG_FUSED_DBG(4);

  const uint32_t idesc = umma_idesc_bf16(128, 128);
  uint32_t mma_batches = 0;                         // MMA batches committed so far (all threads count alike)
  bool dead = false;                                // a barrier timed out (reported in the status word)
  // producer role of this thread: row r of the tile, 16-byte chunk c (8 K values) of the atom
  const int pr_r = tid >> 3, pr_c = tid & 7;
  const uint32_t pr_off = (uint32_t)((pr_r >> 3) * 1024 + (pr_r & 7) * 128 + ((pr_c ^ (pr_r & 7)) << 4));
  // gather role: 8 lanes per node
  const int grp = tid >> 3, sub = tid & 7;

  for (int l = 0; l < p.n_layers; ++l) {
    // ---- layer parameters
    if (tid < 128) {
      sLnG[tid] = p.wb[p.ln_g[l] + tid];
      sLnB[tid] = p.wb[p.ln_b[l] + tid];
      sBias[tid] = p.wb[p.bias[l] + tid];
    }
    __syncthreads();
    // ---- x' = LN(h) W^T into TMEM, one (tile, K atom) batch at a time
    for (int t = 0; t < n_tiles; ++t) {
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
        // A values of this thread: 8 consecutive K of row 128 t + r (the loads are issued before the wait)
        const int row = 128 * t + pr_r;
        float x[8];
        if (row < N) {
          const float4* src4 = reinterpret_cast<const float4*>(p.h + (size_t)(v0 + row) * FUSED_D + 64 * a + 8 * pr_c);
          const float4 q0 = __ldcg(src4), q1 = __ldcg(src4 + 1);
          const float2 ms = sMean[row];
          const float4 g0 = *reinterpret_cast<const float4*>(sLnG + 64 * a + 8 * pr_c);
          const float4 g1 = *reinterpret_cast<const float4*>(sLnG + 64 * a + 8 * pr_c + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sLnB + 64 * a + 8 * pr_c);
          const float4 b1 = *reinterpret_cast<const float4*>(sLnB + 64 * a + 8 * pr_c + 4);
          x[0] = (q0.x - ms.x) * ms.y * g0.x + b0.x; x[1] = (q0.y - ms.x) * ms.y * g0.y + b0.y;
          x[2] = (q0.z - ms.x) * ms.y * g0.z + b0.z; x[3] = (q0.w - ms.x) * ms.y * g0.w + b0.w;
          x[4] = (q1.x - ms.x) * ms.y * g1.x + b1.x; x[5] = (q1.y - ms.x) * ms.y * g1.y + b1.y;
          x[6] = (q1.z - ms.x) * ms.y * g1.z + b1.z; x[7] = (q1.w - ms.x) * ms.y * g1.w + b1.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = 0.0f;
        }
        // the previous batch's MMAs still read the A chunk
        if (mma_batches > 0) mbar_wait_bounded(bar_mma, (mma_batches - 1) & 1, p.status, 0x200, dead);
        // exact 3-way bf16 split by truncation (see gemm_tc.cu)
        uint32_t hb[3][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t b1 = __float_as_uint(x[e]) & 0xFFFF0000u;
          const float r1 = x[e] - __uint_as_float(b1);
          const uint32_t b2 = __float_as_uint(r1) & 0xFFFF0000u;
          const float r2 = r1 - __uint_as_float(b2);
          hb[0][e] = b1; hb[1][e] = b2; hb[2][e] = __float_as_uint(r2);
        }
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) {
          uint4 o;
          o.x = __byte_perm(hb[sp][0], hb[sp][1], 0x7632);
          o.y = __byte_perm(hb[sp][2], hb[sp][3], 0x7632);
          o.z = __byte_perm(hb[sp][4], hb[sp][5], 0x7632);
          o.w = __byte_perm(hb[sp][6], hb[sp][7], 0x7632);
          *reinterpret_cast<uint4*>(pA + sp * FUSED_A_SPLIT + pr_off) = o;
        }
        fence_proxy_async();
        GG_FUSED_DBG(100 + 10 * l + 2 * t + a);
        __syncthreads();
        GG_FUSED_DBG(200 + 10 * l + 2 * t + a);
        if (tid == 0) {
          if (t == 0 && a == 0) mbar_wait_bounded(bar_w, l & 1, p.status, 0x100, dead);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(128 * t);
          const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
          uint32_t acc = a;                            // the first atom starts the accumulator
#pragma unroll
          for (int tt = 5; tt >= 0; --tt) {              // small terms first
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = umma_desc(sA + pa[tt] * FUSED_A_SPLIT + kk * 32);
              const uint64_t bd = umma_desc(sB + pb[tt] * FUSED_B_SPLIT + a * FUSED_B_ATOM + kk * 32);
              umma_bf16(tmem_d, ad, bd, idesc, acc);
              acc = 1;
            }
          }
          umma_commit(bar_mma);
        }
        ++mma_batches;
      }
    }
    // all MMAs of the layer done: the accumulators are complete, W_l and the A chunk are free
    GG_FUSED_DBG(300 + l);
    mbar_wait_bounded(bar_mma, (mma_batches - 1) & 1, p.status, 0x400, dead);
    tc_fence_after();
    GG_FUSED_DBG(400 + l);
    if (tid == 0 && l + 1 < p.n_layers) {
      mbar_expect_tx(bar_w, FUSED_B_BYTES);
      bulk_g2s(sB, p.tc_blob + p.w_off[l + 1], FUSED_B_BYTES, bar_w);
    }
    const float jkw = p.jkw[l];
    // ---- aggregation + epilogue, a quarter of the channels at a time
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      if (warp < 4 * n_tiles) {
        const int tile = warp >> 2, lq = warp & 3;
        const int row = 128 * tile + 32 * lq + lane;
        uint32_t rr[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * lq) << 16) + (uint32_t)(128 * tile + 32 * q), rr);
        tmem_ld_wait();
        if (row < N) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(pX + xq_offset(row, c)) = make_uint4(rr[4 * c], rr[4 * c + 1], rr[4 * c + 2], rr[4 * c + 3]);
        }
      }
      GG_FUSED_DBG(500 + 10 * l + q);
      tc_fence_before();
      __syncthreads();
      GG_FUSED_DBG(600 + 10 * l + q);
      // warp-uniform trip count (the 8-lane groups of a warp shuffle with the full mask): groups
      // past the last node run the iteration with no edges and no memory traffic
      for (int vb = 0; vb < N; vb += FUSED_THREADS / 8) {
        const int v = vb + grp;
        const bool act = v < N;
        const int vc = act ? v : 0;
        const float dv = sDinv[vc];
        const float4 self = *reinterpret_cast<const float4*>(pX + xq_offset(vc, sub));
        const float dd = dv * dv;
        float4 acc = make_float4(dd * self.x, dd * self.y, dd * self.z, dd * self.w);
        const int e0 = act ? sRow[vc] : 0, e1 = act ? sRow[vc + 1] : 0;
        for (int e = e0; e < e1; e += 4) {
          int s4[4];
          float w4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = e + u < e1;
            s4[u] = ok ? (int)sSrc[e + u] : vc;
            w4[u] = (ok && s4[u] != vc) ? dv * sDinv[s4[u]] : 0.0f;
          }
          float4 r4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) r4[u] = *reinterpret_cast<const float4*>(pX + xq_offset(s4[u], sub));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc.x = fmaf(w4[u], r4[u].x, acc.x); acc.y = fmaf(w4[u], r4[u].y, acc.y);
            acc.z = fmaf(w4[u], r4[u].z, acc.z); acc.w = fmaf(w4[u], r4[u].w, acc.w);
          }
        }
        const int col = 32 * q + 4 * sub;
        const size_t go = (size_t)(v0 + vc) * FUSED_D + col;
        float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) {
          const float4 b4 = *reinterpret_cast<const float4*>(sBias + col);
          const float4 gt = __ldg(reinterpret_cast<const float4*>(p.gate + go));
          hv = __ldcg(reinterpret_cast<const float4*>(p.h + go));
          float4 zv = __ldcg(reinterpret_cast<const float4*>(p.z + go));
          hv.x += gelu_erf_f((acc.x + b4.x) * gt.x); hv.y += gelu_erf_f((acc.y + b4.y) * gt.y);
          hv.z += gelu_erf_f((acc.z + b4.z) * gt.z); hv.w += gelu_erf_f((acc.w + b4.w) * gt.w);
          zv.x += jkw * hv.x; zv.y += jkw * hv.y; zv.z += jkw * hv.z; zv.w += jkw * hv.w;
          __stcg(reinterpret_cast<float4*>(p.h + go), hv);
          __stcg(reinterpret_cast<float4*>(p.z + go), zv);
        }
        float s1 = (hv.x + hv.y) + (hv.z + hv.w);
        float s2 = (hv.x * hv.x + hv.y * hv.y) + (hv.z * hv.z + hv.w * hv.w);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (act && sub == 0) {
          double2 st = sStat[v];
          st.x += (double)s1; st.y += (double)s2;
          sStat[v] = st;
        }
      }
      __syncthreads();
    }
    // ---- row statistics of the new h -> next layer's LayerNorm
    for (int i = tid; i < N; i += FUSED_THREADS) {
      const double2 st = sStat[i];
      const double mean = st.x * (1.0 / FUSED_D);
      const double var = fmax(st.y * (1.0 / FUSED_D) - mean * mean, 0.0);
      sMean[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
      sStat[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

size_t gcn_fused_smem_bytes(int node_cap, int edge_cap) {
  const size_t NC8 = (size_t)((node_cap + 7) & ~7);
  return 1024 + FUSED_B_BYTES + FUSED_A_BYTES + NC8 * 128 + NC8 * (16 + 8 + 4) + (NC8 + 8) * 4 + 3 * 128 * 4 + 64 +
         (((size_t)edge_cap * 2 + 15) & ~size_t(15));
}

bool gcn_fused_supported(const gg_context* ctx, int node_cap, int edge_cap) {
  const NetWeights& nw = ctx->net;
  if (!ctx->gcn_fused || ctx->gemm_impl != 1 || nw.D != FUSED_D || nw.n_layers < 1 || nw.n_layers > FUSED_MAX_LAYERS) return false;
  if (node_cap < 1 || node_cap > FUSED_MAX_NODES || edge_cap < 0 || edge_cap > 65535 * 4) return false;
  if (nw.tc_blob == nullptr) return false;
  for (int l = 0; l < nw.n_layers; ++l)
    if (nw.tc_off[GEMM_GCN0 + l] == (size_t)-1) return false;
  return gcn_fused_smem_bytes(node_cap, edge_cap) <= 227 * 1024;
}

int gcn_layers_fused(gg_context* ctx, cudaStream_t st, float* h, float* z, const float* gate,
                     const float2* row_stats, const float* dinv, const int32_t* rowptr, const int32_t* src,
                     const int64_t* graph_off, int n_graphs, int node_cap, int edge_cap) {
  const NetWeights& nw = ctx->net;
  FusedParams p{};
  p.h = h; p.z = z; p.gate = gate; p.row_stats = row_stats; p.dinv = dinv; p.rowptr = rowptr; p.src = src;
  p.graph_off = graph_off; p.tc_blob = reinterpret_cast<const uint8_t*>(nw.tc_blob); p.wb = nw.blob;
  for (int l = 0; l < nw.n_layers; ++l) {
    p.w_off[l] = nw.tc_off[GEMM_GCN0 + l];
    p.ln_g[l] = (unsigned)nw.norm_g[l]; p.ln_b[l] = (unsigned)nw.norm_b[l]; p.bias[l] = (unsigned)nw.gcn_b[l];
    p.jkw[l] = nw.h_jk[l + 1];
  }
  p.n_layers = nw.n_layers; p.node_cap = node_cap; p.edge_cap = edge_cap; p.status = ctx->status_word;
  static volatile int* dbg = getenv("GG_DEBUG_PTR") ? reinterpret_cast<volatile int*>(strtoull(getenv("GG_DEBUG_PTR"), nullptr, 0)) : nullptr;
  p.dbg = dbg;
  const size_t smem = gcn_fused_smem_bytes(node_cap, edge_cap);
  GG_SMEM_ATTR_ONCE(ctx, 50, k_gcn_layers_fused, 227 * 1024);
  GG_LAUNCH(ctx, k_gcn_layers_fused, n_graphs, FUSED_THREADS, smem, st, p);
  return GG_OK;
}

}  // namespace gg
