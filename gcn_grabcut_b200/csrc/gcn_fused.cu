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

// build-time variants (A/B measurements on the GPU): double-buffered A operand,
// per-layer row statistics in registers (node rounds unrolled) instead of shared memory, and the
// 16-instruction GELU instead of erff
#ifndef FUSED_DBUF
#define FUSED_DBUF 0
#endif
#ifndef FUSED_REGSTATS
#define FUSED_REGSTATS 0
#endif
#ifndef FUSED_FASTGELU
#define FUSED_FASTGELU 1
#endif

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

// progress markers (one word per warp in mapped host memory, survive a hang): debug builds only --
// every marker is a system-scope fence
#ifdef GG_FUSED_DEBUG_WAITS
#define GG_FUSED_DBG(v)                                              \
  do {                                                               \
    if (p.dbg && lane == 0) { p.dbg[warp] = (v); __threadfence_system(); } \
  } while (0)
#else
#define GG_FUSED_DBG(v) do { } while (0)
#endif

GG_D float rcp_approx_f(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
GG_D float ex2_approx_f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// GG_FUSED_DEBUG_WAITS: barrier waits give up after ~0.2 s and report in the status word
#ifdef GG_FUSED_DEBUG_WAITS
#define GG_FUSED_WAIT(bar, parity, code) mbar_wait_bounded((bar), (parity), p.status, (code), dead)
#else
#define GG_FUSED_WAIT(bar, parity, code) mbar_wait((bar), (parity))
#endif

// Exact-erf GELU x Phi(x) with Phi from the complementary error function in Abramowitz-Stegun
// form 7.1.26 (|error of erf| <= 1.5e-7, no cancellation on the negative side): 16 instructions
// per value instead of erff's 30 -- the epilogue evaluates 128 of them per node and layer.
GG_D float gelu_erf_f(float x) {
#if !FUSED_FASTGELU
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
#endif
  const float ax = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx_f(fmaf(0.3275911f, ax, 1.0f));
  float pl = fmaf(t, 1.061405429f, -1.453152027f);
  pl = fmaf(t, pl, 1.421413741f);
  pl = fmaf(t, pl, -0.284496736f);
  pl = fmaf(t, pl, 0.254829592f);
  const float half_erfc = 0.5f * t * pl * ex2_approx_f(-1.4426950408889634f * ax * ax);   // Phi(-|x|)
  return x * (x >= 0.0f ? 1.0f - half_erfc : half_erfc);
}

// x' quarter buffer: rows of 32 channels (128 B) at a pitch of 144 B.  "8 lanes read one row" touches
// 128 contiguous bytes; "32 lanes write chunk c of 32 consecutive rows" spreads over 8 bank groups
// (pitch = 36 words), i.e. the minimum of 4 wavefronts for 512 bytes.  Row FUSED_ZERO_ROW(NC8) is
// all zero: self loops and the padding of the 4-edge groups point there, so the inner loop of the
// gather has no weights and no predicates (the d^-1/2 of the source is folded into the rows when
// they are written).
constexpr uint32_t XQ_PITCH = 144;
GG_HD size_t fused_xq_bytes(int nc8) {
  const size_t b = (size_t)(nc8 + 8) * XQ_PITCH;
  return b > FUSED_A_BYTES ? b : FUSED_A_BYTES;
}

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
  const uint32_t sB = base, sA = base + FUSED_B_BYTES, sX = sA + FUSED_A_BYTES;
  unsigned char* pA = smem + FUSED_B_BYTES;
  // x' quarter buffer [NC8 + 1][144 B], last row zero; doubles as the second A buffer (>= 48 KB)
  unsigned char* pX = pA + FUSED_A_BYTES;
  const int NC8 = (NC + 7) & ~7;
  const uint32_t zero_off = (uint32_t)NC8 * XQ_PITCH;
  double2* sStat = reinterpret_cast<double2*>(pX + fused_xq_bytes(NC8));           // [NC8] sum, sum of squares
  float2* sMean = reinterpret_cast<float2*>(sStat + NC8);                          // [NC8] mean, rstd
  float* sDinv = reinterpret_cast<float*>(sMean + NC8);                 // [NC8]
  int* sRow = reinterpret_cast<int*>(sDinv + NC8);                      // [NC8 + 8]
  float* sLnG = reinterpret_cast<float*>(sRow + NC8 + 8);               // [128]
  float* sLnB = sLnG + 128;
  float* sBias = sLnB + 128;
  unsigned char* ctrl_p = reinterpret_cast<unsigned char*>(sBias + 128);   // 64 B: barriers + TMEM pointer
  const uint32_t ctrl = base + (uint32_t)(ctrl_p - smem);
  const uint32_t bar_w = ctrl, bar_mma = ctrl + 8 /* two barriers, one per A buffer */, tmem_slot = ctrl + 24;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(ctrl_p + 24);
  int* scan_scratch = reinterpret_cast<int*>(ctrl_p + 64);             // [40]
  // [edge_cap + 3 NC] byte offset of the source row in pX; every node's list padded to a multiple of 4
  uint16_t* sSrc = reinterpret_cast<uint16_t*>(ctrl_p + 64 + 160);

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
    mbar_init(bar_mma + 8, 1);
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
  // ---- graph slice: d^-1/2, the row statistics of the incoming h, and the CSR with every node's
  // source list padded to a multiple of four entries (sRow = padded row pointers)
  int deg = 0, e_first = 0;
  if (tid < N) {
    e_first = p.rowptr[v0 + tid] - e_base;
    deg = p.rowptr[v0 + tid + 1] - e_base - e_first;
    sDinv[tid] = p.dinv[v0 + tid];
    sMean[tid] = p.row_stats[v0 + tid];
    sStat[tid] = make_double2(0.0, 0.0);
  }
  int padded_total;
  const int p_first = block_exclusive_scan((deg + 3) & ~3, scan_scratch, &padded_total);
  if (tid < N) sRow[tid] = p_first;
  if (tid == 0) sRow[N] = padded_total;
  if (tid < 36) reinterpret_cast<uint32_t*>(pX + zero_off)[tid] = 0u;
  __syncthreads();
  // source rows as byte offsets; a self loop (PyG removes it and adds its own) and the padding read the zero row
  for (int i = tid >> 3; i < N; i += FUSED_THREADS / 8) {
    const int r0 = p.rowptr[v0 + i] - e_base, d_i = p.rowptr[v0 + i + 1] - e_base - r0;
    const int q0 = sRow[i], q1 = sRow[i + 1];
    for (int k = (tid & 7); k < q1 - q0; k += 8) {
      uint32_t off = zero_off;
      if (k < d_i) {
        const int sloc = p.src[e_base + r0 + k] - v0;
        if (sloc != i) off = (uint32_t)sloc * XQ_PITCH;
      }
      sSrc[q0 + k] = (uint16_t)off;
    }
  }
  __syncthreads();
  GG_FUSED_DBG(4);

  const uint32_t idesc = umma_idesc_bf16(128, 128);
  uint32_t mma_cnt[2] = {0, 0};                     // MMA batches committed per A buffer (all threads count alike)
  bool dead = false;                                // a barrier timed out (reported in the status word)
  (void)dead;
  // producer role of this thread: row r of the tile, 16-byte chunk c (8 K values) of the atom
  const int pr_r = tid >> 3, pr_c = tid & 7;
  const uint32_t pr_off = (uint32_t)((pr_r >> 3) * 1024 + (pr_r & 7) * 128 + ((pr_c ^ (pr_r & 7)) << 4));
  // gather role: 8 lanes per node
  const int grp = tid >> 3, sub = tid & 7;

  long long t_mark = clock64(), t_prod = 0, t_drain = 0, t_gather = 0, t_misc = 0;
#define GG_FUSED_LAP(acc_)                                   \
  do {                                                       \
    if (p.dbg) { const long long now_ = clock64(); acc_ += now_ - t_mark; t_mark = now_; } \
  } while (0)
  for (int l = 0; l < p.n_layers; ++l) {
    // ---- layer parameters
    if (tid < 128) {
      sLnG[tid] = p.wb[p.ln_g[l] + tid];
      sLnB[tid] = p.wb[p.ln_b[l] + tid];
      sBias[tid] = p.wb[p.bias[l] + tid];
    }
    __syncthreads();
    // ---- x' = LN(h) W^T into TMEM, one (tile, K atom) batch at a time.  Two A buffers (the second
    // one is the x' quarter buffer, idle during this phase) with one barrier each: batch b + 1 is
    // normalised, split and stored while the tensor core works on batch b, and the rows of h for
    // batch b + 1 are requested before batch b is stored.
    const int n_batches = 2 * n_tiles;
    float4 q0n = make_float4(0.f, 0.f, 0.f, 0.f), q1n = q0n;
    if (pr_r < N) {
      const float4* src4 = reinterpret_cast<const float4*>(p.h + (size_t)(v0 + pr_r) * FUSED_D + 8 * pr_c);
      q0n = __ldcg(src4); q1n = __ldcg(src4 + 1);
    }
#pragma unroll 1
    for (int b = 0; b < n_batches; ++b) {
      const int t = b >> 1, a = b & 1, buf = FUSED_DBUF ? (b & 1) : 0;
      const int row = 128 * t + pr_r;
      const float4 q0 = q0n, q1 = q1n;
      if (b + 1 < n_batches) {
        const int rown = 128 * ((b + 1) >> 1) + pr_r;
        if (rown < N) {
          const float4* src4 =
              reinterpret_cast<const float4*>(p.h + (size_t)(v0 + rown) * FUSED_D + 64 * ((b + 1) & 1) + 8 * pr_c);
          q0n = __ldcg(src4); q1n = __ldcg(src4 + 1);
        }
      }
      float x[8];
      if (row < N) {
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
      // the MMAs of the batch that used this buffer last (two batches ago) must have read it
      if (mma_cnt[buf] > 0) GG_FUSED_WAIT(bar_mma + 8 * buf, (mma_cnt[buf] - 1) & 1, 0x200);
      unsigned char* dstA = (buf ? pX : pA) + pr_off;
#pragma unroll
      for (int sp = 0; sp < 3; ++sp) {
        uint4 o;
        o.x = __byte_perm(hb[sp][0], hb[sp][1], 0x7632);
        o.y = __byte_perm(hb[sp][2], hb[sp][3], 0x7632);
        o.z = __byte_perm(hb[sp][4], hb[sp][5], 0x7632);
        o.w = __byte_perm(hb[sp][6], hb[sp][7], 0x7632);
        *reinterpret_cast<uint4*>(dstA + sp * FUSED_A_SPLIT) = o;
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        if (b == 0) GG_FUSED_WAIT(bar_w, l & 1, 0x100);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(128 * t);
        const uint32_t sAb = buf ? sX : sA;
        const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
        uint32_t acc = a;                            // the first atom starts the accumulator
#pragma unroll
        for (int tt = 5; tt >= 0; --tt) {              // small terms first
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = umma_desc(sAb + pa[tt] * FUSED_A_SPLIT + kk * 32);
            const uint64_t bd = umma_desc(sB + pb[tt] * FUSED_B_SPLIT + a * FUSED_B_ATOM + kk * 32);
            umma_bf16(tmem_d, ad, bd, idesc, acc);
            acc = 1;
          }
        }
        umma_commit(bar_mma + 8 * buf);
      }
      ++mma_cnt[buf];
    }
    // all MMAs of the layer done: the accumulators are complete, W_l and the A buffers are free
    GG_FUSED_WAIT(bar_mma, (mma_cnt[0] - 1) & 1, 0x400);
    if (FUSED_DBUF) GG_FUSED_WAIT(bar_mma + 8, (mma_cnt[1] - 1) & 1, 0x800);
    tc_fence_after();
    if (tid == 0 && l + 1 < p.n_layers) {
      mbar_expect_tx(bar_w, FUSED_B_BYTES);
      bulk_g2s(sB, p.tc_blob + p.w_off[l + 1], FUSED_B_BYTES, bar_w);
    }
    if (FUSED_DBUF && tid < 36) reinterpret_cast<uint32_t*>(pX + zero_off)[tid] = 0u;   // the zero row was part of an A buffer
    const float jkw = p.jkw[l];
    GG_FUSED_LAP(t_prod);
    // ---- aggregation + epilogue, a quarter of the channels at a time
    // per-lane partial sums of the new h (and its square) over the channels this lane owns, one
    // pair per node round; reduced over the 8 lanes of a node once per layer
#if FUSED_REGSTATS
    float st1[FUSED_MAX_NODES / 128], st2[FUSED_MAX_NODES / 128];
#pragma unroll
    for (int j = 0; j < FUSED_MAX_NODES / 128; ++j) st1[j] = st2[j] = 0.0f;
#endif
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      if (warp < 4 * n_tiles) {
        const int tile = warp >> 2, lq = warp & 3;
        const int row = 128 * tile + 32 * lq + lane;
        uint32_t rr[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * lq) << 16) + (uint32_t)(128 * tile + 32 * q), rr);
        tmem_ld_wait();
        if (row < N) {
          const float ds = sDinv[row];                   // the source's d^-1/2, applied once per row
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(pX + (uint32_t)row * XQ_PITCH + 16 * c) =
                make_float4(ds * __uint_as_float(rr[4 * c]), ds * __uint_as_float(rr[4 * c + 1]),
                            ds * __uint_as_float(rr[4 * c + 2]), ds * __uint_as_float(rr[4 * c + 3]));
        }
      }
      tc_fence_before();
      __syncthreads();
      GG_FUSED_LAP(t_drain);
      const int col = 32 * q + 4 * sub;
      const float4 b4 = *reinterpret_cast<const float4*>(sBias + col);
      const unsigned char* xs = pX + 16 * sub;
#if FUSED_REGSTATS
      // software pipeline over the node rounds: the row slices of gate / h / z of round j + 1 are
      // requested before the epilogue of round j (per-SM L2 latency, not bandwidth, bounds this phase)
      constexpr int NR = FUSED_MAX_NODES / 128;
      float4 gt_n = make_float4(0.f, 0.f, 0.f, 0.f), hv_n = gt_n, zv_n = gt_n;
      if (grp < N) {
        const size_t go0 = (size_t)(v0 + grp) * FUSED_D + col;
        gt_n = __ldg(reinterpret_cast<const float4*>(p.gate + go0));
        hv_n = __ldcg(reinterpret_cast<const float4*>(p.h + go0));
        zv_n = __ldcg(reinterpret_cast<const float4*>(p.z + go0));
      }
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int v = 128 * j + grp;
        if (v < N) {
          const size_t go = (size_t)(v0 + v) * FUSED_D + col;
          float4 acc = *reinterpret_cast<const float4*>(xs + (uint32_t)v * XQ_PITCH);     // self loop: d_v^-1/2 x'_v
          const int e0 = sRow[v], e1 = sRow[v + 1];
          for (int e = e0; e < e1; e += 4) {
            const uint2 o = *reinterpret_cast<const uint2*>(sSrc + e);              // four 16-bit row offsets
            const float4 r0 = *reinterpret_cast<const float4*>(xs + (o.x & 0xffffu));
            const float4 r1 = *reinterpret_cast<const float4*>(xs + (o.x >> 16));
            const float4 r2 = *reinterpret_cast<const float4*>(xs + (o.y & 0xffffu));
            const float4 r3 = *reinterpret_cast<const float4*>(xs + (o.y >> 16));
            acc.x += (r0.x + r1.x) + (r2.x + r3.x);
            acc.y += (r0.y + r1.y) + (r2.y + r3.y);
            acc.z += (r0.z + r1.z) + (r2.z + r3.z);
            acc.w += (r0.w + r1.w) + (r2.w + r3.w);
          }
          const float4 gt = gt_n;
          float4 hv = hv_n, zv = zv_n;
          if (j + 1 < NR && v + 128 < N) {
            const size_t gn = go + (size_t)128 * FUSED_D;
            gt_n = __ldg(reinterpret_cast<const float4*>(p.gate + gn));
            hv_n = __ldcg(reinterpret_cast<const float4*>(p.h + gn));
            zv_n = __ldcg(reinterpret_cast<const float4*>(p.z + gn));
          }
          const float dv = sDinv[v];
          hv.x += gelu_erf_f(fmaf(dv, acc.x, b4.x) * gt.x); hv.y += gelu_erf_f(fmaf(dv, acc.y, b4.y) * gt.y);
          hv.z += gelu_erf_f(fmaf(dv, acc.z, b4.z) * gt.z); hv.w += gelu_erf_f(fmaf(dv, acc.w, b4.w) * gt.w);
          zv.x += jkw * hv.x; zv.y += jkw * hv.y; zv.z += jkw * hv.z; zv.w += jkw * hv.w;
          __stcg(reinterpret_cast<float4*>(p.h + go), hv);
          __stcg(reinterpret_cast<float4*>(p.z + go), zv);
          st1[j] += (hv.x + hv.y) + (hv.z + hv.w);
          st2[j] += (hv.x * hv.x + hv.y * hv.y) + (hv.z * hv.z + hv.w * hv.w);
        }
      }
#else
      // warp-uniform trip count (the 8-lane groups of a warp shuffle with the full mask): groups past
      // the last node run the round with no edges and no memory traffic
      for (int vb = 0; vb < N; vb += FUSED_THREADS / 8) {
        const int v = vb + grp;
        const bool act = v < N;
        const int vc = act ? v : 0;
        const size_t go = (size_t)(v0 + vc) * FUSED_D + col;
        // the row slices of gate / h / z are requested before the gather and consumed after it
        float4 gt = make_float4(0.f, 0.f, 0.f, 0.f), hv = gt, zv = gt;
        if (act) {
          gt = __ldg(reinterpret_cast<const float4*>(p.gate + go));
          hv = __ldcg(reinterpret_cast<const float4*>(p.h + go));
          zv = __ldcg(reinterpret_cast<const float4*>(p.z + go));
        }
        float4 acc = *reinterpret_cast<const float4*>(xs + (uint32_t)vc * XQ_PITCH);     // self loop: d_v^-1/2 x'_v
        const int e0 = act ? sRow[vc] : 0, e1 = act ? sRow[vc + 1] : 0;
        for (int e = e0; e < e1; e += 4) {
          const uint2 o = *reinterpret_cast<const uint2*>(sSrc + e);              // four 16-bit row offsets
          const float4 r0 = *reinterpret_cast<const float4*>(xs + (o.x & 0xffffu));
          const float4 r1 = *reinterpret_cast<const float4*>(xs + (o.x >> 16));
          const float4 r2 = *reinterpret_cast<const float4*>(xs + (o.y & 0xffffu));
          const float4 r3 = *reinterpret_cast<const float4*>(xs + (o.y >> 16));
          acc.x += (r0.x + r1.x) + (r2.x + r3.x);
          acc.y += (r0.y + r1.y) + (r2.y + r3.y);
          acc.z += (r0.z + r1.z) + (r2.z + r3.z);
          acc.w += (r0.w + r1.w) + (r2.w + r3.w);
        }
        if (act) {
          const float dv = sDinv[vc];
          hv.x += gelu_erf_f(fmaf(dv, acc.x, b4.x) * gt.x); hv.y += gelu_erf_f(fmaf(dv, acc.y, b4.y) * gt.y);
          hv.z += gelu_erf_f(fmaf(dv, acc.z, b4.z) * gt.z); hv.w += gelu_erf_f(fmaf(dv, acc.w, b4.w) * gt.w);
          zv.x += jkw * hv.x; zv.y += jkw * hv.y; zv.z += jkw * hv.z; zv.w += jkw * hv.w;
          __stcg(reinterpret_cast<float4*>(p.h + go), hv);
          __stcg(reinterpret_cast<float4*>(p.z + go), zv);
        } else {
          hv = make_float4(0.f, 0.f, 0.f, 0.f);
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
#endif
      __syncthreads();
      GG_FUSED_LAP(t_gather);
    }
#if FUSED_REGSTATS
    // ---- row statistics of the new h -> next layer's LayerNorm (all lanes shuffle, lane 0 of a node writes)
#pragma unroll
    for (int j = 0; j < FUSED_MAX_NODES / 128; ++j) {
      float s1 = st1[j], s2 = st2[j];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const int v = 128 * j + grp;
      if (sub == 0 && v < N) {
        const double mean = (double)s1 * (1.0 / FUSED_D);
        const double var = fmax((double)s2 * (1.0 / FUSED_D) - mean * mean, 0.0);
        sMean[v] = make_float2((float)mean, rsqrtf((float)var + 1e-5f));
      }
    }
#else
    // ---- row statistics of the new h -> next layer's LayerNorm
    for (int i = tid; i < N; i += FUSED_THREADS) {
      const double2 st = sStat[i];
      const double mean = st.x * (1.0 / FUSED_D);
      const double var = fmax(st.y * (1.0 / FUSED_D) - mean * mean, 0.0);
      sMean[i] = make_float2((float)mean, rsqrtf((float)var + 1e-5f));
      sStat[i] = make_double2(0.0, 0.0);
    }
#endif
    __syncthreads();
    GG_FUSED_LAP(t_misc);
  }
  if (p.dbg && tid == 0 && blockIdx.x == 0) {
    p.dbg[32] = (int)(t_prod >> 4); p.dbg[33] = (int)(t_drain >> 4); p.dbg[34] = (int)(t_gather >> 4);
    p.dbg[35] = (int)(t_misc >> 4);
    __threadfence_system();
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
  return 1024 + FUSED_B_BYTES + FUSED_A_BYTES + fused_xq_bytes((int)NC8) + NC8 * (16 + 8 + 4) + (NC8 + 8) * 4 + 3 * 128 * 4 + 64 +
         160 + ((((size_t)edge_cap + 3 * NC8) * 2 + 15) & ~size_t(15));
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
