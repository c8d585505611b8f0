// tcgen05 tensor-core path for the dense per-node / per-edge transforms of ResGCNNet.
//
//   C[M,N] (+)= act(A[M,K] W[N,K]^T + bias)        M ~ 10^4..10^5 rows, N,K in {64,128}
//
// fp32 semantics on bf16 tensor cores: every fp32 operand is split EXACTLY into three bf16
// terms by truncation (x = x1 + x2 + x3, 8 mantissa bits each) and the six products with i+j <= 4
// (x1w1, x1w2, x2w1, x2w2, x1w3, x3w1) are accumulated in fp32 in TMEM -- the dropped terms
// are O(2^-24) relative, i.e. fp32-class accuracy, at 6 MMA passes.
//
// Kernel structure (one persistent CTA per SM, 768 threads, warp-specialised):
//   * the weight operand W is pre-split and pre-swizzled ON THE HOST at gg_load_weights
//     time into the exact shared-memory image (canonical K-major SWIZZLE_128B atoms); each CTA
//     pulls it in once with a single 1-D bulk TMA copy (cp.async.bulk + mbarrier tx-count);
//   * producer warps 0-15 (8 rows each), per 128-row tile: load the fp32 A rows (coalesced float4; the loads of
//     tile i+1 are in flight while MMA(i) runs), run the fused prologue, split the values and store
//     the three bf16 images with the 128B swizzle applied by hand; fence.proxy.async; then one
//     producer thread issues the 6 x K/16 tcgen05.mma (M=128, N, K=16) into one of TWO TMEM
//     accumulators and commits to that accumulator's mbarrier;
//   * epilogue warps 16-23 wait for the commit, read the accumulator back with tcgen05.ld (32 lanes
//     x 32 columns per instruction), apply bias / activation / accumulate, store fp32 rows and
//     hand the accumulator back (mbarrier) -- the epilogue of tile i overlaps the producers'
//     work on tile i+1 and MMA(i+1).
// Measured with clock64 stamps (config B, 4 tiles per CTA, ~20 000 cycles per tile): A loads + prologue
// 9-12 k cycles (every CTA bursts at once), split + store 2.5 k, the tcgen05.mma issue ~5 k (the issuing
// thread stalls on the tensor pipe's queue), epilogue 9-16 k with erf-GELU (now the 16-instruction
// evaluation).  Two restructurings were measured SLOWER (0.165 against 0.123 ms for the three plain
// transforms of config B): a 25th warp that only issues the MMAs (800 threads cap the kernel at 72
// registers -- warps are allocated in fours -- and the producers spill), and requesting the next tile's
// rows right after the split, before the MMA issue.
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (bulk TMA).
#include <cuda_bf16.h>

#include "common.cuh"
#include "resgcn.cuh"
#include "tc_ptx.cuh"

namespace gg {

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 768;       // 16 producer warps + 8 epilogue warps
constexpr int TC_PRODUCERS = 512;

// PRO: how the A operand is produced.  0: rows of A as they are.  1: LayerNorm over the row
// (eps 1e-5, affine ln_g/ln_b), optionally after scaling the row by gvec[node_graph[row]]
// (K must be 128).  2: first edge-encoder layer, A[row][k] = GELU(w0[k,:5] . attr[row,:5] + b0[k])
// computed on the fly from the 5-d edge attributes (model.py:124-126).
template <int ACT, int PRO>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_gemm(const float* __restrict__ A, const uint8_t* __restrict__ Bimg, const float* __restrict__ bias,
          float* __restrict__ C, const int* __restrict__ m_ptr, int N, int K, int lda, int ldc,
          int accumulate, const TcPrologue pro) {
  extern __shared__ unsigned char tc_smem_raw[];
  const int M = *m_ptr;
  const int n_tiles = (M + TC_BM - 1) / TC_BM;
  if ((int)blockIdx.x >= n_tiles) return;

  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = tc_smem_raw + (base - raw);
  const int n_atoms = K >> 6;
  const uint32_t a_split = (uint32_t)n_atoms * TC_BM * 128;       // bytes of one A split image
  const uint32_t b_split = (uint32_t)n_atoms * N * 128;
  const uint32_t sA = base, sB = base + 3 * a_split;
  unsigned char* pA = smem;
  const uint32_t ctrl = sB + 3 * b_split;                         // barriers + tmem pointer
  // ctrl: +0 weights landed, +8/+16 MMA done (accumulator 0/1), +24/+32 accumulator drained, +40 TMEM ptr
  const uint32_t bar_b = ctrl, bar_mma0 = ctrl + 8, bar_free0 = ctrl + 24, tmem_slot = ctrl + 40;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(smem + 3 * a_split + 3 * b_split + 40);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t acc_cols = N <= 32 ? 32 : (N <= 64 ? 64 : 128);   // columns of one accumulator
  const uint32_t tmem_cols = 2 * acc_cols;

  if (tid == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_mma0, 1);
    mbar_init(bar_mma0 + 8, 1);
    mbar_init(bar_free0, TC_THREADS / 32 - TC_PRODUCERS / 32);    // one arrival per epilogue warp
    mbar_init(bar_free0 + 8, TC_THREADS / 32 - TC_PRODUCERS / 32);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  float* s_bias = reinterpret_cast<float*>(smem + 3 * a_split + 3 * b_split + 64);   // [N], zeros if no bias
  for (int i = tid; i < N; i += TC_THREADS) s_bias[i] = bias ? __ldg(bias + i) : 0.0f;
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar_b, 3 * b_split);
    bulk_g2s(sB, Bimg, 3 * b_split, bar_b);
  }
  const uint32_t idesc = umma_idesc_bf16(TC_BM, N);
  const int k4 = K >> 2;                                           // float4 per row

  if (warp < TC_PRODUCERS / 32) {
    // =========================================================================== producers
    constexpr int RPW = TC_BM / (TC_PRODUCERS / 32);          // rows per warp
    // Work items are (row, float4 chunk): a warp covers 32/k4 rows per pass (1 for K = 128,
    // 2 for K = 64), so that all lanes are busy for both widths.
    const int rpp = 32 / k4;                                 // rows per pass
    const int sub = lane / k4, ch = lane - sub * k4;        // row within the pass, chunk within the row
    const int n_pass = RPW / rpp;
    float w0[4][5], b0[4];
    if (PRO == 2) {
      // edge attributes (5 floats per row) -> 4 hidden units per lane
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = 4 * ch + u;
        b0[u] = __ldg(pro.b0 + kk);
#pragma unroll
        for (int j = 0; j < 5; ++j) w0[u][j] = __ldg(pro.w0 + kk * 5 + j);
      }
    }
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * TC_BM;
      // ---- A tile: fp32 rows (or the fused prologue's values) in registers.  All the global
      // loads of the warp are issued before the first use (one latency, not 16), and before
      // waiting for the previous tile's MMA, which is still reading the shared A images.
      float4 v[RPW];
      if (PRO == 2) {
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < n_pass) {
            const int row = row0 + warp * RPW + i * rpp + sub;
            if (row < M) {
              float a5[5];
              const size_t arow = pro.row_index ? (size_t)__ldg(pro.row_index + row) : (size_t)row;
#pragma unroll
              for (int j = 0; j < 5; ++j) a5[j] = __ldg(A + arow * 5 + j);
              float o[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float sacc = b0[u];
#pragma unroll
                for (int j = 0; j < 5; ++j) sacc = fmaf(w0[u][j], a5[j], sacc);
                o[u] = pro.relu ? fmaxf(sacc, 0.0f) : gelu_erf_tc(sacc);
              }
              v[i] = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < n_pass) {
            const int row = row0 + warp * RPW + i * rpp + sub;
            if (row < M) v[i] = __ldg(reinterpret_cast<const float4*>(A + (size_t)row * lda) + ch);
          }
        }
        if (PRO == 1) {
          // fused LayerNorm (K == 128: one float4 per lane covers the row, rpp == 1)
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(pro.ln_g) + lane);
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(pro.ln_b) + lane);
          if (pro.row_stats != nullptr && pro.gvec == nullptr) {
            // row statistics were written by the kernel that produced A: no reductions here
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
              const int row = row0 + warp * RPW + i;
              const float2 ms = row < M ? __ldg(pro.row_stats + row) : make_float2(0.f, 0.f);
              const float4 x4 = v[i];
              v[i] = make_float4((x4.x - ms.x) * ms.y * g4.x + b4.x, (x4.y - ms.x) * ms.y * g4.y + b4.y,
                                 (x4.z - ms.x) * ms.y * g4.z + b4.z, (x4.w - ms.x) * ms.y * g4.w + b4.w);
            }
          } else {
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            const int row = row0 + warp * RPW + i;
            float4 x4 = v[i];
            if (pro.gvec != nullptr && row < M) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(pro.gvec + (size_t)pro.node_graph[row] * K) + lane);
              x4.x *= s4.x; x4.y *= s4.y; x4.z *= s4.z; x4.w *= s4.w;
            }
            float sum = (x4.x + x4.y) + (x4.z + x4.w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / (float)K;
            const float dx = x4.x - mean, dy = x4.y - mean, dz = x4.z - mean, dw = x4.w - mean;
            float sq = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            const float rstd = 1.0f / sqrtf(sq / (float)K + 1e-5f);
            v[i] = make_float4(dx * rstd * g4.x + b4.x, dy * rstd * g4.y + b4.y,
                               dz * rstd * g4.z + b4.z, dw * rstd * g4.w + b4.w);
          }
          }
        }
      }
      // ---- the shared A images are free once the previous tile's MMA has completed
      if (it > 0) mbar_wait(bar_mma0 + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
      // exact 3-way split by truncation: x = x1 + x2 + x3 with x1 = top 8 mantissa bits of x,
      // x2 = top 8 bits of the (exact) remainder, x3 = what is left (<= 8 bits): ALU-only.
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        if (i < n_pass) {
          const int r = warp * RPW + i * rpp + sub;
          const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
          uint32_t hb[3][4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t b1 = __float_as_uint(x[e]) & 0xFFFF0000u;
            const float r1 = x[e] - __uint_as_float(b1);
            const uint32_t b2 = __float_as_uint(r1) & 0xFFFF0000u;
            const float r2 = r1 - __uint_as_float(b2);
            hb[0][e] = b1; hb[1][e] = b2; hb[2][e] = __float_as_uint(r2);
          }
          const uint32_t off = sw128_offset(r, 4 * ch, TC_BM);
#pragma unroll
          for (int sp = 0; sp < 3; ++sp) {
            const uint32_t lo = __byte_perm(hb[sp][0], hb[sp][1], 0x7632);   // {hi16(e0), hi16(e1)}
            const uint32_t hi = __byte_perm(hb[sp][2], hb[sp][3], 0x7632);
            *reinterpret_cast<uint2*>(pA + sp * a_split + off) = make_uint2(lo, hi);
          }
        }
      }
      fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor core
      producers_sync();

      // ---- MMA: one thread issues 6 x (K/16) tcgen05.mma into accumulator it & 1
      if (tid == 0) {
        const uint32_t s_ = it & 1, u_ = it >> 1;
        if (it == 0) mbar_wait(bar_b, 0);
        if (u_ > 0) mbar_wait(bar_free0 + 8 * s_, (u_ - 1) & 1);     // the epilogue has drained it
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + s_ * acc_cols;
        const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
        uint32_t acc = 0;
#pragma unroll
        for (int t = 5; t >= 0; --t) {     // small terms first
          for (int a = 0; a < n_atoms; ++a) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = umma_desc(sA + pa[t] * a_split + a * (TC_BM * 128) + kk * 32);
              const uint64_t bd = umma_desc(sB + pb[t] * b_split + a * (N * 128) + kk * 32);
              umma_bf16(tmem_d, ad, bd, idesc, acc);
              acc = 1;
            }
          }
        }
        umma_commit(bar_mma0 + 8 * s_);
      }
    }
  } else {
    // =========================================================================== epilogue
    // TMEM -> registers -> bias / activation / accumulate -> global.  Warp w reads TMEM lanes
    // 32 (w % 4) ..; the two warps of a lane quarter take one half of the columns each.
    const int ew = warp - TC_PRODUCERS / 32;
    const int q = ew & 3, hh = ew >> 2;
    const int ncol_w = N >> 1;                        // columns per warp
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t s_ = it & 1, u_ = it >> 1;
      const int row = tile * TC_BM + 32 * q + lane;
      mbar_wait(bar_mma0 + 8 * s_, u_ & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < ncol_w; c0 += 32) {
        const int col0 = hh * ncol_w + c0;
        uint32_t rr[32];
        tmem_ld32(tmem_base + s_ * acc_cols + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, rr);
        tmem_ld_wait();
        if (c0 + 32 >= ncol_w) {          // last read of this accumulator: hand it back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_free0 + 8 * s_);
        }
        if (row < M) {
          float* crow = C + (size_t)row * ldc + col0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 o;
            float* op = reinterpret_cast<float*>(&o);
            float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
            if (accumulate) prev = *reinterpret_cast<const float4*>(crow + 4 * i);
            const float* pp = reinterpret_cast<const float*>(&prev);
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + col0 + 4 * i);
            const float* bp = reinterpret_cast<const float*>(&b4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float vv = __uint_as_float(rr[4 * i + u]) + bp[u] + pp[u];
              if (ACT == 1) vv = gelu_fast_tc(vv);
              if (ACT == 2) vv = sigmoid_fast_tc(vv);
              op[u] = vv;
            }
            *reinterpret_cast<float4*>(crow + 4 * i) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ----------------------------------------------------------------------------- host side
static size_t tc_image_bytes(int N, int K) { return (size_t)3 * (K / 64) * N * 128; }

// image of the N x K block of W (row stride ldw) that starts at W
static void pack_weight(const float* W, int ldw, int N, int K, uint8_t* img) {
  const size_t split = (size_t)(K / 64) * N * 128;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float r = W[(size_t)n * ldw + k];
      const uint32_t off = sw128_offset(n, k, N);
      for (int s = 0; s < 3; ++s) {
        uint32_t u;
        memcpy(&u, &r, 4);
        u &= 0xFFFF0000u;                       // truncation split: exact, 3 x 8 mantissa bits
        const unsigned short bits = (unsigned short)(u >> 16);
        memcpy(img + s * split + off, &bits, 2);
        float part;
        memcpy(&part, &u, 4);
        r -= part;
      }
    }
}

// Widths of 256 are handled as 2 x 2 blocks of 128 (one launch per block, the second K block
// accumulates into C): the kernel itself sees N, K in {64, 128} and row strides lda / ldc.
static bool tc_shape_ok(int N, int K) {
  return (K == 64 || K == 128 || K == 256) && (N == 64 || N == 128 || N == 256);
}
static int tc_chunk(int dim) { return dim > 128 ? 128 : dim; }

int gemm_tc_prepare_weights(gg_context* ctx, const std::vector<float>& blob) {
  NetWeights& nw = ctx->net;
  if (nw.tc_blob) { cudaFree(nw.tc_blob); nw.tc_blob = nullptr; }
  nw.tc_bytes = 0;
  nw.tc_off.assign(GEMM_GCN0 + nw.n_layers, (size_t)-1);
  const int D = nw.D, c = nw.c;
  struct Item { int which; size_t w_off; int N, K; };
  std::vector<Item> items = {{GEMM_ENC2, nw.ee2_w, c, c}, {GEMM_GATE, nw.eg_w, D, c},
                             {GEMM_SAGE_L, nw.sage_wl, D, D}, {GEMM_SAGE_R, nw.sage_wr, D, D},
                             {GEMM_FUSE, nw.fuse_w, D, D}};
  for (int l = 0; l < nw.n_layers; ++l) items.push_back({GEMM_GCN0 + l, nw.gcn_w[l], D, D});
  size_t total = 0;
  for (auto& it : items)
    if (tc_shape_ok(it.N, it.K)) {
      const int Nc = tc_chunk(it.N), Kc = tc_chunk(it.K);
      nw.tc_off[it.which] = total;
      total += (size_t)(it.N / Nc) * (it.K / Kc) * ((tc_image_bytes(Nc, Kc) + 1023) & ~size_t(1023));
    }
  if (total == 0) return GG_OK;
  std::vector<uint8_t> host(total, 0);
  for (auto& it : items)
    if (tc_shape_ok(it.N, it.K)) {
      const int Nc = tc_chunk(it.N), Kc = tc_chunk(it.K), nK = it.K / Kc;
      const size_t one = (tc_image_bytes(Nc, Kc) + 1023) & ~size_t(1023);
      for (int nb = 0; nb < it.N / Nc; ++nb)
        for (int kb = 0; kb < nK; ++kb)
          pack_weight(blob.data() + it.w_off + (size_t)nb * Nc * it.K + (size_t)kb * Kc, it.K, Nc, Kc,
                      host.data() + nw.tc_off[it.which] + (size_t)(nb * nK + kb) * one);
    }
  GG_CUDA_OK(cudaMalloc(&nw.tc_blob, total));
  GG_CUDA_OK(cudaMemcpy(nw.tc_blob, host.data(), total, cudaMemcpyHostToDevice));
  nw.tc_bytes = total;
  return GG_OK;
}

bool gemm_tc_supported(const gg_context* ctx, int which, int N, int K) {
  const NetWeights& nw = ctx->net;
  return nw.tc_blob != nullptr && which >= 0 && which < (int)nw.tc_off.size() &&
         nw.tc_off[which] != (size_t)-1 && tc_shape_ok(N, K);
}

static int gemm_tc_block(gg_context* ctx, cudaStream_t st, const uint8_t* img, const float* A, const float* bias,
                         float* C, const int* m_ptr, long long m_cap, int N, int K, int lda, int ldc, int act,
                         int accumulate, const TcPrologue& pro) {
  const size_t smem = (size_t)3 * (K / 64) * TC_BM * 128 + tc_image_bytes(N, K) + 64 + 512 + 1024;
  const size_t smem_max = (size_t)3 * 2 * TC_BM * 128 + tc_image_bytes(128, 128) + 64 + 512 + 1024;   // K = N = 128
  const int tiles_cap = ceil_div(m_cap, TC_BM);
  const int grid = tiles_cap < ctx->sm_count ? tiles_cap : ctx->sm_count;
#define GG_TC_CASE(ACT_, PRO_, BIT_)                                                               \
  if (act == ACT_ && pro.mode == PRO_) {                                                          \
    GG_SMEM_ATTR_ONCE(ctx, BIT_, (k_tc_gemm<ACT_, PRO_>), smem_max);                              \
    GG_LAUNCH(ctx, (k_tc_gemm<ACT_, PRO_>), grid, TC_THREADS, smem, st, A, img, bias, C, m_ptr, N, \
              K, lda, ldc, accumulate, pro);                                                      \
    return GG_OK;                                                                                 \
  }
  GG_TC_CASE(0, 0, 16) GG_TC_CASE(1, 0, 17) GG_TC_CASE(2, 0, 18)
  GG_TC_CASE(0, 1, 19) GG_TC_CASE(1, 1, 20) GG_TC_CASE(0, 2, 21) GG_TC_CASE(2, 2, 22)
#undef GG_TC_CASE
  set_error("gemm_tc: unsupported act/prologue combination %d/%d", act, pro.mode);
  return GG_ERR_INVALID;
}

size_t tc_image_bytes_padded(int N, int K) { return (tc_image_bytes(N, K) + 1023) & ~size_t(1023); }
void tc_pack_weight(const float* W, int ldw, int N, int K, unsigned char* img) { pack_weight(W, ldw, N, K, img); }
int gemm_tc_image(gg_context* ctx, cudaStream_t st, const unsigned char* img_dev, const float* A, const float* bias,
                  float* C, const int* m_ptr, long long m_cap, int N, int K, int lda, int ldc, int act, int accumulate,
                  const TcPrologue& pro) {
  GG_REQUIRE((N == 64 || N == 128) && (K == 64 || K == 128), "gemm_tc_image: N, K must be 64 or 128");
  return gemm_tc_block(ctx, st, img_dev, A, bias, C, m_ptr, m_cap, N, K, lda, ldc, act, accumulate, pro);
}

int gemm_tc(gg_context* ctx, cudaStream_t st, int which, const float* A, const float* bias, float* C,
            const int* m_ptr, long long m_cap, int N, int K, int act, int accumulate,
            const TcPrologue* prologue) {
  const NetWeights& nw = ctx->net;
  const uint8_t* img0 = reinterpret_cast<const uint8_t*>(nw.tc_blob) + nw.tc_off[which];
  TcPrologue pro{};
  if (prologue) pro = *prologue;
  const int Nc = tc_chunk(N), Kc = tc_chunk(K), nN = N / Nc, nK = K / Kc;
  GG_REQUIRE(pro.mode == 0 || (pro.mode == 1 && K == 128) || (pro.mode == 2 && K == 64),
             "gemm_tc: unsupported prologue for K=%d", K);
  const size_t one = (tc_image_bytes(Nc, Kc) + 1023) & ~size_t(1023);
  const int lda = pro.mode == 2 ? 5 : K;
  for (int nb = 0; nb < nN; ++nb)
    for (int kb = 0; kb < nK; ++kb) {
      const bool last = kb == nK - 1;
      // bias and activation belong to the finished sum: the last K block applies them
      GG_TRY(gemm_tc_block(ctx, st, img0 + (size_t)(nb * nK + kb) * one, A + (size_t)kb * Kc,
                           last && bias ? bias + (size_t)nb * Nc : nullptr, C + (size_t)nb * Nc, m_ptr, m_cap, Nc,
                           Kc, lda, N, last ? act : 0, (kb > 0 || accumulate) ? 1 : 0, pro));
    }
  return GG_OK;
}

}  // namespace gg
