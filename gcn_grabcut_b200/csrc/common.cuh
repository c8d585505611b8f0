// Shared host/device plumbing for the gcn_grabcut_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gcn_grabcut_b200.h"

#define GG_HD __host__ __device__ __forceinline__
#define GG_D __device__ __forceinline__

namespace gg {

// ----------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

#define GG_CUDA_OK(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      gg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return GG_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define GG_REQUIRE(cond, ...)                  \
  do {                                         \
    if (!(cond)) {                             \
      gg::set_error(__VA_ARGS__);              \
      return GG_ERR_INVALID;                   \
    }                                          \
  } while (0)

#define GG_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != GG_OK) return _s; \
  } while (0)

// device status bits (gg_check_device_status)
enum : int { ST_LABEL_RANGE = 1, ST_PAIR_TABLE = 2, ST_EDGE_CAP = 4, ST_KNN_CAP = 8 };

// ----------------------------------------------------------------------------- arena
// A growable device buffer carved by a bump pointer.  Growth (cudaMalloc) only happens
// when a call needs more than any previous call; steady state performs no allocation.
struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  int reserve(size_t bytes);
  void reset() { off = 0; }
  bool overflowed = false;   // a take() went past the reservation (workspace-size bug): sticky
  template <typename T>
  T* take(size_t n) {
    size_t b = (n * sizeof(T) + 255) & ~size_t(255);
    if (off + b > cap) {       // never hand out memory past the reservation; the entry point fails
      overflowed = true;
      return reinterpret_cast<T*>(base);
    }
    T* p = reinterpret_cast<T*>(base + off);
    off += b;
    return p;
  }
  static size_t padded(size_t n, size_t elem) { return (n * elem + 255) & ~size_t(255); }
  void release();
};

// ----------------------------------------------------------------------------- network
struct NetWeights {
  bool loaded = false;
  int D = 0, n_layers = 0, q = 0, c = 0;
  float* blob = nullptr;  // all parameters, device, fp32
  size_t blob_floats = 0;
  // offsets (in floats) into blob
  size_t jk_w;  // softmax(jk_logits) [n+2]
  size_t bn_scale, bn_shift;  // folded eval BatchNorm: x*scale+shift [19]
  size_t w_in, b_in, ln_in_g, ln_in_b;
  size_t pb0_w, pb0_b, pb2_w, pb2_b;
  size_t ee0_w, ee0_b, ee2_w, ee2_b, eg_ln_g, eg_ln_b, eg_w, eg_b;
  std::vector<size_t> gcn_w, gcn_b, norm_g, norm_b;
  size_t sage_wl, sage_bl, sage_wr, sage_ln_g, sage_ln_b;
  size_t attn_w, attn_b, cmp_w, cmp_b, exp_w, exp_b;
  size_t fuse_ln_g, fuse_ln_b, fuse_w, fuse_b, head_w, head_b;
  std::vector<float> h_jk;   // host copy of softmax(jk_logits)
  // tcgen05 operand images (bf16 split, canonical K-major SWIZZLE_128B), see gemm_tc.cuh
  void* tc_blob = nullptr;
  size_t tc_bytes = 0;
  std::vector<size_t> tc_off;  // byte offset of each GEMM's operand image (by GEMM id), -1 = none
};

// GCNTrimapNet / GATTrimapNet parameters (variants.cu): one blob, tensors in the order of gg_variant_weights
struct VariantWeights {
  bool loaded = false;
  int kind = 0, D = 0, n_layers = 0, heads = 0;
  float* blob = nullptr;
  size_t blob_floats = 0;
  std::vector<size_t> off;
  // tcgen05 operand images of the per-layer EdgeInjectionLayer.proj.2 weights (hidden 64 / 128)
  unsigned char* tc_blob = nullptr;
  size_t tc_stride = 0;      // bytes between consecutive layers' images
  int* d_rows = nullptr;     // device int: row count of the edge transform (the tensor-core kernel reads M on the device)
};

}  // namespace gg

struct gg_context {
  int device = 0;
  int sm_count = 148;
  int cc_major = 0, cc_minor = 0;
  gg::Arena arena;        // device-pointer entry points (caller's stream)
  gg::Arena host_arena;   // gg_trimap_path_host chunk workspaces
  gg::NetWeights net;
  gg::VariantWeights variant;
  int* d_status = nullptr;   // [16] words: 0 last call, 1 sticky (host path), 8.. per sub-batch
  int* status_word = nullptr;  // word the kernels being enqueued right now report into
  double* d_lin = nullptr;   // sRGB linearisation table (256 doubles), built in gg_create
  double* d_coord = nullptr;  // k_coord_tables of the last (H, W) seen
  int coord_H = 0, coord_W = 0;
  int64_t launches = 0;
  uint64_t attr_done = 0;   // one bit per kernel whose max-dynamic-smem attribute is already set
  int gemm_impl = 1;      // 0 = SIMT fp32 (validation), 1 = tcgen05 bf16x3
  int rs_direct = 1;      // k_region_stats hand-over: 1 = RED.F64 per run, 0 = per-warp smem table
  int knn_legacy = 0;     // 1 = per-lane top-k kernel k_knn<K> for every size (it otherwise serves N > 2048 only)
  int gcn_fused = 1;      // 1 = per-graph fused residual GCN blocks where they apply (gcn_fused.cu), 0 = layer-wise
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  cudaStream_t s_sub[4] = {nullptr, nullptr, nullptr, nullptr};   // concurrent sub-batch streams
  int n_sub = 2;
  bool stagger = false;
  std::vector<cudaEvent_t> ev;
  // gg_trimap_path_host_submit / _wait: chunk slots rotate across calls
  static constexpr int MAX_TICKETS = 8;
  cudaEvent_t ticket_ev[MAX_TICKETS] = {};
  bool ticket_open[MAX_TICKETS] = {};
  int* h_ticket_status = nullptr;   // pinned [MAX_TICKETS]: sticky device status as of each ticket's completion
  int tickets_open = 0;
  long long chunk_seq = 0;
  size_t slot_bytes = 0;
  bool slot_used[3] = {false, false, false};
  // per-kernel CUDA-event timing (gg_profile_enable / gg_profile_report)
  bool prof_on = false;
  struct ProfRec { const char* name; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_pool;
  size_t prof_pool_used = 0;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per context and kernel (bit id)
#define GG_SMEM_ATTR_ONCE(ctx, bit, kernel, bytes)                                              \
  do {                                                                                          \
    if (!((ctx)->attr_done >> (bit) & 1ull)) {                                                  \
      GG_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                      (int)(bytes)));                                           \
      (ctx)->attr_done |= 1ull << (bit);                                                        \
    }                                                                                           \
  } while (0)

namespace gg {
void prof_begin(gg_context* ctx, const char* name, cudaStream_t st);
void prof_end(gg_context* ctx, cudaStream_t st);
// Start of a status epoch for the kernels about to be enqueued.  The per-sub-batch / per-chunk
// words are cleared here and OR-ed into a sticky word when their work is done; the main word
// d_status[0] is sticky itself -- only gg_check_device_status reads and clears it, so that a bit
// set by one call (e.g. gg_coo_to_csr) survives the calls that follow it (gg_resgcn_forward).
static inline int status_epoch(gg_context* ctx, cudaStream_t st) {
  if (ctx->status_word != ctx->d_status) {
    cudaError_t e = cudaMemsetAsync(ctx->status_word, 0, sizeof(int), st);
    if (e != cudaSuccess) { set_error("status_epoch: %s", cudaGetErrorString(e)); return GG_ERR_CUDA; }
  }
  return GG_OK;
}
}

#define GG_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)        \
  do {                                                                \
    if ((ctx)->prof_on) gg::prof_begin((ctx), #kernel, (stream));     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);       \
    if ((ctx)->prof_on) gg::prof_end((ctx), (stream));                \
    (ctx)->launches++;                                                \
    GG_CUDA_OK(cudaGetLastError());                                   \
  } while (0)

namespace gg {

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------- device utils
GG_D float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
GG_D double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
GG_D float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
GG_D float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
GG_D int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions through shared memory (blockDim.x multiple of 32, <= 1024).
// `scratch` needs 32 elements; every thread gets the result.
template <typename T, typename Op>
GG_D T block_reduce(T v, T ident, Op op, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : ident;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}

struct OpAdd { template <typename T> GG_D T operator()(T a, T b) const { return a + b; } };
struct OpMaxF { GG_D float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMinF { GG_D float operator()(float a, float b) const { return fminf(a, b); } };
struct OpMaxI { GG_D int operator()(int a, int b) const { return max(a, b); } };

// Block-wide exclusive scan of one int per thread; returns the exclusive prefix and the
// block total through `total`.  `scratch` needs 33 ints.
GG_D int block_exclusive_scan(int v, int* scratch, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < nw) ? scratch[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    scratch[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) scratch[32] = winc;
  }
  __syncthreads();
  *total = scratch[32];
  return scratch[wid] + inc - v;
}

}  // namespace gg
