// Internal interface of resgcn.cu / gemm_tc.cu
#pragma once
#include "common.cuh"

namespace gg {

// offsets (in floats) of the small parameter tensors inside NetWeights::blob, by value to kernels
struct NetOffsets {
  int D, q, c;
  unsigned jk_w, bn_scale, bn_shift, w_in, b_in, ln_in_g, ln_in_b;
  unsigned pb0_w, pb0_b, pb2_w, pb2_b, ee0_w, ee0_b, ee2_b, eg_ln_g, eg_ln_b;
  unsigned attn_w, attn_b, cmp_w, cmp_b, exp_w, exp_b, head_w, head_b;
};

// identifiers of the dense per-node / per-edge transforms (select the packed tensor-core operand)
enum : int { GEMM_ENC2 = 0, GEMM_GATE = 1, GEMM_SAGE_L = 2, GEMM_SAGE_R = 3, GEMM_FUSE = 4, GEMM_GCN0 = 5 };

int load_weights(gg_context* ctx, const gg_resgcn_weights* w);
size_t resgcn_workspace_bytes(const NetWeights& nw, long long node_cap, long long edge_cap, int n_graphs);
// graph_node_cap / graph_edge_cap: upper bounds of the nodes / directed edges of ONE graph when the
// caller knows them (0 = unknown); they enable the per-graph fused kernels (gcn_fused.cu).
int resgcn_forward(gg_context* ctx, Arena& ar, const float* x, const int32_t* rowptr,
                   const int32_t* src, const int32_t* eid, const float* edge_attr,
                   const int64_t* graph_off, int n_graphs, long long node_cap, long long edge_cap,
                   float* logits, float* probs, cudaStream_t st, int graph_node_cap = 0,
                   int graph_edge_cap = 0);

// gcn_fused.cu: all residual GCN blocks in one kernel, one CTA per graph
constexpr int FUSED_MAX_LAYERS = 16;
constexpr int FUSED_MAX_NODES = 384;
bool gcn_fused_supported(const gg_context* ctx, int node_cap, int edge_cap);
int gcn_layers_fused(gg_context* ctx, cudaStream_t st, float* h, float* z, const float* gate,
                     const float2* row_stats, const float* dinv, const int32_t* rowptr, const int32_t* src,
                     const int64_t* graph_off, int n_graphs, int node_cap, int edge_cap);
int coo_to_csr(gg_context* ctx, Arena& ar, const int64_t* ei, long long E, long long N,
               int32_t* rowptr, int32_t* src, int32_t* eid, cudaStream_t st);

// C[M,N] (+)= act(A[M,K] W[N,K]^T + bias); M read on the device from *m_ptr (<= m_cap).
// act: 0 none, 1 GELU(erf), 2 sigmoid.
int gemm_simt(gg_context* ctx, cudaStream_t st, const float* A, const float* W, const float* bias,
              float* C, const int* m_ptr, long long m_cap, int N, int K, int act, int accumulate);
int gemm(gg_context* ctx, cudaStream_t st, int which, const float* A, const float* W,
         const float* bias, float* C, const int* m_ptr, long long m_cap, int N, int K, int act,
         int accumulate);

// How gemm_tc produces its A operand (fused prologues), see k_tc_gemm.
struct TcPrologue {
  int mode = 0;                      // 0 plain, 1 LayerNorm(row [* gvec[node_graph]]), 2 edge-encoder layer 1
  const float* ln_g = nullptr;
  const float* ln_b = nullptr;
  const float* gvec = nullptr;       // [n_graphs, K] optional per-graph row scale (mode 1)
  const int* node_graph = nullptr;   // [M]
  const float2* row_stats = nullptr; // [M] optional (mean, rstd) of every row, precomputed by the producer of A (mode 1, no gvec)
  const float* w0 = nullptr;         // [K,5]  (mode 2)
  const float* b0 = nullptr;         // [K]
  const int* row_index = nullptr;    // mode 2: attribute row of output row r is row_index[r] (e.g. CSR position -> edge id)
  int relu = 0;                      // mode 2: ReLU instead of GELU after the first layer (EdgeInjectionLayer, model.py:152-155)
};

// gemm_tc.cu: tcgen05 path
int gemm_tc_prepare_weights(gg_context* ctx, const std::vector<float>& blob);
bool gemm_tc_supported(const gg_context* ctx, int which, int N, int K);
// One launch over a caller-owned operand image (tc_pack_weight layout; N, K in {64, 128}); used by variants.cu.
size_t tc_image_bytes_padded(int N, int K);
void tc_pack_weight(const float* W, int ldw, int N, int K, unsigned char* img);
int gemm_tc_image(gg_context* ctx, cudaStream_t st, const unsigned char* img_dev, const float* A, const float* bias,
                  float* C, const int* m_ptr, long long m_cap, int N, int K, int lda, int ldc, int act, int accumulate,
                  const TcPrologue& pro);
int gemm_tc(gg_context* ctx, cudaStream_t st, int which, const float* A, const float* bias, float* C,
            const int* m_ptr, long long m_cap, int N, int K, int act, int accumulate,
            const TcPrologue* prologue = nullptr);

}  // namespace gg
