// Internal interface of variants.cu (GCNTrimapNet / GATTrimapNet)
#pragma once
#include "common.cuh"

namespace gg {

int variant_load_weights(gg_context* ctx, const gg_variant_weights* w);
size_t variant_workspace_bytes(const VariantWeights& vw, long long n_nodes, long long n_edges);
int variant_forward(gg_context* ctx, Arena& ar, int kind, const float* x, const int32_t* rowptr, const int32_t* src,
                    const int32_t* eid, const float* edge_attr, const int64_t* graph_off, int n_graphs, long long n_nodes,
                    long long n_edges, float* logits, float* probs, cudaStream_t st);

}  // namespace gg
