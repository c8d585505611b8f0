// Internal interface of graph_build.cu
#pragma once
#include "common.cuh"

namespace gg {

// per-region statistics, struct-of-arrays [B][ST_FIELDS][node_cap] (float32)
enum : int {
  ST_COUNT = 0, ST_SAFE,
  ST_MEAN_L, ST_MEAN_A, ST_MEAN_B,
  ST_STD_L, ST_STD_A, ST_STD_B,
  ST_MEAN_H, ST_MEAN_S, ST_MEAN_V,
  ST_CY, ST_CX, ST_BND, ST_AREA, ST_MGRAD, ST_MGRADN,
  ST_PCY, ST_PCX, ST_BORDER,
  ST_FIELDS
};

size_t graph_workspace_bytes(int B, int H, int W, const gg_graph_config& cfg);

// gray_out (optional): receives the device pointer of the uint8 grey plane [B,H,W] that
// the builder produced inside `ar` (re-used by the trimap stage).
int build_graphs(gg_context* ctx, Arena& ar, const uint8_t* bgr, const int32_t* labels, int B,
                 int H, int W, const gg_graph_config& cfg, const gg_graph_out& out,
                 cudaStream_t st, const uint8_t** gray_out);

int pixel_planes(gg_context* ctx, Arena& ar, const uint8_t* bgr, int B, int H, int W, float* lab,
                 float* hsv, float* gray, float* grad, cudaStream_t st);

// derive_trimap_labels / fg_ratio (dataset.py:175-205, 239-249) for a batch of label maps + masks
int region_labels(gg_context* ctx, Arena& ar, const int32_t* labels, const uint8_t* mask,
                  const int64_t* node_off, int B, int H, int W, long long node_cap_total,
                  double fg_thr, double bg_thr, float* fg_ratio, long long* y, cudaStream_t st);

// compute_auto_prior(segments, lab, ...) (graph_builder.py:357-444) from a caller-supplied float32 Lab
// plane; prior [B*node_cap,3] (dense rows), n_nodes_minus_1 [B] optional (= max label per image)
size_t auto_prior_workspace_bytes(int B, int node_cap);
int auto_prior(gg_context* ctx, Arena& ar, const int32_t* labels, const float* lab, int B, int H, int W, int node_cap,
               double centre_sigma, double contrast_sigma, float* prior, int32_t* n_nodes, cudaStream_t st);

// float32 fast paths of pixel_math.cuh vs the IEEE intrinsics; mismatches[4] (see k_selftest_math)
int selftest_math(gg_context* ctx, Arena& ar, long long* mismatches, cudaStream_t st);

}  // namespace gg
