// Internal interface of slic.cu
#pragma once
#include "common.cuh"

namespace gg {

size_t slic_workspace_bytes(int B, int H, int W, int n_segments);
// number of cluster centres skimage's regular grid places for this shape (the label count is close to it)
int slic_nominal_segments(int H, int W, int n_segments);
// labels [B,H,W] int32, contiguous 0..n_labels[b]-1 in raster order of first appearance; n_labels optional
int slic_labels(gg_context* ctx, Arena& ar, const uint8_t* bgr, int B, int H, int W, int n_segments,
                double compactness, double sigma, int max_iter, int32_t* labels, int32_t* n_labels, cudaStream_t st);

}  // namespace gg
