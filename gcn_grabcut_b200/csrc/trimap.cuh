// Internal interface of trimap.cu
#pragma once
#include "common.cuh"

namespace gg {

size_t trimap_workspace_bytes(int B, int H, int W, bool need_gray);

// gray_in may be NULL (then the grey plane is derived from bgr inside `ar`).
int refine_trimap(gg_context* ctx, Arena& ar, const uint8_t* bgr, const uint8_t* gray_in,
                  const int32_t* labels, const float* probs, const int64_t* node_off, int B, int H,
                  int W, int radius, float eps, float thr_fg, float thr_bg, uint8_t* trimap,
                  float* p_bg, float* p_fg, cudaStream_t st, int node_cap_hint = 0);
// node_cap_hint: upper bound of the labels of one image when the caller knows it (0 = unknown); sizes the
// shared (p_bg, p_fg) table of the first filter stage.

int guided_filter_plane(gg_context* ctx, Arena& ar, const float* guide, const float* src, int H,
                        int W, int radius, float eps, float* out, cudaStream_t st);

int project_trimap(gg_context* ctx, const int32_t* labels, const float* probs,
                   const int64_t* node_off, int B, int H, int W, float thr_fg, float thr_bg,
                   uint8_t* trimap, cudaStream_t st);

// _seed_from_prior (pipeline.py:149-186) in place on device trimaps; x = node_input rows [SN,19].
size_t seed_workspace_bytes(int B, long long node_cap_total);
int seed_from_prior(gg_context* ctx, Arena& ar, uint8_t* trimap, const int32_t* labels, const float* x,
                    const int64_t* node_off, int B, int H, int W, long long node_cap_total,
                    double seed_frac, cudaStream_t st);

// GrabCut.run_with_trimap guards (grabcut.py:127-140) in place; degenerate [B] optional.
int grabcut_guards(gg_context* ctx, Arena& ar, uint8_t* trimap, int B, int H, int W, int32_t* degenerate,
                   cudaStream_t st);

// clean_mask (pipeline.py:189-227), batched: mask / out uint8 {0,1} [B,H,W] (may alias).
size_t clean_workspace_bytes(int B, int H, int W);
int clean_masks(gg_context* ctx, Arena& ar, const uint8_t* mask, uint8_t* out, int B, int H, int W,
                double min_area_ratio, int keep_largest, cudaStream_t st);

}  // namespace gg
