/*
 * gcn_grabcut_b200.h -- C ABI of the B200-native trimap path of GCN-GrabCut
 *
 *     label map + BGR image --> attributed region graph --> ResGCNNet posterior --> trimap
 *
 * The reference (HanielUlises/GCN-GrabCut v0.3.0) is pure Python and has no FFI of its
 * own; its boundary for this path is the Python API consumed by
 * src/gcn_grabcut/pipeline.py:298-321 and inference.py:75-98.  Each entry point below
 * names the reference interface it replaces.  The Python mirror of that interface
 * (gcn_grabcut_b200/{graph_builder,model,pipeline}.py) binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, a negative gg_status on
 *     failure; gg_last_error() returns a thread-local message for the last failure.
 *   - "dev" pointers are device pointers owned by the caller (e.g. torch data_ptr());
 *     "host" pointers are host pointers (pinned memory recommended).
 *   - all device work is enqueued on the caller's stream (a cudaStream_t passed as
 *     void*; NULL = legacy default stream); device-pointer entry points never
 *     synchronise the host.  Sizes that are only known on the device (node / edge
 *     counts) are returned in device arrays; the caller reads them back when needed.
 *   - images of one call share H and W; H*W < 2^24, H,W >= 2; labels are int32 in
 *     [0, node_cap) and need not be dense (an absent label is a region of area 0,
 *     exactly as in the reference: graph_builder.py:158, :194-195).
 *   - there is no CPU fallback: every entry point fails with GG_ERR_CUDA when no
 *     sm_100 device is usable.
 *   - threading contract: a handle is SINGLE-THREADED and its device-pointer calls must be
 *     STREAM-ORDERED -- all of them share one workspace arena, one status word and a few
 *     events, none of which is locked.  Consecutive calls on one handle must be enqueued on the
 *     same stream (or on streams the caller has ordered with events); use one handle per host
 *     thread / per concurrently used stream.  The whole-path entry points fork onto internal
 *     streams and join back into the caller's stream themselves.
 */
#ifndef GCN_GRABCUT_B200_H
#define GCN_GRABCUT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_ABI_VERSION 4

#define GG_N_IMAGE_FEATS 16 /* graph_builder.py:73 */
#define GG_N_PRIOR_FEATS 3  /* graph_builder.py:74 */
#define GG_N_NODE_FEATS 19  /* graph_builder.py:76 */
#define GG_N_EDGE_FEATS 5   /* graph_builder.py:77 */
#define GG_N_CLASSES 3      /* model.py:62-64  [BG, UNK, FG] */

typedef enum gg_status {
  GG_OK = 0,
  GG_ERR_INVALID = -1,   /* bad argument */
  GG_ERR_CUDA = -2,      /* CUDA runtime error / no usable device */
  GG_ERR_CAPACITY = -3,  /* a label >= node_cap, or an edge / pair table overflowed */
  GG_ERR_STATE = -4      /* e.g. forward before gg_load_weights */
} gg_status;

typedef struct gg_context* gg_handle;

/* ------------------------------------------------------------------ lifetime / errors */
int gg_abi_version(void);
const char* gg_last_error(void);
int gg_create(gg_handle* out, int device);
void gg_destroy(gg_handle h);

/* Runtime options.  "gemm_impl": 1 = tcgen05 tensor-core transforms (default),
 * 0 = SIMT fp32 transforms (validation of the tensor-core path; same device, same API).
 * "n_sub": number of concurrent sub-batches (internal streams) the whole-path entry points
 * cut a batch into, 1..4 (default 2; env GG_SUBBATCH).
 * "gcn_fused": 1 = run the residual GCN blocks as one per-graph kernel (x' on chip) where it
 * applies (hidden 128, graphs of <= 384 regions, tcgen05 transforms) and the batch has at least 24
 * graphs (default), 2 = also for smaller batches, 0 = layer-wise kernels (validation of the fused
 * kernel; env GG_GCN_UNFUSED).
 * "knn_legacy": 1 = non-local neighbours by the per-lane top-k kernel for every graph size (it otherwise
 * serves graphs of more than 2048 regions only; validation, env GG_KNN_LEGACY), 0 = selection kernels
 * (default). */
int gg_set_option(gg_handle h, const char* key, int value);

/* Device-side status word of the device-pointer entry points: bit0 label / edge index out of
 * range (>= node_cap, >= n_nodes), bit1 adjacency table overflow, bit2 pair / edge capacity
 * overflow.  The word is STICKY: every kernel ORs into it and only this call reads and clears
 * it, so one check after a run of calls (gg_coo_to_csr + gg_resgcn_forward, or many
 * gg_trimap_path_device steps) reports an overflow of any of them.  The device-pointer entry
 * points never synchronise, so they cannot report these conditions themselves: a batch whose
 * status is non-zero has clamped labels / dropped edges and its outputs are NOT valid.
 * Returns GG_ERR_CAPACITY when a bit is set.  Synchronises the given stream. */
int gg_check_device_status(gg_handle h, void* stream, int* status_bits);

/* ------------------------------------------------------------------ graph construction
 * Replaces GraphBuilder(image, cfg).build() minus SLIC, i.e. graph_builder.py:142-154
 * (pixel planes), :190-226 (_region_statistics), :228-255 (_assemble_node_features),
 * :257-350 (_compute_edges, _pair_features, _nonlocal_pairs) and :357-454
 * (compute_auto_prior), for a batch of B images at once.
 *
 * SuperpixelGraphConfig fields honoured (graph_builder.py:64-71): connectivity (4|8),
 * n_nonlocal (0..32).  n_segments / compactness / sigma / use_lab belong to SLIC, the
 * input producer.
 */
typedef struct gg_graph_config {
  int32_t connectivity; /* 4 or 8 */
  int32_t n_nonlocal;   /* k nearest colour neighbours per node, 0 = off */
  int32_t node_cap;     /* capacity per image: every label must be < node_cap */
  int32_t pair_cap;     /* capacity per image for undirected pairs (adjacency + non-local);
                           0 = choose (8 * node_cap) */
} gg_graph_config;

/* Batched ragged output, all device pointers.  Image b owns nodes
 * [node_off[b], node_off[b+1]) and directed edges [edge_off[b], edge_off[b+1]).
 * Rows are laid out exactly as the reference's SuperpixelGraph fields
 * (graph_builder.py:80-91), so slicing by the offsets yields the per-image arrays.
 * Capacities: node arrays B*node_cap rows, edge arrays 2*B*pair_cap rows.
 * Optional pointers may be NULL. */
typedef struct gg_graph_out {
  int32_t* n_nodes;     /* [B]      n_nodes = max label + 1      (graph_builder.py:158) */
  int32_t* n_edges;     /* [B]      directed edge count E        (graph_builder.py:171) */
  int64_t* node_off;    /* [B+1]    exclusive prefix of n_nodes */
  int64_t* edge_off;    /* [B+1]    exclusive prefix of n_edges */
  float* x;             /* [SN,19]  node_input(): 16 image features | 3 prior (…:93-98) */
  int64_t* edge_index;  /* [2,EC]   row 0 = src, row 1 = dst, image-LOCAL ids, reference COO order
                                    [adj lo->hi | nl lo->hi | adj hi->lo | nl hi->lo] (…:299-306);
                                    EC = edge_index_stride = 2*B*pair_cap */
  float* edge_attr;     /* [SE,5]   (…:309-322) */
  float* centroids;     /* [SN,2]   node_centroids [y,x] (optional) */
  float* areas;         /* [SN]     node_areas           (optional) */
  /* destination-sorted CSR of the same directed edges, GLOBAL node / edge ids, consumed
   * by gg_resgcn_forward: for node v, entries [csr_rowptr[v], csr_rowptr[v+1]). */
  int32_t* csr_rowptr;  /* [SN+1] (entries beyond the last node repeat the total) */
  int32_t* csr_src;     /* [SE]   global source node id */
  int32_t* csr_eid;     /* [SE]   global edge id (row of edge_attr) */
  int32_t* n_adj_pairs; /* [B]    undirected adjacency pairs (optional) */
  int32_t* n_nl_pairs;  /* [B]    undirected non-local pairs (optional) */
  int32_t* shared_cnt;  /* [B*pair_cap] shared-boundary pixel counts s_ij of the adjacency
                                    pairs of image b at [b*pair_cap, …) in sorted order (optional) */
} gg_graph_out;

int gg_build_graphs(gg_handle h, const uint8_t* bgr_dev /*[B,H,W,3]*/,
                    const int32_t* labels_dev /*[B,H,W]*/, int B, int H, int W,
                    const gg_graph_config* cfg, const gg_graph_out* out, void* stream);

/* GraphBuilder._compute_superpixels (graph_builder.py:177-188): skimage.segmentation.slic(lab,
 * n_segments, compactness, sigma, start_label=0) restated on the device for a batch of BGR uint8
 * images -- CIELAB, skimage's global min-max rescale, Gaussian pre-smoothing, its second rgb2lab of
 * 3-channel inputs, k-means on the regular grid (max_iter rounds, 0 = 10), connectivity enforcement
 * with min_size = half a nominal superpixel.  labels_dev [B,H,W] int32: contiguous 0..N-1 per
 * image, every label used, in raster order of first appearance; n_labels_dev [B] optional.
 * scikit-image cannot be run here, so label-for-label parity with it is unpinned; the restatement
 * oracle/slic_port.py and segmentation-quality measures are the gate (tests). */
int gg_slic(gg_handle h, const uint8_t* bgr_dev, int B, int H, int W, int n_segments, double compactness,
            double sigma, int max_iter, int32_t* labels_dev, int32_t* n_labels_dev, void* stream);

/* Per-pixel planes of GraphBuilder.__init__ (graph_builder.py:142-154): _lab [B,H,W,3],
 * _hsv [B,H,W,3], _gray [B,H,W], _grad [B,H,W], all float32; any output may be NULL. */
int gg_pixel_planes(gg_handle h, const uint8_t* bgr_dev, int B, int H, int W, float* lab_dev,
                    float* hsv_dev, float* gray_dev, float* grad_dev, void* stream);

/* compute_auto_prior(segments, lab, centre_sigma, contrast_sigma) (graph_builder.py:357-444) with a
 * caller-supplied float32 CIELAB plane lab_dev [B,H,W,3] (the builder itself derives Lab from BGR):
 * prior_dev [B*node_cap,3] = [fg-ness, bg-ness, ambiguity], image b's region i at row
 * b*node_cap + i; label_max_dev [B] (optional) receives max label per image (n_nodes - 1). */
int gg_auto_prior(gg_handle h, const int32_t* labels_dev, const float* lab_dev, int B, int H, int W,
                  int node_cap, double centre_sigma, double contrast_sigma, float* prior_dev,
                  int32_t* label_max_dev, void* stream);

/* ------------------------------------------------------------------ trimap network
 * Replaces ResGCNNet.load_state_dict / forward / predict_probs in eval mode
 * (model.py:449-546) including the PyG GCNConv / SAGEConv layers it calls.
 * Weight pointers are HOST float32 arrays in the reference's state-dict layout
 * (row-major [out,in] for *.weight); names follow the checkpoint keys
 * (inference.py:76-89).  gcn_* / norm_* are arrays of n_layers pointers.
 */
typedef struct gg_resgcn_weights {
  int32_t hidden;   /* D  = input_proj.0.weight.shape[0] */
  int32_t n_layers; /* n  = number of gcn_layers.*.bias keys */
  const float* jk_logits;                       /* [n+2] */
  const float *in_norm_weight, *in_norm_bias, *in_norm_mean, *in_norm_var; /* [19] */
  const float *input_proj_0_weight, *input_proj_0_bias;   /* [D,19], [D] */
  const float *input_proj_1_weight, *input_proj_1_bias;   /* LayerNorm [D] */
  const float *prior_booster_0_weight, *prior_booster_0_bias; /* [q,3], [q]; q = max(D/4,8) */
  const float *prior_booster_2_weight, *prior_booster_2_bias; /* [D,q], [D] */
  const float *edge_enc_0_weight, *edge_enc_0_bias;       /* [c,5], [c];  c = max(D/2,8) */
  const float *edge_enc_2_weight, *edge_enc_2_bias;       /* [c,c], [c] */
  const float *edge_gate_0_weight, *edge_gate_0_bias;     /* LayerNorm [c] */
  const float *edge_gate_1_weight, *edge_gate_1_bias;     /* [D,c], [D] */
  const float* const* gcn_lin_weight;  /* n x [D,D]  gcn_layers.i.lin.weight */
  const float* const* gcn_bias;        /* n x [D]    gcn_layers.i.bias */
  const float* const* norm_weight;     /* n x [D]    norms.i.weight */
  const float* const* norm_bias;       /* n x [D] */
  const float *sage_lin_l_weight, *sage_lin_l_bias, *sage_lin_r_weight; /* [D,D],[D],[D,D] */
  const float *sage_norm_weight, *sage_norm_bias;         /* [D] */
  const float *ctx_attn_weight, *ctx_attn_bias;           /* [1,D], [1] */
  const float *ctx_compress_weight, *ctx_compress_bias;   /* [D/2,D], [D/2] */
  const float *ctx_expand_weight, *ctx_expand_bias;       /* [D,D/2], [D] */
  const float *fuse_0_weight, *fuse_0_bias;               /* LayerNorm [D] */
  const float *fuse_1_weight, *fuse_1_bias;               /* [D,D], [D] */
  const float *head_weight, *head_bias;                   /* [3,D], [3] */
} gg_resgcn_weights;

int gg_load_weights(gg_handle h, const gg_resgcn_weights* w);

/* COO (2,E) int64 edge list in any order -> destination-sorted CSR (rows sorted by
 * source id, ties by edge id).  Lets forward() accept what the reference's
 * forward(data) accepts (model.py:508-514).  edge_index row stride = E. */
int gg_coo_to_csr(gg_handle h, const int64_t* edge_index_dev, int64_t n_edges, int64_t n_nodes,
                  int32_t* csr_rowptr_dev /*[n_nodes+1]*/, int32_t* csr_src_dev /*[E]*/,
                  int32_t* csr_eid_dev /*[E]*/, void* stream);

/* forward over a batch of graphs.  node_cap_total / edge_cap_total bound the row counts
 * (the true totals are read on the device from graph_off[n_graphs] and
 * csr_rowptr[total nodes]).  graph_off: int64 [n_graphs+1] node offsets per graph
 * (= gg_graph_out.node_off; a single graph is {0, N}).
 * logits_dev / probs_dev: [node_cap_total,3]; either may be NULL. */
int gg_resgcn_forward(gg_handle h, const float* x_dev, const int32_t* csr_rowptr_dev,
                      const int32_t* csr_src_dev, const int32_t* csr_eid_dev,
                      const float* edge_attr_dev, const int64_t* graph_off_dev, int n_graphs,
                      int64_t node_cap_total, int64_t edge_cap_total, float* logits_dev,
                      float* probs_dev, void* stream);

/* ------------------------------------------------------------------ network variants
 * build_model("gcn" | "gat") of the reference (model.py:593-620): GCNTrimapNet (model.py:239-316:
 * ResGCNBlock = GCNConv + BatchNorm + ReLU + identity skip + EdgeInjectionLayer, concat head) and
 * GATTrimapNet (model.py:323-414: GATv2Conv with edge features + LayerNorm + GELU +
 * EdgeInjectionLayer, skip projection, GlobalContextModule, head), eval mode.
 * tensors: HOST float32 arrays in the reference's state-dict layout, in this order
 * (BatchNorm = weight, bias, running_mean, running_var; D = hidden, n = n_layers, H = n_heads):
 *   gcn: in_norm.norm BatchNorm[19] | input_proj.0.{weight [D,19], bias} | input_proj.1 BatchNorm[D] |
 *        per block: conv.bias, conv.lin.weight [D,D], bn BatchNorm[D], edge_inject.proj.0.{weight [D,5], bias},
 *        edge_inject.proj.2.{weight [D,D], bias} | head.0.{weight [D,(n+1)D], bias} | head.1 BatchNorm[D] |
 *        head.4.{weight [D/2,D], bias} | head.6.{weight [3,D/2], bias}                         (20 + 10 n tensors)
 *   gat: in_norm.norm BatchNorm[19] | input_proj.0.{weight, bias} | input_proj.1.{weight, bias} (LayerNorm) |
 *        per layer: convs.att [1,H,D/H], convs.bias, convs.lin_l.{weight [D,D], bias}, convs.lin_r.{weight, bias},
 *        convs.lin_edge.weight [D,5], lns.{weight, bias}, edge_gates.proj.0.{weight [D,5], bias},
 *        edge_gates.proj.2.{weight [D,D], bias} | skip_proj.weight [D,D] | ctx.attn.{weight [1,D], bias [1]} |
 *        ctx.compress.{weight [D/2,D], bias} | ctx.expand.{weight [D,D/2], bias} | head.0.{weight [D,D], bias} |
 *        head.3.{weight [3,D], bias}                                                            (19 + 13 n tensors)
 * hidden: multiple of 32 in [32, 256]; n_heads must divide 32.  A handle holds one variant at a time,
 * next to the ResGCNNet weights of gg_load_weights. */
enum { GG_VARIANT_GCN = 1, GG_VARIANT_GAT = 2 };

typedef struct gg_variant_weights {
  int32_t variant;   /* GG_VARIANT_GCN | GG_VARIANT_GAT */
  int32_t hidden;
  int32_t n_layers;
  int32_t n_heads;   /* gat only */
  int32_t n_tensors;
  int32_t reserved;
  const float* const* tensors; /* [n_tensors] host pointers */
  const int64_t* numel;        /* [n_tensors] element counts (validated against the layout above) */
} gg_variant_weights;

int gg_variant_load_weights(gg_handle h, const gg_variant_weights* w);

/* forward(data) / predict_probs of the loaded variant over a batch of graphs: same graph arguments as
 * gg_resgcn_forward (dst-sorted CSR from gg_coo_to_csr or gg_build_graphs), but n_nodes / n_edges are
 * the exact row counts.  logits_dev / probs_dev: [n_nodes,3]; either may be NULL. */
int gg_variant_forward(gg_handle h, int variant, const float* x_dev, const int32_t* csr_rowptr_dev,
                       const int32_t* csr_src_dev, const int32_t* csr_eid_dev, const float* edge_attr_dev,
                       const int64_t* graph_off_dev, int n_graphs, int64_t n_nodes, int64_t n_edges,
                       float* logits_dev, float* probs_dev, void* stream);

/* ------------------------------------------------------------------ region -> pixel projection
 * gg_refine_trimap replaces refine_trimap (pipeline.py:103-146): guided filter of the BG
 * and FG posteriors under the grey image (pipeline.py:71-100, cv2.blur numerics: float64
 * window sums, BORDER_REFLECT_101) and the threshold rule of eq. 27.
 * gg_project_trimap replaces predict_trimap / _probs_to_trimap (model.py:548-557,
 * :623-678): node labels gathered through the label map (no filter).
 * probs_dev: [SN,3] rows indexed by node_off[b] + label; labels with no row
 * (label >= n_nodes[b]) read as zeros (project_to_pixels padding, model.py:655-661)
 * resp. GC_PR_BGD (model.py:672-677).  node_off: int64 [B+1].
 */
int gg_refine_trimap(gg_handle h, const uint8_t* bgr_dev, const int32_t* labels_dev,
                     const float* probs_dev, const int64_t* node_off_dev, int B, int H, int W,
                     int radius, float eps, float thr_fg, float thr_bg,
                     uint8_t* trimap_dev /*[B,H,W]*/, float* p_bg_dev /*optional [B,H,W]*/,
                     float* p_fg_dev /*optional*/, void* stream);

int gg_project_trimap(gg_handle h, const int32_t* labels_dev, const float* probs_dev,
                      const int64_t* node_off_dev, int B, int H, int W, float thr_fg,
                      float thr_bg, uint8_t* trimap_dev, void* stream);

/* ------------------------------------------------------------------ training-data labels
 * derive_trimap_labels(segments, gt_mask, fg_threshold, bg_threshold) and the fg_ratio tensor of
 * prepare_sample (dataset.py:175-205, 239-249), batched: gt_mask_dev uint8 [B,H,W] (> 0 =
 * foreground).  Per node (rows as gg_graph_out.x, image b at node_off[b]):
 *   fg_ratio = #foreground pixels / max(#pixels, 1)   (float64 quotient, stored as float32)
 *   y        = 2 (FG) if fg_ratio >= fg_threshold; 0 (BG) if fg_ratio <= 1 - bg_threshold (BG is
 *              assigned second and wins); 1 (UNK) otherwise and for regions without pixels.
 * Either output may be NULL.  A label outside [0, n_nodes[b]) sets the device status word. */
int gg_region_labels(gg_handle h, const int32_t* labels_dev, const uint8_t* gt_mask_dev,
                     const int64_t* node_off_dev, int B, int H, int W, int64_t node_cap_total,
                     double fg_threshold, double bg_threshold, float* fg_ratio_dev /*[SN]*/,
                     int64_t* y_dev /*[SN]*/, void* stream);

/* _seed_from_prior(trimap, graph, seed_frac) (pipeline.py:149-186), batched and in place: an image
 * whose trimap has no foreground label (1, 3) gets its max(1, round(seed_frac * n_nodes)) regions
 * with the largest foreground prior (node_input column 16) set to GC_PR_FGD; likewise the
 * background prior (column 17) and GC_PR_BGD when no background label (0, 2) exists.  Equal prior
 * values: the larger region index is taken first.  x_dev = gg_graph_out.x ([SN,19]). */
int gg_seed_from_prior(gg_handle h, uint8_t* trimap_dev, const int32_t* labels_dev, const float* x_dev,
                       const int64_t* node_off_dev, int B, int H, int W, int64_t node_cap_total,
                       double seed_frac, void* stream);

/* The guards of GrabCut.run_with_trimap (grabcut.py:127-140), batched and in place: an image without
 * a GC_FGD (1) pixel gets its GC_PR_FGD (3) pixels promoted to GC_FGD, likewise GC_PR_BGD (2) ->
 * GC_BGD (0).  degenerate_dev [B] (optional): 1 where a side is still missing afterwards -- the
 * reference then skips cv2.grabCut and returns the trimap's own labelling. */
int gg_grabcut_guards(gg_handle h, uint8_t* trimap_dev, int B, int H, int W, int32_t* degenerate_dev,
                      void* stream);

/* clean_mask(mask, min_area_ratio, keep_largest) (pipeline.py:189-227), batched: 8-connected
 * components of the binary masks uint8 {0,1} [B,H,W]; components smaller than
 * min_area_ratio * H * W are dropped (if none survives the largest one is kept), or only the
 * largest one is kept; among equally large components the first in raster order wins, as
 * cv2.connectedComponentsWithStats + argmax does.  out_dev may alias mask_dev.  The caller applies
 * the reference's early returns (empty mask; min_area_ratio <= 0 and not keep_largest). */
int gg_clean_masks(gg_handle h, const uint8_t* mask_dev, uint8_t* out_dev, int B, int H, int W,
                   double min_area_ratio, int keep_largest, void* stream);

/* guided_filter(guide, src, radius, eps) on single float32 planes (pipeline.py:71-100). */
int gg_guided_filter(gg_handle h, const float* guide_dev, const float* src_dev, int H, int W,
                     int radius, float eps, float* out_dev, void* stream);

/* ------------------------------------------------------------------ whole path, host buffers
 * The call pipeline.segment() makes between cv2.imread and cv2.grabCut
 * (pipeline.py:298-321), batched: HOST images + label maps in, HOST trimaps out, with the
 * host<->device copies inside the call (chunked and overlapped with the kernels on
 * internal streams).  Synchronous.  Optional outputs may be NULL.
 */
typedef struct gg_path_config {
  gg_graph_config graph;
  int32_t radius;      /* guided-filter radius (pipeline.py:274; default 8) */
  float eps;           /* pipeline.py:75 (1e-3) */
  float thr_fg;        /* pipeline.py:268 (0.55) */
  float thr_bg;
  int32_t edge_aware;  /* 1: refine_trimap, 0: predict_trimap (pipeline.py:312-321) */
  int32_t chunk;       /* images per pipelined chunk; 0 = choose */
  int32_t label_bytes; /* host entry points only: 4 (or 0) = int32 label maps as the reference holds them
                          (graph_builder.py:188); 2 = uint16 label maps -- a compact transport for
                          PCIe-bound streaming (5 instead of 7 bytes per pixel), widened on the device */
  int32_t reserved;
  double seed_frac;    /* > 0: repair one-sided trimaps like _seed_from_prior (pipeline.py:149-186,
                          called by segment() with 0.1); 0 = leave the trimap as predicted */
  /* Superpixels on the device (GraphBuilder._compute_superpixels, graph_builder.py:177-188): with
   * slic_segments > 0 the label-map argument of the whole-path entry points may be NULL -- the
   * maps are produced by gg_slic from the images (3 instead of 7 bytes per pixel cross PCIe) -- and
   * graph.node_cap must leave room for the label count (about slic_segments; every label
   * >= node_cap is reported through the status word). */
  int32_t slic_segments;     /* SuperpixelGraphConfig.n_segments, 0 = label maps are supplied */
  int32_t slic_iters;        /* k-means iterations, 0 = 10 (skimage max_num_iter) */
  float slic_compactness;    /* SuperpixelGraphConfig.compactness, 0 = 10 */
  float slic_sigma;          /* SuperpixelGraphConfig.sigma (Gaussian pre-smoothing), < 0 = 1 */
} gg_path_config;

int gg_trimap_path_host(gg_handle h, const uint8_t* bgr_host, const void* labels_host /*int32 or uint16 [B,H,W]*/, int B,
                        int H, int W, const gg_path_config* cfg, uint8_t* trimap_host /*[B,H,W]*/,
                        int32_t* n_nodes_host /*optional [B]*/, int32_t* n_edges_host /*optional [B]*/);

/* Asynchronous form of the same call, for streaming: _submit enqueues the copies and kernels
 * of one batch and returns a ticket at once; _wait blocks until that batch's trimaps (and
 * counts) are in the host buffers, which must stay valid (and should be pinned) until then.
 * Up to 8 calls may be in flight; the chunk slots rotate across calls, so the copy-in of
 * batch n+1 overlaps the kernels of batch n.  gg_trimap_path_host == _submit + _wait.
 * The device status is sticky over the calls in flight: _wait reports GG_ERR_CAPACITY if any
 * batch submitted so far overflowed. */
int gg_trimap_path_host_submit(gg_handle h, const uint8_t* bgr_host, const void* labels_host,
                               int B, int H, int W, const gg_path_config* cfg,
                               uint8_t* trimap_host, int32_t* n_nodes_host,
                               int32_t* n_edges_host, int* ticket);
int gg_trimap_path_host_wait(gg_handle h, int ticket);

/* Same path with inputs and outputs resident on the device (caller's stream, no host sync). */
int gg_trimap_path_device(gg_handle h, const uint8_t* bgr_dev, const int32_t* labels_dev, int B,
                          int H, int W, const gg_path_config* cfg, uint8_t* trimap_dev,
                          float* probs_dev /*optional [B*node_cap,3]*/,
                          int64_t* node_off_dev /*optional [B+1]*/, void* stream);

/* Device self test of the float32 fast paths the pixel kernels use instead of the IEEE
 * division / square-root sequences (csrc/pixel_math.cuh), against __fdiv_rn / __fsqrt_rn:
 * mismatches[0] sqrt of every integer < 2^24 (Sobel magnitudes), [1] saturation quotients
 * d/max, [2] hue quotients p/(6d) (both exhaustive over uint8 colours), [3] normalised gradient
 * g/(gmax+1e-6) for every Sobel magnitude and 64 image maxima each.  All four must be 0 for the
 * per-pixel planes to be the reference's float32 values (graph_builder.py:142-154). */
int gg_selftest_math(gg_handle h, int64_t* mismatches /*[4]*/);

/* Number of kernels of this library launched by this handle so far (bench bookkeeping). */
int64_t gg_kernel_launch_count(gg_handle h);

/* Per-kernel timing with CUDA events recorded on the launching stream around every kernel
 * of this library.  enable=1 starts a fresh session, enable=0 stops recording.
 * gg_profile_report synchronises the device and writes one line per kernel,
 * "name,launches,total_ms", sorted by total time, into buf (NUL-terminated). */
int gg_profile_enable(gg_handle h, int enable);
int gg_profile_report(gg_handle h, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* GCN_GRABCUT_B200_H */
