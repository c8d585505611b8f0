// C ABI of libgcn_grabcut_b200.so (see include/gcn_grabcut_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "graph_build.cuh"
#include "pixel_math.cuh"
#include "resgcn.cuh"
#include "slic.cuh"
#include "trimap.cuh"
#include "variants.cuh"

namespace gg {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int Arena::reserve(size_t bytes) {
  off = 0;
  overflowed = false;
  if (bytes <= cap) return GG_OK;
  if (base) {
    GG_CUDA_OK(cudaDeviceSynchronize());   // the old buffer may still be in use by queued work
    GG_CUDA_OK(cudaFree(base));
    base = nullptr;
    cap = 0;
  }
  const size_t want = bytes + bytes / 8 + (1u << 20);
  GG_CUDA_OK(cudaMalloc((void**)&base, want));
  cap = want;
  return GG_OK;
}

void Arena::release() {
  if (base) cudaFree(base);
  base = nullptr;
  cap = off = 0;
}

// an entry point whose kernels were handed workspace past the reservation fails (the estimate in
// *_workspace_bytes disagrees with the takes: a bug, reported instead of corrupting memory)
static int arena_checked(Arena& ar, int rc, const char* what) {
  if (ar.overflowed) {
    ar.overflowed = false;
    set_error("%s: internal workspace overflow (reserved %zu bytes)", what, ar.cap);
    return GG_ERR_INVALID;
  }
  return rc;
}

static cudaEvent_t prof_event(gg_context* ctx) {
  if (ctx->prof_pool_used == ctx->prof_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->prof_pool.push_back(e);
  }
  return ctx->prof_pool[ctx->prof_pool_used++];
}
void prof_begin(gg_context* ctx, const char* name, cudaStream_t st) {
  if (*name == '(') ++name;          // GG_LAUNCH((k<a, b>), ...) stringifies with the parenthesis
  gg_context::ProfRec r{name, prof_event(ctx), prof_event(ctx)};
  cudaEventRecord(r.e0, st);
  ctx->prof.push_back(r);
}
void prof_end(gg_context* ctx, cudaStream_t st) { cudaEventRecord(ctx->prof.back().e1, st); }

// uint16 label maps (compact host transport, gg_path_config.label_bytes == 2) -> the int32 maps
// the kernels read
__global__ void __launch_bounds__(256)
k_widen_labels(const uint16_t* __restrict__ in, int32_t* __restrict__ out, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n && ((uintptr_t)(in + i) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i);
    int4 a = make_int4(v.x & 0xffff, v.x >> 16, v.y & 0xffff, v.y >> 16);
    int4 b = make_int4(v.z & 0xffff, v.z >> 16, v.w & 0xffff, v.w >> 16);
    *reinterpret_cast<int4*>(out + i) = a;
    *reinterpret_cast<int4*>(out + i + 4) = b;
  } else {
    for (size_t k = i; k < n && k < i + 8; ++k) out[k] = in[k];
  }
}

__global__ void k_status_or(const int* __restrict__ word, int* __restrict__ sticky) {
  if (*word) atomicOr(sticky, *word);
}

static gg_graph_config norm_cfg(const gg_graph_config& c) {
  gg_graph_config o = c;
  if (o.pair_cap <= 0) o.pair_cap = 8 * o.node_cap;
  return o;
}

// workspace slice of the whole path for `B` images (graph + network + trimap), plus the
// ragged graph arrays that the device-pointer API would otherwise receive from the caller.
struct PathBuffers {
  gg_graph_out g;
  float* probs;
};

static size_t path_graph_arrays_bytes(int B, const gg_graph_config& cfg) {
  const size_t SN = (size_t)B * cfg.node_cap, SE = (size_t)2 * B * cfg.pair_cap;
  size_t s = 0;
  s += Arena::padded(B, 4) * 2 + Arena::padded(B + 1, 8) * 2;
  s += Arena::padded(SN * 19, 4) + Arena::padded(SE * 2, 8) + Arena::padded(SE * 5, 4);
  s += Arena::padded(SN + 1, 4) + Arena::padded(SE, 4) * 2;
  s += Arena::padded(SN * 3, 4);
  return s + 1024;
}

static PathBuffers take_path_buffers(Arena& ar, int B, const gg_graph_config& cfg) {
  const size_t SN = (size_t)B * cfg.node_cap, SE = (size_t)2 * B * cfg.pair_cap;
  PathBuffers pb{};
  pb.g.n_nodes = ar.take<int32_t>(B);
  pb.g.n_edges = ar.take<int32_t>(B);
  pb.g.node_off = ar.take<int64_t>(B + 1);
  pb.g.edge_off = ar.take<int64_t>(B + 1);
  pb.g.x = ar.take<float>(SN * 19);
  pb.g.edge_index = ar.take<int64_t>(SE * 2);
  pb.g.edge_attr = ar.take<float>(SE * 5);
  pb.g.csr_rowptr = ar.take<int32_t>(SN + 1);
  pb.g.csr_src = ar.take<int32_t>(SE);
  pb.g.csr_eid = ar.take<int32_t>(SE);
  pb.probs = ar.take<float>(SN * 3);
  return pb;
}

static size_t path_workspace_bytes(gg_context* ctx, int B, int H, int W, const gg_path_config& pc) {
  const gg_graph_config cfg = norm_cfg(pc.graph);
  const long long SN = (long long)B * cfg.node_cap, SE = 2ll * B * cfg.pair_cap;
  const size_t slic = pc.slic_segments > 0 ? slic_workspace_bytes(B, H, W, pc.slic_segments) +
                                                 Arena::padded((size_t)B * H * W, 4) : 0;
  return path_graph_arrays_bytes(B, cfg) + graph_workspace_bytes(B, H, W, cfg) +
         resgcn_workspace_bytes(ctx->net, SN, SE, B) + trimap_workspace_bytes(B, H, W, false) +
         seed_workspace_bytes(B, SN) + slic;
}

static int run_path(gg_context* ctx, Arena& ar, const uint8_t* bgr, const int32_t* labels, int B,
                    int H, int W, const gg_path_config& pc, uint8_t* trimap, float* probs_out,
                    int64_t* node_off_out, int32_t* n_nodes_out, int32_t* n_edges_out,
                    cudaStream_t st, cudaEvent_t graph_done = nullptr);

// The same path with the batch cut into ctx->n_sub contiguous sub-batches that run
// concurrently on internal streams (forked from / joined into the caller's stream).  The
// kernels of the path are latency-bound at one batch per GPU; two independent sub-batches in
// flight fill the idle issue slots.  Outputs that are prefix-summed over the whole batch
// (probs / node_off) force a single batch.
static int run_path_multi(gg_context* ctx, Arena& ar, const uint8_t* bgr, const int32_t* labels, int B,
                          int H, int W, const gg_path_config& pc, uint8_t* trimap, float* probs_out,
                          int64_t* node_off_out, int32_t* n_nodes_out, int32_t* n_edges_out,
                          cudaStream_t st, int first_sub_stream = 0);

// graph build -> network -> trimap for B images whose inputs are on the device.
static int run_path(gg_context* ctx, Arena& ar, const uint8_t* bgr, const int32_t* labels, int B,
                    int H, int W, const gg_path_config& pc, uint8_t* trimap, float* probs_out,
                    int64_t* node_off_out, int32_t* n_nodes_out, int32_t* n_edges_out,
                    cudaStream_t st, cudaEvent_t graph_done) {
  const gg_graph_config cfg = norm_cfg(pc.graph);
  const long long SN = (long long)B * cfg.node_cap, SE = 2ll * B * cfg.pair_cap;
  PathBuffers pb = take_path_buffers(ar, B, cfg);
  if (labels == nullptr) {
    // superpixels on the device: the workspace of SLIC is dead once the label maps exist, the
    // stages that follow reuse it (mark / rewind)
    GG_REQUIRE(pc.slic_segments > 0, "trimap path: no label maps and slic_segments == 0");
    int32_t* lab = ar.take<int32_t>((size_t)B * H * W);
    const size_t mark = ar.off;
    GG_TRY(slic_labels(ctx, ar, bgr, B, H, W, pc.slic_segments, pc.slic_compactness > 0 ? pc.slic_compactness : 10.0,
                       pc.slic_sigma < 0 ? 1.0 : pc.slic_sigma, pc.slic_iters > 0 ? pc.slic_iters : 10, lab, nullptr, st));
    ar.off = mark;
    labels = lab;
  }
  const uint8_t* gray = nullptr;
  GG_TRY(build_graphs(ctx, ar, bgr, labels, B, H, W, cfg, pb.g, st, &gray));
  if (graph_done) GG_CUDA_OK(cudaEventRecord(graph_done, st));
  GG_TRY(resgcn_forward(ctx, ar, pb.g.x, pb.g.csr_rowptr, pb.g.csr_src, pb.g.csr_eid, pb.g.edge_attr,
                        pb.g.node_off, B, SN, SE, nullptr, pb.probs, st, cfg.node_cap, 2 * cfg.pair_cap));
  if (pc.edge_aware) {
    GG_TRY(refine_trimap(ctx, ar, bgr, gray, labels, pb.probs, pb.g.node_off, B, H, W, pc.radius,
                         pc.eps, pc.thr_fg, pc.thr_bg, trimap, nullptr, nullptr, st, cfg.node_cap));
  } else {
    GG_TRY(project_trimap(ctx, labels, pb.probs, pb.g.node_off, B, H, W, pc.thr_fg, pc.thr_bg, trimap, st));
  }
  if (pc.seed_frac > 0.0)
    GG_TRY(seed_from_prior(ctx, ar, trimap, labels, pb.g.x, pb.g.node_off, B, H, W, SN, pc.seed_frac, st));
  if (probs_out)
    GG_CUDA_OK(cudaMemcpyAsync(probs_out, pb.probs, (size_t)SN * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (node_off_out)
    GG_CUDA_OK(cudaMemcpyAsync(node_off_out, pb.g.node_off, (size_t)(B + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (n_nodes_out)
    GG_CUDA_OK(cudaMemcpyAsync(n_nodes_out, pb.g.n_nodes, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (n_edges_out)
    GG_CUDA_OK(cudaMemcpyAsync(n_edges_out, pb.g.n_edges, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return GG_OK;
}

static int run_path_multi(gg_context* ctx, Arena& ar, const uint8_t* bgr, const int32_t* labels, int B,
                          int H, int W, const gg_path_config& pc, uint8_t* trimap, float* probs_out,
                          int64_t* node_off_out, int32_t* n_nodes_out, int32_t* n_edges_out,
                          cudaStream_t st, int first_sub_stream) {
  int S = std::min(ctx->n_sub, B);
  if (probs_out || node_off_out || B < 16) S = 1;
  if (S <= 1) {
    ctx->status_word = ctx->d_status;
    return run_path(ctx, ar, bgr, labels, B, H, W, pc, trimap, probs_out, node_off_out, n_nodes_out,
                    n_edges_out, st);
  }
  const size_t npx = (size_t)H * W;
  const int per = (B + S - 1) / S;
  const size_t slice = path_workspace_bytes(ctx, per, H, W, pc);
  cudaEvent_t ev_fork = ctx->ev[12];
  GG_CUDA_OK(cudaEventRecord(ev_fork, st));
  int rc = GG_OK;
  for (int s = 0; s < S && rc == GG_OK; ++s) {
    const int b0 = s * per, nb = std::min(per, B - b0);
    if (nb <= 0) break;
    cudaStream_t ss = ctx->s_sub[(first_sub_stream + s) & 3];
    GG_CUDA_OK(cudaStreamWaitEvent(ss, ev_fork, 0));
    // stagger: sub-batch s starts its (issue-bound) pixel kernels when sub-batch s-1 has finished
    // its own and moved on to the (latency-bound, low-occupancy) network stage
    if (s > 0 && ctx->stagger) GG_CUDA_OK(cudaStreamWaitEvent(ss, ctx->ev[17 + s - 1], 0));
    Arena sub;
    sub.base = ar.base + ar.off + (size_t)s * slice;
    sub.cap = slice;
    ctx->status_word = ctx->d_status + 8 + s;
    rc = run_path(ctx, sub, bgr + (size_t)b0 * npx * 3, labels ? labels + (size_t)b0 * npx : nullptr, nb, H, W, pc,
                  trimap + (size_t)b0 * npx, nullptr, nullptr, n_nodes_out ? n_nodes_out + b0 : nullptr,
                  n_edges_out ? n_edges_out + b0 : nullptr, ss, ctx->ev[17 + s]);
    if (sub.overflowed) ar.overflowed = true;
    if (rc == GG_OK) {
      GG_LAUNCH(ctx, k_status_or, 1, 1, 0, ss, ctx->d_status + 8 + s, ctx->d_status);
      GG_CUDA_OK(cudaEventRecord(ctx->ev[13 + s], ss));
      GG_CUDA_OK(cudaStreamWaitEvent(st, ctx->ev[13 + s], 0));
    }
  }
  ctx->status_word = ctx->d_status;
  ar.off += (size_t)S * slice;
  return rc;
}

static size_t path_multi_workspace_bytes(gg_context* ctx, int B, int H, int W, const gg_path_config& pc) {
  const int S = std::max(1, std::min(ctx->n_sub, B));
  const int per = (B + S - 1) / S;
  return std::max(path_workspace_bytes(ctx, B, H, W, pc), (size_t)S * path_workspace_bytes(ctx, per, H, W, pc));
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_abi_version(void) { return GG_ABI_VERSION; }
const char* gg_last_error(void) { return g_err; }

int gg_create(gg_handle* out, int device) {
  GG_REQUIRE(out != nullptr, "gg_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("gg_create: no CUDA device (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return GG_ERR_CUDA;
  }
  GG_REQUIRE(device >= 0 && device < n, "gg_create: device %d out of range (%d devices)", device, n);
  GG_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GG_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("gg_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
              device, prop.major, prop.minor);
    return GG_ERR_CUDA;
  }
  gg_context* c = new gg_context();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  GG_CUDA_OK(cudaMalloc((void**)&c->d_status, 64));
  GG_CUDA_OK(cudaMemset(c->d_status, 0, 64));
  GG_CUDA_OK(cudaHostAlloc((void**)&c->h_ticket_status, gg_context::MAX_TICKETS * sizeof(int), cudaHostAllocDefault));
  memset(c->h_ticket_status, 0, gg_context::MAX_TICKETS * sizeof(int));
  c->status_word = c->d_status;
  double lin[256];
  for (int v = 0; v < 256; ++v) lin[v] = srgb_linear(v);
  GG_CUDA_OK(cudaMalloc((void**)&c->d_lin, sizeof(lin)));
  GG_CUDA_OK(cudaMemcpy(c->d_lin, lin, sizeof(lin), cudaMemcpyHostToDevice));
  GG_CUDA_OK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  GG_CUDA_OK(cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking));
  GG_CUDA_OK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (auto& sst : c->s_sub) GG_CUDA_OK(cudaStreamCreateWithFlags(&sst, cudaStreamNonBlocking));
  if (const char* ns = getenv("GG_SUBBATCH")) c->n_sub = std::max(1, std::min(4, atoi(ns)));
  if (const char* sg = getenv("GG_STAGGER")) c->stagger = atoi(sg) != 0;
  if (getenv("GG_GCN_UNFUSED")) c->gcn_fused = 0;
  c->ev.resize(24);
  for (auto& ev : c->ev) GG_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto& ev : c->ticket_ev) GG_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync));
  const char* impl = getenv("GG_GEMM_IMPL");
  c->gemm_impl = 1;
  if (impl && !strcmp(impl, "simt")) c->gemm_impl = 0;
  *out = c;
  return GG_OK;
}

void gg_destroy(gg_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  h->arena.release();
  h->host_arena.release();
  if (h->net.blob) cudaFree(h->net.blob);
  if (h->net.tc_blob) cudaFree(h->net.tc_blob);
  if (h->variant.blob) cudaFree(h->variant.blob);
  if (h->variant.tc_blob) cudaFree(h->variant.tc_blob);
  if (h->variant.d_rows) cudaFree(h->variant.d_rows);
  if (h->d_status) cudaFree(h->d_status);
  if (h->h_ticket_status) cudaFreeHost(h->h_ticket_status);
  if (h->d_lin) cudaFree(h->d_lin);
  if (h->d_coord) cudaFree(h->d_coord);
  for (auto& ev : h->ev) cudaEventDestroy(ev);
  for (auto& ev : h->ticket_ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : h->prof_pool) cudaEventDestroy(ev);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_run) cudaStreamDestroy(h->s_run);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  for (auto& sst : h->s_sub) if (sst) cudaStreamDestroy(sst);
  delete h;
}

int gg_set_option(gg_handle h, const char* key, int value) {
  GG_REQUIRE(h && key, "gg_set_option: null");
  if (!strcmp(key, "gemm_impl")) { h->gemm_impl = value; return GG_OK; }
  if (!strcmp(key, "knn_legacy")) { h->knn_legacy = value != 0; return GG_OK; }
  if (!strcmp(key, "n_sub")) { h->n_sub = std::max(1, std::min(4, value)); return GG_OK; }
  if (!strcmp(key, "gcn_fused")) { h->gcn_fused = value; return GG_OK; }   // 0 off, 1 auto (>= 24 graphs), 2 always
  set_error("gg_set_option: unknown key %s", key);
  return GG_ERR_INVALID;
}

int gg_check_device_status(gg_handle h, void* stream, int* status_bits) {
  GG_REQUIRE(h && status_bits, "gg_check_device_status: null");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_CUDA_OK(cudaMemcpyAsync(status_bits, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GG_CUDA_OK(cudaMemsetAsync(h->d_status, 0, sizeof(int), (cudaStream_t)stream));   // read-and-clear
  GG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  if (*status_bits) {
    set_error("device status 0x%x:%s%s%s", *status_bits,
              (*status_bits & ST_LABEL_RANGE) ? " label/index out of range (>= node_cap)" : "",
              (*status_bits & ST_PAIR_TABLE) ? " adjacency table overflow" : "",
              (*status_bits & ST_EDGE_CAP) ? " pair/edge capacity exceeded" : "");
    return GG_ERR_CAPACITY;
  }
  return GG_OK;
}

int gg_build_graphs(gg_handle h, const uint8_t* bgr_dev, const int32_t* labels_dev, int B, int H, int W,
                    const gg_graph_config* cfg, const gg_graph_out* out, void* stream) {
  GG_REQUIRE(h && bgr_dev && labels_dev && cfg && out, "gg_build_graphs: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  const gg_graph_config c = norm_cfg(*cfg);
  GG_REQUIRE(c.node_cap > 0 && B > 0 && H > 0 && W > 0, "gg_build_graphs: bad sizes");
  GG_TRY(h->arena.reserve(graph_workspace_bytes(B, H, W, c)));
  return arena_checked(h->arena, build_graphs(h, h->arena, bgr_dev, labels_dev, B, H, W, c, *out, (cudaStream_t)stream, nullptr),
                       "gg_build_graphs");
}

int gg_slic(gg_handle h, const uint8_t* bgr_dev, int B, int H, int W, int n_segments, double compactness, double sigma,
            int max_iter, int32_t* labels_dev, int32_t* n_labels_dev, void* stream) {
  GG_REQUIRE(h && bgr_dev && labels_dev, "gg_slic: null argument");
  GG_REQUIRE(B > 0 && H >= 2 && W >= 2 && n_segments >= 1, "gg_slic: bad sizes");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(slic_workspace_bytes(B, H, W, n_segments)));
  return arena_checked(h->arena, slic_labels(h, h->arena, bgr_dev, B, H, W, n_segments, compactness, sigma,
                                             max_iter > 0 ? max_iter : 10, labels_dev, n_labels_dev, (cudaStream_t)stream),
                       "gg_slic");
}

int gg_pixel_planes(gg_handle h, const uint8_t* bgr_dev, int B, int H, int W, float* lab_dev,
                    float* hsv_dev, float* gray_dev, float* grad_dev, void* stream) {
  GG_REQUIRE(h && bgr_dev, "gg_pixel_planes: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  return pixel_planes(h, h->arena, bgr_dev, B, H, W, lab_dev, hsv_dev, gray_dev, grad_dev, (cudaStream_t)stream);
}

int gg_auto_prior(gg_handle h, const int32_t* labels_dev, const float* lab_dev, int B, int H, int W, int node_cap,
                  double centre_sigma, double contrast_sigma, float* prior_dev, int32_t* label_max_dev, void* stream) {
  GG_REQUIRE(h && labels_dev && lab_dev && prior_dev, "gg_auto_prior: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_REQUIRE(B > 0 && H > 0 && W > 0 && node_cap > 0, "gg_auto_prior: bad sizes");
  GG_TRY(h->arena.reserve(auto_prior_workspace_bytes(B, node_cap)));
  return arena_checked(h->arena, auto_prior(h, h->arena, labels_dev, lab_dev, B, H, W, node_cap, centre_sigma,
                                            contrast_sigma, prior_dev, label_max_dev, (cudaStream_t)stream),
                       "gg_auto_prior");
}

int gg_load_weights(gg_handle h, const gg_resgcn_weights* w) {
  GG_REQUIRE(h && w, "gg_load_weights: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  return load_weights(h, w);
}

int gg_coo_to_csr(gg_handle h, const int64_t* edge_index_dev, int64_t n_edges, int64_t n_nodes,
                  int32_t* csr_rowptr_dev, int32_t* csr_src_dev, int32_t* csr_eid_dev, void* stream) {
  GG_REQUIRE(h && csr_rowptr_dev && (n_edges == 0 || (edge_index_dev && csr_src_dev && csr_eid_dev)),
             "gg_coo_to_csr: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(Arena::padded((size_t)n_nodes, 4) + 4096));
  return arena_checked(h->arena, coo_to_csr(h, h->arena, edge_index_dev, n_edges, n_nodes, csr_rowptr_dev, csr_src_dev,
                                            csr_eid_dev, (cudaStream_t)stream), "gg_coo_to_csr");
}

int gg_resgcn_forward(gg_handle h, const float* x_dev, const int32_t* csr_rowptr_dev,
                      const int32_t* csr_src_dev, const int32_t* csr_eid_dev, const float* edge_attr_dev,
                      const int64_t* graph_off_dev, int n_graphs, int64_t node_cap_total,
                      int64_t edge_cap_total, float* logits_dev, float* probs_dev, void* stream) {
  GG_REQUIRE(h && x_dev && csr_rowptr_dev && graph_off_dev, "gg_resgcn_forward: null argument");
  GG_REQUIRE(logits_dev || probs_dev, "gg_resgcn_forward: no output requested");
  GG_CUDA_OK(cudaSetDevice(h->device));
  if (!h->net.loaded) { set_error("gg_resgcn_forward: call gg_load_weights first"); return GG_ERR_STATE; }
  GG_TRY(h->arena.reserve(resgcn_workspace_bytes(h->net, node_cap_total, edge_cap_total, n_graphs)));
  return arena_checked(h->arena, resgcn_forward(h, h->arena, x_dev, csr_rowptr_dev, csr_src_dev, csr_eid_dev, edge_attr_dev,
                                                graph_off_dev, n_graphs, node_cap_total, edge_cap_total, logits_dev,
                                                probs_dev, (cudaStream_t)stream,
                                                n_graphs == 1 && node_cap_total <= FUSED_MAX_NODES ? (int)node_cap_total : 0,
                                                n_graphs == 1 && edge_cap_total < (1 << 18) ? (int)edge_cap_total : 0),
                       "gg_resgcn_forward");
}

int gg_variant_load_weights(gg_handle h, const gg_variant_weights* w) {
  GG_REQUIRE(h && w, "gg_variant_load_weights: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_CUDA_OK(cudaDeviceSynchronize());          // queued forwards may still read the old blob
  return variant_load_weights(h, w);
}

int gg_variant_forward(gg_handle h, int variant, const float* x_dev, const int32_t* csr_rowptr_dev,
                       const int32_t* csr_src_dev, const int32_t* csr_eid_dev, const float* edge_attr_dev,
                       const int64_t* graph_off_dev, int n_graphs, int64_t n_nodes, int64_t n_edges,
                       float* logits_dev, float* probs_dev, void* stream) {
  GG_REQUIRE(h && x_dev && csr_rowptr_dev && graph_off_dev, "gg_variant_forward: null argument");
  GG_REQUIRE(logits_dev || probs_dev, "gg_variant_forward: no output requested");
  GG_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_graphs >= 1 && n_nodes < (1ll << 31) / 256 * 8 && n_edges < (1ll << 31) / 256,
             "gg_variant_forward: bad sizes");
  GG_REQUIRE(n_edges == 0 || (csr_src_dev && csr_eid_dev && edge_attr_dev), "gg_variant_forward: null edge arrays");
  GG_CUDA_OK(cudaSetDevice(h->device));
  if (!h->variant.loaded || h->variant.kind != variant) {
    set_error("gg_variant_forward: call gg_variant_load_weights for variant %d first", variant);
    return GG_ERR_STATE;
  }
  GG_TRY(h->arena.reserve(variant_workspace_bytes(h->variant, n_nodes, n_edges)));
  return arena_checked(h->arena, variant_forward(h, h->arena, variant, x_dev, csr_rowptr_dev, csr_src_dev, csr_eid_dev,
                                                 edge_attr_dev, graph_off_dev, n_graphs, n_nodes, n_edges, logits_dev,
                                                 probs_dev, (cudaStream_t)stream), "gg_variant_forward");
}

int gg_refine_trimap(gg_handle h, const uint8_t* bgr_dev, const int32_t* labels_dev, const float* probs_dev,
                     const int64_t* node_off_dev, int B, int H, int W, int radius, float eps, float thr_fg,
                     float thr_bg, uint8_t* trimap_dev, float* p_bg_dev, float* p_fg_dev, void* stream) {
  GG_REQUIRE(h && bgr_dev && labels_dev && probs_dev && node_off_dev && trimap_dev, "gg_refine_trimap: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(trimap_workspace_bytes(B, H, W, true)));
  return arena_checked(h->arena, refine_trimap(h, h->arena, bgr_dev, nullptr, labels_dev, probs_dev, node_off_dev, B, H, W,
                                               radius, eps, thr_fg, thr_bg, trimap_dev, p_bg_dev, p_fg_dev,
                                               (cudaStream_t)stream), "gg_refine_trimap");
}

int gg_project_trimap(gg_handle h, const int32_t* labels_dev, const float* probs_dev,
                      const int64_t* node_off_dev, int B, int H, int W, float thr_fg, float thr_bg,
                      uint8_t* trimap_dev, void* stream) {
  GG_REQUIRE(h && labels_dev && probs_dev && node_off_dev && trimap_dev, "gg_project_trimap: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  return project_trimap(h, labels_dev, probs_dev, node_off_dev, B, H, W, thr_fg, thr_bg, trimap_dev,
                        (cudaStream_t)stream);
}

int gg_region_labels(gg_handle h, const int32_t* labels_dev, const uint8_t* gt_mask_dev,
                     const int64_t* node_off_dev, int B, int H, int W, int64_t node_cap_total,
                     double fg_threshold, double bg_threshold, float* fg_ratio_dev, int64_t* y_dev, void* stream) {
  GG_REQUIRE(h && labels_dev && gt_mask_dev && node_off_dev && (fg_ratio_dev || y_dev), "gg_region_labels: null argument");
  GG_REQUIRE(B > 0 && H > 0 && W > 0 && node_cap_total > 0, "gg_region_labels: bad sizes");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(2 * Arena::padded((size_t)node_cap_total, 4) + 1024));
  return arena_checked(h->arena, region_labels(h, h->arena, labels_dev, gt_mask_dev, node_off_dev, B, H, W, node_cap_total,
                                               fg_threshold, bg_threshold, fg_ratio_dev,
                                               reinterpret_cast<long long*>(y_dev), (cudaStream_t)stream),
                       "gg_region_labels");
}

int gg_seed_from_prior(gg_handle h, uint8_t* trimap_dev, const int32_t* labels_dev, const float* x_dev,
                       const int64_t* node_off_dev, int B, int H, int W, int64_t node_cap_total,
                       double seed_frac, void* stream) {
  GG_REQUIRE(h && trimap_dev && labels_dev && x_dev && node_off_dev, "gg_seed_from_prior: null argument");
  GG_REQUIRE(B > 0 && H > 0 && W > 0 && node_cap_total > 0, "gg_seed_from_prior: bad sizes");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(seed_workspace_bytes(B, node_cap_total)));
  return arena_checked(h->arena, seed_from_prior(h, h->arena, trimap_dev, labels_dev, x_dev, node_off_dev, B, H, W,
                                                 node_cap_total, seed_frac, (cudaStream_t)stream), "gg_seed_from_prior");
}

int gg_grabcut_guards(gg_handle h, uint8_t* trimap_dev, int B, int H, int W, int32_t* degenerate_dev, void* stream) {
  GG_REQUIRE(h && trimap_dev, "gg_grabcut_guards: null argument");
  GG_REQUIRE(B > 0 && H > 0 && W > 0, "gg_grabcut_guards: bad sizes");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(Arena::padded((size_t)B, 4) + 512));
  return arena_checked(h->arena, grabcut_guards(h, h->arena, trimap_dev, B, H, W, degenerate_dev, (cudaStream_t)stream),
                       "gg_grabcut_guards");
}

int gg_clean_masks(gg_handle h, const uint8_t* mask_dev, uint8_t* out_dev, int B, int H, int W, double min_area_ratio,
                   int keep_largest, void* stream) {
  GG_REQUIRE(h && mask_dev && out_dev, "gg_clean_masks: null argument");
  GG_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 31), "gg_clean_masks: bad sizes");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(clean_workspace_bytes(B, H, W)));
  return arena_checked(h->arena, clean_masks(h, h->arena, mask_dev, out_dev, B, H, W, min_area_ratio, keep_largest,
                                             (cudaStream_t)stream), "gg_clean_masks");
}

int gg_guided_filter(gg_handle h, const float* guide_dev, const float* src_dev, int H, int W, int radius,
                     float eps, float* out_dev, void* stream) {
  GG_REQUIRE(h && guide_dev && src_dev && out_dev, "gg_guided_filter: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(Arena::padded((size_t)H * W * 2, 4) + 4096));
  return arena_checked(h->arena, guided_filter_plane(h, h->arena, guide_dev, src_dev, H, W, radius, eps, out_dev,
                                                     (cudaStream_t)stream), "gg_guided_filter");
}

int gg_trimap_path_device(gg_handle h, const uint8_t* bgr_dev, const int32_t* labels_dev, int B, int H, int W,
                          const gg_path_config* cfg, uint8_t* trimap_dev, float* probs_dev,
                          int64_t* node_off_dev, void* stream) {
  GG_REQUIRE(h && bgr_dev && cfg && trimap_dev && (labels_dev || cfg->slic_segments > 0),
             "gg_trimap_path_device: null argument");
  GG_CUDA_OK(cudaSetDevice(h->device));
  if (!h->net.loaded) { set_error("gg_trimap_path_device: call gg_load_weights first"); return GG_ERR_STATE; }
  GG_TRY(h->arena.reserve(path_multi_workspace_bytes(h, B, H, W, *cfg)));
  return arena_checked(h->arena, run_path_multi(h, h->arena, bgr_dev, labels_dev, B, H, W, *cfg, trimap_dev, probs_dev,
                                                node_off_dev, nullptr, nullptr, (cudaStream_t)stream),
                       "gg_trimap_path_device");
}

// Host buffers in, host trimaps out.  The batch is cut into chunks; chunk i+1 is copied in
// (stream s_in) and chunk i-1 copied out (s_out) while chunk i runs (s_run / s_sub[3],
// alternating, so that the latency-bound kernels of two chunks overlap).  Three chunk slots
// (inputs + workspace + trimaps) rotate, across calls as well: a submitted call only enqueues
// work, so the copy-in of call n+1 overlaps the kernels of call n and the host<->device link
// stays busy (one call alone pays the fill and drain of the pipeline).
static int host_submit(gg_handle h, const uint8_t* bgr_host, const void* labels_host_v, int B, int H,
                       int W, const gg_path_config* cfg, uint8_t* trimap_host, int32_t* n_nodes_host,
                       int32_t* n_edges_host, int* ticket, size_t chunk_input_bytes) {
  GG_REQUIRE(h && bgr_host && cfg && trimap_host && ticket && (labels_host_v || cfg->slic_segments > 0),
             "gg_trimap_path_host: null argument");
  const bool device_slic = labels_host_v == nullptr;
  GG_REQUIRE(B > 0 && H >= 2 && W >= 2, "gg_trimap_path_host: bad shape");
  GG_REQUIRE(cfg->label_bytes == 0 || cfg->label_bytes == 4 || cfg->label_bytes == 2,
             "gg_trimap_path_host: label_bytes must be 4 (int32) or 2 (uint16)");
  const size_t lbytes = cfg->label_bytes == 2 ? 2 : 4;
  const char* labels_host = reinterpret_cast<const char*>(labels_host_v);
  GG_CUDA_OK(cudaSetDevice(h->device));
  if (!h->net.loaded) { set_error("gg_trimap_path_host_submit: call gg_load_weights first"); return GG_ERR_STATE; }
  if (h->tickets_open >= gg_context::MAX_TICKETS) {
    set_error("gg_trimap_path_host_submit: %d calls already in flight; wait for one first", gg_context::MAX_TICKETS);
    return GG_ERR_STATE;
  }
  const size_t npx = (size_t)H * W;
  // default chunk: sized by pixels (7 B/px whatever the label transport: the kernels, not the copy,
  // set the efficient size), then balanced so that all chunks of the call are equal
  int chunk = cfg->chunk > 0 ? cfg->chunk : std::max(1, std::min(B, (int)(chunk_input_bytes / (npx * 7) + 1)));
  chunk = std::min(chunk, B);
  if (cfg->chunk <= 0) chunk = (B + (B + chunk - 1) / chunk - 1) / ((B + chunk - 1) / chunk);
  const int n_chunks = (B + chunk - 1) / chunk;
  const int n_slots = 3;
  const size_t in_bytes = Arena::padded((size_t)chunk * npx * 3, 1) + Arena::padded((size_t)chunk * npx, 4) +
                          Arena::padded((size_t)chunk * npx, 1) + (lbytes == 2 ? Arena::padded((size_t)chunk * npx, 2) : 0);
  const size_t slot_bytes = in_bytes + path_workspace_bytes(h, chunk, H, W, *cfg) + 4096;
  if (slot_bytes != h->slot_bytes) {
    // a different chunk geometry: let the calls in flight finish, then lay the slots out anew
    GG_CUDA_OK(cudaDeviceSynchronize());
    GG_TRY(h->host_arena.reserve(slot_bytes * n_slots));
    h->slot_bytes = slot_bytes;
    for (bool& u : h->slot_used) u = false;
  }
  char* base = h->host_arena.base;
  cudaEvent_t* ev_in = &h->ev[0];     // [3] input of slot s landed
  cudaEvent_t* ev_run = &h->ev[3];    // [3] compute of slot s done
  cudaEvent_t* ev_out = &h->ev[6];    // [3] output of slot s copied out
  for (int ci = 0; ci < n_chunks; ++ci) {
    const long long seq = h->chunk_seq++;
    const int s = (int)(seq % n_slots), par = (int)(seq & 1);
    const int b0 = ci * chunk, nb = std::min(chunk, B - b0);
    Arena ar;
    ar.base = base + (size_t)s * slot_bytes;
    ar.cap = slot_bytes;
    uint8_t* d_bgr = ar.take<uint8_t>((size_t)chunk * npx * 3);
    int32_t* d_lab = ar.take<int32_t>((size_t)chunk * npx);
    uint8_t* d_tri = ar.take<uint8_t>((size_t)chunk * npx);
    uint16_t* d_lab16 = lbytes == 2 ? ar.take<uint16_t>((size_t)chunk * npx) : nullptr;
    // the slot is free again once its previous trimaps have been copied out
    if (h->slot_used[s]) GG_CUDA_OK(cudaStreamWaitEvent(h->s_in, ev_out[s], 0));
    h->slot_used[s] = true;
    GG_CUDA_OK(cudaMemcpyAsync(d_bgr, bgr_host + (size_t)b0 * npx * 3, (size_t)nb * npx * 3, cudaMemcpyHostToDevice, h->s_in));
    if (!device_slic)
      GG_CUDA_OK(cudaMemcpyAsync(lbytes == 2 ? (void*)d_lab16 : (void*)d_lab, labels_host + (size_t)b0 * npx * lbytes,
                                 (size_t)nb * npx * lbytes, cudaMemcpyHostToDevice, h->s_in));
    GG_CUDA_OK(cudaEventRecord(ev_in[s], h->s_in));
    cudaStream_t rs = par ? h->s_sub[3] : h->s_run;
    GG_CUDA_OK(cudaStreamWaitEvent(rs, ev_in[s], 0));
    if (lbytes == 2 && !device_slic)
      GG_LAUNCH(h, k_widen_labels, ceil_div((long long)nb * npx, 256 * 8), 256, 0, rs, d_lab16, d_lab, (size_t)nb * npx);
    h->status_word = h->d_status + 2 + par;
    h->rs_direct = (lbytes == 2 || device_slic) ? 1 : 0;   // 7 B/px in: copy-bound, keep the L2 atomics low (see build_graphs)
    int st = run_path(h, ar, d_bgr, device_slic ? nullptr : d_lab, nb, H, W, *cfg, d_tri, nullptr, nullptr,
                      n_nodes_host ? n_nodes_host + b0 : nullptr, n_edges_host ? n_edges_host + b0 : nullptr,
                      rs);
    h->status_word = h->d_status;
    h->rs_direct = 1;
    st = arena_checked(ar, st, "gg_trimap_path_host");
    if (st != GG_OK) { cudaDeviceSynchronize(); return st; }
    // accumulate the per-chunk device status into the sticky word
    GG_LAUNCH(h, k_status_or, 1, 1, 0, rs, h->d_status + 2 + par, h->d_status + 1);
    GG_CUDA_OK(cudaEventRecord(ev_run[s], rs));
    GG_CUDA_OK(cudaStreamWaitEvent(h->s_out, ev_run[s], 0));
    GG_CUDA_OK(cudaMemcpyAsync(trimap_host + (size_t)b0 * npx, d_tri, (size_t)nb * npx, cudaMemcpyDeviceToHost, h->s_out));
    GG_CUDA_OK(cudaEventRecord(ev_out[s], h->s_out));
  }
  // s_out has waited for the compute of every chunk of this call, in order: the sticky status
  // word, as it stands when this call's last chunk is done, travels to the ticket's pinned host
  // word on s_out -- _wait never has to touch a stream that later calls are queued on
  int t = 0;
  while (h->ticket_open[t]) ++t;
  GG_CUDA_OK(cudaMemcpyAsync(h->h_ticket_status + t, h->d_status + 1, sizeof(int), cudaMemcpyDeviceToHost, h->s_out));
  GG_CUDA_OK(cudaEventRecord(h->ticket_ev[t], h->s_out));
  h->ticket_open[t] = true;
  h->tickets_open++;
  *ticket = t;
  return GG_OK;
}

// Blocks until the call that returned `ticket` has delivered its trimaps.  The device status
// is sticky over the calls in flight: a capacity error is reported by the first wait after it.
int gg_trimap_path_host_wait(gg_handle h, int ticket) {
  GG_REQUIRE(h && ticket >= 0 && ticket < gg_context::MAX_TICKETS && h->ticket_open[ticket],
             "gg_trimap_path_host_wait: unknown ticket");
  GG_CUDA_OK(cudaSetDevice(h->device));
  h->ticket_open[ticket] = false;
  h->tickets_open--;
  GG_CUDA_OK(cudaEventSynchronize(h->ticket_ev[ticket]));
  const int dev_bits = h->h_ticket_status[ticket];
  if (dev_bits) {
    GG_CUDA_OK(cudaDeviceSynchronize());
    GG_CUDA_OK(cudaMemset(h->d_status + 1, 0, sizeof(int)));
    set_error("gg_trimap_path_host: device status 0x%x (label >= node_cap or pair capacity exceeded)", dev_bits);
    return GG_ERR_CAPACITY;
  }
  return GG_OK;
}

// Default chunk sizes (cfg->chunk == 0), measured on B200 at 320x480: a streamed call keeps the
// pipeline full across calls and prefers large chunks (kernel efficiency: ~128 images, 138 MB of
// input); a single synchronous call pays the fill and drain itself and prefers ~64 images.
int gg_trimap_path_host_submit(gg_handle h, const uint8_t* bgr_host, const void* labels_host, int B, int H,
                               int W, const gg_path_config* cfg, uint8_t* trimap_host, int32_t* n_nodes_host,
                               int32_t* n_edges_host, int* ticket) {
  return host_submit(h, bgr_host, labels_host, B, H, W, cfg, trimap_host, n_nodes_host, n_edges_host, ticket,
                     (size_t)137 << 20);
}

int gg_trimap_path_host(gg_handle h, const uint8_t* bgr_host, const void* labels_host, int B, int H, int W,
                        const gg_path_config* cfg, uint8_t* trimap_host, int32_t* n_nodes_host,
                        int32_t* n_edges_host) {
  int ticket = -1;
  GG_TRY(host_submit(h, bgr_host, labels_host, B, H, W, cfg, trimap_host, n_nodes_host, n_edges_host, &ticket,
                     (size_t)68 << 20));
  return gg_trimap_path_host_wait(h, ticket);
}

int gg_selftest_math(gg_handle h, int64_t* mismatches) {
  GG_REQUIRE(h && mismatches, "gg_selftest_math: null");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_TRY(h->arena.reserve(4096));
  long long m[4] = {0, 0, 0, 0};
  GG_TRY(selftest_math(h, h->arena, m, nullptr));
  for (int i = 0; i < 4; ++i) mismatches[i] = m[i];
  return GG_OK;
}

int64_t gg_kernel_launch_count(gg_handle h) { return h ? h->launches : 0; }

int gg_profile_enable(gg_handle h, int enable) {
  GG_REQUIRE(h, "gg_profile_enable: null");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_CUDA_OK(cudaDeviceSynchronize());
  h->prof_on = enable != 0;
  h->prof.clear();
  h->prof_pool_used = 0;
  return GG_OK;
}

// Text report "kernel,launches,total_ms\n" per kernel, sorted by total time (descending).
int gg_profile_report(gg_handle h, char* buf, size_t cap) {
  GG_REQUIRE(h && buf && cap > 0, "gg_profile_report: null");
  GG_CUDA_OK(cudaSetDevice(h->device));
  GG_CUDA_OK(cudaDeviceSynchronize());
  struct Acc { const char* name; long n; double ms; };
  std::vector<Acc> acc;
  for (auto& r : h->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    bool found = false;
    for (auto& a : acc)
      if (!strcmp(a.name, r.name)) { a.n++; a.ms += ms; found = true; break; }
    if (!found) acc.push_back({r.name, 1, (double)ms});
  }
  std::sort(acc.begin(), acc.end(), [](const Acc& a, const Acc& b) { return a.ms > b.ms; });
  size_t off = 0;
  buf[0] = 0;
  for (auto& a : acc) {
    int w = snprintf(buf + off, cap - off, "%s,%ld,%.6f\n", a.name, a.n, a.ms);
    if (w < 0 || (size_t)w >= cap - off) break;
    off += (size_t)w;
  }
  return GG_OK;
}

}  // extern "C"
