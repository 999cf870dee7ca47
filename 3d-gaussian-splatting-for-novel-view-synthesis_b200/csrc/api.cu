// C ABI of libb200gs (see include/b200gs.h for the contract and the reference interfaces replaced).
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* what) {
  g_last_error = what;
  return code;
}
int fail(int code, const std::string& what) { return fail(code, what.c_str()); }
int fail_cuda(cudaError_t e, const char* where) {
  g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
  return B200GS_ERR_CUDA;
}
#define CU(expr)                                         \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return fail_cuda(_e, #expr);  \
  } while (0)

// --- optional per-region CUDA-event profiling (bench.py's per-kernel roofline table) ---------------------
enum Region { R_PREPROCESS_FWD = 0, R_DEPTH_SORT, R_SCAN, R_EMIT, R_TILE_SORT, R_SPLIT, R_BLEND_FWD, R_BLEND_BWD,
              R_PREPROCESS_BWD, R_EVAL_SH, R_BUILD_SIGMA, R_EVAL_SH_BWD, R_BUILD_SIGMA_BWD, R_LOSS_FWD, R_LOSS_BWD, R_ADAM, R_CLIP, R_PEER_STEP, R_PEER_ALLREDUCE, R_COMPACT, R_BAND_SELECT, R_ROUTE, R_GATHER, R_BARRIER, R_ROUTE_REC, R_COUNT };
const char* kRegionNames[R_COUNT] = {"preprocess_fwd", "depth_sort", "scan", "emit_super", "super_sort", "split_tiles",
                                     "blend_fwd", "blend_bwd", "preprocess_bwd", "evaluate_sh", "build_sigma",
                                     "evaluate_sh_bwd", "build_sigma_bwd", "l1_ssim_fwd", "l1_ssim_bwd", "adam_step", "clip_grad_norm", "peer_adam_step", "peer_allreduce", "compact_keys", "band_select", "route_slice", "gather_routed", "peer_barrier", "route_records"};
struct ProfRec { int region; cudaEvent_t a, b; };
struct Profiler {
  std::mutex mu;
  bool on = false;
  std::vector<ProfRec> recs;
};
Profiler g_prof;
std::atomic<unsigned long long> g_launches{0};

struct ProfScope {
  int region; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(int region_, cudaStream_t s_, int launches) : region(region_), s(s_) {
    g_launches.fetch_add((unsigned long long)launches, std::memory_order_relaxed);
    if (!g_prof.on) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
    cudaEventRecord(a, s);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, s);
    std::lock_guard<std::mutex> lock(g_prof.mu);
    g_prof.recs.push_back({region, a, b});
  }
};
#define PCU(region, launches, expr)             \
  do {                                          \
    ProfScope _scope(region, s, launches);      \
    CU(expr);                                   \
  } while (0)

bool env_is(const char* name, const char* value) {
  const char* v = getenv(name);
  return v && !strcmp(v, value);
}

int tile_bits(int n_tiles) {
  int b = 1;
  while ((1 << b) < n_tiles) ++b;
  return b;
}

int make_params(const b200gs_camera* cam, gs::RenderParams& rp) {
  if (!cam || !cam->c2w) return fail(B200GS_ERR_ARG, "camera or c2w is null");
  if (cam->tile != B200GS_TILE) return fail(B200GS_ERR_TILE, "only T=16 is supported");
  if (cam->H <= 0 || cam->W <= 0) return fail(B200GS_ERR_ARG, "H and W must be positive");
  if (cam->W > 65535 * 16 || cam->H > 65535 * 16) return fail(B200GS_ERR_ARG, "image too large");
  gs::fill_render_params(rp, cam->H, cam->W, cam->fx, cam->fy, cam->cx, cam->cy, cam->near_plane, cam->far_plane,
                         cam->pix_guard, cam->min_conis, cam->chi_square_clip, cam->alpha_max, cam->alpha_cutoff);
  if (cam->tile_row_end > 0 || cam->tile_row_begin > 0) {
    rp.row_begin = cam->tile_row_begin < 0 ? 0 : cam->tile_row_begin;
    rp.row_end = cam->tile_row_end > rp.tiles_y ? rp.tiles_y : cam->tile_row_end;
    if (rp.row_end < rp.row_begin) rp.row_end = rp.row_begin;
  }
  return B200GS_OK;
}

int make_gauss(const b200gs_gaussians* g, gs::GaussIn& o) {
  if (!g) return fail(B200GS_ERR_ARG, "gaussians is null");
  if (g->n < 0) return fail(B200GS_ERR_ARG, "n < 0");
  if (g->n > 0) {
    if (!g->pos || !g->opacity_raw) return fail(B200GS_ERR_ARG, "pos / opacity_raw missing");
    const bool raw_cov = g->scale_raw && g->q_raw;
    if (!raw_cov && !g->sigma) return fail(B200GS_ERR_ARG, "need (scale_raw, q_raw) or sigma");
    const bool raw_sh = g->f_dc && g->f_rest;
    if (!raw_sh && !g->color) return fail(B200GS_ERR_ARG, "need (f_dc, f_rest) or color");
    if (raw_cov && (reinterpret_cast<uintptr_t>(g->q_raw) & 15u))
      return fail(B200GS_ERR_ARG, "q_raw must be 16-byte aligned");
  }
  o.n = g->n; o.pos = g->pos; o.opacity_raw = g->opacity_raw;
  const bool raw_cov = g->scale_raw && g->q_raw;
  o.scale_raw = raw_cov ? g->scale_raw : nullptr;
  o.q_raw = raw_cov ? g->q_raw : nullptr;
  o.sigma = raw_cov ? nullptr : g->sigma;
  const bool raw_sh = g->f_dc && g->f_rest;
  o.f_dc = raw_sh ? g->f_dc : nullptr;
  o.f_rest = raw_sh ? g->f_rest : nullptr;
  o.color = raw_sh ? nullptr : g->color;
  return B200GS_OK;
}

bool is_band(const gs::RenderParams& rp) { return rp.row_begin > 0 || rp.row_end < rp.tiles_y; }
// frames whose depth-sorted ids are a compacted prefix counted on the device (stats->n_sorted)
bool is_compacted(const gs::RenderParams& rp, const b200gs_camera* cam) {
  return is_band(rp) || (cam->flags & B200GS_CAM_ROUTED) != 0;
}

// Supertile grid of the rows this frame renders: super_x columns, rows [super_y0, super_y0 + super_y).  A band numbers
// its supertiles from its own first row, so that the binning sort handles ceil(log2(#band supertiles)) bits (one radix
// pass instead of two for an eighth of a 4K frame) and the split kernels launch for the band's supertiles only.
int super_dims(const gs::RenderParams& rp, int& super_x, int& super_y0, int& super_y) {
  super_x = gs::ceil_div(rp.tiles_x, gs::kSuperX);
  if (rp.row_end <= rp.row_begin) { super_y0 = 0; super_y = 1; return super_x; }
  super_y0 = rp.row_begin / gs::kSuperY;
  super_y = (rp.row_end - 1) / gs::kSuperY - super_y0 + 1;
  return super_x * super_y;
}

// where the supertile sort leaves its result: pass 0 writes the *_alt buffers, so odd pass counts end there
struct SortedSuper { const uint32_t* keys; const uint32_t* vals; };
SortedSuper sorted_super(void* isect_ws, const gs::IsectLayout& IL, int n_super_tiles) {
  const int passes = gs::sort_passes(0, tile_bits(n_super_tiles)).num;
  SortedSuper s;
  if (passes % 2 == 0) { s.keys = gs::ws_ptr<uint32_t>(isect_ws, IL.keys); s.vals = gs::ws_ptr<uint32_t>(isect_ws, IL.vals); }
  else { s.keys = gs::ws_ptr<uint32_t>(isect_ws, IL.keys_alt); s.vals = gs::ws_ptr<uint32_t>(isect_ws, IL.vals_alt); }
  return s;
}

// --- debug export kernels -------------------------------------------------------------------------
__global__ void export_kernel(int n, const float4* rec0, const float4* rec1, const float4* rec2,
                              const uint32_t* depth_key, const uint2* rect, const uint32_t* radius_in,
                              const uint32_t* super_touched, float* xy,
                              float* depth, float* conic, float* opacity, float* color, int32_t* radius,
                              int32_t* rect_out, int32_t* tiles_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool vis = depth_key[i] != gs::kCulledKey;
  const float4 a = vis ? rec0[i] : make_float4(0, 0, 0, 0), b = vis ? rec1[i] : make_float4(0, 0, 0, 0),
               c = vis ? rec2[i] : make_float4(0, 0, 0, 0);
  const float inv = 1.f / gs::kBlendExpScale;     // the records hold the conic scaled by -log2(e)/2
  if (xy) { xy[2 * i] = a.x; xy[2 * i + 1] = a.y; }
  if (depth) depth[i] = vis ? __uint_as_float(depth_key[i]) : -1.f;
  if (conic) { conic[3 * i] = a.z * inv; conic[3 * i + 1] = 0.5f * a.w * inv; conic[3 * i + 2] = b.x * inv; }
  if (opacity) opacity[i] = vis ? exp2f(b.y) : 0.f;
  if (color) { color[3 * i] = b.w; color[3 * i + 1] = c.x; color[3 * i + 2] = c.y; }
  if (radius) radius[i] = vis ? (int32_t)radius_in[i] : 0;
  const uint2 r = vis ? rect[i] : make_uint2(0, 0);
  if (rect_out) {
    rect_out[4 * i] = r.x & 0xFFFF; rect_out[4 * i + 1] = r.x >> 16;
    rect_out[4 * i + 2] = r.y & 0xFFFF; rect_out[4 * i + 3] = r.y >> 16;
  }
  // tiles touched = area of the (band-clipped) tile rect; 0 when the band clipping left nothing
  if (tiles_out)
    tiles_out[i] = !vis ? -1 : (super_touched[i] ? (int32_t)(((r.x >> 16) - (r.x & 0xFFFF) + 1) * ((r.y >> 16) - (r.y & 0xFFFF) + 1)) : 0);
}

// fp32 [0,1] image -> uint8, exactly what the reference scripts do on the host after the download:
// (img.cpu().numpy() * 255).astype(np.uint8)  (render_trained.py:357, inference.py:117) - multiply in fp32, truncate.
__global__ void image_to_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n && (reinterpret_cast<uintptr_t>(img) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
    const float4 v = *reinterpret_cast<const float4*>(img + i);
    uchar4 o;
    o.x = (uint8_t)(v.x * 255.f); o.y = (uint8_t)(v.y * 255.f); o.z = (uint8_t)(v.z * 255.f); o.w = (uint8_t)(v.w * 255.f);
    *reinterpret_cast<uchar4*>(out + i) = o;
  } else {
    for (size_t k = i; k < n && k < i + 4; ++k) out[k] = (uint8_t)(img[k] * 255.f);
  }
}

__global__ void copy_u32_kernel(const uint32_t* src, int32_t* dst, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (int32_t)src[i];
}
__global__ void copy_ranges_kernel(const uint2* src, int32_t* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { dst[2 * i] = (int32_t)src[i].x; dst[2 * i + 1] = (int32_t)src[i].y; }
}

// --- frame statistics to the host -----------------------------------------------------------------
// A 64-byte cudaMemcpyAsync would queue behind whatever the device-to-host copy engine is doing - typically the
// 25 MB image of the previous frame, for ~0.45 ms - and stall the stream it sits in (measured: the binning half of
// a pipelined frame went from 280 us to 560 us as soon as finished frames were streaming to the host).  Pinned
// host memory is device-addressable (UVA), so a one-warp kernel stores the counters there directly.
__global__ void stats_to_host_kernel(const b200gs_frame_stats* __restrict__ src, volatile uint32_t* __restrict__ dst) {
  if (threadIdx.x < sizeof(b200gs_frame_stats) / 4) dst[threadIdx.x] = reinterpret_cast<const uint32_t*>(src)[threadIdx.x];
  __threadfence_system();
}

cudaError_t publish_stats(const b200gs_frame_stats* stats, b200gs_frame_stats* stats_host, cudaStream_t s) {
  static thread_local const void* last_checked = nullptr;
  static thread_local bool last_mapped = false;
  if (stats_host != last_checked) {
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, stats_host);
    last_mapped = (e == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr);
    if (e != cudaSuccess) cudaGetLastError();
    last_checked = stats_host;
  }
  if (!last_mapped) return cudaMemcpyAsync(stats_host, stats, sizeof(b200gs_frame_stats), cudaMemcpyDeviceToHost, s);
  stats_to_host_kernel<<<1, 32, 0, s>>>(stats, reinterpret_cast<volatile uint32_t*>(stats_host));
  return cudaGetLastError();
}

// --- host-buffer path cache ------------------------------------------------------------------------
struct HostPathCache {
  std::mutex mu;
  void* bufs[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t caps[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t stream = nullptr;
  b200gs_frame_stats* stats_pinned = nullptr;
  cudaError_t ensure(int slot, size_t bytes) {
    if (caps[slot] >= bytes) return cudaSuccess;
    if (bufs[slot]) cudaFree(bufs[slot]);
    bufs[slot] = nullptr; caps[slot] = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&bufs[slot], want);
    if (e == cudaSuccess) caps[slot] = want;
    return e;
  }
};
HostPathCache g_host;

}  // namespace

extern "C" {

int b200gs_abi_version(void) { return B200GS_ABI_VERSION; }
const char* b200gs_last_error(void) { return g_last_error.c_str(); }

int b200gs_workspace_sizes(int32_t n, int32_t H, int32_t W, uint32_t isect_capacity, b200gs_sizes* out) {
  if (!out || n < 0 || H <= 0 || W <= 0) return fail(B200GS_ERR_ARG, "bad arguments to workspace_sizes");
  out->frame_bytes = gs::frame_layout(n, H, W).total;
  out->isect_bytes = gs::isect_layout(isect_capacity).total;
  return B200GS_OK;
}

int b200gs_build_sigma(int32_t n, const float* scale_raw, const float* q_raw, float* sigma_out, void* stream) {
  if (n < 0 || (n > 0 && (!scale_raw || !q_raw || !sigma_out))) return fail(B200GS_ERR_ARG, "build_sigma: null");
  if (reinterpret_cast<uintptr_t>(q_raw) & 15u) return fail(B200GS_ERR_ARG, "q_raw must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_BUILD_SIGMA, 1, gs::launch_build_sigma(n, scale_raw, q_raw, sigma_out, s));
  return B200GS_OK;
}

int b200gs_build_sigma_backward(int32_t n, const float* scale_raw, const float* q_raw, const float* grad_sigma,
                                float* grad_scale_raw, float* grad_q_raw, void* stream) {
  if (n < 0 || (n > 0 && (!scale_raw || !q_raw || !grad_sigma || !grad_scale_raw || !grad_q_raw)))
    return fail(B200GS_ERR_ARG, "build_sigma_backward: null");
  if ((reinterpret_cast<uintptr_t>(q_raw) & 15u) || (reinterpret_cast<uintptr_t>(grad_q_raw) & 15u))
    return fail(B200GS_ERR_ARG, "q_raw / grad_q_raw must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_BUILD_SIGMA_BWD, 1, gs::launch_build_sigma_bwd(n, scale_raw, q_raw, grad_sigma, grad_scale_raw, grad_q_raw, s));
  return B200GS_OK;
}

int b200gs_evaluate_sh(int32_t n, const float* f_dc, const float* f_rest, const float* points, const float* c2w,
                       float* color_out, void* stream) {
  if (n < 0 || (n > 0 && (!f_dc || !f_rest || !points || !c2w || !color_out)))
    return fail(B200GS_ERR_ARG, "evaluate_sh: null");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_EVAL_SH, 1, gs::launch_eval_sh(n, f_dc, f_rest, points, c2w, color_out, s));
  return B200GS_OK;
}

int b200gs_evaluate_sh_backward(int32_t n, const float* f_dc, const float* f_rest, const float* points,
                                const float* c2w, const float* grad_color, float* grad_f_dc, float* grad_f_rest,
                                float* grad_points, void* stream) {
  if (n < 0 || (n > 0 && (!f_dc || !f_rest || !points || !c2w || !grad_color || !grad_f_dc || !grad_f_rest ||
                          !grad_points)))
    return fail(B200GS_ERR_ARG, "evaluate_sh_backward: null");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_EVAL_SH_BWD, 1, gs::launch_eval_sh_bwd(n, f_dc, f_rest, points, c2w, grad_color, grad_f_dc, grad_f_rest, grad_points, s));
  return B200GS_OK;
}

size_t b200gs_loss_workspace_bytes(int32_t n_img, int32_t H, int32_t W, int32_t with_grad) {
  if (n_img <= 0 || H <= 0 || W <= 0) return 0;
  return gs::loss_workspace_bytes(n_img, H, W, with_grad != 0);
}

int b200gs_l1_ssim_forward(const float* pred, const float* target, int32_t n_img, int32_t H, int32_t W,
                           double lambda_l1, double lambda_ssim, void* workspace, size_t workspace_bytes,
                           int32_t with_grad, float* out3, void* stream) {
  if (!pred || !target || !workspace || !out3 || n_img <= 0 || H <= 0 || W <= 0)
    return fail(B200GS_ERR_ARG, "l1_ssim_forward: null or empty argument");
  if (workspace_bytes < gs::loss_workspace_bytes(n_img, H, W, with_grad != 0))
    return fail(B200GS_ERR_WORKSPACE, "l1_ssim_forward: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_LOSS_FWD, 1, gs::launch_l1_ssim_fwd(pred, target, n_img, H, W, (float)lambda_l1, (float)lambda_ssim, workspace,
                                            with_grad != 0, out3, s));
  return B200GS_OK;
}

int b200gs_l1_ssim_backward(const float* pred, const float* target, int32_t n_img, int32_t H, int32_t W,
                            double lambda_l1, double lambda_ssim, const void* workspace, size_t workspace_bytes,
                            const float* grad_total, float* grad_pred, void* stream) {
  if (!pred || !target || !workspace || !grad_pred || n_img <= 0 || H <= 0 || W <= 0)
    return fail(B200GS_ERR_ARG, "l1_ssim_backward: null or empty argument");
  if (workspace_bytes < gs::loss_workspace_bytes(n_img, H, W, true))
    return fail(B200GS_ERR_WORKSPACE, "l1_ssim_backward: workspace too small (forward must run with with_grad)");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_LOSS_BWD, 1, gs::launch_l1_ssim_bwd(pred, target, n_img, H, W, (float)lambda_l1, (float)lambda_ssim, workspace,
                                            grad_total, grad_pred, s));
  return B200GS_OK;
}

int b200gs_image_to_u8(const float* image, uint8_t* out, size_t numel, void* stream) {
  if (numel > 0 && (!image || !out)) return fail(B200GS_ERR_ARG, "image_to_u8: null");
  if (numel == 0) return B200GS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t threads = (numel + 3) / 4;
  image_to_u8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(image, out, numel);
  CU(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return B200GS_OK;
}

int b200gs_adam_step(const b200gs_adam_tensor* tensors, int32_t n_tensors, double beta1, double beta2, double eps,
                     void* stream) {
  if (n_tensors < 0 || (n_tensors > 0 && !tensors)) return fail(B200GS_ERR_ARG, "adam_step: null table");
  for (int i = 0; i < n_tensors; ++i) {
    const b200gs_adam_tensor& t = tensors[i];
    if (t.numel < 0 || t.step < 1) return fail(B200GS_ERR_ARG, "adam_step: numel < 0 or step < 1");
    if (t.numel > 0 && (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq)) return fail(B200GS_ERR_ARG, "adam_step: null tensor");
  }
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_ADAM, 1, gs::launch_adam_step(tensors, n_tensors, beta1, beta2, eps, s));
  return B200GS_OK;
}

size_t b200gs_clip_workspace_bytes(int64_t numel) { return gs::clip_workspace_bytes(numel > 0 ? numel : 0); }

int b200gs_clip_grad_norm(float* grad, int64_t numel, double max_norm, void* workspace, size_t workspace_bytes,
                          float* total_norm_out, void* stream) {
  if (numel < 0 || (numel > 0 && (!grad || !workspace))) return fail(B200GS_ERR_ARG, "clip_grad_norm: null");
  if (numel > 0 && workspace_bytes < gs::clip_workspace_bytes(numel)) return fail(B200GS_ERR_WORKSPACE, "clip_grad_norm: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_CLIP, 2, gs::launch_clip_grad_norm(grad, numel, max_norm, workspace, total_norm_out, s));
  return B200GS_OK;
}

size_t b200gs_clip_workspace_bytes_multi(const int64_t* numel, int32_t n_tensors) {
  if (!numel || n_tensors <= 0) return gs::clip_workspace_bytes(0);
  return gs::clip_workspace_bytes_multi(numel, n_tensors);
}

int b200gs_clip_grad_norm_multi(float* const* grads, const int64_t* numel, int32_t n_tensors, double max_norm,
                                void* workspace, size_t workspace_bytes, float* total_norm_out, void* stream) {
  if (n_tensors < 0 || (n_tensors > 0 && (!grads || !numel))) return fail(B200GS_ERR_ARG, "clip_grad_norm_multi: null table");
  int live = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (numel[i] < 0 || (numel[i] > 0 && !grads[i])) return fail(B200GS_ERR_ARG, "clip_grad_norm_multi: numel < 0 or null tensor");
    live += numel[i] > 0;
  }
  if (live > B200GS_CLIP_MAX_TENSORS) return fail(B200GS_ERR_ARG, "clip_grad_norm_multi: more than B200GS_CLIP_MAX_TENSORS tensors");
  if (live > 0 && (!workspace || workspace_bytes < gs::clip_workspace_bytes_multi(numel, n_tensors)))
    return fail(B200GS_ERR_WORKSPACE, "clip_grad_norm_multi: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_CLIP, 2, gs::launch_clip_grad_norm_multi(grads, numel, n_tensors, max_norm, workspace, total_norm_out, s));
  return B200GS_OK;
}

size_t b200gs_densify_workspace_bytes(int32_t n) { return gs::densify_workspace_bytes(n); }

int b200gs_densify_plan(int32_t n, const float* opacity_raw, const float* scale_raw, const float* pos_grad,
                        double opacity_threshold, double max_grad, double scale_threshold, void* workspace,
                        size_t workspace_bytes, uint32_t* counts_host, void* stream) {
  if (n < 0 || !workspace || (n > 0 && (!opacity_raw || !scale_raw))) return fail(B200GS_ERR_ARG, "densify_plan: null argument");
  if (workspace_bytes < gs::densify_workspace_bytes(n)) return fail(B200GS_ERR_WORKSPACE, "densify_plan: workspace too small");
  CU(gs::launch_densify_plan(n, opacity_raw, scale_raw, pos_grad, opacity_threshold, max_grad, scale_threshold, workspace,
                             counts_host, (cudaStream_t)stream));
  g_launches.fetch_add(n > 0 ? 4 : 0, std::memory_order_relaxed);
  return B200GS_OK;
}

int b200gs_densify_apply(int32_t n, const void* workspace, size_t workspace_bytes, const float* const* in6,
                         float* const* out6, const float* noise, void* stream) {
  if (n < 0 || !workspace || !in6 || !out6) return fail(B200GS_ERR_ARG, "densify_apply: null argument");
  if (workspace_bytes < gs::densify_workspace_bytes(n)) return fail(B200GS_ERR_WORKSPACE, "densify_apply: workspace too small");
  if (n > 0)
    for (int t = 0; t < 6; ++t)
      if (!in6[t] || !out6[t]) return fail(B200GS_ERR_ARG, "densify_apply: null tensor");
  CU(gs::launch_densify_apply(n, workspace, in6, out6, noise, (cudaStream_t)stream));
  g_launches.fetch_add(n > 0 ? 6 : 0, std::memory_order_relaxed);
  return B200GS_OK;
}

int b200gs_peer_layout_compute(const int64_t* numel, int32_t n_tensors, int32_t world, b200gs_peer_layout* out) {
  if (gs::peer_layout_compute(numel, n_tensors, world, out) != 0)
    return fail(B200GS_ERR_ARG, "peer_layout_compute: need 0 <= n_tensors <= 8, 1 <= world <= 16, numel >= 0");
  return B200GS_OK;
}

size_t b200gs_peer_area_bytes(const b200gs_peer_layout* layout) {
  return layout ? (size_t)B200GS_PEER_CTRL_BYTES + 8 * (size_t)layout->flat_total : 0;
}

static int check_peer_group(const b200gs_peer_group* g, const char* who) {
  if (!g || g->world < 1 || g->world > B200GS_MAX_PEERS || g->rank < 0 || g->rank >= g->world)
    return fail(B200GS_ERR_ARG, std::string(who) + ": bad peer group");
  for (int q = 0; q < g->world; ++q)
    if (!g->area[q]) return fail(B200GS_ERR_ARG, std::string(who) + ": unmapped peer area");
  return B200GS_OK;
}

int b200gs_peer_barrier(const b200gs_peer_group* group, uint32_t* epoch, void* stream) {
  if (int rc = check_peer_group(group, "peer_barrier")) return rc;
  if (!epoch) return fail(B200GS_ERR_ARG, "peer_barrier: null epoch");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_BARRIER, 1, gs::launch_peer_barrier(group, ++*epoch, s));
  return B200GS_OK;
}

static int peer_step_common(const b200gs_peer_group* group, const b200gs_peer_layout* layout,
                            const b200gs_peer_tensor* tensors, int32_t n_tensors, bool adam, float* m, float* v,
                            double beta1, double beta2, double eps, double max_norm, int32_t write_grads,
                            uint32_t* epoch, float* total_norm_out, void* stream, const char* who) {
  if (int rc = check_peer_group(group, who)) return rc;
  if (!layout || !epoch || n_tensors < 0 || n_tensors > B200GS_PEER_MAX_TENSORS || (n_tensors > 0 && !tensors))
    return fail(B200GS_ERR_ARG, std::string(who) + ": null argument or too many tensors");
  for (int i = 0; i < n_tensors; ++i) {
    const b200gs_peer_tensor& t = tensors[i];
    if (t.numel < 0 || (adam && t.step < 1)) return fail(B200GS_ERR_ARG, std::string(who) + ": numel < 0 or step < 1");
    if (layout->offset[i] + t.numel > layout->flat_total || (layout->offset[i] & 31))
      return fail(B200GS_ERR_ARG, std::string(who) + ": tensor does not match the layout");
    if (!adam && t.numel > 0 && !t.grad) return fail(B200GS_ERR_ARG, std::string(who) + ": null gradient");
  }
  if (adam && layout->shard_total > 0 && (!m || !v)) return fail(B200GS_ERR_ARG, std::string(who) + ": null moment shard");
  cudaStream_t s = (cudaStream_t)stream;
  int launches = 0;
  {
    ProfScope scope(adam ? R_PEER_STEP : R_PEER_ALLREDUCE, s, 0);
    CU(gs::launch_peer_step(group, layout, tensors, n_tensors, adam, m, v, beta1, beta2, eps, max_norm, write_grads,
                            epoch, total_norm_out, s, &launches));
  }
  g_launches.fetch_add((unsigned long long)launches, std::memory_order_relaxed);
  return B200GS_OK;
}

int b200gs_peer_adam_step(const b200gs_peer_group* group, const b200gs_peer_layout* layout,
                          const b200gs_peer_tensor* tensors, int32_t n_tensors, float* exp_avg_shard,
                          float* exp_avg_sq_shard, double beta1, double beta2, double eps, double max_norm,
                          int32_t write_grads, uint32_t* epoch, float* total_norm_out, void* stream) {
  return peer_step_common(group, layout, tensors, n_tensors, true, exp_avg_shard, exp_avg_sq_shard, beta1, beta2, eps,
                          max_norm, write_grads, epoch, total_norm_out, stream, "peer_adam_step");
}

int b200gs_peer_allreduce(const b200gs_peer_group* group, const b200gs_peer_layout* layout,
                          const b200gs_peer_tensor* tensors, int32_t n_tensors, uint32_t* epoch, void* stream) {
  return peer_step_common(group, layout, tensors, n_tensors, false, nullptr, nullptr, 0.9, 0.999, 1e-8, 0.0, 1, epoch,
                          nullptr, stream, "peer_allreduce");
}

int b200gs_render_project(const b200gs_gaussians* g, const b200gs_camera* cam, void* frame_ws, size_t frame_bytes,
                          b200gs_frame_stats* stats_host, void* stream) {
  gs::RenderParams rp;
  gs::GaussIn gi;
  int rc = make_params(cam, rp);
  if (rc) return rc;
  rc = make_gauss(g, gi);
  if (rc) return rc;
  if (!frame_ws) return fail(B200GS_ERR_ARG, "frame_ws is null");
  const gs::FrameLayout L = gs::frame_layout(gi.n, rp.H, rp.W);
  if (frame_bytes < L.total) return fail(B200GS_ERR_WORKSPACE, "frame workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  b200gs_frame_stats* stats = gs::ws_ptr<b200gs_frame_stats>(frame_ws, L.header);
  CU(cudaMemsetAsync(stats, 0, sizeof(b200gs_frame_stats), s));
  if (gi.n > 0) {
    bool hist_done = false;
    CU(gs::radix_sort_prepare(gs::ws_ptr<void>(frame_ws, L.scratch), L.scratch_bytes, (uint32_t)gi.n, 0, rp.key_bits, s));
    uint32_t* hist = gs::radix_sort_hist(gs::ws_ptr<void>(frame_ws, L.scratch));
    // the sorted (key, id) pairs always end in (sort_key_alt2, order); the radix passes ping-pong between that pair and
    // (sort_key_alt, order_alt), so with an odd number of passes the source of the first pass sits in the second pair
    const gs::DepthKeyPlan kp = gs::depth_key_plan(rp);
    const bool odd = (kp.sp.num & 1) != 0;
    uint32_t* dst_k = gs::ws_ptr<uint32_t>(frame_ws, L.sort_key_alt2);
    uint32_t* dst_v = gs::ws_ptr<uint32_t>(frame_ws, L.order);
    uint32_t* tmp_k = gs::ws_ptr<uint32_t>(frame_ws, L.sort_key_alt);
    uint32_t* tmp_v = gs::ws_ptr<uint32_t>(frame_ws, L.order_alt);
    uint32_t* src_k = odd ? tmp_k : dst_k;         // where a band's compacted keys go (the first pass writes the other pair)
    uint32_t* src_v = odd ? tmp_v : dst_v;
    uint32_t* first_k = odd ? dst_k : tmp_k;
    uint32_t* first_v = odd ? dst_v : tmp_v;
    uint32_t* cand_k = gs::ws_ptr<uint32_t>(frame_ws, L.cand_key);
    uint32_t* cand_v = gs::ws_ptr<uint32_t>(frame_ws, L.cand_id);
    char* band_scratch = gs::ws_ptr<char>(frame_ws, L.band_scratch);
    bool band_route = false;
    if (is_band(rp)) {
      // a band of tile rows, raw parameters: select the candidates with a cheap test, project only those
      // (flag words in `offsets`, unused by the fused binning)
      PCU(R_BAND_SELECT, 2, gs::launch_band_select(gi, cam->c2w, rp, frame_ws, L, gs::ws_ptr<uint32_t>(frame_ws, L.offsets),
                                                   cand_v, &stats->n_candidates, band_scratch, L.band_scratch_half,
                                                   &band_route, s));
      if (band_route)
        PCU(R_PREPROCESS_FWD, 1, gs::launch_band_project(gi, cam->c2w, rp, frame_ws, L, cand_v, &stats->n_candidates, cand_k,
                                                         hist, s));
      hist_done = band_route;
    }
    if (!band_route)
      PCU(R_PREPROCESS_FWD, 1, gs::launch_preprocess_fwd(gi, cam->c2w, rp, frame_ws, L, s, hist, &hist_done));
    // S8: global depth order.  depth_key -> (sort_key_alt2, order) after the radix passes; ties keep index order.
    int in_a = 0;
    if (is_band(rp)) {
      // a band keeps a fraction of the Gaussians: compact the live keys (stable) and sort those only
      if (band_route)
        PCU(R_COMPACT, 1, gs::launch_compact_keys(cand_k, cand_v, (uint32_t)gi.n, &stats->n_candidates, src_k, src_v,
                                                  &stats->n_sorted, band_scratch + L.band_scratch_half,
                                                  L.band_scratch_half, s));
      else
        PCU(R_COMPACT, 1, gs::launch_compact_keys(gs::ws_ptr<uint32_t>(frame_ws, L.depth_key), nullptr, (uint32_t)gi.n,
                                                  nullptr, src_k, src_v, &stats->n_sorted,
                                                  band_scratch + L.band_scratch_half, L.band_scratch_half, s));
      PCU(R_DEPTH_SORT, kp.sp.num + (hist_done ? 0 : 1),
          gs::launch_radix_sort(src_k, src_v, src_k, src_v, first_k, first_v, (uint32_t)gi.n, &stats->n_sorted, 0, rp.key_bits,
                                gs::ws_ptr<void>(frame_ws, L.scratch), L.scratch_bytes, &in_a, s, hist_done, kp.base,
                                kp.max_key, (cam->flags & B200GS_CAM_OVERLAPPED) != 0));
    } else {
      PCU(R_DEPTH_SORT, kp.sp.num + (hist_done ? 0 : 1),
          gs::launch_radix_sort(gs::ws_ptr<uint32_t>(frame_ws, L.depth_key), nullptr, src_k, src_v, first_k, first_v,
                                (uint32_t)gi.n, nullptr, 0, rp.key_bits, gs::ws_ptr<void>(frame_ws, L.scratch),
                                L.scratch_bytes, &in_a, s, hist_done, kp.base, kp.max_key,
                                (cam->flags & B200GS_CAM_OVERLAPPED) != 0));
    }
    if ((in_a != 0) != !odd) return fail(B200GS_ERR_ARG, "internal: depth sort result buffer");
  }
  if (stats_host) CU(publish_stats(stats, stats_host, s));
  return B200GS_OK;
}

static_assert(sizeof(b200gs_route) == 12 + 17 * 4 + 16 * 8 + 8 + 8, "b200gs_route layout (b200gs/_lib.py::Route mirrors it)");

static int check_route(const b200gs_route* r, const char* who) {
  if (!r || r->world < 1 || r->world > B200GS_MAX_PEERS || r->rank < 0 || r->rank >= r->world || r->seg_capacity == 0)
    return fail(B200GS_ERR_ARG, std::string(who) + ": bad route (world, rank or seg_capacity)");
  if ((unsigned long long)r->world * r->seg_capacity > 0x7FFFFFFFull)
    return fail(B200GS_ERR_ARG, std::string(who) + ": world * seg_capacity does not fit 31 bits");
  for (int q = 0; q < r->world; ++q) {
    if (!r->band_ws[q]) return fail(B200GS_ERR_ARG, std::string(who) + ": unmapped band workspace");
    if (r->band_row[q + 1] < r->band_row[q] || r->band_row[q] < 0)
      return fail(B200GS_ERR_ARG, std::string(who) + ": band rows must be non-decreasing");
  }
  return B200GS_OK;
}

int b200gs_route_project_slice(const b200gs_gaussians* g_slice, const b200gs_camera* cam, void* slice_ws,
                               size_t slice_bytes, const b200gs_route* route, void* stream) {
  gs::RenderParams rp;
  gs::GaussIn gi;
  int rc = make_params(cam, rp);
  if (rc) return rc;
  rc = make_gauss(g_slice, gi);
  if (rc) return rc;
  if ((rc = check_route(route, "route_project_slice"))) return rc;
  if (is_band(rp)) return fail(B200GS_ERR_ARG, "route_project_slice: the camera must describe the full frame");
  if ((uint32_t)gi.n > route->seg_capacity) return fail(B200GS_ERR_CAPACITY, "route_project_slice: slice larger than a segment");
  if (!slice_ws) return fail(B200GS_ERR_ARG, "slice_ws is null");
  const gs::FrameLayout SL = gs::frame_layout(gi.n, rp.H, rp.W);
  const gs::FrameLayout BL = gs::frame_layout((int)(route->world * route->seg_capacity), rp.H, rp.W);
  if (slice_bytes < SL.total) return fail(B200GS_ERR_WORKSPACE, "slice workspace too small");
  if (route->band_ws_bytes < BL.total) return fail(B200GS_ERR_WORKSPACE, "band workspace too small");
  // the counters of the routing pass live in the slice workspace's (unused) sort sections [order, grad_acc)
  const size_t scratch_bytes = SL.grad_acc - SL.order;
  if (scratch_bytes < gs::route_scratch_bytes(gi.n, route->world)) return fail(B200GS_ERR_WORKSPACE, "route scratch");
  cudaStream_t s = (cudaStream_t)stream;
  CU(cudaMemsetAsync(gs::ws_ptr<b200gs_frame_stats>(slice_ws, SL.header), 0, sizeof(b200gs_frame_stats), s));
  if (gi.n > 0)       // an empty slice (more ranks than 32-entry groups) still reports its zero totals to every band
    PCU(R_PREPROCESS_FWD, 1, gs::launch_preprocess_fwd(gi, cam->c2w, rp, slice_ws, SL, s, nullptr, nullptr));
  const bool split = (route->flags & B200GS_ROUTE_RECORDS_LATER) != 0;
  PCU(R_ROUTE, 3, gs::launch_route_slice(gi.n, slice_ws, SL, route, BL, gs::ws_ptr<void>(slice_ws, SL.order),
                                         scratch_bytes, split ? gs::kRouteMeta : gs::kRouteAll, s));
  return B200GS_OK;
}

int b200gs_route_records(int32_t n_slice, const b200gs_camera* cam, void* slice_ws, size_t slice_bytes,
                         const b200gs_route* route, void* stream) {
  gs::RenderParams rp;
  int rc = make_params(cam, rp);
  if (rc) return rc;
  if ((rc = check_route(route, "route_records"))) return rc;
  if (n_slice < 0 || (uint32_t)n_slice > route->seg_capacity) return fail(B200GS_ERR_ARG, "route_records: bad slice length");
  if (!slice_ws) return fail(B200GS_ERR_ARG, "slice_ws is null");
  const gs::FrameLayout SL = gs::frame_layout(n_slice, rp.H, rp.W);
  const gs::FrameLayout BL = gs::frame_layout((int)(route->world * route->seg_capacity), rp.H, rp.W);
  if (slice_bytes < SL.total) return fail(B200GS_ERR_WORKSPACE, "slice workspace too small");
  if (route->band_ws_bytes < BL.total) return fail(B200GS_ERR_WORKSPACE, "band workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_ROUTE_REC, 1, gs::launch_route_slice(n_slice, slice_ws, SL, route, BL,
                                             gs::ws_ptr<void>(slice_ws, SL.order), SL.grad_acc - SL.order,
                                             gs::kRouteRecords, s));
  return B200GS_OK;
}

int b200gs_render_project_routed(const b200gs_camera* cam, const b200gs_route* route, void* frame_ws,
                                 size_t frame_bytes, b200gs_frame_stats* stats_host, void* stream) {
  gs::RenderParams rp;
  int rc = make_params(cam, rp);
  if (rc) return rc;
  if ((rc = check_route(route, "render_project_routed"))) return rc;
  if (!(cam->flags & B200GS_CAM_ROUTED)) return fail(B200GS_ERR_ARG, "render_project_routed: camera without B200GS_CAM_ROUTED");
  if (!frame_ws) return fail(B200GS_ERR_ARG, "frame_ws is null");
  const int n = (int)(route->world * route->seg_capacity);
  const gs::FrameLayout L = gs::frame_layout(n, rp.H, rp.W);
  if (frame_bytes < L.total) return fail(B200GS_ERR_WORKSPACE, "frame workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  b200gs_frame_stats* stats = gs::ws_ptr<b200gs_frame_stats>(frame_ws, L.header);
  CU(cudaMemsetAsync(stats, 0, sizeof(b200gs_frame_stats), s));
  CU(gs::radix_sort_prepare(gs::ws_ptr<void>(frame_ws, L.scratch), L.scratch_bytes, (uint32_t)n, 0, rp.key_bits, s));
  uint32_t* hist = gs::radix_sort_hist(gs::ws_ptr<void>(frame_ws, L.scratch));
  const gs::DepthKeyPlan kp = gs::depth_key_plan(rp);
  const bool odd = (kp.sp.num & 1) != 0;            // see b200gs_render_project: the result ends in (sort_key_alt2, order)
  uint32_t* dst_k = gs::ws_ptr<uint32_t>(frame_ws, L.sort_key_alt2);
  uint32_t* dst_v = gs::ws_ptr<uint32_t>(frame_ws, L.order);
  uint32_t* tmp_k = gs::ws_ptr<uint32_t>(frame_ws, L.sort_key_alt);
  uint32_t* tmp_v = gs::ws_ptr<uint32_t>(frame_ws, L.order_alt);
  uint32_t* src_k = odd ? tmp_k : dst_k;
  uint32_t* src_v = odd ? tmp_v : dst_v;
  uint32_t* first_k = odd ? dst_k : tmp_k;
  uint32_t* first_v = odd ? dst_v : tmp_v;
  PCU(R_GATHER, 1, gs::launch_gather_routed(route->world, route->seg_capacity, frame_ws, L, src_k, src_v, hist, kp, s));
  int in_a = 0;
  PCU(R_DEPTH_SORT, kp.sp.num,
      gs::launch_radix_sort(src_k, src_v, src_k, src_v, first_k, first_v, (uint32_t)n, &stats->n_sorted, 0, rp.key_bits,
                            gs::ws_ptr<void>(frame_ws, L.scratch), L.scratch_bytes, &in_a, s, /*hist_ready=*/true, kp.base,
                            kp.max_key, (cam->flags & B200GS_CAM_OVERLAPPED) != 0));
  if ((in_a != 0) != !odd) return fail(B200GS_ERR_ARG, "internal: depth sort result buffer");
  if (stats_host) CU(publish_stats(stats, stats_host, s));
  return B200GS_OK;
}

int b200gs_render_rasterize(const b200gs_camera* cam, int32_t n, void* frame_ws, size_t frame_bytes, void* isect_ws,
                            size_t isect_bytes, uint32_t isect_capacity, float* image_out,
                            b200gs_frame_stats* stats_host, void* stream) {
  return b200gs_render_rasterize_ev(cam, n, frame_ws, frame_bytes, isect_ws, isect_bytes, isect_capacity, image_out,
                                    stats_host, nullptr, stream);
}

int b200gs_render_rasterize_ev(const b200gs_camera* cam, int32_t n, void* frame_ws, size_t frame_bytes, void* isect_ws,
                               size_t isect_bytes, uint32_t isect_capacity, float* image_out,
                               b200gs_frame_stats* stats_host, void* stats_event, void* stream) {
  return b200gs_render_rasterize_split(cam, n, frame_ws, frame_bytes, isect_ws, isect_bytes, isect_capacity, image_out,
                                       stats_host, stats_event, stream, stream);
}

int b200gs_render_rasterize_split(const b200gs_camera* cam, int32_t n, void* frame_ws, size_t frame_bytes, void* isect_ws,
                                  size_t isect_bytes, uint32_t isect_capacity, float* image_out,
                                  b200gs_frame_stats* stats_host, void* stats_event, void* stream,
                                  void* blend_stream) {
  gs::RenderParams rp;
  int rc = make_params(cam, rp);
  if (rc) return rc;
  if (!frame_ws || !image_out || n < 0) return fail(B200GS_ERR_ARG, "rasterize: null argument");
  const gs::FrameLayout L = gs::frame_layout(n, rp.H, rp.W);
  const gs::IsectLayout IL = gs::isect_layout(isect_capacity);
  if (frame_bytes < L.total) return fail(B200GS_ERR_WORKSPACE, "frame workspace too small");
  if (!isect_ws || isect_bytes < IL.total) return fail(B200GS_ERR_WORKSPACE, "isect workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  b200gs_frame_stats* stats = gs::ws_ptr<b200gs_frame_stats>(frame_ws, L.header);
  int super_x, super_y0, super_y;
  const int n_super_tiles = super_dims(rp, super_x, super_y0, super_y);
  uint32_t* keys = gs::ws_ptr<uint32_t>(isect_ws, IL.keys);
  uint32_t* vals = gs::ws_ptr<uint32_t>(isect_ws, IL.vals);
  uint32_t* lists = gs::ws_ptr<uint32_t>(isect_ws, IL.lists);
  // a frame that overflowed a speculative capacity is rasterized again with exact buffers: start clean
  CU(cudaMemsetAsync(&stats->overflow, 0, sizeof(uint32_t), s));
  if (is_compacted(rp, cam))     // the split kernels only visit the band's supertiles: every other tile has an empty list
    CU(cudaMemsetAsync(gs::ws_ptr<uint2>(frame_ws, L.ranges), 0, (size_t)rp.tiles_x * rp.tiles_y * sizeof(uint2), s));
  PCU(R_EMIT, 1, gs::launch_scan_emit_super(n, is_compacted(rp, cam) ? &stats->n_sorted : nullptr, gs::ws_ptr<uint32_t>(frame_ws, L.order), gs::ws_ptr<uint32_t>(frame_ws, L.super_touched),
                                            gs::ws_ptr<uint2>(frame_ws, L.rect), super_x, super_y0, isect_capacity, keys, vals, stats,
                                            tile_bits(n_super_tiles), gs::ws_ptr<void>(isect_ws, IL.scratch), IL.scratch_bytes,
                                            gs::ws_ptr<void>(frame_ws, L.scratch), L.scratch_bytes, s));
  // every counter (I, V, pair count, overflow) is final here: hand them to the host now, so that it can
  // decide about a capacity overflow while the rest of the frame is still running
  if (stats_host) CU(publish_stats(stats, stats_host, s));
  if (stats_event) CU(cudaEventRecord((cudaEvent_t)stats_event, s));
  int in_a = 0;
  PCU(R_TILE_SORT, gs::sort_passes(0, tile_bits(n_super_tiles)).num,
      gs::launch_radix_sort(keys, vals, keys, vals, gs::ws_ptr<uint32_t>(isect_ws, IL.keys_alt),
                            gs::ws_ptr<uint32_t>(isect_ws, IL.vals_alt), isect_capacity, &stats->n_super, 0,
                            tile_bits(n_super_tiles), gs::ws_ptr<void>(isect_ws, IL.scratch), IL.scratch_bytes, &in_a, s,
                            /*hist_ready=*/true, 0, 0xFFFFFFFFu, (cam->flags & B200GS_CAM_OVERLAPPED) != 0));
  const SortedSuper ss = sorted_super(isect_ws, IL, n_super_tiles);
  {
    ProfScope _scope(R_SPLIT, s, 2);
    CU(gs::launch_split_super(false, ss.keys, ss.vals, gs::ws_ptr<uint2>(frame_ws, L.rect), isect_capacity, stats, super_x,
                              super_y0, super_y, rp.tiles_x, rp.tiles_y, gs::ws_ptr<uint32_t>(frame_ws, L.tile_count),
                              gs::ws_ptr<uint2>(frame_ws, L.ranges), lists, s));
    CU(gs::launch_split_super(true, ss.keys, ss.vals, gs::ws_ptr<uint2>(frame_ws, L.rect), isect_capacity, stats, super_x,
                              super_y0, super_y, rp.tiles_x, rp.tiles_y, gs::ws_ptr<uint32_t>(frame_ws, L.tile_count),
                              gs::ws_ptr<uint2>(frame_ws, L.ranges), lists, s));
  }
  // pixels of tiles outside this rank's band are not touched; the whole image is zeroed first so that
  // "pixels in empty tiles stay 0" (render.py:318) also holds for bands
  // frame pipelining: the blend may run on another stream (ordered after the binning by an event), so that the
  // binning of the NEXT frame - queued on `stream` right away - overlaps it
  if (blend_stream != stream) {
    cudaEvent_t binned;
    CU(cudaEventCreateWithFlags(&binned, cudaEventDisableTiming));
    CU(cudaEventRecord(binned, s));
    CU(cudaStreamWaitEvent((cudaStream_t)blend_stream, binned, 0));
    CU(cudaEventDestroy(binned));      // released once the recorded work completes
    s = (cudaStream_t)blend_stream;
  }
  if (!is_band(rp) || (cam->flags & B200GS_CAM_KEEP_OUTSIDE_BAND)) {
    // every pixel is written by the blend kernel / the pixels outside the band belong to somebody else
  } else {
    CU(cudaMemsetAsync(image_out, 0, (size_t)rp.H * rp.W * 3 * sizeof(float), s));
  }
  PCU(R_BLEND_FWD, 1, gs::launch_blend_fwd(rp, frame_ws, L, lists, image_out, s, blend_stream != stream,
                                           (cam->flags & B200GS_CAM_KEEP_OUTSIDE_BAND) != 0 && !env_is("B200GS_ROW_STORES", "0")));
  return B200GS_OK;
}

int b200gs_render_backward(const b200gs_gaussians* g, const b200gs_camera* cam, void* frame_ws, size_t frame_bytes,
                           void* isect_ws, size_t isect_bytes, uint32_t isect_capacity, const float* grad_image,
                           const b200gs_grads* grads, void* stream) {
  gs::RenderParams rp;
  gs::GaussIn gi;
  int rc = make_params(cam, rp);
  if (rc) return rc;
  rc = make_gauss(g, gi);
  if (rc) return rc;
  if (!frame_ws || !isect_ws || !grad_image || !grads) return fail(B200GS_ERR_ARG, "backward: null argument");
  const gs::FrameLayout L = gs::frame_layout(gi.n, rp.H, rp.W);
  const gs::IsectLayout IL = gs::isect_layout(isect_capacity);
  if (frame_bytes < L.total) return fail(B200GS_ERR_WORKSPACE, "frame workspace too small");
  if (isect_bytes < IL.total) return fail(B200GS_ERR_WORKSPACE, "isect workspace too small");
  gs::GaussGrad gg;
  gg.pos = grads->pos; gg.opacity_raw = grads->opacity_raw; gg.scale_raw = grads->scale_raw; gg.q_raw = grads->q_raw;
  gg.sigma = grads->sigma; gg.f_dc = grads->f_dc; gg.f_rest = grads->f_rest; gg.color = grads->color;
  if (gi.n > 0) {
    if (!gg.pos || !gg.opacity_raw) return fail(B200GS_ERR_ARG, "backward: grads.pos / opacity_raw missing");
    if (gi.scale_raw ? (!gg.scale_raw || !gg.q_raw) : !gg.sigma) return fail(B200GS_ERR_ARG, "backward: covariance grads missing");
    if (gi.f_dc ? (!gg.f_dc || !gg.f_rest) : !gg.color) return fail(B200GS_ERR_ARG, "backward: colour grads missing");
    if (gi.scale_raw && (reinterpret_cast<uintptr_t>(gg.q_raw) & 15u))
      return fail(B200GS_ERR_ARG, "grads.q_raw must be 16-byte aligned");
  }
  cudaStream_t s = (cudaStream_t)stream;
  PCU(R_BLEND_BWD, 1, gs::launch_blend_bwd(rp, frame_ws, L, gs::ws_ptr<uint32_t>(isect_ws, IL.lists), grad_image, gi.n, s));
  PCU(R_PREPROCESS_BWD, 1, gs::launch_preprocess_bwd(gi, gg, cam->c2w, rp, frame_ws, L, s));
  return B200GS_OK;
}

int b200gs_render_host(const b200gs_gaussians* gh, const b200gs_camera* cam, const float* c2w_host, float* image_host,
                       b200gs_frame_stats* stats_out) {
  if (!gh || !cam || !c2w_host || !image_host) return fail(B200GS_ERR_ARG, "render_host: null argument");
  std::lock_guard<std::mutex> lock(g_host.mu);
  if (!g_host.stream) CU(cudaStreamCreateWithFlags(&g_host.stream, cudaStreamNonBlocking));
  if (!g_host.stats_pinned) CU(cudaMallocHost(&g_host.stats_pinned, sizeof(b200gs_frame_stats)));
  cudaStream_t s = g_host.stream;
  const size_t n = (size_t)(gh->n > 0 ? gh->n : 0);
  b200gs_gaussians gd;
  memset(&gd, 0, sizeof(gd));
  gd.n = gh->n;
  struct Up { const float* src; const float** dst; size_t floats; int slot; };
  const Up ups[] = {
      {gh->pos, &gd.pos, n * 3, 0},         {gh->opacity_raw, &gd.opacity_raw, n, 1},
      {gh->scale_raw, &gd.scale_raw, n * 3, 2}, {gh->q_raw, &gd.q_raw, n * 4, 3},
      {gh->sigma, &gd.sigma, n * 9, 4},     {gh->f_dc, &gd.f_dc, n * 3, 5},
      {gh->f_rest, &gd.f_rest, n * 45, 6},  {gh->color, &gd.color, n * 3, 7},
  };
  for (const Up& u : ups) {
    if (!u.src || u.floats == 0) continue;
    CU(g_host.ensure(u.slot, u.floats * 4));
    CU(cudaMemcpyAsync(g_host.bufs[u.slot], u.src, u.floats * 4, cudaMemcpyHostToDevice, s));
    *u.dst = reinterpret_cast<const float*>(g_host.bufs[u.slot]);
  }
  CU(g_host.ensure(8, 64));
  CU(cudaMemcpyAsync(g_host.bufs[8], c2w_host, 64, cudaMemcpyHostToDevice, s));
  b200gs_camera cd = *cam;
  cd.c2w = reinterpret_cast<const float*>(g_host.bufs[8]);
  b200gs_sizes sz;
  int rc = b200gs_workspace_sizes(gd.n, cd.H, cd.W, 0, &sz);
  if (rc) return rc;
  CU(g_host.ensure(9, sz.frame_bytes));
  rc = b200gs_render_project(&gd, &cd, g_host.bufs[9], g_host.caps[9], g_host.stats_pinned, s);
  if (rc) return rc;
  CU(cudaStreamSynchronize(s));
  const uint32_t isect = g_host.stats_pinned->n_isect;
  rc = b200gs_workspace_sizes(gd.n, cd.H, cd.W, isect, &sz);
  if (rc) return rc;
  CU(g_host.ensure(10, sz.isect_bytes));
  const size_t img_bytes = (size_t)cd.H * cd.W * 3 * sizeof(float);
  CU(g_host.ensure(11, img_bytes));
  rc = b200gs_render_rasterize(&cd, gd.n, g_host.bufs[9], g_host.caps[9], g_host.bufs[10], g_host.caps[10], isect,
                               reinterpret_cast<float*>(g_host.bufs[11]), g_host.stats_pinned, s);
  if (rc) return rc;
  CU(cudaMemcpyAsync(image_host, g_host.bufs[11], img_bytes, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (stats_out) *stats_out = *g_host.stats_pinned;
  if (g_host.stats_pinned->overflow) return fail(B200GS_ERR_CAPACITY, "intersection list overflow");
  return B200GS_OK;
}

int b200gs_debug_export(int32_t n, const void* frame_ws, size_t frame_bytes, int32_t H, int32_t W, float* xy,
                        float* depth, float* conic, float* opacity, float* color, int32_t* radius, int32_t* rect,
                        int32_t* tiles_touched, int32_t* depth_order, void* stream) {
  if (!frame_ws || n < 0) return fail(B200GS_ERR_ARG, "debug_export: null");
  const gs::FrameLayout L = gs::frame_layout(n, H, W);
  if (frame_bytes < L.total) return fail(B200GS_ERR_WORKSPACE, "frame workspace too small");
  if (n == 0) return B200GS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  export_kernel<<<gs::ceil_div(n, 256), 256, 0, s>>>(
      n, gs::ws_ptr<float4>(frame_ws, L.rec0), gs::ws_ptr<float4>(frame_ws, L.rec1), gs::ws_ptr<float4>(frame_ws, L.rec2),
      gs::ws_ptr<uint32_t>(frame_ws, L.depth_key), gs::ws_ptr<uint2>(frame_ws, L.rect),
      gs::ws_ptr<uint32_t>(frame_ws, L.radius), gs::ws_ptr<uint32_t>(frame_ws, L.super_touched), xy, depth, conic, opacity, color, radius, rect, tiles_touched);
  CU(cudaGetLastError());
  if (depth_order) {
    copy_u32_kernel<<<gs::ceil_div(n, 256), 256, 0, s>>>(gs::ws_ptr<uint32_t>(frame_ws, L.order), depth_order, (uint32_t)n);
    CU(cudaGetLastError());
  }
  return B200GS_OK;
}

int b200gs_debug_export_lists(const void* frame_ws, size_t frame_bytes, const void* isect_ws, size_t isect_bytes,
                              uint32_t isect_capacity, int32_t n, int32_t H, int32_t W, int32_t* list_tile,
                              int32_t* list_id, uint32_t count, int32_t* ranges, void* stream) {
  if (!frame_ws || !isect_ws) return fail(B200GS_ERR_ARG, "debug_export_lists: null");
  const gs::FrameLayout L = gs::frame_layout(n, H, W);
  const gs::IsectLayout IL = gs::isect_layout(isect_capacity);
  if (frame_bytes < L.total || isect_bytes < IL.total) return fail(B200GS_ERR_WORKSPACE, "workspace too small");
  if (count > isect_capacity) return fail(B200GS_ERR_ARG, "count > capacity");
  cudaStream_t s = (cudaStream_t)stream;
  const int n_tiles = gs::ceil_div(W, gs::kTile) * gs::ceil_div(H, gs::kTile);
  if (count) {
    if (list_tile) CU(gs::launch_fill_list_tiles(gs::ws_ptr<uint2>(frame_ws, L.ranges), n_tiles, count, list_tile, s));
    if (list_id) copy_u32_kernel<<<(count + 255) / 256, 256, 0, s>>>(gs::ws_ptr<uint32_t>(isect_ws, IL.lists), list_id, count);
  }
  if (ranges) copy_ranges_kernel<<<gs::ceil_div(n_tiles, 256), 256, 0, s>>>(gs::ws_ptr<uint2>(frame_ws, L.ranges), ranges, n_tiles);
  CU(cudaGetLastError());
  return B200GS_OK;
}

int b200gs_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (ProfRec& r : g_prof.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.recs.clear();
  g_prof.on = on != 0;
  return B200GS_OK;
}

int b200gs_profile_collect(float* ms_out, int32_t* calls_out, int32_t max_regions) {
  if (!ms_out || !calls_out || max_regions < R_COUNT) return fail(B200GS_ERR_ARG, "profile_collect: need room for all regions");
  CU(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (int i = 0; i < max_regions; ++i) { ms_out[i] = 0.f; calls_out[i] = 0; }
  for (ProfRec& r : g_prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { ms_out[r.region] += ms; calls_out[r.region] += 1; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.recs.clear();
  return R_COUNT;
}

const char* b200gs_profile_region_name(int32_t id) { return (id >= 0 && id < R_COUNT) ? kRegionNames[id] : ""; }
unsigned long long b200gs_kernel_launch_count(void) { return g_launches.load(); }

size_t b200gs_scan_scratch_bytes(uint32_t n) { return gs::scan_scratch_bytes(n); }
size_t b200gs_sort_scratch_bytes(uint32_t n) { return gs::sort_scratch_bytes(n); }

int b200gs_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total_out, void* scratch,
                              size_t scratch_bytes, void* stream) {
  if (n > 0 && (!in || !out || !scratch)) return fail(B200GS_ERR_ARG, "scan: null");
  if (scratch_bytes < gs::scan_scratch_bytes(n)) return fail(B200GS_ERR_WORKSPACE, "scan scratch too small");
  CU(gs::launch_exclusive_scan(in, nullptr, out, n, total_out, scratch, scratch_bytes, (cudaStream_t)stream));
  return B200GS_OK;
}

int b200gs_radix_sort_pairs(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, uint32_t n,
                            int begin_bit, int end_bit, void* scratch, size_t scratch_bytes, void* stream) {
  if (n > 0 && (!keys_in || !vals_in || !keys_out || !vals_out || !scratch)) return fail(B200GS_ERR_ARG, "sort: null");
  if (end_bit <= begin_bit || end_bit > 32 || begin_bit < 0) return fail(B200GS_ERR_ARG, "sort: bad bit range");
  if (scratch_bytes < gs::sort_scratch_bytes(n)) return fail(B200GS_ERR_WORKSPACE, "sort scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  int in_a = 0;
  // ping-pong a = *_in, b = *_out: odd pass counts end in *_out, even ones in *_in (then copy over)
  CU(gs::launch_radix_sort(keys_in, vals_in, keys_in, vals_in, keys_out, vals_out, n, nullptr, begin_bit, end_bit,
                           scratch, scratch_bytes, &in_a, s));
  if (in_a && n) {
    CU(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  }
  return B200GS_OK;
}

}  // extern "C"
