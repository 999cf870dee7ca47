/*
 * b200gs - C ABI of the B200-native (sm_100a) differentiable Gaussian-splat rasterizer.
 *
 * This is the drop-in boundary for the reference's render path.  The reference
 * (ashu1069/3D-Gaussian-Splatting-for-Novel-View-Synthesis) has no FFI: the path sits behind three
 * Python callables.  Each entry point below names the reference interface it replaces:
 *
 *   gaussian_splatting/gaussian.py:71            build_sigma_from_params(scale_raw, q_raw)
 *   gaussian_splatting/spherical_harmonics.py:70 evaluate_sh(f_dc, f_rest, points, c2w)
 *   gaussian_splatting/render.py:62-64           render(pos, color, opacity_raw, sigma, c2w, H, W,
 *                                                       fx, fy, cx, cy, near, far, pix_guard, T,
 *                                                       min_conis, chi_square_clip, alpha_max,
 *                                                       alpha_cutoff)
 *   (autograd of the three above)                scripts/train.py:530  total_loss.backward()
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all buffers are caller-owned (the Python host allocates them with torch's caching allocator);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream unless
 *     stated otherwise;
 *   - return value: 0 = OK, negative = B200GS_ERR_* (no exceptions cross this boundary);
 *   - fp32 arithmetic throughout ("dtype": "f32"); tile size is fixed at 16x16 (the reference's
 *     callers never pass another T).
 */
#ifndef B200GS_H
#define B200GS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GS_ABI_VERSION 1
#define B200GS_TILE 16

enum {
  B200GS_OK = 0,
  B200GS_ERR_ARG = -1,       /* null / inconsistent argument */
  B200GS_ERR_WORKSPACE = -2, /* workspace too small */
  B200GS_ERR_CUDA = -3,      /* a CUDA runtime call failed; see b200gs_last_error() */
  B200GS_ERR_TILE = -4,      /* T != 16 */
  B200GS_ERR_CAPACITY = -5   /* intersection list did not fit isect_capacity (result invalid) */
};

/* The Gaussian set.  `pos` and `opacity_raw` are always required.  Shape/colour come either from
 * the raw parameters (fused path: scale_raw+q_raw and/or f_dc+f_rest non-null) or precomputed
 * (sigma / color non-null) exactly as the reference's render() receives them. */
typedef struct b200gs_gaussians {
  int32_t n;
  const float* pos;         /* [n,3] */
  const float* opacity_raw; /* [n]   */
  const float* scale_raw;   /* [n,3]  log-scales            (or NULL -> use sigma) */
  const float* q_raw;       /* [n,4]  quaternion (x,y,z,w)  (or NULL -> use sigma) */
  const float* sigma;       /* [n,3,3] precomputed covariance (used when scale_raw==NULL) */
  const float* f_dc;        /* [n,3]  SH band 0             (or NULL -> use color) */
  const float* f_rest;      /* [n,45] SH bands 1-3, channel-major (or NULL -> use color) */
  const float* color;       /* [n,3]  precomputed RGB (used when f_dc==NULL) */
} b200gs_gaussians;

/* Camera + the keyword arguments of render.py:62-64 (doubles, because the reference receives Python
 * floats and casts them to fp32 only at the point of use; the library does the same). */
typedef struct b200gs_camera {
  const float* c2w; /* DEVICE pointer, 16 floats row-major (render.py: c2w is a device tensor) */
  int32_t H, W;
  double fx, fy, cx, cy;
  double near_plane, far_plane, pix_guard;
  double min_conis, chi_square_clip, alpha_max, alpha_cutoff;
  int32_t tile;          /* must be 16 */
  int32_t tile_row_begin; /* render only tile rows [begin, end) - tile-row sharding; 0,0 = all */
  int32_t tile_row_end;
  int32_t flags;         /* B200GS_CAM_* bits */
} b200gs_camera;
/* A band normally zeroes the pixels outside its tile rows (render.py:318: untouched pixels are 0).  With this bit
 * the rasterizer writes the band's pixels only and leaves the rest of image_out alone - image_out may then be the
 * frame buffer of another GPU (peer-mapped memory), into which every rank stores its own band. */
#define B200GS_CAM_KEEP_OUTSIDE_BAND 1
/* The frame's binning kernels are queued beside other work of the caller (frame pipelining: the previous frame's blend
 * is still running): the radix sorts then use their small block shape (256 threads, 20 K registers, 38 KB of shared
 * memory), which finds room on a busy SM, instead of one large block per SM. */
#define B200GS_CAM_OVERLAPPED 2
/* The frame workspace already holds the band's splat records, routed there by b200gs_route_project_slice (set by
 * b200gs_render_project_routed's callers for the rasterize call of the same frame): the rasterizer then takes the
 * number of live entries from the device-side counter whatever the band's extent. */
#define B200GS_CAM_ROUTED 4

/* Gradients written by b200gs_render_backward.  Non-null members are OVERWRITTEN (dense, zero for
 * culled Gaussians).  Members for the path not in use must be NULL. */
typedef struct b200gs_grads {
  float* pos;         /* [n,3] */
  float* opacity_raw; /* [n]   */
  float* scale_raw;   /* [n,3] */
  float* q_raw;       /* [n,4] */
  float* sigma;       /* [n,3,3] */
  float* f_dc;        /* [n,3] */
  float* f_rest;      /* [n,45] */
  float* color;       /* [n,3] */
} b200gs_grads;

typedef struct b200gs_sizes {
  size_t frame_bytes; /* per-Gaussian + per-tile + per-pixel state (must survive until backward) */
  size_t isect_bytes; /* per-intersection lists for `isect_capacity` entries */
} b200gs_sizes;

/* Counters the device leaves in the first 64 bytes of the frame workspace. */
typedef struct b200gs_frame_stats {
  uint32_t n_isect;    /* I: total (tile, Gaussian) intersections */
  uint32_t n_visible;  /* V: Gaussians that survive every cull (a band frame may drop Gaussians that cannot touch its
                          rows before the culls: its V, n_in_frustum and I describe the band, not the frame) */
  uint32_t overflow;   /* 1 if I exceeded isect_capacity in rasterize */
  uint32_t n_in_frustum; /* survivors of S1-S7 (before the on-screen test; render.py:235 raises when
                            this is > 0 but n_visible == 0) */
  uint32_t n_super;    /* (supertile, Gaussian) pairs: the number of keys the binning sort handles */
  uint32_t n_sorted;   /* band frames: Gaussians whose tile rect meets the band (the keys the depth sort handles) */
  uint32_t n_candidates; /* band frames: Gaussians that passed the cheap band test and were projected */
  uint32_t reserved[9];
} b200gs_frame_stats;

int b200gs_abi_version(void);
const char* b200gs_last_error(void);

/* Workspace sizes for n Gaussians, an HxW frame and room for isect_capacity intersections. */
int b200gs_workspace_sizes(int32_t n, int32_t H, int32_t W, uint32_t isect_capacity, b200gs_sizes* out);

/* gaussian.py:71-127.  sigma_out [n,3,3]. */
int b200gs_build_sigma(int32_t n, const float* scale_raw, const float* q_raw, float* sigma_out, void* stream);
/* autograd of the above: grad_sigma [n,3,3] -> grad_scale_raw [n,3], grad_q_raw [n,4] (overwritten). */
int b200gs_build_sigma_backward(int32_t n, const float* scale_raw, const float* q_raw,
                                const float* grad_sigma, float* grad_scale_raw, float* grad_q_raw,
                                void* stream);

/* spherical_harmonics.py:70-166.  color_out [n,3]. */
int b200gs_evaluate_sh(int32_t n, const float* f_dc, const float* f_rest, const float* points,
                       const float* c2w, float* color_out, void* stream);
/* autograd of the above: grad_color [n,3] -> grad_f_dc [n,3], grad_f_rest [n,45], grad_points [n,3]. */
int b200gs_evaluate_sh_backward(int32_t n, const float* f_dc, const float* f_rest, const float* points,
                                const float* c2w, const float* grad_color, float* grad_f_dc,
                                float* grad_f_rest, float* grad_points, void* stream);

/* render.py:104-258 + 305-315 (S1-S11, S15) fused with Sigma/SH when given raw parameters, then the
 * global depth sort (S8) and the prefix sum over tile counts.  Leaves I in the frame header; when
 * `stats_host` (pinned host memory) is non-null the header is also copied there asynchronously. */
int b200gs_render_project(const b200gs_gaussians* g, const b200gs_camera* cam, void* frame_ws,
                          size_t frame_bytes, b200gs_frame_stats* stats_host, void* stream);

/* render.py:260-410 (S12-S17): emit (tile, id) pairs in depth order, stable radix sort by tile,
 * tile ranges, per-tile front-to-back blend.  image_out [H,W,3], clamped to [0,1].
 * If I > isect_capacity the overflow flag in the frame header is set and the image is invalid. */
int b200gs_render_rasterize(const b200gs_camera* cam, int32_t n, void* frame_ws, size_t frame_bytes,
                            void* isect_ws, size_t isect_bytes, uint32_t isect_capacity,
                            float* image_out, b200gs_frame_stats* stats_host, void* stream);

/* Same call; additionally records `stats_event` (a cudaEvent_t, may be NULL) on `stream` right after the
 * binning scan, i.e. as soon as every counter of the frame statistics (I, V, pair count, overflow) is
 * final and on its way to `stats_host`.  A host that sizes the lists speculatively waits for this event
 * only - not for the end of the frame - to learn whether they fitted, and can queue the next frame while
 * the sort, split and blend kernels of this one are still running. */
int b200gs_render_rasterize_ev(const b200gs_camera* cam, int32_t n, void* frame_ws, size_t frame_bytes,
                               void* isect_ws, size_t isect_bytes, uint32_t isect_capacity,
                               float* image_out, b200gs_frame_stats* stats_host, void* stats_event,
                               void* stream);

/* Same again with the blend on its own stream: binning runs on `stream`, the blend kernel on `blend_stream`, ordered
 * by an event.  A host rendering a sequence of independent frames (an orbit: scripts/render_trained.py:333-358) queues
 * the next frame's project + binning on `stream` (ideally a high-priority stream) while this frame's blend is still
 * running on `blend_stream`; image_out is valid in `blend_stream` order.  The workspaces must stay untouched until
 * the blend has finished. */
int b200gs_render_rasterize_split(const b200gs_camera* cam, int32_t n, void* frame_ws, size_t frame_bytes,
                                  void* isect_ws, size_t isect_bytes, uint32_t isect_capacity,
                                  float* image_out, b200gs_frame_stats* stats_host, void* stats_event,
                                  void* stream, void* blend_stream);

/* Backward of project+rasterize (scripts/train.py:530): grad_image [H,W,3] -> b200gs_grads.
 * Needs the frame/isect workspaces exactly as the forward left them. */
int b200gs_render_backward(const b200gs_gaussians* g, const b200gs_camera* cam, void* frame_ws,
                           size_t frame_bytes, void* isect_ws, size_t isect_bytes,
                           uint32_t isect_capacity, const float* grad_image, const b200gs_grads* grads,
                           void* stream);

/* One-call forward with HOST buffers (pageable or pinned): uploads the Gaussian set and the pose,
 * allocates its own device workspaces, renders, downloads the image.  Synchronous.  This is the call
 * `bench.py` times as "e2e".  c2w_host: 16 floats.  image_host: [H,W,3]. */
int b200gs_render_host(const b200gs_gaussians* g_host, const b200gs_camera* cam, const float* c2w_host,
                       float* image_host, b200gs_frame_stats* stats_out);

/* Introspection for parity tests: copies of device-side per-Gaussian results.  All optional outputs
 * are DEVICE pointers of n elements (xy: [n,2], conic: [n,3], rect: [n,4] int32 tile rect
 * tu0,tu1,tv0,tv1), valid after b200gs_render_project. */
int b200gs_debug_export(int32_t n, const void* frame_ws, size_t frame_bytes, int32_t H, int32_t W,
                        float* xy, float* depth, float* conic, float* opacity, float* color,
                        int32_t* radius, int32_t* rect, int32_t* tiles_touched, int32_t* depth_order,
                        void* stream);
/* After rasterize: sorted (tile, Gaussian id) list and per-tile ranges [tiles,2]. */
int b200gs_debug_export_lists(const void* frame_ws, size_t frame_bytes, const void* isect_ws,
                              size_t isect_bytes, uint32_t isect_capacity, int32_t n, int32_t H, int32_t W,
                              int32_t* list_tile, int32_t* list_id, uint32_t count, int32_t* ranges,
                              void* stream);

/* gaussian_splatting/losses.py:158-185  compute_loss(pred, target, lambda_l1, lambda_ssim) - the training loss
 * that consumes the rendered image (scripts/train.py:511): mean |pred - target| and 1 - mean SSIM (11x11
 * Gaussian window, sigma 1.5, zero padding, per channel; losses.py:27-155), fused into one kernel.
 * pred / target: [n_img, H, W, 3] fp32.  out3 (device, 3 floats) = (l1, ssim_loss, lambda_l1 l1 + lambda_ssim
 * ssim_loss).  with_grad != 0 also leaves the three partial-derivative maps in the workspace that
 * b200gs_l1_ssim_backward convolves into grad_pred [n_img,H,W,3] = d(total)/d(pred) * (*grad_total)
 * (grad_total: device scalar, NULL = 1).  The workspace (b200gs_loss_workspace_bytes) is caller-owned and
 * must be the same buffer in both calls. */
size_t b200gs_loss_workspace_bytes(int32_t n_img, int32_t H, int32_t W, int32_t with_grad);
int b200gs_l1_ssim_forward(const float* pred, const float* target, int32_t n_img, int32_t H, int32_t W,
                           double lambda_l1, double lambda_ssim, void* workspace, size_t workspace_bytes,
                           int32_t with_grad, float* out3, void* stream);
int b200gs_l1_ssim_backward(const float* pred, const float* target, int32_t n_img, int32_t H, int32_t W,
                            double lambda_l1, double lambda_ssim, const void* workspace, size_t workspace_bytes,
                            const float* grad_total, float* grad_pred, void* stream);

/* Frame sink (scripts/render_trained.py:357, scripts/inference.py:117): the reference downloads the fp32 image
 * (12 B/pixel) and converts it on the host with (img * 255).astype(uint8); this does the same conversion (fp32
 * multiply, truncation) on the device so that 3 B/pixel cross PCIe.  image: [numel] fp32 in [0,1]; out: [numel] u8. */
int b200gs_image_to_u8(const float* image, uint8_t* out, size_t numel, void* stream);

/* scripts/train.py:394-401,538  torch.optim.Adam over the six parameter groups (per-group lr, eps = 1e-15):
 * one launch updates every tensor of the table in place (param, exp_avg, exp_avg_sq), reading each array
 * once.  Arithmetic of torch/optim/adam.py with weight_decay = 0, amsgrad = False; `step` is the 1-based
 * step count of that tensor AFTER this update (bias corrections are evaluated on the host in double). */
typedef struct b200gs_adam_tensor {
  float* param;        /* [numel] updated in place */
  const float* grad;   /* [numel] */
  float* exp_avg;      /* [numel] first moment, updated in place */
  float* exp_avg_sq;   /* [numel] second moment, updated in place */
  int64_t numel;
  double lr;
  int32_t step;
  int32_t reserved;
} b200gs_adam_tensor;
int b200gs_adam_step(const b200gs_adam_tensor* tensors, int32_t n_tensors, double beta1, double beta2, double eps,
                     void* stream);

/* scripts/train.py:536  torch.nn.utils.clip_grad_norm_(model.pos, max_norm): total_norm = ||grad||_2 (written to
 * total_norm_out, device, may be NULL), grad *= min(1, max_norm / (total_norm + 1e-6)) in place - without a host
 * sync.  workspace: b200gs_clip_workspace_bytes(numel) bytes, caller-owned. */
size_t b200gs_clip_workspace_bytes(int64_t numel);
int b200gs_clip_grad_norm(float* grad, int64_t numel, double max_norm, void* workspace, size_t workspace_bytes,
                          float* total_norm_out, void* stream);
/* The same for SEVERAL gradient tensors, as torch.nn.utils.clip_grad_norm_(parameters, max_norm) takes them: the norm
 * is the joint L2 norm of all tensors (torch/nn/utils/clip_grad.py: the norm of the per-tensor norms), one coefficient
 * scales every tensor.  At most B200GS_CLIP_MAX_TENSORS non-empty tensors per call.  workspace:
 * b200gs_clip_workspace_bytes_multi(numel, n_tensors) bytes, caller-owned. */
#define B200GS_CLIP_MAX_TENSORS 16
size_t b200gs_clip_workspace_bytes_multi(const int64_t* numel, int32_t n_tensors);
int b200gs_clip_grad_norm_multi(float* const* grads, const int64_t* numel, int32_t n_tensors, double max_norm,
                                void* workspace, size_t workspace_bytes, float* total_norm_out, void* stream);

/* scripts/train.py:89-195  GaussianModel.densify_and_prune (+ _prune_points / _split_points / _clone_points), called
 * every 100 iterations (train.py:544-557): prune rows with sigmoid(opacity_raw) < opacity_threshold; of the kept rows
 * with ||pos_grad||_2 > max_grad, append a displaced, shrunk copy of the large ones (max exp(scale_raw) >
 * scale_threshold: pos + (noise * exp(scale_raw)) * 0.1, scale_raw - 0.5) and an identical copy of the small ones.
 * plan():  flags + three scans into `workspace` (b200gs_densify_workspace_bytes(n) bytes, caller-owned); the counts
 *          {n_keep, n_split, n_clone} are copied to counts_host (3 x uint32, pinned or pageable) on `stream`.
 * apply(): in[t] -> out[t] for the six tensors in the order pos[3] opacity_raw[1] f_dc[3] f_rest[45] scale_raw[3] q_raw[4];
 *          out[t] has n_keep + n_split + n_clone rows (kept rows, then split copies, then clones, each in order);
 *          noise: [n_split,3] standard-normal samples (the reference draws them with torch.randn_like). */
size_t b200gs_densify_workspace_bytes(int32_t n);
int b200gs_densify_plan(int32_t n, const float* opacity_raw, const float* scale_raw, const float* pos_grad,
                        double opacity_threshold, double max_grad, double scale_threshold, void* workspace,
                        size_t workspace_bytes, uint32_t* counts_host, void* stream);
int b200gs_densify_apply(int32_t n, const void* workspace, size_t workspace_bytes, const float* const* in6,
                         float* const* out6, const float* noise, void* stream);

/* ---- Data-parallel optimizer step over NVLink peer memory (one process per GPU) --------------------------------
 * The data-parallel form of scripts/train.py:530-538 (backward -> SUM of the six gradient tensors over the ranks ->
 * clip_grad_norm_(model.pos, 1.0) -> optim.Adam.step()); the reference itself is single-GPU (scripts/train.py:285-291
 * only prints the GPU count).  Every rank owns 1/world of every tensor: one kernel reads the owned gradient
 * elements from every rank's staging buffer over NVLink (reduce-scatter), clips, updates its shard of the Adam
 * moments and stores the new parameter values into every rank's parameter buffer (all-gather).
 *
 * Each rank allocates one peer-visible "area" of b200gs_peer_area_bytes() bytes (zero-filled before first use),
 *     [ control block: B200GS_PEER_CTRL_BYTES ][ parameters: flat_total floats ][ gradient staging: flat_total floats ]
 * and the host side maps every rank's area into every process (CUDA VMM / IPC) and passes the mapped base
 * pointers in rank order.  Tensor t lives at float offset layout.offset[t] of the flat buffers; rank r owns its
 * elements [min(numel, r*per[t]), min(numel, (r+1)*per[t])) and keeps their moments at shard_offset[t] of its two
 * shard arrays (shard_total floats each).  All ranks must issue the same sequence of peer calls; *epoch is a
 * host-side counter (start at 0) that the calls advance by the number of cross-GPU barriers they used. */
#define B200GS_MAX_PEERS 16
#define B200GS_PEER_MAX_TENSORS 8
#define B200GS_PEER_CTRL_BYTES 4096
typedef struct b200gs_peer_layout {
  int64_t offset[B200GS_PEER_MAX_TENSORS];
  int64_t per[B200GS_PEER_MAX_TENSORS];
  int64_t shard_offset[B200GS_PEER_MAX_TENSORS];
  int64_t flat_total, shard_total;
} b200gs_peer_layout;
typedef struct b200gs_peer_group {
  int32_t world, rank;
  void* area[B200GS_MAX_PEERS];   /* area[q]: rank q's area as mapped into THIS process (area[rank] is local) */
  void* multicast;                /* NVLS multicast mapping of all the areas (multimem.ld_reduce / multimem.st go
                                     through it), or NULL: plain peer loads and stores */
} b200gs_peer_group;
typedef struct b200gs_peer_tensor {
  float* grad;      /* this rank's local gradient [numel] (NULL: contributes zeros); receives the reduced gradient
                       when write_grads != 0 / from b200gs_peer_allreduce */
  int64_t numel;
  double lr;
  int32_t step;     /* 1-based step count after this update */
  int32_t clip;     /* != 0: member of the set whose joint L2 norm is clipped to max_norm */
} b200gs_peer_tensor;
/* Host arithmetic only (no device work). */
int b200gs_peer_layout_compute(const int64_t* numel, int32_t n_tensors, int32_t world, b200gs_peer_layout* out);
size_t b200gs_peer_area_bytes(const b200gs_peer_layout* layout);
/* Cross-GPU barrier on `stream` (every rank must call it with the same epoch sequence). */
int b200gs_peer_barrier(const b200gs_peer_group* group, uint32_t* epoch, void* stream);
/* Fused reduce-scatter + clip + Adam + all-gather.  exp_avg_shard / exp_avg_sq_shard: [layout.shard_total] floats,
 * local.  max_norm <= 0 disables clipping; total_norm_out (device, may be NULL) receives the joint norm. */
int b200gs_peer_adam_step(const b200gs_peer_group* group, const b200gs_peer_layout* layout,
                          const b200gs_peer_tensor* tensors, int32_t n_tensors, float* exp_avg_shard,
                          float* exp_avg_sq_shard, double beta1, double beta2, double eps, double max_norm,
                          int32_t write_grads, uint32_t* epoch, float* total_norm_out, void* stream);
/* SUM all-reduce of the gradient tensors over peer memory (reduce-scatter + all-gather in one kernel, no NCCL);
 * every tensors[t].grad is replaced by the sum over the ranks. */
int b200gs_peer_allreduce(const b200gs_peer_group* group, const b200gs_peer_layout* layout,
                          const b200gs_peer_tensor* tensors, int32_t n_tensors, uint32_t* epoch, void* stream);

/* ---- Tile-row bands with the per-Gaussian work divided over the ranks ("sort-middle") --------------------------------
 * The reference projects every Gaussian once (render.py:104-258) and then loops over independent tiles
 * (render.py:325-399).  With one process per GPU, rank r projects the Gaussians [r*N/p, (r+1)*N/p) and writes each
 * survivor's splat record, depth key and tile rect (clamped to the band) into the frame workspace of every rank whose
 * band of tile rows the rect meets - over peer-mapped memory, in index order, into segment r of that workspace
 * (seg_capacity entries per source rank, >= the largest slice, so nothing can overflow).  After a cross-GPU barrier
 * (b200gs_peer_barrier) each rank continues its band with b200gs_render_project_routed + b200gs_render_rasterize*
 * (camera flags B200GS_CAM_ROUTED, usually B200GS_CAM_KEEP_OUTSIDE_BAND with image_out = the root's frame buffer).
 * Band workspaces are laid out by b200gs_workspace_sizes(world * seg_capacity, H, W, ...); a routed entry's id is its
 * position there.  Per-tile lists, and therefore the pixels, are those of the one-GPU frame bit for bit. */
typedef struct b200gs_route {
  int32_t world, rank;
  uint32_t seg_capacity;                    /* entries per (source rank, band) segment */
  int32_t band_row[B200GS_MAX_PEERS + 1];   /* band q = tile rows [band_row[q], band_row[q+1]) */
  void* band_ws[B200GS_MAX_PEERS];          /* band q's frame workspace as mapped into THIS process */
  size_t band_ws_bytes;
  uint32_t flags, reserved;                 /* B200GS_ROUTE_* */
} b200gs_route;
/* b200gs_route_project_slice writes only what the destinations need BEFORE their blend (depth key, tile rect,
 * supertile count: 16 of the 64 bytes per entry); the splat records follow with b200gs_route_records, which the caller
 * puts on a second stream so that they cross NVLink beside the destinations' depth sort.  Two barriers then: one after
 * b200gs_route_project_slice (before b200gs_render_project_routed), one after b200gs_route_records (before
 * b200gs_render_rasterize*). */
#define B200GS_ROUTE_RECORDS_LATER 1
/* Source role: project `g_slice` (this rank's slice; cam = the FULL frame, no tile-row range) into the private
 * `slice_ws` (b200gs_workspace_sizes(g_slice->n, H, W, ...) bytes) and route the survivors. */
int b200gs_route_project_slice(const b200gs_gaussians* g_slice, const b200gs_camera* cam, void* slice_ws,
                               size_t slice_bytes, const b200gs_route* route, void* stream);
/* Second half of the source role when route->flags has B200GS_ROUTE_RECORDS_LATER: writes the splat records of the
 * slice projected by the last b200gs_route_project_slice on `slice_ws` (n_slice entries, same camera, same route) to the
 * positions that call assigned.  Any stream ordered after that call. */
int b200gs_route_records(int32_t n_slice, const b200gs_camera* cam, void* slice_ws, size_t slice_bytes,
                         const b200gs_route* route, void* stream);
/* Destination role, after the barrier: gathers the routed entries of this band (cam = the band's camera: tile-row
 * range + B200GS_CAM_ROUTED) and depth-sorts them; the frame then continues with b200gs_render_rasterize* called with
 * n = world * seg_capacity and the same camera.  Statistics: V = entries routed here, I = their intersections. */
int b200gs_render_project_routed(const b200gs_camera* cam, const b200gs_route* route, void* frame_ws,
                                 size_t frame_bytes, b200gs_frame_stats* stats_host, void* stream);

/* Per-region CUDA-event profiling (bench.py's per-kernel table).  enable(1) starts recording an event
 * pair around every kernel group launched through this library; collect() synchronises the device,
 * sums the elapsed milliseconds and the number of calls per region, clears the records and returns the
 * number of regions.  b200gs_kernel_launch_count() counts every kernel this library has launched. */
int b200gs_profile_enable(int on);
int b200gs_profile_collect(float* ms_out, int32_t* calls_out, int32_t max_regions);
const char* b200gs_profile_region_name(int32_t id);
unsigned long long b200gs_kernel_launch_count(void);

/* Stand-alone primitives (exported for unit tests of the integer stages). */
int b200gs_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total_out,
                              void* scratch, size_t scratch_bytes, void* stream);
size_t b200gs_scan_scratch_bytes(uint32_t n);
/* Stable LSD radix sort of (key, value) pairs on key bits [begin_bit, end_bit).  Results end up in
 * keys_out/vals_out; keys_in/vals_in are clobbered. */
int b200gs_radix_sort_pairs(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                            uint32_t n, int begin_bit, int end_bit, void* scratch, size_t scratch_bytes,
                            void* stream);
size_t b200gs_sort_scratch_bytes(uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* B200GS_H */
