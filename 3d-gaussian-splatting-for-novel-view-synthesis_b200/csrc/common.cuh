// Shared device helpers + the workspace layout used by every stage.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <atomic>

#include "gs_math.cuh"
#include "../../include/b200gs.h"

namespace gs {

constexpr int kTile = 16;
constexpr int kSuperX = 8, kSuperY = 4;   // a supertile is 8x4 tiles (128x64 px): 32 tiles = one 32-bit mask
constexpr uint32_t kCulledKey = 0xFFFFFFFFu;

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Function attributes (the opt-in to more than 48 KB of dynamic shared memory, the shared-memory carve-out) belong to the
// (kernel, device) pair, not to the process: a process that renders on cuda:0 and then on cuda:1 has to set them on both.
// `mask` is a per-call-site bit set of device ordinals; f() runs before the bit is published, so a second host thread
// can at worst repeat the (idempotent) attribute call, never launch ahead of it.
template <class F>
inline void once_per_device(std::atomic<uint64_t>& mask, F&& f) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (mask.load(std::memory_order_acquire) & bit) return;
  f();
  mask.fetch_or(bit, std::memory_order_release);
}

// ---- streaming loads / stores (read-once data must not pollute L1) ---------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- workspace layouts -----------------------------------------------------------------------------
// Frame workspace: header | per-Gaussian | per-tile | per-pixel.   Everything 256-byte aligned.
struct FrameLayout {
  size_t header;          // b200gs_frame_stats (64 B) + scan/sort bookkeeping
  size_t rec0, rec1, rec2;  // float4[n] each: (u,v,c*A11,c*2*A12) (c*A22,log2 op,gate,r) (g,b,ext_u,ext_v), see preprocess.cu
  size_t depth_key;       // u32[n]  float bits of z, 0xFFFFFFFF when culled
  size_t rect;            // uint2[n] packed u16 tile rect: x = tu0 | tu1<<16, y = tv0 | tv1<<16
  size_t radius;          // u32[n]  ceil(2.5 sqrt(lambda_max)) (introspection only)
  size_t super_touched;   // u32[n]  number of supertiles the tile rect overlaps
  size_t sort_key_alt;    // u32[n]  depth-sort ping-pong buffers (depth_key itself stays intact)
  size_t sort_key_alt2;   // u32[n]
  size_t order;           // u32[n]  Gaussian ids in depth order (sorted values)
  size_t order_alt;       // u32[n]
  size_t cand_key;        // u32[n]  band frames: depth keys / ids of the band's candidates (before the compaction)
  size_t cand_id;         // u32[n]
  size_t offsets;         // u32[n]  exclusive scan of super_touched in depth order
  size_t grad_acc;        // float[n*12] blend-backward accumulators (Mx,My,Mxx,Mxy,Myy,M0,r,g,b,+pad): moments of dL/dq
  size_t ranges;          // uint2[tiles] (start,end) per tile
  size_t tile_count;      // u32[supertiles*4*32] entries per tile and list quarter, supertile-major; + u32[supertiles*4] totals
  size_t final_T;         // float[P]
  size_t n_contrib;       // u32[P]  (#list entries consumed) | channel-overflow bits 29..31
  size_t band_scratch;    // look-back state of the band kernels: [0, half) band_select, [half, 2 half) compact_keys
  size_t band_scratch_half;
  size_t scratch;         // scan + sort scratch (sized for max(n, isect) users at call time)
  size_t scratch_bytes;
  size_t total;
};

struct IsectLayout {
  size_t keys, keys_alt;  // u32[cap] supertile ids of the (supertile, Gaussian) pairs
  size_t vals, vals_alt;  // u32[cap] Gaussian ids of the pairs
  size_t lists;           // u32[cap] per-tile depth-sorted Gaussian ids (grouped by supertile)
  size_t scratch;         // sort scratch for cap entries
  size_t scratch_bytes;
  size_t total;
};

// device-visible header (first bytes of the frame workspace)
struct FrameHeader {
  b200gs_frame_stats stats;   // 64 bytes
};
// routed band frames (route.cu): per source rank (entries, intersections) of the rank's segment, uint2[B200GS_MAX_PEERS],
// inside the 256-byte header section behind the statistics
constexpr size_t kRouteInOffset = 64;
static_assert(kRouteInOffset + B200GS_MAX_PEERS * 8 <= 256, "route_in must fit the header section");

// Radix-sort pass plan (scan_sort.cu).  Digits are 8 bits wide, or 9 when that saves a pass (27-bit depth keys:
// 3 passes); the global digit histograms a producer kernel may fill are laid out [pass][kSortMaxRadix].
constexpr int kSortMaxPasses = 4;
constexpr int kSortMaxRadix = 512;
struct SortPasses {
  int num;
  int digit_bits;                  // 8 or 9: which instance of the pass kernel runs
  int shift[kSortMaxPasses];
  int bits[kSortMaxPasses];
};
SortPasses sort_passes(int begin_bit, int end_bit);

// The depth sort's view of a depth key: sort key = min(float_bits(z) - base, max_key) (culled marker -> max_key), on the
// key_bits low bits (RenderParams::key_base / key_bits).  The producers of the keys (preprocess kernels) use the same
// plan to accumulate the digit histograms of all passes; the first radix pass applies the transform on load, so the
// depth_key array itself keeps the raw float bits.
struct DepthKeyPlan {
  uint32_t base, max_key;
  SortPasses sp;
};
inline DepthKeyPlan depth_key_plan(const RenderParams& rp) {
  DepthKeyPlan kp;
  kp.base = rp.key_base;
  kp.max_key = rp.key_bits >= 32 ? 0xFFFFFFFFu : ((1u << rp.key_bits) - 1u);
  kp.sp = sort_passes(0, rp.key_bits);
  return kp;
}
__device__ __forceinline__ uint32_t depth_sort_key(uint32_t raw, uint32_t base, uint32_t max_key) {
  const uint32_t k = raw - base;
  return k < max_key ? k : max_key;
}
// s_dh: [kSortMaxPasses][kSortMaxRadix] shared-memory counters
__device__ __forceinline__ void depth_hist_add(uint32_t* s_dh, const DepthKeyPlan& kp, uint32_t raw_key) {
  const uint32_t k = depth_sort_key(raw_key, kp.base, kp.max_key);
#pragma unroll
  for (int p = 0; p < kSortMaxPasses; ++p)
    if (p < kp.sp.num) atomicAdd(&s_dh[p * kSortMaxRadix + ((k >> kp.sp.shift[p]) & ((1u << kp.sp.bits[p]) - 1u))], 1u);
}

size_t scan_scratch_bytes(uint32_t n);
size_t sort_scratch_bytes(uint32_t n);
size_t scan_emit_scratch_bytes(uint32_t n);

inline FrameLayout frame_layout(int n, int H, int W) {
  FrameLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t N = (size_t)(n > 0 ? n : 1);
  const size_t tiles = (size_t)ceil_div(W, kTile) * ceil_div(H, kTile);
  const size_t P = (size_t)H * W;
  L.header = take(256);
  L.rec0 = take(N * 16); L.rec1 = take(N * 16); L.rec2 = take(N * 16);
  L.depth_key = take(N * 4);
  L.rect = take(N * 8);
  L.radius = take(N * 4);
  L.super_touched = take(N * 4);
  L.sort_key_alt = take(N * 4);
  L.sort_key_alt2 = take(N * 4);
  L.order = take(N * 4);
  L.order_alt = take(N * 4);
  L.cand_key = take(N * 4);
  L.cand_id = take(N * 4);
  L.offsets = take(N * 4);
  L.grad_acc = take(N * 12 * 4);
  L.ranges = take(tiles * 8);
  // per supertile and list quarter (kSplitParts = 4): 32 per-tile counts, then one total per (supertile, quarter)
  L.tile_count = take((size_t)ceil_div(ceil_div(W, kTile), kSuperX) * ceil_div(ceil_div(H, kTile), kSuperY) * (32 + 1) * 4 * 4);
  L.final_T = take(P * 4);
  L.n_contrib = take(P * 4);
  L.band_scratch_half = align_up(512 + (N / 1024 + 2) * 8, 256);
  L.band_scratch = take(2 * L.band_scratch_half);
  size_t a = scan_scratch_bytes((uint32_t)N);
  const size_t b = sort_scratch_bytes((uint32_t)N), c = scan_emit_scratch_bytes((uint32_t)N);
  if (c > a) a = c;
  L.scratch_bytes = a > b ? a : b;
  L.scratch = take(L.scratch_bytes);
  L.total = off;
  return L;
}

inline IsectLayout isect_layout(uint32_t cap) {
  IsectLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t C = cap > 0 ? cap : 1;
  L.keys = take(C * 4); L.keys_alt = take(C * 4);
  L.vals = take(C * 4); L.vals_alt = take(C * 4);
  L.lists = take(C * 4);
  L.scratch_bytes = sort_scratch_bytes((uint32_t)C);
  L.scratch = take(L.scratch_bytes);
  L.total = off;
  return L;
}

template <typename T>
__host__ __device__ inline T* ws_ptr(void* base, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + off);
}
template <typename T>
__host__ __device__ inline const T* ws_ptr(const void* base, size_t off) {
  return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + off);
}

// ---- stage entry points (host functions implemented in the .cu files) --------------------------------
struct GaussIn {       // device pointers, mirrors b200gs_gaussians
  int n;
  const float *pos, *opacity_raw, *scale_raw, *q_raw, *sigma, *f_dc, *f_rest, *color;
};
struct GaussGrad {
  float *pos, *opacity_raw, *scale_raw, *q_raw, *sigma, *f_dc, *f_rest, *color;
};

// depth_hist (optional): [4][256] zeroed counters; when the kernel variant can, it accumulates the digit
// histograms of the depth keys there and sets *hist_done (the depth sort then skips its histogram pass).
cudaError_t launch_preprocess_fwd(const GaussIn& g, const float* c2w, const RenderParams& rp, void* frame_ws,
                                  const FrameLayout& L, cudaStream_t s, uint32_t* depth_hist = nullptr,
                                  bool* hist_done = nullptr);   // depth_hist: [kSortMaxPasses][kSortMaxRadix], plan = depth_key_plan(rp)
cudaError_t radix_sort_prepare(void* scratch, size_t scratch_bytes, uint32_t n, int begin_bit, int end_bit, cudaStream_t s);
uint32_t* radix_sort_hist(void* scratch);
cudaError_t launch_preprocess_bwd(const GaussIn& g, const GaussGrad& gg, const float* c2w, const RenderParams& rp,
                                  void* frame_ws, const FrameLayout& L, cudaStream_t s);
cudaError_t launch_build_sigma(int n, const float* scale_raw, const float* q_raw, float* sigma, cudaStream_t s);
cudaError_t launch_build_sigma_bwd(int n, const float* scale_raw, const float* q_raw, const float* g_sigma,
                                   float* g_scale, float* g_q, cudaStream_t s);
cudaError_t launch_eval_sh(int n, const float* f_dc, const float* f_rest, const float* pts, const float* c2w,
                           float* color, cudaStream_t s);
cudaError_t launch_eval_sh_bwd(int n, const float* f_dc, const float* f_rest, const float* pts, const float* c2w,
                               const float* g_color, float* g_dc, float* g_rest, float* g_pts, cudaStream_t s);

cudaError_t launch_exclusive_scan(const uint32_t* in, const uint32_t* gather_idx, uint32_t* out, uint32_t n,
                                  uint32_t* total_out, void* scratch, size_t scratch_bytes, cudaStream_t s);
cudaError_t launch_radix_sort(const uint32_t* keys_src, const uint32_t* vals_src, uint32_t* keys_a,
                              uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n,
                              const uint32_t* n_dev, int begin_bit, int end_bit, void* scratch,
                              size_t scratch_bytes, int* result_in_a, cudaStream_t s, bool hist_ready = false,
                              uint32_t key_sub = 0, uint32_t key_max = 0xFFFFFFFFu, bool narrow = false);
// narrow: small blocks (256 threads, 20 K registers, 38 KB) that fit beside a running blend kernel (frame pipeline);
// otherwise 512-thread blocks of 7168 keys, one per SM at 1M keys
// (key_sub, key_max): the first pass reads keys as min(key - key_sub, key_max) (depth keys, see DepthKeyPlan); the
// ping-pong buffers then hold the transformed keys
// n_dev (optional): device-side count of live entries at the front of `order` (band frames compact their keys)
cudaError_t launch_scan_emit_super(int n, const uint32_t* n_dev, const uint32_t* order, const uint32_t* super_touched,
                                   const uint2* rect, int super_x, int super_y0, uint32_t capacity, uint32_t* keys,
                                   uint32_t* vals, b200gs_frame_stats* stats, int sort_bits, void* sort_scratch,
                                   size_t sort_scratch_bytes_, void* se_scratch, size_t se_scratch_bytes,
                                   cudaStream_t s);
// (key, id) of every key != kCulledKey in index order -> out_keys / out_ids, their number -> *count_out (device)
// (n_dev: optional device-side count of the valid keys, n then being the host's upper bound; ids_in: optional ids to
// carry instead of the key's position)
size_t compact_scratch_bytes(uint32_t n);
cudaError_t launch_compact_keys(const uint32_t* keys, const uint32_t* ids_in, uint32_t n, const uint32_t* n_dev,
                                uint32_t* out_keys, uint32_t* out_ids, uint32_t* count_out, void* scratch,
                                size_t scratch_bytes, cudaStream_t s);
size_t band_select_scratch_bytes(int n);
// flag_words: u32[ceil(n / 128) * 4] (one bit per Gaussian)
cudaError_t launch_band_select(const GaussIn& g, const float* c2w, const RenderParams& rp, void* frame_ws,
                               const FrameLayout& L, uint32_t* flag_words, uint32_t* cand_ids, uint32_t* n_cand,
                               void* select_scratch, size_t select_scratch_bytes, bool* used, cudaStream_t s);
cudaError_t launch_band_project(const GaussIn& g, const float* c2w, const RenderParams& rp, void* frame_ws,
                                const FrameLayout& L, const uint32_t* cand_ids, const uint32_t* n_cand, uint32_t* cand_key,
                                uint32_t* depth_hist, cudaStream_t s);

cudaError_t launch_emit_super(int n, const uint32_t* order, const uint32_t* offsets, const uint32_t* super_touched,
                              const uint2* rect, int super_x, uint32_t capacity, uint32_t* keys, uint32_t* vals,
                              b200gs_frame_stats* stats, cudaStream_t s);
// supertile rows [super_y0, super_y0 + super_y) (a band only bins, sorts and splits its own rows; ids are band-local)
cudaError_t launch_split_super(bool write, const uint32_t* keys, const uint32_t* vals, const uint2* rect,
                               uint32_t capacity, const b200gs_frame_stats* stats, int super_x, int super_y0, int super_y,
                               int tiles_x, int tiles_y, uint32_t* tile_count, uint2* ranges, uint32_t* lists,
                               cudaStream_t s);
cudaError_t launch_fill_list_tiles(const uint2* ranges, int n_tiles, uint32_t count, int32_t* list_tile,
                                   cudaStream_t s);

cudaError_t launch_adam_step(const b200gs_adam_tensor* tensors, int n_tensors, double beta1, double beta2, double eps,
                             cudaStream_t s);
size_t clip_workspace_bytes(long long numel);
size_t clip_workspace_bytes_multi(const int64_t* numel, int n_tensors);
cudaError_t launch_clip_grad_norm_multi(float* const* grads, const int64_t* numel, int n_tensors, double max_norm, void* ws,
                                        float* total_norm_out, cudaStream_t s);
cudaError_t launch_clip_grad_norm(float* grad, long long numel, double max_norm, void* ws, float* total_norm_out,
                                  cudaStream_t s);

// densify.cu: prune / split / clone as stream compaction
size_t densify_workspace_bytes(int n);
cudaError_t launch_densify_plan(int n, const float* opacity_raw, const float* scale_raw, const float* pos_grad,
                                double thr_op, double max_grad, double thr_scale, void* ws, uint32_t* counts_host,
                                cudaStream_t s);
cudaError_t launch_densify_apply(int n, const void* ws, const float* const in[6], float* const out[6], const float* noise,
                                 cudaStream_t s);

// route.cu: tile-row bands with the per-Gaussian work divided over the ranks (project a slice, route the records)
size_t route_scratch_bytes(int n, int world);
enum { kRouteAll = 0, kRouteMeta = 1, kRouteRecords = 2 };
cudaError_t launch_route_slice(int n, const void* slice_ws, const FrameLayout& SL, const b200gs_route* route,
                               const FrameLayout& BL, void* scratch, size_t scratch_bytes, int what, cudaStream_t s);
cudaError_t launch_gather_routed(int world, uint32_t seg_cap, void* band_ws, const FrameLayout& BL, uint32_t* out_keys,
                                 uint32_t* out_ids, uint32_t* depth_hist, const DepthKeyPlan& kp, cudaStream_t s);

// peer.cu: data-parallel optimizer step over NVLink peer memory
int peer_layout_compute(const int64_t* numel, int n_tensors, int world, b200gs_peer_layout* out);
cudaError_t launch_peer_barrier(const b200gs_peer_group* g, uint32_t epoch, cudaStream_t s);
cudaError_t launch_peer_step(const b200gs_peer_group* g, const b200gs_peer_layout* L, const b200gs_peer_tensor* tensors,
                             int n_tensors, bool adam, float* m_shard, float* v_shard, double beta1, double beta2,
                             double eps, double max_norm, int write_grads, uint32_t* epoch, float* total_norm_out,
                             cudaStream_t s, int* launches);

size_t loss_workspace_bytes(int n_img, int H, int W, bool with_grad);
cudaError_t launch_l1_ssim_fwd(const float* pred, const float* target, int n_img, int H, int W, float lambda_l1,
                               float lambda_ssim, void* ws, bool with_grad, float* out3, cudaStream_t s);
cudaError_t launch_l1_ssim_bwd(const float* pred, const float* target, int n_img, int H, int W, float lambda_l1,
                               float lambda_ssim, const void* ws, const float* grad_total, float* grad_pred,
                               cudaStream_t s);

cudaError_t launch_blend_fwd(const RenderParams& rp, const void* frame_ws, const FrameLayout& L, const uint32_t* vals,
                             float* image, cudaStream_t s, bool overlapped = false, bool row_stores = false);
cudaError_t launch_blend_bwd(const RenderParams& rp, void* frame_ws, const FrameLayout& L, const uint32_t* vals,
                             const float* image_grad, int n, cudaStream_t s);

}  // namespace gs
