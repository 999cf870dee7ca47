// Tile-row bands, sort-middle: every rank projects its SLICE of the Gaussians once and routes each survivor's splat
// record to the rank(s) whose band of tile rows its tile rect meets, over peer-mapped memory (SURVEY.md section 8e,
// BASELINE.json configs[4]; tiles are independent in the reference's loop, render.py:325-399, and so are Gaussians in its
// projection, render.py:104-258).
//
// The band renderer of round 1 / early round 2 had every rank look at ALL N Gaussians (band_select + band_project):
// ~350 of a band's ~600 us did not shrink with the number of ranks.  Here the per-Gaussian work is divided by the number
// of ranks as well:
//   source role   preprocess_fwd on Gaussians [rank * N/p, (rank+1) * N/p) into a private slice workspace (full-frame
//                 parameters: the same kernel, the same bits as a one-GPU frame), then
//                 route_count -> route_scan -> route_write: for every band b, the survivors whose tile rect meets b are
//                 written IN INDEX ORDER into segment `rank` of band b's workspace (records, depth key, rect clamped to
//                 the band, supertile count), and the segment's entry / intersection totals into b's header;
//   (flag barrier: every segment of every band has landed)
//   dest role     gather_routed: (key, position) pairs of the p segment prefixes, in segment order - rank order, and
//                 index order inside a segment, i.e. global index order, so the stable depth sort breaks ties exactly as
//                 the one-GPU frame does - plus the depth-digit histograms; then the band frame continues as before
//                 (depth sort, pair emission, supertile sort, split, blend into the root's frame buffer).
// A routed entry's "Gaussian id" is its position in the band's workspace; nothing downstream needs the global id (forward
// only).  Segments have room for a whole slice (seg_cap >= slice size), so nothing can overflow.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace gs {

constexpr int kRouteThreads = 512;
constexpr int kRouteWarps = kRouteThreads / 32;
constexpr int kRouteScanThreads = 1024;

struct RouteParams {                       // by-value kernel argument
  int world, rank;
  uint32_t seg_cap;
  int row[B200GS_MAX_PEERS + 1];           // band b = tile rows [row[b], row[b+1])
  char* ws[B200GS_MAX_PEERS];              // band workspaces (peer-mapped), all with the layout below
  size_t rec0, rec1, rec2, depth_key, rect, super_touched, route_in;
};

struct RouteHit { bool hit; int lo, hi; };
__device__ __forceinline__ RouteHit route_test(bool live, int tv0, int tv1, int r0, int r1) {
  RouteHit h;
  h.lo = max(tv0, r0);
  h.hi = min(tv1, r1 - 1);
  h.hit = live && h.hi >= h.lo;
  return h;
}

// per (band, block of kRouteThreads consecutive slice entries): how many entries go to the band, and their tile
// intersections there
__global__ void __launch_bounds__(kRouteThreads) route_count_kernel(uint32_t n, const uint32_t* __restrict__ depth_key,
                                                                    const uint2* __restrict__ rect, RouteParams rp,
                                                                    uint32_t n_blocks, uint32_t* __restrict__ counts,
                                                                    uint32_t* __restrict__ tiles) {
  __shared__ uint32_t s_c[B200GS_MAX_PEERS][kRouteWarps], s_t[B200GS_MAX_PEERS][kRouteWarps];
  const uint32_t i = blockIdx.x * kRouteThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool live = i < n && depth_key[i] != kCulledKey;
  uint2 rc = make_uint2(0u, 0u);
  if (live) rc = rect[i];
  const int tu0 = rc.x & 0xFFFF, tu1 = rc.x >> 16, tv0 = rc.y & 0xFFFF, tv1 = rc.y >> 16;
  for (int b = 0; b < rp.world; ++b) {
    const RouteHit h = route_test(live, tv0, tv1, rp.row[b], rp.row[b + 1]);
    const uint32_t m = __ballot_sync(0xffffffffu, h.hit);
    const uint32_t t = m ? __reduce_add_sync(0xffffffffu, h.hit ? (uint32_t)((tu1 - tu0 + 1) * (h.hi - h.lo + 1)) : 0u) : 0u;
    if (lane == 0) { s_c[b][warp] = __popc(m); s_t[b][warp] = t; }
  }
  __syncthreads();
  if (threadIdx.x < rp.world) {
    const int b = threadIdx.x;
    uint32_t c = 0, t = 0;
#pragma unroll
    for (int w = 0; w < kRouteWarps; ++w) { c += s_c[b][w]; t += s_t[b][w]; }
    counts[(size_t)b * n_blocks + blockIdx.x] = c;
    tiles[(size_t)b * n_blocks + blockIdx.x] = t;
  }
}

// one block per band: exclusive scan of the band's per-block counts in place, totals into the band's header
__global__ void __launch_bounds__(kRouteScanThreads) route_scan_kernel(uint32_t n_blocks, uint32_t* __restrict__ counts,
                                                                       const uint32_t* __restrict__ tiles, RouteParams rp) {
  __shared__ uint32_t s_c[kRouteScanThreads / 32], s_t[kRouteScanThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* c = counts + (size_t)b * n_blocks;
  const uint32_t* t = tiles + (size_t)b * n_blocks;
  const uint32_t per = (n_blocks + kRouteScanThreads - 1) / kRouteScanThreads;
  const uint32_t i0 = min((uint32_t)tid * per, n_blocks), i1 = min(i0 + per, n_blocks);
  uint32_t sum = 0, tsum = 0;
  for (uint32_t i = i0; i < i1; ++i) { sum += c[i]; tsum += t[i]; }
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += v;
  }
  const uint32_t wt = __reduce_add_sync(0xffffffffu, tsum);
  if (lane == 31) s_c[warp] = inc;
  if (lane == 0) s_t[warp] = wt;
  __syncthreads();
  uint32_t warp_off = 0, total = 0, ttotal = 0;
  for (int w = 0; w < kRouteScanThreads / 32; ++w) {
    const uint32_t v = s_c[w];
    if (w < warp) warp_off += v;
    total += v;
    ttotal += s_t[w];
  }
  uint32_t run = warp_off + (inc - sum);
  for (uint32_t i = i0; i < i1; ++i) { const uint32_t v = c[i]; c[i] = run; run += v; }
  if (tid == 0)      // this rank's segment of band b: entries, intersections (one 8-byte store, remote for b != rank)
    *reinterpret_cast<uint2*>(rp.ws[b] + rp.route_in + (size_t)rp.rank * sizeof(uint2)) = make_uint2(total, ttotal);
}

// Variant with per-warp runs (no staging): B200GS_ROUTE_WRITE=warp, kept for the comparison recorded in DESIGN.md.
__global__ void __launch_bounds__(kRouteThreads) route_write_warp_kernel(uint32_t n, const uint32_t* __restrict__ depth_key,
                                                                         const uint2* __restrict__ rect,
                                                                         const float4* __restrict__ rec0,
                                                                         const float4* __restrict__ rec1,
                                                                         const float4* __restrict__ rec2, RouteParams rp,
                                                                         uint32_t n_blocks, const uint32_t* __restrict__ base) {
  __shared__ uint32_t s_c[B200GS_MAX_PEERS][kRouteWarps];
  const uint32_t i = blockIdx.x * kRouteThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t key = kCulledKey;
  if (i < n) key = depth_key[i];
  const bool live = key != kCulledKey;
  uint2 rc = make_uint2(0u, 0u);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
  if (live) { rc = rect[i]; a0 = rec0[i]; a1 = rec1[i]; a2 = rec2[i]; }
  const int tu0 = rc.x & 0xFFFF, tu1 = rc.x >> 16, tv0 = rc.y & 0xFFFF, tv1 = rc.y >> 16;
  for (int b = 0; b < rp.world; ++b) {
    const uint32_t m = __ballot_sync(0xffffffffu, route_test(live, tv0, tv1, rp.row[b], rp.row[b + 1]).hit);
    if (lane == 0) s_c[b][warp] = __popc(m);
  }
  __syncthreads();
  for (int b = 0; b < rp.world; ++b) {
    const RouteHit h = route_test(live, tv0, tv1, rp.row[b], rp.row[b + 1]);
    const uint32_t m = __ballot_sync(0xffffffffu, h.hit);
    if (!h.hit) continue;
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += s_c[b][w];
    const size_t pos = (size_t)rp.rank * rp.seg_cap + base[(size_t)b * n_blocks + blockIdx.x] + before + __popc(m & lt);
    char* w = rp.ws[b];
    reinterpret_cast<float4*>(w + rp.rec0)[pos] = a0;
    reinterpret_cast<float4*>(w + rp.rec1)[pos] = a1;
    reinterpret_cast<float4*>(w + rp.rec2)[pos] = a2;
    reinterpret_cast<uint32_t*>(w + rp.depth_key)[pos] = key;
    reinterpret_cast<uint2*>(w + rp.rect)[pos] = make_uint2(rc.x, (uint32_t)h.lo | ((uint32_t)h.hi << 16));
    reinterpret_cast<uint32_t*>(w + rp.super_touched)[pos] =
        (uint32_t)((tu1 / kSuperX - tu0 / kSuperX + 1) * (h.hi / kSuperY - h.lo / kSuperY + 1));
  }
}

// The entries of one block are first compacted into shared memory, grouped by band and in index order inside a band,
// and then copied out as contiguous runs - ~kRouteThreads / world entries, i.e. around a kilobyte per array, band and
// block - because consecutive Gaussians of a slice land in different bands and NVLink wants large write requests
// (per-warp compaction gave 64-byte runs: 133 us for the routing pass at 8 ranks against 49 us into local memory).
// A block whose entries (counted once per band they meet) exceed the staging capacity goes band by band instead.
constexpr int kRouteStage = 768;           // staged entries per block: 48 KB of dynamic shared memory
constexpr size_t kRouteStageBytes = (size_t)kRouteStage * (3 * 16 + 8 + 4 + 4);

struct RouteStage {
  float4 *r0, *r1, *r2;
  uint2* rect;
  uint32_t *key, *sup;
};

// META: depth key, tile rect, supertile count (16 B per entry - all the destination needs until its blend);
// REC: the three float4 of the splat record (48 B per entry - what the blend reads).
template <bool META, bool REC>
__device__ __forceinline__ void route_copy_out(const RouteStage& st, const RouteParams& rp, int b, size_t pos, uint32_t k0,
                                               uint32_t k) {
  char* w = rp.ws[b];
  if (REC) {
    reinterpret_cast<float4*>(w + rp.rec0)[pos] = st.r0[k0 + k];
    reinterpret_cast<float4*>(w + rp.rec1)[pos] = st.r1[k0 + k];
    reinterpret_cast<float4*>(w + rp.rec2)[pos] = st.r2[k0 + k];
  }
  if (META) {
    reinterpret_cast<uint32_t*>(w + rp.depth_key)[pos] = st.key[k0 + k];
    reinterpret_cast<uint2*>(w + rp.rect)[pos] = st.rect[k0 + k];
    reinterpret_cast<uint32_t*>(w + rp.super_touched)[pos] = st.sup[k0 + k];
  }
}

template <bool META, bool REC>
__global__ void __launch_bounds__(kRouteThreads) route_write_kernel(uint32_t n, const uint32_t* __restrict__ depth_key,
                                                                    const uint2* __restrict__ rect,
                                                                    const float4* __restrict__ rec0,
                                                                    const float4* __restrict__ rec1,
                                                                    const float4* __restrict__ rec2, RouteParams rp,
                                                                    uint32_t n_blocks, const uint32_t* __restrict__ base) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ uint32_t s_c[B200GS_MAX_PEERS][kRouteWarps];
  __shared__ uint32_t s_off[B200GS_MAX_PEERS + 1];      // first staged slot of band b (single-pass route)
  RouteStage st;
  st.r0 = reinterpret_cast<float4*>(s_dyn);
  st.r1 = st.r0 + kRouteStage;
  st.r2 = st.r1 + kRouteStage;
  st.rect = reinterpret_cast<uint2*>(st.r2 + kRouteStage);
  st.key = reinterpret_cast<uint32_t*>(st.rect + kRouteStage);
  st.sup = st.key + kRouteStage;
  const int tid = threadIdx.x;
  const uint32_t i = blockIdx.x * kRouteThreads + tid;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t key = kCulledKey;
  if (i < n) key = depth_key[i];
  const bool live = key != kCulledKey;
  uint2 rc = make_uint2(0u, 0u);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
  if (live) {
    rc = rect[i];
    if (REC) { a0 = rec0[i]; a1 = rec1[i]; a2 = rec2[i]; }
  }
  const int tu0 = rc.x & 0xFFFF, tu1 = rc.x >> 16, tv0 = rc.y & 0xFFFF, tv1 = rc.y >> 16;
  for (int b = 0; b < rp.world; ++b) {
    const uint32_t m = __ballot_sync(0xffffffffu, route_test(live, tv0, tv1, rp.row[b], rp.row[b + 1]).hit);
    if (lane == 0) s_c[b][warp] = __popc(m);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (int b = 0; b < rp.world; ++b) {
      s_off[b] = run;
#pragma unroll
      for (int w = 0; w < kRouteWarps; ++w) run += s_c[b][w];
    }
    s_off[rp.world] = run;
  }
  __syncthreads();
  const uint32_t total = s_off[rp.world];
  if (total == 0) return;
  const bool single = total <= (uint32_t)kRouteStage;
  for (int b = 0; b < rp.world; ++b) {
    const uint32_t off_b = s_off[b], cnt = s_off[b + 1] - off_b;
    if (cnt == 0) continue;                // block-uniform
    const RouteHit h = route_test(live, tv0, tv1, rp.row[b], rp.row[b + 1]);
    const uint32_t m = __ballot_sync(0xffffffffu, h.hit);
    if (h.hit) {
      uint32_t k = (single ? off_b : 0u) + __popc(m & lt);
      for (int w = 0; w < warp; ++w) k += s_c[b][w];
      if (REC) { st.r0[k] = a0; st.r1[k] = a1; st.r2[k] = a2; }
      if (META) {
        st.key[k] = key;
        st.rect[k] = make_uint2(rc.x, (uint32_t)h.lo | ((uint32_t)h.hi << 16));
        st.sup[k] = (uint32_t)((tu1 / kSuperX - tu0 / kSuperX + 1) * (h.hi / kSuperY - h.lo / kSuperY + 1));
      }
    }
    if (single) continue;
    __syncthreads();                       // band by band: cnt <= kRouteThreads <= kRouteStage
    const size_t pos = (size_t)rp.rank * rp.seg_cap + base[(size_t)b * n_blocks + blockIdx.x];
    for (uint32_t k = tid; k < cnt; k += kRouteThreads) route_copy_out<META, REC>(st, rp, b, pos + k, 0u, k);
    __syncthreads();
  }
  if (!single) return;
  __syncthreads();
  for (uint32_t k = tid; k < total; k += kRouteThreads) {
    int b = 0;
    while (s_off[b + 1] <= k) ++b;
    const size_t pos = (size_t)rp.rank * rp.seg_cap + base[(size_t)b * n_blocks + blockIdx.x] + (k - s_off[b]);
    route_copy_out<META, REC>(st, rp, b, pos, 0u, k);
  }
}

static uint32_t route_blocks(int n) { return (uint32_t)(((n > 0 ? n : 1) + kRouteThreads - 1) / kRouteThreads); }

size_t route_scratch_bytes(int n, int world) { return 2 * (size_t)world * route_blocks(n) * sizeof(uint32_t); }

static RouteParams make_route_params(const b200gs_route* r, const FrameLayout& BL) {
  RouteParams p;
  p.world = r->world; p.rank = r->rank; p.seg_cap = r->seg_capacity;
  for (int q = 0; q <= B200GS_MAX_PEERS; ++q) p.row[q] = q <= r->world ? r->band_row[q] : r->band_row[r->world];
  for (int q = 0; q < B200GS_MAX_PEERS; ++q) p.ws[q] = q < r->world ? reinterpret_cast<char*>(r->band_ws[q]) : nullptr;
  p.rec0 = BL.rec0; p.rec1 = BL.rec1; p.rec2 = BL.rec2; p.depth_key = BL.depth_key; p.rect = BL.rect;
  p.super_touched = BL.super_touched; p.route_in = BL.header + kRouteInOffset;
  return p;
}

// slice_ws: the slice's frame workspace after launch_preprocess_fwd (layout SL, n entries); BL: layout of the band
// workspaces; scratch: route_scratch_bytes(n, world) bytes of the slice workspace.
// what: kRouteAll = count + scan + everything written; kRouteMeta = count + scan + keys / rects / supertile counts only
// (the records follow with kRouteRecords, which reuses the scanned offsets in `scratch` - typically on another stream,
// beside the destination's depth sort).
template <bool META, bool REC>
static cudaError_t launch_route_write(uint32_t un, const void* slice_ws, const FrameLayout& SL, const RouteParams& p,
                                      uint32_t n_blocks, const uint32_t* counts, cudaStream_t s) {
  static std::atomic<uint64_t> attr_set{0};
  once_per_device(attr_set, [] {
    cudaFuncSetAttribute(route_write_kernel<META, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRouteStageBytes);
  });
  route_write_kernel<META, REC><<<(int)n_blocks, kRouteThreads, kRouteStageBytes, s>>>(
      un, ws_ptr<uint32_t>(slice_ws, SL.depth_key), ws_ptr<uint2>(slice_ws, SL.rect), ws_ptr<float4>(slice_ws, SL.rec0),
      ws_ptr<float4>(slice_ws, SL.rec1), ws_ptr<float4>(slice_ws, SL.rec2), p, n_blocks, counts);
  return cudaGetLastError();
}

cudaError_t launch_route_slice(int n, const void* slice_ws, const FrameLayout& SL, const b200gs_route* route,
                               const FrameLayout& BL, void* scratch, size_t scratch_bytes, int what, cudaStream_t s) {
  if (scratch_bytes < route_scratch_bytes(n, route->world)) return cudaErrorInvalidValue;
  const RouteParams p = make_route_params(route, BL);
  const uint32_t n_blocks = route_blocks(n);
  uint32_t* counts = reinterpret_cast<uint32_t*>(scratch);
  uint32_t* tiles = counts + (size_t)route->world * n_blocks;
  const uint32_t un = (uint32_t)(n > 0 ? n : 0);
  const int grid = (int)n_blocks;
  if (what == kRouteRecords) return launch_route_write<false, true>(un, slice_ws, SL, p, n_blocks, counts, s);
  route_count_kernel<<<grid, kRouteThreads, 0, s>>>(un, ws_ptr<uint32_t>(slice_ws, SL.depth_key),
                                                    ws_ptr<uint2>(slice_ws, SL.rect), p, n_blocks, counts, tiles);
  route_scan_kernel<<<route->world, kRouteScanThreads, 0, s>>>(n_blocks, counts, tiles, p);
  if (what == kRouteMeta) return launch_route_write<true, false>(un, slice_ws, SL, p, n_blocks, counts, s);
  const char* variant = getenv("B200GS_ROUTE_WRITE");
  if (variant && !strcmp(variant, "warp")) {
    route_write_warp_kernel<<<grid, kRouteThreads, 0, s>>>(un, ws_ptr<uint32_t>(slice_ws, SL.depth_key),
                                                           ws_ptr<uint2>(slice_ws, SL.rect), ws_ptr<float4>(slice_ws, SL.rec0),
                                                           ws_ptr<float4>(slice_ws, SL.rec1), ws_ptr<float4>(slice_ws, SL.rec2),
                                                           p, n_blocks, counts);
    return cudaGetLastError();
  }
  return launch_route_write<true, true>(un, slice_ws, SL, p, n_blocks, counts, s);
}

// ---- destination side -------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kGatherItems = 8;
constexpr int kGatherTile = kGatherThreads * kGatherItems;

__global__ void __launch_bounds__(kGatherThreads) gather_routed_kernel(int world, uint32_t seg_cap,
                                                                       const uint2* __restrict__ route_in,
                                                                       const uint32_t* __restrict__ depth_key,
                                                                       uint32_t* __restrict__ out_keys,
                                                                       uint32_t* __restrict__ out_ids,
                                                                       b200gs_frame_stats* __restrict__ stats,
                                                                       uint32_t* __restrict__ depth_hist, DepthKeyPlan kp) {
  __shared__ uint32_t s_dh[kSortMaxPasses * kSortMaxRadix];
  const int seg = blockIdx.y, tid = threadIdx.x;
  uint32_t off = 0, cnt = 0, total = 0, ttotal = 0;
  for (int q = 0; q < world; ++q) {
    const uint2 v = route_in[q];
    const uint32_t c = v.x < seg_cap ? v.x : seg_cap;
    if (q < seg) off += c;
    if (q == seg) cnt = c;
    total += c;
    ttotal += v.y;
  }
  if (blockIdx.x == 0 && seg == 0 && tid == 0) {
    stats->n_sorted = total;
    stats->n_visible = total;
    stats->n_in_frustum = total;
    stats->n_isect = ttotal;
  }
  const uint32_t j0 = blockIdx.x * kGatherTile;
  if (j0 >= cnt) return;
  for (int i = tid; i < kSortMaxPasses * kSortMaxRadix; i += kGatherThreads) s_dh[i] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kGatherItems; ++k) {
    const uint32_t j = j0 + k * kGatherThreads + tid;
    if (j < cnt) {
      const uint32_t pos = (uint32_t)seg * seg_cap + j;
      const uint32_t key = depth_key[pos];
      out_keys[off + j] = key;
      out_ids[off + j] = pos;
      depth_hist_add(s_dh, kp, key);
    }
  }
  __syncthreads();
  for (int i = tid; i < kSortMaxPasses * kSortMaxRadix; i += kGatherThreads) {
    const uint32_t c = s_dh[i];
    if (c) atomicAdd(&depth_hist[i], c);
  }
}

cudaError_t launch_gather_routed(int world, uint32_t seg_cap, void* band_ws, const FrameLayout& BL, uint32_t* out_keys,
                                 uint32_t* out_ids, uint32_t* depth_hist, const DepthKeyPlan& kp, cudaStream_t s) {
  dim3 grid((seg_cap + kGatherTile - 1) / kGatherTile, world);      // blocks behind a segment's entries leave at once
  gather_routed_kernel<<<grid, kGatherThreads, 0, s>>>(world, seg_cap,
                                                       ws_ptr<uint2>(band_ws, BL.header + kRouteInOffset),
                                                       ws_ptr<uint32_t>(band_ws, BL.depth_key), out_keys, out_ids,
                                                       ws_ptr<b200gs_frame_stats>(band_ws, BL.header), depth_hist, kp);
  return cudaGetLastError();
}

}  // namespace gs
