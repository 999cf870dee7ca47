// Tile binning, two levels.
//
// Replaces (reference): render.py:260-303 - dense-mask expansion (S12), the composite (tile, depth-rank)
// sort (S13) and unique_consecutive + cumsum (S14).  Integer work, bit-exact by construction.
//
// The reference sorts one key per (tile, Gaussian) intersection.  Here the expansion is hierarchical:
//   1. emit one pair per (SUPERTILE, Gaussian) in depth order - a supertile is 8x4 tiles (128x64 px), so
//      there are ~3.7x fewer pairs than intersections at the headline workload;
//   2. one stable radix pass on the supertile id (<= 8 bits up to 2040x... see api.cu) groups them; every
//      supertile's list is still in depth order;
//   3. one CTA per supertile streams its list and splits it 32 ways into the per-tile lists with warp
//      ballots (a stable multi-split: rank of an entry in tile t = number of earlier entries covering t),
//      first counting (tile ranges), then writing.
// The per-tile depth-sorted lists that come out are exactly the reference's; only their placement in the
// list buffer is grouped by supertile instead of by ascending tile id (the blend only needs ranges[t]).
#include <stdlib.h>

#include "common.cuh"

namespace gs {

constexpr int kEmitBlock = 256;
constexpr int kSplitParts = 4;     // CTAs per supertile of a whole frame: each takes a contiguous part of the supertile's list
constexpr int kSplitPartsMax = 16; // a band has few supertiles with the same long lists: more parts (launch_split_super)
constexpr int kSuperTiles = kSuperX * kSuperY;   // 32: one bit per tile in a 32-bit mask

// One thread per depth rank r.  Gaussian id = order[r]; its supertile pairs go to [offsets[r], +count).
__global__ void __launch_bounds__(kEmitBlock) emit_super_kernel(int n, const uint32_t* __restrict__ order,
                                                                const uint32_t* __restrict__ offsets,
                                                                const uint32_t* __restrict__ super_touched,
                                                                const uint2* __restrict__ rect, int super_x,
                                                                uint32_t capacity, uint32_t* __restrict__ keys,
                                                                uint32_t* __restrict__ vals,
                                                                b200gs_frame_stats* __restrict__ stats) {
  const int r = blockIdx.x * kEmitBlock + threadIdx.x;
  if (r == 0 && stats->n_isect > capacity) stats->overflow = 1u;   // the per-tile lists would not fit either
  if (r >= n) return;
  const uint32_t id = order[r];
  const uint32_t cnt = super_touched[id];
  if (cnt == 0) return;
  uint32_t off = offsets[r];
  if (off + cnt > capacity || off + cnt < off) {   // does not fit: flag it, never write out of bounds
    stats->overflow = 1u;
    return;
  }
  const uint2 rc = rect[id];
  const int sx0 = (rc.x & 0xFFFF) / kSuperX, sx1 = (rc.x >> 16) / kSuperX;
  const int sy0 = (rc.y & 0xFFFF) / kSuperY, sy1 = (rc.y >> 16) / kSuperY;
  for (int sy = sy0; sy <= sy1; ++sy)
    for (int sx = sx0; sx <= sx1; ++sx) {
      keys[off] = (uint32_t)(sy * super_x + sx);
      vals[off] = id;
      ++off;
    }
}

cudaError_t launch_emit_super(int n, const uint32_t* order, const uint32_t* offsets, const uint32_t* super_touched,
                              const uint2* rect, int super_x, uint32_t capacity, uint32_t* keys, uint32_t* vals,
                              b200gs_frame_stats* stats, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  emit_super_kernel<<<ceil_div(n, kEmitBlock), kEmitBlock, 0, s>>>(n, order, offsets, super_touched, rect, super_x,
                                                                  capacity, keys, vals, stats);
  return cudaGetLastError();
}

// first index i in sorted keys[0, count) with keys[i] >= v
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ keys, uint32_t count, uint32_t v) {
  uint32_t lo = 0, hi = count;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// bit (ty*8 + tx) set for every tile of this supertile that the Gaussian's (already clipped) tile rect covers
__device__ __forceinline__ uint32_t tile_mask(uint2 rc, int sx, int sy) {
  const int tu0 = rc.x & 0xFFFF, tu1 = rc.x >> 16, tv0 = rc.y & 0xFFFF, tv1 = rc.y >> 16;
  const int x0 = max(tu0 - sx * kSuperX, 0), x1 = min(tu1 - sx * kSuperX, kSuperX - 1);
  const int y0 = max(tv0 - sy * kSuperY, 0), y1 = min(tv1 - sy * kSuperY, kSuperY - 1);
  if (x1 < x0 || y1 < y0) return 0u;
  const uint32_t row = ((1u << (x1 - x0 + 1)) - 1u) << x0;     // 8 bits
  uint32_t m = 0;
  for (int y = y0; y <= y1; ++y) m |= row << (y * kSuperX);
  return m;
}

// 32x32 bit-matrix transpose across a warp: lane L gives word m_L, lane t receives the word whose bit L
// is bit t of m_L - i.e. lane t ends up with the ballot of "entry covers tile t" (5 shuffles instead of
// 32 votes).
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t m, int lane) {
#define GS_TSTEP(J, MASK)                                                                         \
  {                                                                                               \
    const uint32_t x = __shfl_xor_sync(0xffffffffu, m, J);                                        \
    m = (lane & J) ? ((m & ~(MASK)) | ((x & ~(MASK)) >> J)) : ((m & (MASK)) | ((x & (MASK)) << J)); \
  }
  GS_TSTEP(16, 0x0000FFFFu)
  GS_TSTEP(8, 0x00FF00FFu)
  GS_TSTEP(4, 0x0F0F0F0Fu)
  GS_TSTEP(2, 0x33333333u)
  GS_TSTEP(1, 0x55555555u)
#undef GS_TSTEP
  return m;
}

constexpr int kSplitPer = 4;                                  // entries per thread and iteration

// One CTA per supertile.  WRITE == false: count the entries of each of its 32 tiles -> tile_count.
// WRITE == true: derive the tile ranges from tile_count and write the per-tile lists.
// kSplitThreads: CTA size (256; a 128-thread variant - 512 entries per iteration for list quarters of ~650 entries -
// was measured and is slower, see launch_split_super).
template <bool WRITE, int kSplitThreads>
__global__ void __launch_bounds__(kSplitThreads) split_super_kernel(const uint32_t* __restrict__ keys,
                                                                    const uint32_t* __restrict__ vals,
                                                                    const uint2* __restrict__ rect,
                                                                    uint32_t capacity,
                                                                    const b200gs_frame_stats* __restrict__ stats,
                                                                    int super_x, int super_y0, int tiles_x, int tiles_y,
                                                                    int parts, uint32_t* __restrict__ tile_count,
                                                                    uint32_t* __restrict__ part_total,
                                                                    uint2* __restrict__ ranges,
                                                                    uint32_t* __restrict__ lists) {
  constexpr int kSplitWarps = kSplitThreads / 32;
  constexpr int kSplitChunk = kSplitThreads * kSplitPer;        // entries per iteration
  constexpr int kSplitRows = kSplitWarps * kSplitPer;           // warp-chunks per iteration, in list order
  __shared__ uint32_t s_wc[kSplitRows][kSuperTiles];    // per-warp-chunk tile counts, then exclusive offsets
  __shared__ uint32_t s_run[kSuperTiles];               // entries already placed per tile
  __shared__ uint32_t s_base[kSuperTiles];              // start of each tile's list in `lists`
  __shared__ uint32_t s_red[kSplitWarps];
  __shared__ uint32_t s_seg[2];
  // s = band-local supertile id (the sort key); (sx, sy) = its position in the frame
  const int s = blockIdx.x / parts, part = blockIdx.x % parts, sx = s % super_x, sy = super_y0 + s / super_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t count = stats->n_super;
  if (count > capacity || stats->overflow) count = 0;
  if (tid < 2) s_seg[tid] = lower_bound_u32(keys, count, (uint32_t)s + tid);
  if (tid < kSuperTiles) s_run[tid] = 0;
  if (WRITE) {
    // start of this supertile's region = number of intersections of all earlier supertiles = sum of the per-CTA
    // totals the counting pass left behind (one word per CTA: summing the per-tile counts here instead made every
    // CTA read up to 128 words per earlier supertile - 67 M loads per headline frame)
    uint32_t acc = 0;
    for (int i = tid; i < s * parts; i += kSplitThreads) acc += part_total[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      uint32_t base = 0;
#pragma unroll
      for (int w = 0; w < kSplitWarps; ++w) base += s_red[w];
      uint32_t c = 0, before = 0;     // whole-supertile count of tile `lane`, and the part of it in earlier parts
#pragma unroll 4
      for (int q = 0; q < parts; ++q) {
        const uint32_t v = tile_count[(s * parts + q) * kSuperTiles + lane];
        c += v;
        if (q < part) before += v;
      }
      s_run[lane] = before;
      uint32_t inc = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      const uint32_t start = base + inc - c;
      s_base[lane] = start;
      const int gx = sx * kSuperX + (lane & (kSuperX - 1)), gy = sy * kSuperY + (lane / kSuperX);
      if (part == 0 && gx < tiles_x && gy < tiles_y) ranges[gy * tiles_x + gx] = make_uint2(start, start + c);
    }
  }
  __syncthreads();
  // this CTA's quarter of the supertile's list (boundaries are multiples of 32 so warp-chunks stay whole)
  const uint32_t seg_lo = s_seg[0], seg_hi = s_seg[1];
  const uint32_t per = ((seg_hi - seg_lo + (uint32_t)parts - 1) / (uint32_t)parts + 31u) & ~31u;
  const uint32_t lo = min(seg_lo + part * per, seg_hi), hi = min(lo + per, seg_hi);
  const uint32_t lane_lt = (1u << lane) - 1u;
  uint32_t my_total = 0;   // WRITE == false: lane t of every warp accumulates the count of tile t
  for (uint32_t chunk = lo; chunk < hi; chunk += kSplitChunk) {
    uint32_t id[kSplitPer], m[kSplitPer], T[kSplitPer];
#pragma unroll
    for (int k = 0; k < kSplitPer; ++k) {
      const uint32_t e = chunk + k * kSplitThreads + tid;     // list order: (k, warp, lane)
      id[k] = (e < hi) ? vals[e] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kSplitPer; ++k) {
      const uint32_t e = chunk + k * kSplitThreads + tid;
      m[k] = (e < hi) ? tile_mask(rect[id[k]], sx, sy) : 0u;
      T[k] = warp_bit_transpose(m[k], lane);                  // lane t: which lanes' entries cover tile t
    }
    if (!WRITE) {
#pragma unroll
      for (int k = 0; k < kSplitPer; ++k) my_total += __popc(T[k]);
    } else {
#pragma unroll
      for (int k = 0; k < kSplitPer; ++k) s_wc[k * kSplitWarps + warp][lane] = __popc(T[k]);
      __syncthreads();
      // exclusive offsets over the warp-chunks (list order) for every tile + running totals
      if (warp == 0) {
        uint32_t run = s_run[lane];
#pragma unroll 8
        for (int r = 0; r < kSplitRows; ++r) {
          const uint32_t c = s_wc[r][lane];
          s_wc[r][lane] = run;
          run += c;
        }
        s_run[lane] = run;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kSplitPer; ++k) {
        uint32_t rest = m[k];
        while (__any_sync(0xffffffffu, rest != 0u)) {
          const int t = rest ? (__ffs(rest) - 1) : 0;
          const uint32_t b = __shfl_sync(0xffffffffu, T[k], t);     // ballot of tile t within this warp-chunk
          if (rest) {
            lists[s_base[t] + s_wc[k * kSplitWarps + warp][t] + __popc(b & lane_lt)] = id[k];
            rest &= rest - 1u;
          }
        }
      }
      __syncthreads();   // s_wc is rewritten by the next iteration
    }
  }
  if (!WRITE) {
    if (tid < kSuperTiles) s_run[tid] = 0;
    __syncthreads();
    if (my_total) atomicAdd(&s_run[lane], my_total);
    __syncthreads();
    if (tid < kSuperTiles) {
      const uint32_t c = s_run[tid];
      tile_count[(s * parts + part) * kSuperTiles + tid] = c;
      const uint32_t tot = __reduce_add_sync(0xffffffffu, c);      // warp 0 = the 32 tiles
      if (tid == 0) part_total[s * parts + part] = tot;
    }
  }
}

cudaError_t launch_split_super(bool write, const uint32_t* keys, const uint32_t* vals, const uint2* rect,
                               uint32_t capacity, const b200gs_frame_stats* stats, int super_x, int super_y0, int super_y,
                               int tiles_x, int tiles_y, uint32_t* tile_count, uint2* ranges, uint32_t* lists,
                               cudaStream_t s) {
  // CTAs per supertile: 4 for a whole frame (2 040 CTAs at 1080p).  A band of tile rows has few supertiles whose lists
  // are as long as ever - an eighth of a 4K frame: 150 supertiles x ~7 000 pairs, 600 CTAs walking two to three 1024-entry
  // chunks each, 72 us for count + write - so it gets more parts, up to what the workspace section (sized for the whole
  // frame's supertiles x 4) holds.  Measured on that band (tools/routed_probe.py): 4 parts 73 us, 8 parts 56 us, 13 parts
  // 62 us (every CTA of the write pass sums the totals of all CTAs before it) - hence one wave of resident CTAs.
  const int n_super = super_x * super_y;
  if (n_super <= 0) return cudaSuccess;
  const int full = ceil_div(tiles_x, kSuperX) * ceil_div(tiles_y, kSuperY);
  int parts = kSplitParts;
  static const int parts_env = getenv("B200GS_SPLIT_PARTS") ? atoi(getenv("B200GS_SPLIT_PARTS")) : 0;
  const int room = kSplitParts * full / n_super;
  int want = parts_env > 0 ? parts_env : 1200 / n_super;
  if (want > kSplitPartsMax) want = kSplitPartsMax;
  if (want > room) want = room;
  if (want > parts) parts = want;
  const int grid = n_super * parts;
  uint32_t* part_total = tile_count + (size_t)grid * kSuperTiles;     // one word per CTA, behind the per-tile counts
  // measured on the headline frame (count + write): 256 threads 64.9 us, 128 threads 81.7 us
  static const int threads = getenv("B200GS_SPLIT_THREADS") ? atoi(getenv("B200GS_SPLIT_THREADS")) : 256;
#define GS_SPLIT_ARGS keys, vals, rect, capacity, stats, super_x, super_y0, tiles_x, tiles_y, parts, tile_count, part_total, ranges, lists
  if (threads == 256) {
    if (write) split_super_kernel<true, 256><<<grid, 256, 0, s>>>(GS_SPLIT_ARGS);
    else split_super_kernel<false, 256><<<grid, 256, 0, s>>>(GS_SPLIT_ARGS);
  } else {
    if (write) split_super_kernel<true, 128><<<grid, 128, 0, s>>>(GS_SPLIT_ARGS);
    else split_super_kernel<false, 128><<<grid, 128, 0, s>>>(GS_SPLIT_ARGS);
  }
#undef GS_SPLIT_ARGS
  return cudaGetLastError();
}

// Parity introspection: tile id of every list entry (the lists are grouped by supertile, so the tile ids
// are not stored anywhere).
__global__ void fill_list_tiles_kernel(const uint2* __restrict__ ranges, int n_tiles, uint32_t count,
                                       int32_t* __restrict__ list_tile) {
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const uint2 r = ranges[t];
  for (uint32_t i = r.x + threadIdx.x; i < r.y && i < count; i += blockDim.x) list_tile[i] = t;
}

cudaError_t launch_fill_list_tiles(const uint2* ranges, int n_tiles, uint32_t count, int32_t* list_tile,
                                   cudaStream_t s) {
  if (n_tiles <= 0 || count == 0) return cudaSuccess;
  fill_list_tiles_kernel<<<n_tiles, 64, 0, s>>>(ranges, n_tiles, count, list_tile);
  return cudaGetLastError();
}

}  // namespace gs
