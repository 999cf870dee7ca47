// Tile binning: emit (tile id, Gaussian id) pairs in depth order and extract per-tile ranges.
//
// Replaces (reference): render.py:260-281 (dense-mask expansion, S12) and render.py:300-303
// (unique_consecutive + cumsum, S14).  Integer work, HBM-bound, bit-exact by construction.
#include "common.cuh"

namespace gs {

constexpr int kEmitBlock = 256;

// One thread per depth rank r.  Gaussian id = order[r]; its pairs go to [offsets[r], offsets[r]+count).
// Emitting in depth-rank order is what lets the following sort be a stable sort on the tile id alone
// (the reference sorts the composite key tile*(V+1)+rank, render.py:289-292).
__global__ void __launch_bounds__(kEmitBlock) emit_pairs_kernel(int n, const uint32_t* __restrict__ order,
                                                                const uint32_t* __restrict__ offsets,
                                                                const uint32_t* __restrict__ tiles_touched,
                                                                const uint2* __restrict__ rect, int tiles_x,
                                                                uint32_t capacity, uint32_t* __restrict__ keys,
                                                                uint32_t* __restrict__ vals,
                                                                b200gs_frame_stats* __restrict__ stats) {
  const int r = blockIdx.x * kEmitBlock + threadIdx.x;
  if (r >= n) return;
  const uint32_t id = order[r];
  const uint32_t cnt = tiles_touched[id];
  if (cnt == 0) return;
  uint32_t off = offsets[r];
  if (off + cnt > capacity || off + cnt < off) {   // does not fit: flag it, never write out of bounds
    stats->overflow = 1u;
    return;
  }
  const uint2 rc = rect[id];
  const int tu0 = rc.x & 0xFFFF, tu1 = rc.x >> 16, tv0 = rc.y & 0xFFFF, tv1 = rc.y >> 16;
  for (int ty = tv0; ty <= tv1; ++ty)
    for (int tx = tu0; tx <= tu1; ++tx) {
      keys[off] = (uint32_t)(ty * tiles_x + tx);
      vals[off] = id;
      ++off;
    }
}

cudaError_t launch_emit_pairs(int n, const uint32_t* order, const uint32_t* offsets, const uint32_t* tiles_touched,
                              const uint2* rect, int tiles_x, uint32_t capacity, uint32_t* keys, uint32_t* vals,
                              b200gs_frame_stats* stats, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  emit_pairs_kernel<<<ceil_div(n, kEmitBlock), kEmitBlock, 0, s>>>(n, order, offsets, tiles_touched, rect, tiles_x,
                                                                  capacity, keys, vals, stats);
  return cudaGetLastError();
}

// ranges[t] = (first index, one-past-last index) of tile t in the sorted key list; (0,0) when empty.
__global__ void __launch_bounds__(kEmitBlock) tile_ranges_kernel(const uint32_t* __restrict__ keys,
                                                                 uint32_t capacity,
                                                                 const b200gs_frame_stats* __restrict__ stats,
                                                                 uint2* __restrict__ ranges) {
  uint32_t count = stats->n_isect;
  if (count > capacity || stats->overflow) count = 0;
  const uint32_t stride = gridDim.x * kEmitBlock;
  for (uint32_t i = blockIdx.x * kEmitBlock + threadIdx.x; i < count; i += stride) {
    const uint32_t t = keys[i];
    if (i == 0 || keys[i - 1] != t) ranges[t].x = i;
    if (i + 1 == count || keys[i + 1] != t) ranges[t].y = i + 1;
  }
}

cudaError_t launch_tile_ranges(const uint32_t* keys, uint32_t capacity, const b200gs_frame_stats* stats,
                               uint2* ranges, int n_tiles, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)n_tiles * sizeof(uint2), s);
  if (e != cudaSuccess) return e;
  if (capacity == 0) return cudaSuccess;
  int grid = (int)((capacity + kEmitBlock - 1) / kEmitBlock);
  if (grid > 148 * 16) grid = 148 * 16;
  tile_ranges_kernel<<<grid, kEmitBlock, 0, s>>>(keys, capacity, stats, ranges);
  return cudaGetLastError();
}

}  // namespace gs
