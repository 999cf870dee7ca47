// Integer stages of the binning pipeline: single-pass exclusive scan (decoupled look-back) and a
// stable onesweep LSD radix sort of (u32 key, u32 value) pairs.  Hand-written; no CUB.
//
// Replaces (reference): torch.argsort (render.py:211), torch.sort on composite keys (render.py:292),
// cumsum (render.py:302).  Roofline: HBM (streaming reads, scattered-but-run-coalesced writes).
//
// Both kernels hand out "virtual" block ids through an atomic ticket so that block v only ever waits
// on blocks < v that are already resident: the look-back cannot deadlock whatever the hardware's
// block scheduling order is.
#include <stdlib.h>

#include "common.cuh"

namespace gs {

// =================================================================================================
// Exclusive scan
// =================================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;   // 4096

// status word: [63:32] flag (0 = empty, 1 = tile aggregate, 2 = inclusive prefix), [31:0] value
__device__ __forceinline__ void st_status64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_status64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_status32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

size_t scan_scratch_bytes(uint32_t n) {
  const size_t tiles = (n + kScanTile - 1) / kScanTile + 1;
  return 256 + tiles * 8;
}

// Decoupled look-back executed by one warp: posts this tile's aggregate, sums the aggregates of the
// predecessors (32 per step, lane l inspects tile j - l) down to the nearest inclusive prefix, posts the
// inclusive prefix and returns the exclusive one (valid in every lane).
__device__ __forceinline__ uint32_t scan_lookback(unsigned long long* status, uint32_t tile, uint32_t tile_sum, int lane) {
  uint32_t prefix = 0;
  if (tile == 0) {
    if (lane == 0) st_status64(status + tile, (2ull << 32) | tile_sum);
    return 0;
  }
  if (lane == 0) st_status64(status + tile, (1ull << 32) | tile_sum);
  int j = (int)tile - 1;
  while (true) {
    const int idx = j - lane;
    const unsigned long long s = (idx >= 0) ? ld_status64(status + idx) : (2ull << 32);  // virtual prefix 0
    const uint32_t flag = (uint32_t)(s >> 32);
    const unsigned ready = __ballot_sync(0xffffffffu, flag != 0);
    const unsigned pref = __ballot_sync(0xffffffffu, flag == 2);
    const int p = pref ? (__ffs(pref) - 1) : 32;                  // nearest inclusive prefix
    const unsigned need = (p >= 31) ? 0xffffffffu : ((2u << p) - 1u);
    if ((ready & need) != need) { __nanosleep(20); continue; }    // someone before it is not posted yet
    uint32_t c = (lane <= p) ? (uint32_t)s : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    prefix += c;
    if (p < 32) break;
    j -= 32;
  }
  if (lane == 0) st_status64(status + tile, (2ull << 32) | (uint32_t)(prefix + tile_sum));
  return prefix;
}

// out[i] = sum_{j<i} in[gather ? gather[j] : j];  *total_out = sum of everything.
__global__ void __launch_bounds__(kScanThreads) exclusive_scan_kernel(const uint32_t* __restrict__ in,
                                                                      const uint32_t* __restrict__ gather,
                                                                      uint32_t* __restrict__ out, uint32_t n,
                                                                      uint32_t* __restrict__ total_out,
                                                                      uint32_t* ticket,
                                                                      unsigned long long* status) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * kScanTile + tid * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint32_t idx = base + k;
    uint32_t x = 0;
    if (idx < n) x = gather ? in[gather[idx]] : in[idx];
    v[k] = sum;          // exclusive within the thread
    sum += x;
  }
  // warp inclusive scan of thread sums
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t warp_off = 0, tile_sum = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) warp_off += t;
    tile_sum += t;
  }
  if (warp == 0) {
    const uint32_t prefix = scan_lookback(status, tile, tile_sum, lane);
    if (lane == 0) {
      s_prefix = prefix;
      if (total_out && tile == (n - 1) / kScanTile) *total_out = prefix + tile_sum;
    }
  }
  __syncthreads();
  const uint32_t off = s_prefix + warp_off + (inc - sum);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint32_t idx = base + k;
    if (idx < n) out[idx] = off + v[k];
  }
}

cudaError_t launch_exclusive_scan(const uint32_t* in, const uint32_t* gather_idx, uint32_t* out, uint32_t n,
                                  uint32_t* total_out, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  if (scratch_bytes < scan_scratch_bytes(n)) return cudaErrorInvalidValue;
  if (n == 0) {
    if (total_out) return cudaMemsetAsync(total_out, 0, 4, s);
    return cudaSuccess;
  }
  cudaError_t e = cudaMemsetAsync(scratch, 0, scan_scratch_bytes(n), s);
  if (e != cudaSuccess) return e;
  uint32_t* ticket = reinterpret_cast<uint32_t*>(scratch);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(scratch) + 256);
  const int grid = (int)((n + kScanTile - 1) / kScanTile);
  exclusive_scan_kernel<<<grid, kScanThreads, 0, s>>>(in, gather_idx, out, n, total_out, ticket, status);
  return cudaGetLastError();
}

// =================================================================================================
// Stable compaction of the live depth keys (tile-row bands: most Gaussians miss the band, so the depth sort
// should not carry them): (key, id) of every key != culled, in index order, and their number.
// One pass, decoupled look-back on the per-block counts.
// =================================================================================================
constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 16;
constexpr int kCompactTile = kCompactThreads * kCompactItems;   // 4096 keys per block

size_t compact_scratch_bytes(uint32_t n) { return 256 + ((size_t)(n + kCompactTile - 1) / kCompactTile + 1) * 8; }

__device__ __forceinline__ uint32_t eff_count(uint32_t n_host, const uint32_t* n_dev);

__global__ void __launch_bounds__(kCompactThreads) compact_keys_kernel(const uint32_t* __restrict__ keys,
                                                                       const uint32_t* __restrict__ ids_in,
                                                                       uint32_t n_host, const uint32_t* __restrict__ n_dev,
                                                                       uint32_t* __restrict__ out_keys,
                                                                       uint32_t* __restrict__ out_ids,
                                                                       uint32_t* __restrict__ count_out,
                                                                       uint32_t* ticket, unsigned long long* status) {
  __shared__ uint32_t s_warp[kCompactThreads / 32];
  __shared__ uint32_t s_tile, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = eff_count(n_host, n_dev);
  if (n == 0) { if (blockIdx.x == 0 && tid == 0) *count_out = 0; return; }
  if ((uint32_t)blockIdx.x * (uint32_t)kCompactTile >= n) return;       // surplus CTAs (grid sized for the upper bound)
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * kCompactTile + tid * kCompactItems;     // thread-contiguous: index order is kept
  uint32_t k[kCompactItems];
  if (base + kCompactItems <= n) {
#pragma unroll
    for (int q = 0; q < kCompactItems / 4; ++q) {
      const uint4 v = reinterpret_cast<const uint4*>(keys + base)[q];
      k[4 * q] = v.x; k[4 * q + 1] = v.y; k[4 * q + 2] = v.z; k[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kCompactItems; ++i) k[i] = (base + i < n) ? keys[base + i] : kCulledKey;
  }
  uint32_t cnt = 0;
#pragma unroll
  for (int i = 0; i < kCompactItems; ++i) cnt += (k[i] != kCulledKey) ? 1u : 0u;
  uint32_t inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t warp_off = 0, tile_sum = 0;
#pragma unroll
  for (int w = 0; w < kCompactThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) warp_off += t;
    tile_sum += t;
  }
  if (warp == 0) {
    const uint32_t prefix = scan_lookback(status, tile, tile_sum, lane);
    if (lane == 0) {
      s_prefix = prefix;
      if (tile == (n - 1) / kCompactTile) *count_out = prefix + tile_sum;
    }
  }
  __syncthreads();
  uint32_t off = s_prefix + warp_off + (inc - cnt);
#pragma unroll
  for (int i = 0; i < kCompactItems; ++i)
    if (k[i] != kCulledKey) {
      out_keys[off] = k[i];
      out_ids[off] = ids_in ? ids_in[base + i] : base + i;
      ++off;
    }
}

cudaError_t launch_compact_keys(const uint32_t* keys, const uint32_t* ids_in, uint32_t n, const uint32_t* n_dev,
                                uint32_t* out_keys, uint32_t* out_ids, uint32_t* count_out, void* scratch,
                                size_t scratch_bytes, cudaStream_t s) {
  if (n == 0) return cudaMemsetAsync(count_out, 0, 4, s);
  if (scratch_bytes < compact_scratch_bytes(n)) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(scratch, 0, compact_scratch_bytes(n), s);
  if (e != cudaSuccess) return e;
  uint32_t* ticket = reinterpret_cast<uint32_t*>(scratch);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(scratch) + 256);
  compact_keys_kernel<<<(n + kCompactTile - 1) / kCompactTile, kCompactThreads, 0, s>>>(keys, ids_in, n, n_dev, out_keys,
                                                                                      out_ids, count_out, ticket, status);
  return cudaGetLastError();
}

// =================================================================================================
// Onesweep radix sort (8-bit digits)
// =================================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;   // 4096 keys per block
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRadix = 256;
constexpr int kMaxPasses = 4;

// scratch: [0,64): tickets u32[4] | [256, 256+4*256*4): global histograms | then per-pass status arrays
static size_t sort_blocks(uint32_t n) { return (n + kSortTile - 1) / kSortTile; }
size_t sort_scratch_bytes(uint32_t n) {
  return 256 + (size_t)kMaxPasses * kRadix * 4 + (size_t)kMaxPasses * (sort_blocks(n) + 1) * kRadix * 4;
}

struct SortPasses {
  int num;
  int shift[kMaxPasses];
  int bits[kMaxPasses];
};

__device__ __forceinline__ uint32_t eff_count(uint32_t n_host, const uint32_t* n_dev) {
  if (n_dev) { const uint32_t d = *n_dev; return d < n_host ? d : n_host; }
  return n_host;
}

// One read of the keys builds the digit histograms of every pass.
__global__ void __launch_bounds__(kSortThreads) sort_histogram_kernel(const uint32_t* __restrict__ keys,
                                                                      uint32_t n_host,
                                                                      const uint32_t* __restrict__ n_dev,
                                                                      SortPasses sp, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t s_hist[kMaxPasses][kRadix];
  const uint32_t n = eff_count(n_host, n_dev);
  for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kSortThreads) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * kSortThreads;
  for (uint32_t i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride) {
    const uint32_t k = keys[i];
#pragma unroll
    for (int p = 0; p < kMaxPasses; ++p)
      if (p < sp.num) atomicAdd(&s_hist[p][(k >> sp.shift[p]) & ((1u << sp.bits[p]) - 1u)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kSortThreads) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

// exclusive scan of each pass's 256-bin histogram, in place (one block per pass)
__global__ void __launch_bounds__(kRadix) sort_scan_hist_kernel(uint32_t* __restrict__ ghist) {
  __shared__ uint32_t s[kRadix];
  uint32_t* h = ghist + blockIdx.x * kRadix;
  const int t = threadIdx.x;
  const uint32_t mine = h[t];
  s[t] = mine;
  __syncthreads();
  for (int d = 1; d < kRadix; d <<= 1) {
    const uint32_t add = (t >= d) ? s[t - d] : 0u;
    __syncthreads();
    s[t] += add;
    __syncthreads();
  }
  h[t] = s[t] - mine;
}

// One onesweep pass.  288 threads: warps 0-7 (256 threads) rank and move 4096 keys, warp 8 runs the
// decoupled look-back for all 256 digits (8 digits per lane, 4 predecessors per digit in flight) WHILE
// the others do the expensive stable ranking and the shared-memory reorder, so the look-back latency -
// which a synchronised wave of blocks pays in full, wave after wave - is off the critical path.
constexpr int kSortBlock = kSortThreads + 32;
#ifndef ONESWEEP_MIN_BLOCKS
#define ONESWEEP_MIN_BLOCKS 3      // 72 registers (96 with 2): a smaller CTA finds room sooner next to the blend CTAs of the previous frame
#endif
constexpr int kLbBatch = 4;

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(kSortBlock, ONESWEEP_MIN_BLOCKS) onesweep_pass_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, uint32_t n_host, const uint32_t* __restrict__ n_dev, int shift, int bits,
    const uint32_t* __restrict__ ghist /*[256] digit counts of this pass*/, uint32_t* __restrict__ status, uint32_t* ticket) {
  __shared__ uint32_t s_warp_hist[kSortWarps][kRadix + 1];
  __shared__ uint32_t s_keys[kSortTile];
  __shared__ uint32_t s_vals[kSortTile];
  __shared__ uint32_t s_hist[kRadix];          // digit counts of this block (early counts)
  __shared__ uint32_t s_excl[kRadix];          // digit counts of all earlier blocks (look-back result)
  __shared__ uint32_t s_gbase[kRadix];         // exclusive scan of the global digit histogram
  __shared__ uint32_t s_digit_start[kRadix];
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_vbid;

  const uint32_t n = eff_count(n_host, n_dev);
  // The grid is sized for the host's upper bound; with the count on the device (band frames, supertile pairs) most CTAs
  // may have nothing to do.  Exactly ceil(n / tile) CTAs must work and ANY may (virtual ids come from the ticket), so
  // the surplus leaves before taking a ticket or clearing shared memory.
  if ((uint32_t)blockIdx.x * (uint32_t)kSortTile >= n) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_lb = warp == kSortWarps;
  if (tid == 0) s_vbid = atomicAdd(ticket, 1u);
  for (int i = tid; i < kSortWarps * (kRadix + 1); i += kSortBlock) (&s_warp_hist[0][0])[i] = 0;
  if (tid < kRadix) s_hist[tid] = 0;
  __syncthreads();
  const uint32_t vbid = s_vbid;
  const uint32_t base = vbid * (uint32_t)kSortTile;
  if (base >= n) return;   // uniform for the whole block
  const uint32_t mask = (1u << bits) - 1u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const bool full = base + (uint32_t)kSortTile <= n;

  // ---- load (warp-striped: item i of lane l sits at warp_base + 32 i + l) + early digit counts ------------
  const uint32_t warp_base = base + warp * (32 * kSortItems);
  uint32_t key[kSortItems];
  uint32_t val[kSortItems];
  if (!is_lb) {
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      key[i] = (full || idx < n) ? keys_in[idx] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      val[i] = vals_in ? ((full || idx < n) ? vals_in[idx] : 0u) : idx;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      if (full || idx < n) atomicAdd(&s_hist[(key[i] >> shift) & mask], 1u);
    }
  }
  __syncthreads();

  if (is_lb) {
    // ---- global digit offsets: exclusive scan of the 256-bin histogram (lane owns digits 8 lane .. 8 lane + 7) ---
    {
      uint32_t h[8], run = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { h[k] = ghist[8 * lane + k]; run += h[k]; }
      uint32_t inc = run;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      uint32_t acc = inc - run;
#pragma unroll
      for (int k = 0; k < 8; ++k) { s_gbase[8 * lane + k] = acc; acc += h[k]; }
    }
    // ---- decoupled look-back: lane handles digits lane + 32 k ------------------------------------------------
    uint32_t cnt[8], excl[8];
    int j[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cnt[k] = s_hist[lane + 32 * k];
      excl[k] = 0;
      j[k] = (int)vbid - 1;
      st_status32(status + (size_t)vbid * kRadix + lane + 32 * k, ((vbid == 0 ? 2u : 1u) << 30) | cnt[k]);
    }
    unsigned open = (vbid == 0) ? 0u : 0xFFu;
    while (open) {
      uint32_t sv[8][kLbBatch];
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int b = 0; b < kLbBatch; ++b) {
          const int idx = j[k] - b;
          sv[k][b] = ((open >> k) & 1u) && idx >= 0 ? ld_status32(status + (size_t)idx * kRadix + lane + 32 * k)
                                                    : (2u << 30);   // closed chain / before block 0: prefix 0
        }
      bool progressed = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if ((open >> k) & 1u) {
          int used = 0;
          bool found = false;
#pragma unroll
          for (int b = 0; b < kLbBatch; ++b) {
            const uint32_t flag = sv[k][b] >> 30;
            if (!found && used == b && flag != 0) {
              excl[k] += sv[k][b] & 0x3FFFFFFFu;
              used = b + 1;
              found = (flag == 2);
            }
          }
          j[k] -= used;
          progressed |= used > 0;
          if (found) open &= ~(1u << k);
        }
      }
      if (!progressed) __nanosleep(40);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (vbid != 0) st_status32(status + (size_t)vbid * kRadix + lane + 32 * k, (2u << 30) | (excl[k] + cnt[k]));
      s_excl[lane + 32 * k] = excl[k];
    }
  } else {
    // ---- stable ranking (see the comment on match.any + atomics below) -------------------------------------
    // Per round (one key per lane): `match.any` gives the set of lanes holding the same digit ("peers");
    // the lowest peer adds the group size to the warp's digit counter with a shared-memory atomic that
    // returns the running count.  Rounds only depend on each other through those atomics (same-address
    // atomics of one warp retire in program order), so the 16 rounds pipeline.
    uint32_t rank[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      const bool valid = full || idx < n;
      // invalid tail keys get digit 256: they only match each other and never touch a counter
      const uint32_t d = valid ? ((key[i] >> shift) & mask) : (uint32_t)kRadix;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      uint32_t old = 0;
      if (valid && (peers & lane_lt) == 0u) old = atomicAdd(&s_warp_hist[warp][d], (uint32_t)__popc(peers));
      old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
      rank[i] = old + __popc(peers & lane_lt);
    }
    named_bar_sync(1, kSortThreads);
    // ---- per digit (thread d): exclusive offsets across warps, start of the digit in the sorted block -------
    {
      uint32_t cnt = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t t = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = cnt;
        cnt += t;
      }
      uint32_t inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) s_scan[warp] = inc;
      named_bar_sync(1, kSortThreads);
      uint32_t woff = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) woff += (w < warp) ? s_scan[w] : 0u;
      s_digit_start[tid] = woff + inc - cnt;
    }
    named_bar_sync(1, kSortThreads);
    // ---- scatter into shared memory in sorted order -----------------------------------------------------------
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      if (full || idx < n) {
        const uint32_t d = (key[i] >> shift) & mask;
        const uint32_t pos = s_digit_start[d] + s_warp_hist[warp][d] + rank[i];
        s_keys[pos] = key[i];
        s_vals[pos] = val[i];
      }
    }
  }
  __syncthreads();   // reorder done AND look-back done

  // ---- stream out in digit runs: global position = gbase[d] + earlier blocks' count + offset within the digit ----
  if (tid < kRadix) s_hist[tid] = s_gbase[tid] + s_excl[tid] - s_digit_start[tid];   // s_hist reused as delta
  __syncthreads();
  const uint32_t in_block = min((uint32_t)kSortTile, n - base);
  for (uint32_t p = tid; p < in_block; p += kSortBlock) {
    const uint32_t k = s_keys[p];
    const uint32_t dst = s_hist[(k >> shift) & mask] + p;
    keys_out[dst] = k;
    vals_out[dst] = s_vals[p];
  }
}

static SortPasses make_passes(int begin_bit, int end_bit) {
  SortPasses sp;
  sp.num = (end_bit - begin_bit + 7) / 8;
  for (int p = 0; p < kMaxPasses; ++p) { sp.shift[p] = 0; sp.bits[p] = 8; }
  for (int p = 0; p < sp.num; ++p) {
    sp.shift[p] = begin_bit + 8 * p;
    sp.bits[p] = (end_bit - sp.shift[p]) < 8 ? (end_bit - sp.shift[p]) : 8;
  }
  return sp;
}

// =================================================================================================
// Fused binning level 1: exclusive scan of the supertile counts in depth order (decoupled look-back),
// emission of the (supertile id, Gaussian id) pairs at the scanned offsets, AND the digit histograms the
// following radix sort needs - one kernel, one pass over the depth-ordered Gaussians.
// Replaces (reference): render.py:260-281 (expansion) + the implicit offsets of repeat_interleave.
// =================================================================================================
constexpr int kSeThreads = 256;
constexpr int kSeItems = 4;
constexpr int kSeTile = kSeThreads * kSeItems;   // 1024 depth ranks per block

__global__ void __launch_bounds__(kSeThreads) scan_emit_super_kernel(
    int n_host, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ order,
    const uint32_t* __restrict__ super_touched,
    const uint2* __restrict__ rect, int super_x, int super_y0, uint32_t capacity, uint32_t* __restrict__ keys,
    uint32_t* __restrict__ vals, b200gs_frame_stats* __restrict__ stats, uint32_t* ticket,
    unsigned long long* status, SortPasses sp, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t s_hist[kMaxPasses][kRadix];
  __shared__ uint32_t s_warp[kSeThreads / 32];
  __shared__ uint32_t s_tile, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = (int)eff_count((uint32_t)n_host, n_dev);   // band frames: only the compacted prefix of `order` is live
  if (blockIdx.x == 0 && tid == 0 && stats->n_isect > capacity) stats->overflow = 1u;   // per-tile lists would not fit
  if ((uint32_t)blockIdx.x * (uint32_t)kSeTile >= (uint32_t)n) return;   // surplus CTAs leave before taking a ticket
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kMaxPasses * kRadix; i += kSeThreads) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t r0 = tile * kSeTile + tid * kSeItems;
  uint32_t id[kSeItems], cnt[kSeItems], excl[kSeItems];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kSeItems; ++k) {
    const uint32_t r = r0 + k;
    id[k] = (r < (uint32_t)n) ? order[r] : 0u;
    cnt[k] = (r < (uint32_t)n) ? super_touched[id[k]] : 0u;
    excl[k] = sum;
    sum += cnt[k];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t warp_off = 0, tile_sum = 0;
#pragma unroll
  for (int w = 0; w < kSeThreads / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < warp) warp_off += t;
    tile_sum += t;
  }
  if (warp == 0) {
    const uint32_t prefix = scan_lookback(status, tile, tile_sum, lane);
    if (lane == 0) {
      s_prefix = prefix;
      if (tile == (uint32_t)(n - 1) / kSeTile) stats->n_super = prefix + tile_sum;
    }
  }
  __syncthreads();
  const uint32_t base = s_prefix + warp_off + (inc - sum);
#pragma unroll
  for (int k = 0; k < kSeItems; ++k) {
    if (cnt[k] == 0) continue;
    uint32_t off = base + excl[k];
    if (off + cnt[k] > capacity || off + cnt[k] < off) {   // does not fit: flag it, never write out of bounds
      stats->overflow = 1u;
      continue;
    }
    const uint2 rc = rect[id[k]];
    const int sx0 = (rc.x & 0xFFFF) / kSuperX, sx1 = (rc.x >> 16) / kSuperX;
    const int sy0 = (rc.y & 0xFFFF) / kSuperY, sy1 = (rc.y >> 16) / kSuperY;
    for (int sy = sy0; sy <= sy1; ++sy)
      for (int sx = sx0; sx <= sx1; ++sx) {
        const uint32_t key = (uint32_t)((sy - super_y0) * super_x + sx);    // supertile id, local to the band
        keys[off] = key;
        vals[off] = id[k];
        ++off;
#pragma unroll
        for (int p = 0; p < kMaxPasses; ++p)
          if (p < sp.num) atomicAdd(&s_hist[p][(key >> sp.shift[p]) & ((1u << sp.bits[p]) - 1u)], 1u);
      }
  }
  __syncthreads();
  for (int i = tid; i < kMaxPasses * kRadix; i += kSeThreads) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

size_t scan_emit_scratch_bytes(uint32_t n) { return 256 + ((size_t)(n + kSeTile - 1) / kSeTile + 1) * 8; }

// Zeroes `sort_scratch` (tickets, histograms, look-back status of the sort that follows) and `se_scratch`,
// runs the fused kernel; the sort must then be launched with hist_ready = true on the same scratch.
cudaError_t launch_scan_emit_super(int n, const uint32_t* n_dev, const uint32_t* order, const uint32_t* super_touched,
                                   const uint2* rect, int super_x, int super_y0, uint32_t capacity, uint32_t* keys,
                                   uint32_t* vals, b200gs_frame_stats* stats, int sort_bits, void* sort_scratch,
                                   size_t sort_scratch_bytes_, void* se_scratch, size_t se_scratch_bytes,
                                   cudaStream_t s) {
  if (sort_scratch_bytes_ < sort_scratch_bytes(capacity) || se_scratch_bytes < scan_emit_scratch_bytes((uint32_t)(n > 0 ? n : 1)))
    return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(sort_scratch, 0, sort_scratch_bytes(capacity), s);
  if (e != cudaSuccess) return e;
  if (n <= 0) return cudaSuccess;
  e = cudaMemsetAsync(se_scratch, 0, scan_emit_scratch_bytes((uint32_t)n), s);
  if (e != cudaSuccess) return e;
  uint32_t* ghist = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sort_scratch) + 256);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(se_scratch);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(se_scratch) + 256);
  const SortPasses sp = make_passes(0, sort_bits);
  scan_emit_super_kernel<<<(n + kSeTile - 1) / kSeTile, kSeThreads, 0, s>>>(n, n_dev, order, super_touched, rect, super_x,
                                                                          super_y0, capacity, keys, vals, stats, ticket,
                                                                          status, sp, ghist);
  return cudaGetLastError();
}

// Zeroes the sort scratch so that a producer kernel can accumulate the digit histograms (radix_sort_hist)
// before launch_radix_sort(..., hist_ready = true).
cudaError_t radix_sort_prepare(void* scratch, size_t scratch_bytes, uint32_t n, cudaStream_t s) {
  if (scratch_bytes < sort_scratch_bytes(n)) return cudaErrorInvalidValue;
  return cudaMemsetAsync(scratch, 0, sort_scratch_bytes(n), s);
}
uint32_t* radix_sort_hist(void* scratch) { return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(scratch) + 256); }

// Sorts on key bits [begin_bit, end_bit) with ceil(bits/8) passes.  Pass 0 reads (keys_src, vals_src) and
// writes (keys_b, vals_b); later passes ping-pong b -> a -> b ...  `*result_in_a` tells where the sorted
// data ended up (a for an even pass count, b for an odd one).  keys_src may alias keys_a (then the
// source is clobbered by pass 1) or be a separate read-only buffer.  vals_src == nullptr means
// "values are 0..n-1".  n is the host-side upper bound used for grid sizing; n_dev (optional) holds the
// real count on the device.
cudaError_t launch_radix_sort(const uint32_t* keys_src, const uint32_t* vals_src, uint32_t* keys_a,
                              uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n,
                              const uint32_t* n_dev, int begin_bit, int end_bit, void* scratch,
                              size_t scratch_bytes, int* result_in_a, cudaStream_t s, bool hist_ready) {
  if (end_bit <= begin_bit || end_bit - begin_bit > 8 * kMaxPasses) return cudaErrorInvalidValue;
  if (scratch_bytes < sort_scratch_bytes(n)) return cudaErrorInvalidValue;
  const SortPasses sp = make_passes(begin_bit, end_bit);
  if (result_in_a) *result_in_a = (sp.num % 2 == 0) ? 1 : 0;
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  if (!hist_ready) {     // otherwise the producer of the keys zeroed the scratch and filled the histograms
    e = cudaMemsetAsync(scratch, 0, sort_scratch_bytes(n), s);
    if (e != cudaSuccess) return e;
  }
  uint32_t* tickets = reinterpret_cast<uint32_t*>(scratch);
  uint32_t* ghist = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(scratch) + 256);
  uint32_t* status0 = ghist + kMaxPasses * kRadix;
  const size_t nblk = sort_blocks(n);
  int hgrid = (int)((n + kSortThreads * 8 - 1) / (kSortThreads * 8));
  if (hgrid > 148 * 8) hgrid = 148 * 8;
  if (!hist_ready) sort_histogram_kernel<<<hgrid, kSortThreads, 0, s>>>(keys_src, n, n_dev, sp, ghist);
  const uint32_t *ki = keys_src, *vi = vals_src;
  uint32_t *ko = keys_b, *vo = vals_b;
  for (int p = 0; p < sp.num; ++p) {
    uint32_t* status = status0 + (size_t)p * (nblk + 1) * kRadix;
    onesweep_pass_kernel<<<(int)nblk, kSortBlock, 0, s>>>(ki, vi, ko, vo, n, n_dev, sp.shift[p], sp.bits[p],
                                                         ghist + p * kRadix, status, tickets + p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ki = ko; vi = vo;
    if (ko == keys_b) { ko = keys_a; vo = vals_a; } else { ko = keys_b; vo = vals_b; }
  }
  return cudaSuccess;
}

}  // namespace gs
