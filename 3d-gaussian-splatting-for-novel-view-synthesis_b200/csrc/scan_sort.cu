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
#include <string.h>

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
constexpr int kMaxRadix = kSortMaxRadix;     // 9-bit digits (onesweep_pass2_kernel<9>); histograms and status rows are laid out for it
constexpr int kMaxPasses = kSortMaxPasses;
constexpr int kSortMaxGridRows = 480;   // = kSortMaxGrid (persistent blocks per launch at most)

// scratch: [0,64): tickets u32[4] | [256, 256+4*512*4): global histograms (512 bins per pass) | then per-pass status arrays
// look-back status rows a launch may use: one per tile - at most kSortMaxGrid while the tiles are smaller than 4096 keys
static size_t sort_blocks(uint32_t n) {
  const size_t b = (n + kSortTile - 1) / kSortTile;
  return b > (size_t)kSortMaxGridRows ? b : (size_t)kSortMaxGridRows;
}
size_t sort_scratch_bytes(uint32_t n) {
  return 256 + (size_t)kMaxPasses * kMaxRadix * 4 + (size_t)kMaxPasses * (sort_blocks(n) + 1) * kMaxRadix * 4;
}
// the part of it a sort with this pass plan touches (what has to be zeroed before it)
static size_t sort_scratch_used(uint32_t n, const SortPasses& sp) {
  return 256 + (size_t)kMaxPasses * kMaxRadix * 4 + (size_t)sp.num * (sort_blocks(n) + 1) * ((size_t)4 << sp.digit_bits);
}

__device__ __forceinline__ uint32_t eff_count(uint32_t n_host, const uint32_t* n_dev) {
  if (n_dev) { const uint32_t d = *n_dev; return d < n_host ? d : n_host; }
  return n_host;
}

// One read of the keys builds the digit histograms of every pass.
__global__ void __launch_bounds__(kSortThreads) sort_histogram_kernel(const uint32_t* __restrict__ keys,
                                                                      uint32_t n_host,
                                                                      const uint32_t* __restrict__ n_dev,
                                                                      SortPasses sp, uint32_t* __restrict__ ghist,
                                                                      uint32_t key_sub, uint32_t key_max) {
  __shared__ uint32_t s_hist[kMaxPasses][kMaxRadix];
  const uint32_t n = eff_count(n_host, n_dev);
  for (int i = threadIdx.x; i < kMaxPasses * kMaxRadix; i += kSortThreads) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * kSortThreads;
  for (uint32_t i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride) {
    const uint32_t k = depth_sort_key(keys[i], key_sub, key_max);
#pragma unroll
    for (int p = 0; p < kMaxPasses; ++p)
      if (p < sp.num) atomicAdd(&s_hist[p][(k >> sp.shift[p]) & ((1u << sp.bits[p]) - 1u)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMaxPasses * kMaxRadix; i += kSortThreads) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

// -------------------------------------------------------------------------------------------------
// One onesweep pass.  288 threads + MATCH.ANY ranking + shared-memory atomics + a look-back warp walking four
// predecessors per round trip was the first form (20.7 us per pass of 1M pairs); ncu's source view and per-block phase
// stamps (tools/sortlab/sort_trace.cu) showed where it went, and this form answers each item:
//   * half of all stall samples sat behind the 16 MATCH.ANY of the ranking (4.0 us for a block alone on its SM): the
//     peers of a key now come from one ballot per digit bit (2.9 us; bound by the SM's VOTE rate, ~4 cycles each),
//   * shared-memory atomics retire at two cycles per active lane: the per-warp digit counters are bumped with a plain
//     load + store by the lowest peer (leaders hold distinct digits, rounds are ordered by __syncwarp), and the block's
//     digit counts fall out of the per-warp counters after the ranking - no separate early-count atomics,
//   * every thread runs the look-back of its own digit(s), 16 predecessors per round trip (coalesced 128-byte reads of
//     the status rows), and spins on ONE status word with a sleep while its nearest predecessor has not posted
//     (a block spinning at full rate takes the issue slots of the block it is waiting for when they share an SM),
//   * digits may be 9 bits wide (512 bins, two per thread): depth keys of a [near, far] frustum need 27 bits, i.e. three
//     passes instead of four,
//   * the blocks are PERSISTENT and the tile size is a run-time value: the launch has a fixed number of blocks per SM, the
//     n keys (n may live on the device) are cut into ceil(n / blocks) keys each - every SM ranks the same number of keys,
//     whatever n is (245 fixed tiles of 4096 put two blocks on 97 SMs and one on 51; a grid sized for a host-side
//     upper bound of n left three working blocks on some SMs and none on others: 20 us of ranking there, everybody
//     else spinning) - and a block takes another ticket when there are more tiles than blocks.
// -------------------------------------------------------------------------------------------------
#ifndef SORT_LB
#define SORT_LB 16
#endif
constexpr int kLb2 = SORT_LB;    // predecessors inspected per look-back round trip and digit
#ifndef ONESWEEP_MIN_BLOCKS
#define ONESWEEP_MIN_BLOCKS 3      // narrow shape: 80 registers, three blocks per SM: a smaller block finds room sooner next to the blend blocks of the previous frame
#endif
constexpr int kSortMaxGrid = 480;     // persistent blocks per launch at most (the look-back status rows are sized for it)

#ifdef SORT_TRACE      // tools/sortlab: per-block phase stamps (SM clock) of thread 0
__device__ unsigned long long g_sort_trace[4096 * 12];
__device__ uint32_t g_sort_trace_sink;
#define TRACE_STAMP(k) do { if (tid == 0 && vbid < 4096) g_sort_trace[vbid * 12 + (k)] = clock64(); } while (0)
#else
#define TRACE_STAMP(k) do { } while (0)
#endif

template <int BITS, int THREADS, int ITEMS>
constexpr size_t pass2_smem_bytes() {
  return (size_t)(2 * THREADS * ITEMS + (1 << BITS)) * 4 + (size_t)(THREADS / 32) * ((1 << BITS) + 2) * 2;
}

// keys per tile for n keys on `grid` persistent blocks: even shares, whole rounds of the block (a multiple of THREADS),
// at least 4 rounds, at most ITEMS
template <int THREADS, int ITEMS>
__host__ __device__ __forceinline__ uint32_t pass2_tile(uint32_t n, uint32_t grid) {
  uint32_t items = ((n + grid - 1) / grid + THREADS - 1) / THREADS;
  items = items < 4u ? 4u : (items > (uint32_t)ITEMS ? (uint32_t)ITEMS : items);
  return items * THREADS;
}

template <int BITS, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS, (THREADS == 256 ? ONESWEEP_MIN_BLOCKS : 1)) onesweep_pass2_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, uint32_t n_host, const uint32_t* __restrict__ n_dev, int shift, int bits,
    const uint32_t* __restrict__ ghist /*[1 << BITS] digit counts of this pass*/, uint32_t* __restrict__ status,
    uint32_t* ticket, uint32_t key_sub, uint32_t key_max) {
  constexpr int R = 1 << BITS;
  constexpr int WARPS = THREADS / 32;
  constexpr int DPT = (R + THREADS - 1) / THREADS; // digits per thread (threads beyond the last digit own none)
  constexpr int kRow = R + 2;                      // 16-bit counters per warp row (the pad staggers the rows' banks)
  // per-warp digit counters, 16 bits each (a warp holds <= 512 keys, a block <= 7168).  After the ranking they are
  // rewritten as the position in the sorted block at which the warp's keys of that digit start.
  extern __shared__ __align__(16) uint32_t s_dyn[];       // pass2_smem_bytes<BITS, THREADS, ITEMS>()
  uint32_t* const s_keys = s_dyn;                          // [THREADS * ITEMS]
  uint32_t* const s_vals = s_dyn + THREADS * ITEMS;        // [THREADS * ITEMS]
  uint32_t* const s_delta = s_dyn + 2 * THREADS * ITEMS;   // [R] global position of a digit's run minus its position in the sorted block
  uint16_t (*const s_wh)[kRow] = reinterpret_cast<uint16_t (*)[kRow]>(s_dyn + 2 * THREADS * ITEMS + R);   // [WARPS][kRow]
  __shared__ uint2 s_scan[WARPS];
  __shared__ uint32_t s_vbid;

  const uint32_t n = eff_count(n_host, n_dev);
  if (n == 0) return;
  const uint32_t tile = pass2_tile<THREADS, ITEMS>(n, gridDim.x);
  const uint32_t items = tile / THREADS;                   // rounds of this launch (uniform)
  const uint32_t nblk = (n + tile - 1) / tile;
  if (blockIdx.x >= nblk) return;                          // more blocks than tiles (tiny n): leave before taking a ticket
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mask = (1u << bits) - 1u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  // global digit counts of this pass (the same for every tile of the block)
  uint32_t ghist_own[DPT];
#pragma unroll
  for (int k = 0; k < DPT; ++k) ghist_own[k] = (DPT * tid + k < R) ? ghist[DPT * tid + k] : 0u;

  for (;;) {
    if (tid == 0) s_vbid = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * kRow / 2; i += THREADS) reinterpret_cast<uint32_t*>(&s_wh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t vbid = s_vbid;
    if (vbid >= nblk) return;                              // uniform for the whole block
    const uint32_t base = vbid * tile;
    const bool full = base + tile <= n;
#ifdef SORT_TRACE
    if (tid == 0 && vbid < 4096) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_sort_trace[vbid * 12 + 10] = gt; g_sort_trace[vbid * 12 + 11] = blockIdx.x; uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_sort_trace[vbid * 12 + 9] = sm; }
#endif
    TRACE_STAMP(0);

    // ---- load (warp-striped: item i of lane l sits at warp_base + 32 i + l) -----------------------------------
    const uint32_t warp_base = base + warp * (32 * items);
    uint32_t key[ITEMS], val[ITEMS], rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      key[i] = ((uint32_t)i < items && (full || idx < n)) ? depth_sort_key(keys_in[idx], key_sub, key_max) : 0xFFFFFFFFu;
    }
#ifdef SORT_TRACE
    { uint32_t x = 0;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) x ^= key[i];
      if (x == 0xDEADBEEFu) g_sort_trace_sink = 1; }
    TRACE_STAMP(1);      // keys arrived
#endif

    // ---- stable ranking: peers of a key = lanes holding the same digit ------------------------------------------
    // One ballot per digit bit (bits above the pass's width are zero in every lane and change nothing).  This loop is
    // bound by the SM's ballot rate - one VOTE per ~3.7 cycles per SM, measured; REDUX.OR shares the unit, MATCH.ANY is
    // slower still - so a round costs ~33 cycles of the SM whatever the number of warps: 14 rounds x 16 warps = 3.8 us.
    // (Forming all masks first and updating the counters in a second loop was measured: 4.8 us - there is no latency
    // chain to break.)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if ((uint32_t)i >= items) break;                                      // uniform
      const uint32_t idx = warp_base + i * 32 + lane;
      const bool valid = full || idx < n;
      const uint32_t d = (key[i] >> shift) & mask;
      uint32_t peers = full ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int b = 0; b < BITS; ++b) {
        const bool bit = (d & (1u << b)) != 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= m ^ (bit ? 0u : 0xffffffffu);
      }
      // The lowest peer bumps the warp's counter of the digit.  In one round the leaders hold distinct digits, and the
      // rounds of a warp are ordered by __syncwarp, so a plain load + store does it (shared-memory ATOMICS retire at
      // two cycles per active lane: 16 rounds x 8 warps of them were 4 us of a pass).
      uint32_t old = 0;
      if (valid && (peers & lane_lt) == 0u) {
        old = s_wh[warp][d];
        s_wh[warp][d] = (uint16_t)(old + __popc(peers));
      }
      __syncwarp();
      old = __shfl_sync(0xffffffffu, old, (__ffs(peers) - 1) & 31);
      rank[i] = old + __popc(peers & lane_lt);
    }
    TRACE_STAMP(2);      // ranking done
    // the values are only needed for the scatter: their loads are in flight during the digit bookkeeping below
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      val[i] = vals_in ? (((uint32_t)i < items && (full || idx < n)) ? vals_in[idx] : 0u) : idx;
    }
    __syncthreads();
    // ---- per digit (thread t owns digits DPT t ...): the block's count, the start of each warp's keys of the digit
    //      in the sorted block, the start of the digit in the output (scan of the global histogram) ----------------
    uint32_t cnt[DPT], gdelta[DPT];
    {
      uint16_t* const wh16 = &s_wh[0][0];
      uint32_t csum = 0, gsum = 0;
#pragma unroll
      for (int k = 0; k < DPT; ++k) {
        uint32_t c = 0;
        if (DPT * tid + k < R) {
#pragma unroll
          for (int w = 0; w < WARPS; ++w) c += wh16[w * kRow + DPT * tid + k];
          // post this block's digit count for the blocks behind it (block 0: it is the inclusive prefix already)
          st_status32(status + (size_t)vbid * R + DPT * tid + k, ((vbid == 0 ? 2u : 1u) << 30) | c);
        }
        cnt[k] = c;
        csum += c;
        gsum += ghist_own[k];
      }
      uint32_t cinc = csum, ginc = gsum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, cinc, d);
        const uint32_t g = __shfl_up_sync(0xffffffffu, ginc, d);
        if (lane >= d) { cinc += t; ginc += g; }
      }
      if (lane == 31) s_scan[warp] = make_uint2(cinc, ginc);
      __syncthreads();
      uint32_t cw = 0, gw = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const uint2 t = s_scan[w];
        if (w < warp) { cw += t.x; gw += t.y; }
      }
      uint32_t cstart = cw + cinc - csum, gstart = gw + ginc - gsum;
#pragma unroll
      for (int k = 0; k < DPT; ++k) {
        gdelta[k] = gstart - cstart;            // global start of the digit minus its start in the sorted block
        if (DPT * tid + k < R) {
          uint32_t c = cstart;
#pragma unroll
          for (int w = 0; w < WARPS; ++w) {
            const uint32_t t = wh16[w * kRow + DPT * tid + k];
            wh16[w * kRow + DPT * tid + k] = (uint16_t)c;
            c += t;
          }
        }
        cstart += cnt[k];
        gstart += ghist_own[k];
      }
    }
    __syncthreads();
    TRACE_STAMP(3);      // digit bookkeeping done
    // ---- scatter into shared memory in sorted order ------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint32_t idx = warp_base + i * 32 + lane;
      if ((uint32_t)i < items && (full || idx < n)) {
        const uint32_t d = (key[i] >> shift) & mask;
        const uint32_t pos = (uint32_t)s_wh[warp][d] + rank[i];
        s_keys[pos] = key[i];
        s_vals[pos] = val[i];
      }
    }
    TRACE_STAMP(4);      // scatter issued
    // ---- decoupled look-back, one chain per digit, kLb2 predecessors per round trip --------------------------
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
      if (DPT * tid + k >= R) continue;
      uint32_t excl = 0;
      if (vbid != 0) {
        const uint32_t* col = status + DPT * tid + k;
        int j = (int)vbid - 1;
        uint32_t spins = 0;
        while (true) {
          // the nearest predecessor first, alone and with a sleep: a block that polls at full rate takes the issue
          // slots of the very block it waits for when the two share an SM
          uint32_t s0;
          while (((s0 = ld_status32(col + (size_t)j * R)) >> 30) == 0u) {
            __nanosleep(100);
            if (++spins > (1u << 22)) __trap();     // a predecessor never posted: fail loudly instead of hanging the GPU
          }
          excl += s0 & 0x3FFFFFFFu;
          if ((s0 >> 30) == 2u) break;
          --j;
          uint32_t sv[kLb2];
#pragma unroll
          for (int b = 0; b < kLb2; ++b) {
            const int idx = j - b;
            sv[b] = idx >= 0 ? ld_status32(col + (size_t)idx * R) : (2u << 30);   // before block 0: prefix 0
          }
          int used = 0;
          bool found = false;
#pragma unroll
          for (int b = 0; b < kLb2; ++b) {
            const uint32_t flag = sv[b] >> 30;
            if (!found && used == b && flag != 0) {
              excl += sv[b] & 0x3FFFFFFFu;
              used = b + 1;
              found = (flag == 2);
            }
          }
          j -= used;
          if (found) break;
        }
        st_status32(status + (size_t)vbid * R + DPT * tid + k, (2u << 30) | (excl + cnt[k]));
      }
      s_delta[DPT * tid + k] = gdelta[k] + excl;
    }
    TRACE_STAMP(5);      // this thread's look-back done
    __syncthreads();   // reorder done AND look-back done
    TRACE_STAMP(6);

    // ---- stream out in digit runs -------------------------------------------------------------------------------
    const uint32_t in_block = min(tile, n - base);
#pragma unroll 4
    for (uint32_t p = tid; p < in_block; p += THREADS) {
      const uint32_t k = s_keys[p];
      const uint32_t dst = s_delta[(k >> shift) & mask] + p;
      keys_out[dst] = k;
      vals_out[dst] = s_vals[p];
    }
    TRACE_STAMP(7);
    if (nblk <= gridDim.x) return;      // one tile per block: no further ticket to take
    __syncthreads();                    // the shared buffers are reused by the next tile
  }
}

// 8-bit digits unless 9-bit digits need fewer passes; the bits are spread evenly over the passes.
SortPasses sort_passes(int begin_bit, int end_bit) {
  SortPasses sp;
  const int total = end_bit - begin_bit;
  const int n8 = (total + 7) / 8, n9 = (total + 8) / 9;
  static const int force = getenv("B200GS_SORT_DIGIT_BITS") ? atoi(getenv("B200GS_SORT_DIGIT_BITS")) : 0;
  sp.digit_bits = (force == 8 || force == 9) ? force : (n9 < n8 ? 9 : 8);
  sp.num = (total + sp.digit_bits - 1) / sp.digit_bits;
  for (int p = 0; p < kMaxPasses; ++p) { sp.shift[p] = 0; sp.bits[p] = sp.digit_bits; }
  int at = begin_bit;
  for (int p = 0; p < sp.num; ++p) {
    const int left = end_bit - at, passes_left = sp.num - p;
    sp.shift[p] = at;
    sp.bits[p] = (left + passes_left - 1) / passes_left;
    at += sp.bits[p];
  }
  return sp;
}
static SortPasses make_passes(int begin_bit, int end_bit) { return sort_passes(begin_bit, end_bit); }

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
  __shared__ uint32_t s_hist[kMaxPasses][kMaxRadix];
  __shared__ uint32_t s_warp[kSeThreads / 32];
  __shared__ uint32_t s_tile, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = (int)eff_count((uint32_t)n_host, n_dev);   // band frames: only the compacted prefix of `order` is live
  if (blockIdx.x == 0 && tid == 0 && stats->n_isect > capacity) stats->overflow = 1u;   // per-tile lists would not fit
  if ((uint32_t)blockIdx.x * (uint32_t)kSeTile >= (uint32_t)n) return;   // surplus CTAs leave before taking a ticket
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kMaxPasses * kMaxRadix; i += kSeThreads) (&s_hist[0][0])[i] = 0;
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
  for (int i = tid; i < kMaxPasses * kMaxRadix; i += kSeThreads) {
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
  const SortPasses sp = make_passes(0, sort_bits);
  cudaError_t e = cudaMemsetAsync(sort_scratch, 0, sort_scratch_used(capacity, sp), s);
  if (e != cudaSuccess) return e;
  if (n <= 0) return cudaSuccess;
  e = cudaMemsetAsync(se_scratch, 0, scan_emit_scratch_bytes((uint32_t)n), s);
  if (e != cudaSuccess) return e;
  uint32_t* ghist = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sort_scratch) + 256);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(se_scratch);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(se_scratch) + 256);
  scan_emit_super_kernel<<<(n + kSeTile - 1) / kSeTile, kSeThreads, 0, s>>>(n, n_dev, order, super_touched, rect, super_x,
                                                                          super_y0, capacity, keys, vals, stats, ticket,
                                                                          status, sp, ghist);
  return cudaGetLastError();
}

// Zeroes the sort scratch so that a producer kernel can accumulate the digit histograms (radix_sort_hist)
// before launch_radix_sort(..., hist_ready = true).
cudaError_t radix_sort_prepare(void* scratch, size_t scratch_bytes, uint32_t n, int begin_bit, int end_bit, cudaStream_t s) {
  if (scratch_bytes < sort_scratch_bytes(n)) return cudaErrorInvalidValue;
  return cudaMemsetAsync(scratch, 0, sort_scratch_used(n, make_passes(begin_bit, end_bit)), s);
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
                              size_t scratch_bytes, int* result_in_a, cudaStream_t s, bool hist_ready,
                              uint32_t key_sub, uint32_t key_max, bool narrow) {
  if (end_bit <= begin_bit || end_bit - begin_bit > 32) return cudaErrorInvalidValue;
  if (scratch_bytes < sort_scratch_bytes(n)) return cudaErrorInvalidValue;
  const SortPasses sp = make_passes(begin_bit, end_bit);
  if (result_in_a) *result_in_a = (sp.num % 2 == 0) ? 1 : 0;
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  if (!hist_ready) {     // otherwise the producer of the keys zeroed the scratch and filled the histograms
    e = cudaMemsetAsync(scratch, 0, sort_scratch_used(n, sp), s);
    if (e != cudaSuccess) return e;
  }
  uint32_t* tickets = reinterpret_cast<uint32_t*>(scratch);
  uint32_t* ghist = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(scratch) + 256);
  uint32_t* status0 = ghist + kMaxPasses * kMaxRadix;
  int hgrid = (int)((n + kSortThreads * 8 - 1) / (kSortThreads * 8));
  if (hgrid > 148 * 8) hgrid = 148 * 8;
  if (!hist_ready) sort_histogram_kernel<<<hgrid, kSortThreads, 0, s>>>(keys_src, n, n_dev, sp, ghist, key_sub, key_max);
  // Block shape and count.  A pass is latency-bound at frame sizes, and its ranking is bound by the SM's VOTE rate:
  // what counts is how many keys the busiest SM ranks and how many warps share the latencies.  Two shapes, both
  // persistent with a run-time tile size (onesweep_pass2_kernel): "wide" = 512 threads x <= 14 keys, one block per SM
  // (1M keys: 6757 keys per SM, all SMs alike); "narrow" = 256 threads x <= 16 keys, up to three blocks per SM - the
  // shape that finds room beside the blend kernel of the previous frame (frame pipeline, `narrow`) and the one for
  // small inputs.  B200GS_SORT_SHAPE = narrow | wide and B200GS_SORT_BPS = blocks per SM override (measurement only).
  static const int shape_env = [] {
    const char* e = getenv("B200GS_SORT_SHAPE");
    return !e ? -1 : !strcmp(e, "narrow") ? 0 : !strcmp(e, "wide") ? 1 : -1;
  }();
  static const int bps_env = getenv("B200GS_SORT_BPS") ? atoi(getenv("B200GS_SORT_BPS")) : 0;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0) sm_count = 148;
  }
  const bool wide = shape_env >= 0 ? shape_env == 1 : (!narrow && n >= (uint32_t)sm_count * 2048u);
  const int bps = bps_env > 0 ? bps_env : (wide ? 1 : 2);
  const uint32_t min_tile = wide ? 512u * 4u : 256u * 4u;
  uint32_t grid = (uint32_t)(sm_count * bps);
  if (grid > (uint32_t)kSortMaxGrid) grid = (uint32_t)kSortMaxGrid;
  if (grid > (n + min_tile - 1) / min_tile) grid = (n + min_tile - 1) / min_tile;
  const uint32_t *ki = keys_src, *vi = vals_src;
  uint32_t *ko = keys_b, *vo = vals_b;
  for (int p = 0; p < sp.num; ++p) {
    uint32_t* status = status0 + (size_t)p * (sort_blocks(n) + 1) * ((size_t)1 << sp.digit_bits);
    const uint32_t ks = p == 0 ? key_sub : 0u, km = p == 0 ? key_max : 0xFFFFFFFFu;   // later passes read transformed keys
#define GS_PASS2(B, T, I)                                                                                              \
  do {                                                                                                                 \
    static std::atomic<uint64_t> attr_done{0};                                                                         \
    once_per_device(attr_done, [] {                                                                                    \
      cudaFuncSetAttribute(onesweep_pass2_kernel<B, T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                           (int)pass2_smem_bytes<B, T, I>());                                                          \
    });                                                                                                                \
    onesweep_pass2_kernel<B, T, I><<<(int)grid, T, pass2_smem_bytes<B, T, I>(), s>>>(                                  \
        ki, vi, ko, vo, n, n_dev, sp.shift[p], sp.bits[p], ghist + p * kMaxRadix, status, tickets + p, ks, km);        \
  } while (0)
    if (sp.digit_bits == 9) {
      if (wide) GS_PASS2(9, 512, 14);
      else GS_PASS2(9, 256, 16);
    } else {
      if (wide) GS_PASS2(8, 512, 14);
      else GS_PASS2(8, 256, 16);
    }
#undef GS_PASS2
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ki = ko; vi = vo;
    if (ko == keys_b) { ko = keys_a; vo = vals_a; } else { ko = keys_b; vo = vals_b; }
  }
  return cudaSuccess;
}

}  // namespace gs
