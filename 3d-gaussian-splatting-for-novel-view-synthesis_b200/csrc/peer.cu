// Data-parallel optimizer step over NVLink peer memory: gradient reduce-scatter + clip + Adam + parameter
// all-gather in ONE kernel (training row; section 5 "fuse the reduction into the kernel").
//
// Replaces (reference): scripts/train.py:530-538 for one process per GPU - the reference has no multi-GPU path;
// the data-parallel form of its iteration is  backward -> SUM all-reduce of the six gradient tensors ->
// clip_grad_norm_(model.pos, 1.0) -> optim.Adam.step().  Done with NCCL + the single-GPU kernels that is
// 236 MB through a ring, a write of the reduced gradients, and a second pass over them by the optimizer.
// Here every rank owns 1/world of every tensor.  With all ranks' staging and parameter buffers mapped into
// every process (CUDA VMM / IPC handles, exchanged by the host side), the owner of an element
//     reads the element's gradient from every rank's staging buffer over NVLink (p2p loads) and sums them in
//     rank order (the reduce-scatter), applies the clip coefficient, updates ITS shard of the Adam moments
//     (the moments are sharded: 1/world of the optimizer state and of its HBM traffic per GPU), and stores the
//     new parameter value into every rank's parameter buffer (p2p stores: the all-gather).
// The wire traffic is that of an all-reduce (2 (p-1)/p x 236 MB per GPU); the reduced gradient is never
// written unless the caller asks for it (scripts/train.py:544-557 reads pos.grad for densification every 100
// iterations), and the transfer overlaps the optimizer arithmetic element by element.
//
// Cross-GPU ordering is three flag barriers per step (monotonic epochs, release/acquire at system scope):
//   B1 all staging buffers are written, nobody still reads parameters     -> sum-of-squares partials
//   B2 all partials are published                                         -> fused step
//   B3 all parameter stores have landed, staging buffers may be reused
// A barrier that waits longer than ~20 s traps (a dead peer must not hang the GPU).
//
// With an NVLS multicast mapping of the areas (NVSwitch; group.multicast != NULL) the p loads per element become one
// multimem.ld_reduce (the switch sums) and the p stores one multimem.st (the switch replicates): the bytes a GPU puts
// on / takes off its links go from 2 (p-1)/p x S to (1 + 1/p) x S (S = 4 B x elements; through the switch a rank's own
// copy crosses the links too), which pays from 8 ranks up: 0.85 ms against 1.06 ms at N = 1M (2 ranks: 0.84 vs 0.58).
//
// Roofline: NVLink (reads (p-1)/p x 4 B + writes (p-1)/p x 4 B per element per GPU) + HBM (20 B per owned element).
#include "common.cuh"

namespace gs {

constexpr int kPeerThreads = 256;
constexpr int kPeerChunk = kPeerThreads * 4 * 4;            // 4096 elements per block
constexpr int kPeerMaxTensors = B200GS_PEER_MAX_TENSORS;
constexpr int kMaxPeers = B200GS_MAX_PEERS;
constexpr size_t kCtrlBytes = B200GS_PEER_CTRL_BYTES;
// control block of every rank's area
constexpr size_t kCtrlArrive = 0;          // uint32 arrive[kMaxPeers] : arrive[q] = last epoch rank q signalled to me
constexpr size_t kCtrlSq = 256;            // double sq[kMaxPeers]     : sq[q] = sum of squares of rank q's clip slices
constexpr size_t kCtrlAcc = 512;           // double acc; uint32 ticket (local scratch of the sum-of-squares kernel)

struct PeerPtrs {
  char* area[kMaxPeers];
  char* mc;                // NVLS multicast mapping of all areas, or null
  int world, rank;
};

// NVLS: one load returns the element summed over every rank's copy (the switch reduces), one store writes every
// rank's copy.
__device__ __forceinline__ float4 mc_ld_sum_f4(const void* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float mc_ld_sum_f1(const void* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_f4(void* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st_f1(void* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

// MC: false = plain peer loads / stores, true = multimem.ld_reduce + multimem.st through the multicast mapping
template <bool MC>
__device__ __forceinline__ void load_sum4(const PeerPtrs& pp, long long grad_off, long long e4, float4 (&g)[4]) {
  if (MC) {
    const float4* g4 = reinterpret_cast<const float4*>(pp.mc + grad_off) + e4;
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = mc_ld_sum_f4(g4 + k * kPeerThreads + threadIdx.x);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < pp.world; ++q) {              // rank order
      const float4* g4 = reinterpret_cast<const float4*>(pp.area[q] + grad_off) + e4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 x = __ldcs(g4 + k * kPeerThreads + threadIdx.x);
        g[k].x += x.x; g[k].y += x.y; g[k].z += x.z; g[k].w += x.w;
      }
    }
  }
}
template <bool MC>
__device__ __forceinline__ float load_sum1(const PeerPtrs& pp, long long grad_off, long long e) {
  if (MC) return mc_ld_sum_f1(reinterpret_cast<const float*>(pp.mc + grad_off) + e);
  float s = 0.f;
  for (int q = 0; q < pp.world; ++q) s += reinterpret_cast<const float*>(pp.area[q] + grad_off)[e];
  return s;
}
template <bool MC>
__device__ __forceinline__ void store_all4(const PeerPtrs& pp, long long off, long long e4, const float4 (&x)[4]) {
  if (MC) {
    float4* o4 = reinterpret_cast<float4*>(pp.mc + off) + e4;
#pragma unroll
    for (int k = 0; k < 4; ++k) mc_st_f4(o4 + k * kPeerThreads + threadIdx.x, x[k]);
  } else {
    for (int q = 0; q < pp.world; ++q) {
      float4* o4 = reinterpret_cast<float4*>(pp.area[q] + off) + e4;
#pragma unroll
      for (int k = 0; k < 4; ++k) o4[k * kPeerThreads + threadIdx.x] = x[k];
    }
  }
}
template <bool MC>
__device__ __forceinline__ void store_all1(const PeerPtrs& pp, long long off, long long e, float x) {
  if (MC) {
    mc_st_f1(reinterpret_cast<float*>(pp.mc + off) + e, x);
    return;
  }
  for (int q = 0; q < pp.world; ++q) reinterpret_cast<float*>(pp.area[q] + off)[e] = x;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One block; thread q signals rank q and waits for rank q's signal.
__global__ void peer_barrier_kernel(PeerPtrs pp, uint32_t epoch) {
  const int q = threadIdx.x;
  if (q >= pp.world) return;
  __threadfence_system();
  st_release_sys(reinterpret_cast<uint32_t*>(pp.area[q] + kCtrlArrive) + pp.rank, epoch);
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(pp.area[pp.rank] + kCtrlArrive) + q;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    if (clock64() - t0 > 40000000000LL) __trap();          // ~20 s at 2 GHz: a peer died
    __nanosleep(200);
  }
}

// The same barrier signalled THROUGH the multicast mapping: after a kernel that replicated data with multimem.st, the
// arrival flag must not overtake that data on its way through the switch, so it takes the same path - one
// multimem.st.release per rank writes arrive_mc[rank] in every rank's control block.
constexpr size_t kCtrlArriveMc = 1024;     // uint32 arrive_mc[kMaxPeers]
__global__ void peer_barrier_mc_kernel(PeerPtrs pp, uint32_t epoch) {
  const int q = threadIdx.x;
  if (q >= pp.world) return;
  if (q == 0) {
    __threadfence_system();
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    uint32_t* flag = reinterpret_cast<uint32_t*>(pp.mc + kCtrlArriveMc) + pp.rank;
    asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" :: "l"(flag), "r"(epoch) : "memory");
  }
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(pp.area[pp.rank] + kCtrlArriveMc) + q;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    if (clock64() - t0 > 40000000000LL) __trap();
    __nanosleep(200);
  }
}

struct PeerTensor {
  long long offset;        // floats from the start of the flat buffers
  long long begin, end;    // this rank's slice [begin, end) of the tensor
  long long shard_offset;  // floats from the start of the moment shards
  float step_size, bc2_sqrt;
  int chunk_begin;         // first block of this tensor's slice
  int clip;
};
struct PeerTable {
  PeerTensor t[kPeerMaxTensors];
  int n_tensors;
  float beta1, beta2, eps, one_minus_beta1, one_minus_beta2, max_norm;
  long long param_off, grad_off;      // byte offsets of the flat parameter / staging buffers inside an area
  float* m; float* v;                 // this rank's moment shards
  float* total_norm_out;
  int write_grads;
};

__device__ __forceinline__ const PeerTensor& tensor_of_block(const PeerTable& tb) {
  int ti = 0;
#pragma unroll
  for (int k = 1; k < kPeerMaxTensors; ++k)
    if (k < tb.n_tensors && (int)blockIdx.x >= tb.t[k].chunk_begin) ti = k;
  return tb.t[ti];
}

// Sum of squares of the REDUCED gradient over this rank's slices of the clipped tensors; the last block
// publishes the total to every rank's control block.
template <bool MC>
__global__ void __launch_bounds__(kPeerThreads) peer_sqnorm_kernel(const __grid_constant__ PeerPtrs pp,
                                                                   const __grid_constant__ PeerTable tb) {
  __shared__ double s_w[kPeerThreads / 32];
  const PeerTensor& t = tensor_of_block(tb);
  const long long base = t.begin + (long long)((int)blockIdx.x - t.chunk_begin) * kPeerChunk;
  const long long end = min(base + (long long)kPeerChunk, t.end);
  float acc = 0.f;
  if (end - base == kPeerChunk) {                 // slice boundaries are multiples of 4 floats, offsets of 32
    float4 s[4];
    load_sum4<MC>(pp, tb.grad_off, (t.offset + base) >> 2, s);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fmaf(s[k].x, s[k].x, fmaf(s[k].y, s[k].y, fmaf(s[k].z, s[k].z, fmaf(s[k].w, s[k].w, acc))));
  } else {
    for (long long i = base + threadIdx.x; i < end; i += kPeerThreads) {
      const float s = load_sum1<MC>(pp, tb.grad_off, t.offset + i);
      acc = fmaf(s, s, acc);
    }
  }
  double d = (double)acc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) s_w[warp] = d;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double tsum = 0.0;
#pragma unroll
  for (int k = 0; k < kPeerThreads / 32; ++k) tsum += s_w[k];
  char* mine = pp.area[pp.rank];
  double* accp = reinterpret_cast<double*>(mine + kCtrlAcc);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(mine + kCtrlAcc + 8);
  atomicAdd(accp, tsum);
  __threadfence();
  if (atomicAdd(ticket, 1u) != gridDim.x - 1) return;
  __threadfence();
  const double total = *reinterpret_cast<volatile double*>(accp);
  for (int q = 0; q < pp.world; ++q) reinterpret_cast<double*>(pp.area[q] + kCtrlSq)[pp.rank] = total;
  *accp = 0.0;
  *ticket = 0u;
  __threadfence_system();
}

// A rank with no clipped elements still has to publish its (zero) partial.
__global__ void peer_sqnorm_zero_kernel(PeerPtrs pp) {
  if (threadIdx.x < pp.world) reinterpret_cast<double*>(pp.area[threadIdx.x] + kCtrlSq)[pp.rank] = 0.0;
  __threadfence_system();
}

__device__ __forceinline__ float clip_coefficient(const PeerPtrs& pp, const PeerTable& tb, float* norm_out) {
  double total = 0.0;
  const double* sq = reinterpret_cast<const double*>(pp.area[pp.rank] + kCtrlSq);
  for (int q = 0; q < pp.world; ++q) total += sq[q];            // same order on every rank: identical coefficient
  const float norm = (float)sqrt(total);
  if (norm_out) *norm_out = norm;
  const float c = tb.max_norm / (norm + 1e-6f);                 // torch/nn/utils/clip_grad.py
  return c < 1.f ? c : 1.f;
}

__device__ __forceinline__ void adam_update_peer(float& p, float g, float& m, float& v, const PeerTable& tb, const PeerTensor& t) {
  adam_update_f32(p, g, m, v, tb.one_minus_beta1, tb.beta2, tb.one_minus_beta2, tb.eps, t.step_size, t.bc2_sqrt);   // = optim.cu
}

// ADAM = true : reduce-scatter + clip + Adam on the owned slice + all-gather of the new parameters
// ADAM = false: reduce-scatter + all-gather of the reduced gradient (a plain SUM all-reduce over peer memory)
template <bool ADAM, bool MC>
__global__ void __launch_bounds__(kPeerThreads) peer_step_kernel(const __grid_constant__ PeerPtrs pp,
                                                                 const __grid_constant__ PeerTable tb) {
  const PeerTensor& t = tensor_of_block(tb);
  const long long base = t.begin + (long long)((int)blockIdx.x - t.chunk_begin) * kPeerChunk;
  const long long end = min(base + (long long)kPeerChunk, t.end);
  float coef = 1.f;
  if (ADAM && tb.max_norm > 0.f) {
    const float c = clip_coefficient(pp, tb, (blockIdx.x == 0 && threadIdx.x == 0) ? tb.total_norm_out : nullptr);
    if (t.clip) coef = c;
  }
  const bool store_g = !ADAM || tb.write_grads;
  if (end - base == kPeerChunk) {
    const long long e4 = (t.offset + base) >> 2;
    float4 g[4];
    load_sum4<MC>(pp, tb.grad_off, e4, g);
    if (coef != 1.f) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { g[k].x *= coef; g[k].y *= coef; g[k].z *= coef; g[k].w *= coef; }
    }
    if (store_g) store_all4<MC>(pp, tb.grad_off, e4, g);
    if (ADAM) {
      const long long s4 = (t.shard_offset + (base - t.begin)) >> 2;
      float4* m4 = reinterpret_cast<float4*>(tb.m) + s4;
      float4* v4 = reinterpret_cast<float4*>(tb.v) + s4;
      const float4* p4 = reinterpret_cast<const float4*>(pp.area[pp.rank] + tb.param_off) + e4;
      float4 p[4], m[4], v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = k * kPeerThreads + threadIdx.x;
        p[k] = __ldcs(p4 + i); m[k] = __ldcs(m4 + i); v[k] = __ldcs(v4 + i);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        adam_update_peer(p[k].x, g[k].x, m[k].x, v[k].x, tb, t);
        adam_update_peer(p[k].y, g[k].y, m[k].y, v[k].y, tb, t);
        adam_update_peer(p[k].z, g[k].z, m[k].z, v[k].z, tb, t);
        adam_update_peer(p[k].w, g[k].w, m[k].w, v[k].w, tb, t);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = k * kPeerThreads + threadIdx.x;
        __stcs(m4 + i, m[k]); __stcs(v4 + i, v[k]);
      }
      store_all4<MC>(pp, tb.param_off, e4, p);
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += kPeerThreads) {
      float g = load_sum1<MC>(pp, tb.grad_off, t.offset + i);
      g *= coef;
      if (store_g) store_all1<MC>(pp, tb.grad_off, t.offset + i, g);
      if (ADAM) {
        const long long si = t.shard_offset + (i - t.begin);
        float p = reinterpret_cast<const float*>(pp.area[pp.rank] + tb.param_off)[t.offset + i], m = tb.m[si], v = tb.v[si];
        adam_update_peer(p, g, m, v, tb, t);
        tb.m[si] = m; tb.v[si] = v;
        store_all1<MC>(pp, tb.param_off, t.offset + i, p);
      }
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------
static long long round_up_ll(long long x, long long a) { return (x + a - 1) / a * a; }

int peer_layout_compute(const int64_t* numel, int n_tensors, int world, b200gs_peer_layout* out) {
  if (!numel || !out || n_tensors < 0 || n_tensors > kPeerMaxTensors || world < 1 || world > kMaxPeers) return -1;
  long long off = 0, soff = 0;
  for (int t = 0; t < kPeerMaxTensors; ++t) out->offset[t] = out->per[t] = out->shard_offset[t] = 0;
  for (int t = 0; t < n_tensors; ++t) {
    if (numel[t] < 0) return -1;
    const long long per = round_up_ll((numel[t] + world - 1) / world, 4);   // slice length (the last slices may be shorter)
    out->offset[t] = off;
    out->per[t] = per;
    out->shard_offset[t] = soff;
    off += round_up_ll(numel[t], 32);
    soff += round_up_ll(per, 32);
  }
  out->flat_total = off;
  out->shard_total = soff;
  return 0;
}

static PeerPtrs make_ptrs(const b200gs_peer_group* g) {
  PeerPtrs pp;
  for (int q = 0; q < kMaxPeers; ++q) pp.area[q] = q < g->world ? static_cast<char*>(g->area[q]) : nullptr;
  pp.mc = g->world > 1 ? static_cast<char*>(g->multicast) : nullptr;
  pp.world = g->world;
  pp.rank = g->rank;
  return pp;
}

cudaError_t launch_peer_barrier(const b200gs_peer_group* g, uint32_t epoch, cudaStream_t s) {
  peer_barrier_kernel<<<1, 32, 0, s>>>(make_ptrs(g), epoch);
  return cudaGetLastError();
}

// Stages the local gradients, then runs either the fused optimizer step or the plain all-reduce.
// Returns the number of kernels launched through *launches.
cudaError_t launch_peer_step(const b200gs_peer_group* g, const b200gs_peer_layout* L, const b200gs_peer_tensor* tensors,
                             int n_tensors, bool adam, float* m_shard, float* v_shard, double beta1, double beta2,
                             double eps, double max_norm, int write_grads, uint32_t* epoch, float* total_norm_out,
                             cudaStream_t s, int* launches) {
  const PeerPtrs pp = make_ptrs(g);
  PeerTable tb, tc;          // all tensors / the clipped tensors only
  tb.n_tensors = tc.n_tensors = 0;
  tb.beta1 = (float)beta1; tb.beta2 = (float)beta2; tb.eps = (float)eps;
  tb.one_minus_beta1 = (float)(1.0 - beta1); tb.one_minus_beta2 = (float)(1.0 - beta2);
  tb.max_norm = adam ? (float)max_norm : 0.f;
  tb.param_off = (long long)kCtrlBytes;
  tb.grad_off = (long long)kCtrlBytes + 4 * L->flat_total;
  tb.m = m_shard; tb.v = v_shard; tb.total_norm_out = total_norm_out; tb.write_grads = write_grads;
  char* mine = static_cast<char*>(g->area[g->rank]);
  float* stage = reinterpret_cast<float*>(mine + tb.grad_off);
  long long blocks = 0, cblocks = 0;
  cudaError_t e;
  for (int k = 0; k < n_tensors; ++k) {
    const b200gs_peer_tensor& a = tensors[k];
    if (a.numel <= 0) continue;
    // stage: local gradient -> this rank's peer-visible staging buffer
    // (a gradient that already lives in the staging buffer - the render backward wrote it there - is not copied)
    if (a.grad == stage + L->offset[k]) e = cudaSuccess;
    else if (a.grad) e = cudaMemcpyAsync(stage + L->offset[k], a.grad, (size_t)a.numel * 4, cudaMemcpyDeviceToDevice, s);
    else e = cudaMemsetAsync(stage + L->offset[k], 0, (size_t)a.numel * 4, s);
    if (e != cudaSuccess) return e;
    PeerTensor t;
    t.offset = L->offset[k];
    t.begin = std::min<long long>(a.numel, (long long)g->rank * L->per[k]);
    t.end = std::min<long long>(a.numel, t.begin + L->per[k]);
    t.shard_offset = L->shard_offset[k];
    if (adam) {
      const double bc1 = 1.0 - pow(beta1, (double)a.step), bc2 = 1.0 - pow(beta2, (double)a.step);
      t.step_size = (float)(a.lr / bc1);
      t.bc2_sqrt = (float)sqrt(bc2);
    } else {
      t.step_size = 0.f; t.bc2_sqrt = 1.f;
    }
    t.clip = a.clip;
    if (t.end <= t.begin) continue;
    const long long nb = (t.end - t.begin + kPeerChunk - 1) / kPeerChunk;
    t.chunk_begin = (int)blocks;
    tb.t[tb.n_tensors++] = t;
    blocks += nb;
    if (a.clip) {
      t.chunk_begin = (int)cblocks;
      tc.t[tc.n_tensors++] = t;
      cblocks += nb;
    }
  }
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
  int nl = 0;
  if ((e = launch_peer_barrier(g, ++*epoch, s)) != cudaSuccess) return e;                      // B1
  ++nl;
  if (adam && max_norm > 0.0) {
    tc.grad_off = tb.grad_off; tc.param_off = tb.param_off;
    if (cblocks > 0 && pp.mc) peer_sqnorm_kernel<true><<<(unsigned)cblocks, kPeerThreads, 0, s>>>(pp, tc);
    else if (cblocks > 0) peer_sqnorm_kernel<false><<<(unsigned)cblocks, kPeerThreads, 0, s>>>(pp, tc);
    else peer_sqnorm_zero_kernel<<<1, 32, 0, s>>>(pp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_peer_barrier(g, ++*epoch, s)) != cudaSuccess) return e;                    // B2
    nl += 2;
  }
  if (blocks > 0) {
    const unsigned nb = (unsigned)blocks;
    if (adam && pp.mc) peer_step_kernel<true, true><<<nb, kPeerThreads, 0, s>>>(pp, tb);
    else if (adam) peer_step_kernel<true, false><<<nb, kPeerThreads, 0, s>>>(pp, tb);
    else if (pp.mc) peer_step_kernel<false, true><<<nb, kPeerThreads, 0, s>>>(pp, tb);
    else peer_step_kernel<false, false><<<nb, kPeerThreads, 0, s>>>(pp, tb);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++nl;
  }
  if (pp.mc) {                                                                                 // B3
    peer_barrier_mc_kernel<<<1, 32, 0, s>>>(pp, ++*epoch);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  } else if ((e = launch_peer_barrier(g, ++*epoch, s)) != cudaSuccess) return e;
  ++nl;
  if (!adam || write_grads) {          // hand the reduced gradients back to the caller's tensors
    for (int k = 0; k < n_tensors; ++k) {
      const b200gs_peer_tensor& a = tensors[k];
      if (a.numel <= 0 || !a.grad || a.grad == stage + L->offset[k]) continue;
      if ((e = cudaMemcpyAsync(a.grad, stage + L->offset[k], (size_t)a.numel * 4, cudaMemcpyDeviceToDevice, s)) != cudaSuccess)
        return e;
    }
  }
  if (launches) *launches = nl;
  return cudaSuccess;
}

}  // namespace gs
