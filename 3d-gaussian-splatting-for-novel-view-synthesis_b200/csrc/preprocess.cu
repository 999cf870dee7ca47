// Per-Gaussian kernels: fused preprocess forward / backward, stand-alone Sigma and SH kernels.
//
// Replaces (reference paths): gaussian_splatting/gaussian.py:71-127, spherical_harmonics.py:70-166,
// render.py:104-258 + 305-315 (S1-S11, S15) and their autograd.
//
// Roofline: HBM.  Each Gaussian is read once (236 B from raw parameters, 64 B from sigma/color) and a
// 64-B splat record is written for survivors.  Rows of the [n,3] / [n,9] / [n,45] arrays are staged
// through shared memory with 16-byte coalesced streaming loads (row strides of 3, 9 and 45 words are
// odd, so the per-thread reads from shared memory are bank-conflict free).
//
// Compiled with -fmad=false: the projection feeds floor()/ceil() decisions that must not depend on
// the compiler's contraction choices; FMAs are written explicitly (fmaf) where wanted.
#include <stdlib.h>

#include "common.cuh"

namespace gs {

constexpr int kPreBlock = 128;

template <int K>
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, float* __restrict__ dst, int n0,
                                           int count) {
  const float* s = src + (size_t)n0 * K;
  const int total = count * K;
  if ((reinterpret_cast<uintptr_t>(s) & 15u) == 0) {
    const int nvec = total >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < nvec; i += kPreBlock) d4[i] = ld_stream_f4(s4 + i);
    for (int i = (nvec << 2) + threadIdx.x; i < total; i += kPreBlock) dst[i] = ld_stream_f(s + i);
  } else {
    for (int i = threadIdx.x; i < total; i += kPreBlock) dst[i] = ld_stream_f(s + i);
  }
}

// coalesced write-back of K floats per thread through shared memory
template <int K>
__device__ __forceinline__ void unstage_rows(float* __restrict__ dstg, const float* __restrict__ srcs, int n0,
                                             int count) {
  float* d = dstg + (size_t)n0 * K;
  const int total = count * K;
  if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
    const int nvec = total >> 2;
    float4* d4 = reinterpret_cast<float4*>(d);
    const float4* s4 = reinterpret_cast<const float4*>(srcs);
    for (int i = threadIdx.x; i < nvec; i += kPreBlock) d4[i] = s4[i];
    for (int i = (nvec << 2) + threadIdx.x; i < total; i += kPreBlock) d[i] = srcs[i];
  } else {
    for (int i = threadIdx.x; i < total; i += kPreBlock) d[i] = srcs[i];
  }
}

struct FrameView {
  float4 *rec0, *rec1, *rec2;
  uint32_t* depth_key;
  uint2* rect;
  uint32_t* radius;
  uint32_t* super_touched;
  float* grad_acc;
  b200gs_frame_stats* stats;
};

// The record the blend kernels consume, already in the form their per-(pixel, splat) visit wants, so that
// staging a tile's splats into shared memory is a plain asynchronous copy:  with c = -log2(e)/2 the exponent of
//     alpha_raw = op * exp(-q/2) = 2^e        is      e = c*A11*du^2 + c*2*A12*du*dv + c*A22*dv^2 + log2(op)
// and both gates of render.py:362-374 (q <= chi2; min(alpha_raw, alpha_max) >= alpha_cutoff) are  e >= gate,
// gate = max(c*chi2 + log2(op), log2(alpha_cutoff)).
//   rec0 = (u, v, c*A11, c*2*A12)   rec1 = (c*A22, log2(op), gate, r)   rec2 = (g, b, ext_u, ext_v)
__device__ __forceinline__ void write_splat_record(const FrameView& f, int i, const Projection& o, const float rgb[3],
                                                   float eu, float ev, const RenderParams& rp) {
  const float lop = log2f(o.op);
  f.rec0[i] = make_float4(o.u, o.v, kBlendExpScale * o.A11, kBlendExpScale * (2.f * o.A12));
  f.rec1[i] = make_float4(kBlendExpScale * o.A22, lop, fmaxf(rp.chi2c + lop, rp.cut_e), rgb[0]);
  f.rec2[i] = make_float4(rgb[1], rgb[2], eu, ev);
  f.radius[i] = (uint32_t)o.radius;
}

__device__ __forceinline__ void sh_color(const float* coef_dc, const float* coef_rest, const float Y[16],
                                         float rgb[3], float acc_out[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = coef_dc[c] * Y[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) acc = fmaf(coef_rest[15 * c + k - 1], Y[k], acc);
    acc_out[c] = acc;
    rgb[c] = sigmoidf_(acc);
  }
}

// Conservative half-extents of the region where the splat can contribute, for the per-warp culling in the
// blend kernels: { q <= chi2 } intersected with { opacity * exp(-q/2) >= alpha_cutoff }, i.e.
// q <= min(chi2, 2 ln(opacity / alpha_cutoff)).  A splat whose opacity is below the cutoff never contributes.
__device__ __forceinline__ void conic_extent(float A11, float A12, float A22, float chi2, float op, float alpha_cutoff,
                                             float& eu, float& ev) {
  const float detc = A11 * A22 - A12 * A12;
  if (alpha_cutoff > 0.f) chi2 = fminf(chi2, 2.f * logf(op / alpha_cutoff) * 1.0005f + 1e-3f);
  if (!(chi2 > 0.f)) { eu = -1e30f; ev = -1e30f; return; }
  if (detc > 0.f && isfinite(detc)) {
    eu = sqrtf(chi2 * A22 / detc) * 1.001f + 0.01f;
    ev = sqrtf(chi2 * A11 / detc) * 1.001f + 0.01f;
    if (!isfinite(eu)) eu = 1e30f;
    if (!isfinite(ev)) ev = 1e30f;
  } else {
    eu = 1e30f; ev = 1e30f;
  }
}

template <bool RAW_COV, bool RAW_SH>
__global__ void __launch_bounds__(kPreBlock) preprocess_fwd_kernel(GaussIn g, const float* __restrict__ c2w,
                                                                   RenderParams rp, FrameView f) {
  __shared__ __align__(16) float s_pos[kPreBlock * 3];
  __shared__ __align__(16) float s_cov[kPreBlock * (RAW_COV ? 3 : 9)];
  __shared__ __align__(16) float s_col[kPreBlock * 3];
  __shared__ __align__(16) float s_rest[RAW_SH ? kPreBlock * 45 : 4];
  __shared__ float s_c2w[16];
  __shared__ uint32_t s_tiles;

  const int n0 = blockIdx.x * kPreBlock;
  const int count = min(kPreBlock, g.n - n0);
  const int tid = threadIdx.x;
  if (tid < 16) s_c2w[tid] = c2w[tid];
  if (tid == 0) s_tiles = 0;
  stage_rows<3>(g.pos, s_pos, n0, count);
  if (RAW_COV) stage_rows<3>(g.scale_raw, s_cov, n0, count); else stage_rows<9>(g.sigma, s_cov, n0, count);
  if (RAW_SH) { stage_rows<3>(g.f_dc, s_col, n0, count); stage_rows<45>(g.f_rest, s_rest, n0, count); }
  else stage_rows<3>(g.color, s_col, n0, count);
  __syncthreads();
  const int i = n0 + tid;
  bool vis = false, past_s7 = false;
  uint32_t my_tiles = 0;
  if (tid < count) {
    const Pose ps = make_pose(s_c2w);
    const float p[3] = {s_pos[3 * tid], s_pos[3 * tid + 1], s_pos[3 * tid + 2]};
    Cov3 S;
    if (RAW_COV) {
      const float sr[3] = {s_cov[3 * tid], s_cov[3 * tid + 1], s_cov[3 * tid + 2]};
      const float4 q4 = ld_stream_f4(reinterpret_cast<const float4*>(g.q_raw) + i);
      const float q[4] = {q4.x, q4.y, q4.z, q4.w};
      QuatScale qs;
      quat_scale_forward(sr, q, qs);
      float full[9];
      sigma_full(qs, full);
      S = sym_from_full(full);
    } else {
      S = sym_from_full(&s_cov[9 * tid]);
    }
    Projection o;
    vis = project_gaussian(p, S, ld_stream_f(g.opacity_raw + i), ps, rp, o);
    past_s7 = vis || o.offscreen;
    // tile-row sharding: this rank only bins tile rows [row_begin, row_end); a survivor whose rect misses the band
    // is dropped here (culled key, no record, no SH evaluation) but still counts as visible
    int tv0 = 0, tv1 = 0, tiles = 0;
    if (vis) {
      tv0 = max(o.tv0, rp.row_begin); tv1 = min(o.tv1, rp.row_end - 1);
      tiles = (tv1 >= tv0) ? (o.tu1 - o.tu0 + 1) * (tv1 - tv0 + 1) : 0;
    }
    if (!vis || tiles == 0) {
      f.depth_key[i] = kCulledKey;
      f.super_touched[i] = 0;
    } else {
      float rgb[3];
      if (RAW_SH) {
        const ViewDir vd = view_dir(p, ps.cam);
        float Y[16], acc[3];
        sh_basis(vd.d, Y);
        sh_color(&s_col[3 * tid], &s_rest[45 * tid], Y, rgb, acc);
      } else {
        rgb[0] = s_col[3 * tid]; rgb[1] = s_col[3 * tid + 1]; rgb[2] = s_col[3 * tid + 2];
      }
      float eu, ev;
      conic_extent(o.A11, o.A12, o.A22, rp.chi2, o.op, rp.alpha_cutoff, eu, ev);
      write_splat_record(f, i, o, rgb, eu, ev, rp);
      f.depth_key[i] = __float_as_uint(o.z);
      f.rect[i] = make_uint2((uint32_t)o.tu0 | ((uint32_t)o.tu1 << 16), (uint32_t)tv0 | ((uint32_t)tv1 << 16));
      my_tiles = (uint32_t)tiles;
      f.super_touched[i] = (uint32_t)((o.tu1 / kSuperX - o.tu0 / kSuperX + 1) * (tv1 / kSuperY - tv0 / kSuperY + 1));
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, vis);
  const unsigned m7 = __ballot_sync(0xffffffffu, past_s7);
  if ((threadIdx.x & 31) == 0) {
    if (m) atomicAdd(&f.stats->n_visible, (uint32_t)__popc(m));
    if (m7) atomicAdd(&f.stats->n_in_frustum, (uint32_t)__popc(m7));
  }
  // I = sum of tile counts: one global atomic per block
  const uint32_t warp_tiles = __reduce_add_sync(0xffffffffu, my_tiles);
  if ((threadIdx.x & 31) == 0 && warp_tiles) atomicAdd(&s_tiles, warp_tiles);
  __syncthreads();
  if (tid == 0 && s_tiles) atomicAdd(&f.stats->n_isect, s_tiles);
}

// ------------------------------------------------------------------------------------------------
// Blackwell path of the fully-raw forward (the headline route): a persistent kernel, one CTA loop over
// 128-Gaussian chunks, whose six input rows-blocks are fetched with 1-D bulk async copies
// (cp.async.bulk global -> shared, completion on an mbarrier: SASS UBLKCP) into a double buffer, so the
// copy of chunk k+1 overlaps the math of chunk k and no thread spends issue slots on address arithmetic.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct __align__(128) PreStage {
  float rest[kPreBlock * 45];   // 23040 B
  float quat[kPreBlock * 4];    //  2048 B
  float pos[kPreBlock * 3];     //  1536 B
  float scale[kPreBlock * 3];
  float dc[kPreBlock * 3];
  float opac[kPreBlock];        //   512 B
};
constexpr uint32_t kPreStageBytes = kPreBlock * (45 + 4 + 3 + 3 + 3 + 1) * 4;

// Tile-row bands: a conservative "cannot touch rows [row_begin, row_end)" test that needs neither the covariance nor
// the eigenvalues.  Sigma_2D = M Sigma M^T with M = J Rwc, so lambda_max(Sigma_2D) <= |J|_F^2 max_i s_i^2, hence
//   radius = ceil(2.5 sqrt(clamp(lambda_max))) <= 2.5 sqrt(min(max(B, 1e-6), 1e4)) + 2,
//   B = 1.002 max_i s_i^2 ((fx^2 + fy^2)/z^2 + (fx^2 x^2 + fy^2 y^2)/z^4)
// (the 0.2 % and the +2 cover ceil() and every rounding in between).  A Gaussian whose centre row v is further than
// that from the band cannot have a tile there (project_gaussian S9-S11); one that fails the opacity pre-cull or the
// frustum test is dropped as it would be anyway.  ~40 instructions instead of the whole projection for the 7/8 of
// the scene that an eighth of the frame does not see.
__device__ __forceinline__ bool band_cannot_touch(const float p[3], const float sr[3], float opacity_raw, const Pose& ps,
                                                  const RenderParams& rp) {
  const float op = fminf(fmaxf(sigmoidf_(opacity_raw), 0.f), 0.999f);
  if (!(op >= rp.alpha_pre)) return true;
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = ps.r[3 * i] * p[0];
    acc = fmaf(ps.r[3 * i + 1], p[1], acc);
    acc = fmaf(ps.r[3 * i + 2], p[2], acc);
    c[i] = acc + ps.t[i];
  }
  const float x = c[0], y = c[1], z = c[2];
  const float fxx = rp.fx * x, fyy = rp.fy * y;
  const bool vis = (z > 0.f) && (z > rp.near_plane) && (z < rp.far_plane) &&
                   (fxx > z * rp.ulo) && (fxx < z * rp.uhi) && (fyy > z * rp.vlo) && (fyy < z * rp.vhi);
  if (!vis) return true;
  const float v = fyy / z + rp.cy;
  const float smax = fmaxf(__expf(fmaxf(sr[0], fmaxf(sr[1], sr[2]))), 1e-6f);
  const float invz = 1.0f / fmaxf(z, 1e-6f), invz2 = invz * invz;
  const float jf2 = (rp.fx * rp.fx + rp.fy * rp.fy) * invz2 + (fxx * fxx + fyy * fyy) * invz2 * invz2;
  const float B = 1.002f * smax * smax * jf2;
  const float rb = 2.5f * sqrtf(fminf(fmaxf(B, 1e-6f), 1e4f)) + 2.f;      // NaN / inf -> comparisons false -> kept
  return (v + rb < (float)(rp.row_begin * kTile)) || (v - rb >= (float)(rp.row_end * kTile));
}

// ------------------------------------------------------------------------------------------------
// Tile-row bands, raw-parameter route: two kernels instead of the fused one.
//   band_select_kernel   every Gaussian: 28 B in (pos, scale_raw, opacity_raw), the cheap test above, the ids of the
//                        candidates written in index order (block scan + decoupled look-back), depth_key = culled
//   band_project_kernel  one thread per CANDIDATE (dense, no divergence): gathers its Gaussian's rows by id, full
//                        projection, SH evaluation for the band's survivors, splat records; a depth key (or the culled
//                        marker) per candidate, which compact_keys_kernel then compacts together with the ids.
// A band of 1/8 of a 4K frame keeps ~1/8 of the scene: the fused kernel spent 280-330 us there whatever the band (one
// thread per Gaussian: every warp still ran the whole projection for its few live lanes; compacting inside the CTA
// serialised it into one warp per chunk), these two take the time of their bytes.
// ------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 256;
constexpr int kSelItems = 4;
constexpr int kSelTile = kSelThreads * kSelItems;      // 1024 Gaussians per block

// look-back state of band_expand_kernel (one status word per 32768 Gaussians)
size_t band_select_scratch_bytes(int n) { return 256 + ((size_t)(n + 32767) / 32768 + 1) * 8; }

// status word / look-back exactly as in scan_sort.cu (flag << 32 | value; 1 = aggregate, 2 = inclusive prefix)
__device__ __forceinline__ uint32_t sel_lookback(unsigned long long* status, uint32_t tile, uint32_t tile_sum, int lane) {
  auto st = [](unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  };
  auto ld = [](const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
  };
  if (tile == 0) {
    if (lane == 0) st(status, (2ull << 32) | tile_sum);
    return 0;
  }
  if (lane == 0) st(status + tile, (1ull << 32) | tile_sum);
  uint32_t prefix = 0;
  int j = (int)tile - 1;
  while (true) {
    const int idx = j - lane;
    const unsigned long long sv = (idx >= 0) ? ld(status + idx) : (2ull << 32);
    const uint32_t flag = (uint32_t)(sv >> 32);
    const unsigned ready = __ballot_sync(0xffffffffu, flag != 0);
    const unsigned pref = __ballot_sync(0xffffffffu, flag == 2);
    const int pp = pref ? (__ffs(pref) - 1) : 32;
    const unsigned need = (pp >= 31) ? 0xffffffffu : ((2u << pp) - 1u);
    if ((ready & need) != need) { __nanosleep(20); continue; }
    uint32_t c = (lane <= pp) ? (uint32_t)sv : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    prefix += c;
    if (pp < 32) break;
    j -= 32;
  }
  if (lane == 0) st(status + tile, (2ull << 32) | (uint32_t)(prefix + tile_sum));
  return prefix;
}

// Step 1 of the selection: pure streaming, no cross-CTA dependency.  A warp stages 128 rows with coalesced 16-byte
// loads, every lane tests four of them (item k of lane l = row 32 k + l) and the four ballots go out as flag words.
__global__ void __launch_bounds__(kSelThreads) band_select_kernel(GaussIn g, const float* __restrict__ c2w, RenderParams rp,
                                                                  uint32_t* __restrict__ depth_key,
                                                                  uint32_t* __restrict__ flag_words) {
  constexpr int kWarps = kSelThreads / 32;
  constexpr int kPerWarp = 32 * kSelItems;                       // 128 Gaussians per warp and step
  __shared__ __align__(16) float s_pos[kWarps][kPerWarp * 3];
  __shared__ __align__(16) float s_scale[kWarps][kPerWarp * 3];
  __shared__ __align__(16) float s_op[kWarps][kPerWarp];
  __shared__ float s_c2w[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 16) s_c2w[tid] = c2w[tid];
  __syncthreads();
  const Pose ps = make_pose(s_c2w);
  const int n_chunks = (g.n + kPerWarp - 1) / kPerWarp;
  for (int chunk = blockIdx.x * kWarps + warp; chunk < n_chunks; chunk += gridDim.x * kWarps) {
    const int w0 = chunk * kPerWarp;                                // first Gaussian of this warp's chunk
    if (w0 + kPerWarp <= g.n) {                                     // 1536 + 1536 + 512 contiguous bytes
      const float4* p4 = reinterpret_cast<const float4*>(g.pos + (size_t)w0 * 3);
      const float4* s4 = reinterpret_cast<const float4*>(g.scale_raw + (size_t)w0 * 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        reinterpret_cast<float4*>(s_pos[warp])[lane + 32 * k] = ld_stream_f4(p4 + lane + 32 * k);
        reinterpret_cast<float4*>(s_scale[warp])[lane + 32 * k] = ld_stream_f4(s4 + lane + 32 * k);
      }
      reinterpret_cast<float4*>(s_op[warp])[lane] = ld_stream_f4(reinterpret_cast<const float4*>(g.opacity_raw + w0) + lane);
    } else {
      for (int e = lane; e < kPerWarp * 3; e += 32) {
        const bool ok = w0 + e / 3 < g.n;
        s_pos[warp][e] = ok ? g.pos[(size_t)w0 * 3 + e] : 0.f;
        s_scale[warp][e] = ok ? g.scale_raw[(size_t)w0 * 3 + e] : 0.f;
      }
      for (int e = lane; e < kPerWarp; e += 32) s_op[warp][e] = (w0 + e < g.n) ? g.opacity_raw[w0 + e] : -1e30f;
    }
    __syncwarp();
    unsigned mine = 0;
#pragma unroll
    for (int k = 0; k < kSelItems; ++k) {
      const int e = 32 * k + lane;                                   // stride-3 rows: conflict-free
      const float p[3] = {s_pos[warp][3 * e], s_pos[warp][3 * e + 1], s_pos[warp][3 * e + 2]};
      const float sr[3] = {s_scale[warp][3 * e], s_scale[warp][3 * e + 1], s_scale[warp][3 * e + 2]};
      const bool cand = (w0 + e < g.n) && !band_cannot_touch(p, sr, s_op[warp][e], ps, rp);
      const unsigned bal = __ballot_sync(0xffffffffu, cand);
      if (lane == k) mine = bal;
      // every key starts out culled; band_project_kernel overwrites the survivors'
      if (w0 + e < g.n) depth_key[w0 + e] = kCulledKey;
    }
    if (lane < kSelItems) flag_words[chunk * kSelItems + lane] = mine;   // word w covers Gaussians [32 w, 32 w + 32)
    __syncwarp();
  }
}

// Step 2: the flag words (one per 32 Gaussians - 750 KB for 6M) -> candidate ids in index order: popcounts, block scan,
// decoupled look-back, expansion of the set bits.
constexpr int kExpThreads = 256;
constexpr int kExpItems = 4;
constexpr int kExpTile = kExpThreads * kExpItems;       // 1024 words = 32768 Gaussians per block

__global__ void __launch_bounds__(kExpThreads) band_expand_kernel(const uint32_t* __restrict__ flag_words, uint32_t n_words,
                                                                  uint32_t* __restrict__ cand_ids,
                                                                  uint32_t* __restrict__ n_cand_out, uint32_t* ticket,
                                                                  unsigned long long* status) {
  __shared__ uint32_t s_warp[kExpThreads / 32];
  __shared__ uint32_t s_tile, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t w0 = tile * kExpTile + tid * kExpItems;              // thread-contiguous words: index order is kept
  uint32_t w[kExpItems], cnt = 0;
#pragma unroll
  for (int k = 0; k < kExpItems; ++k) {
    w[k] = (w0 + k < n_words) ? flag_words[w0 + k] : 0u;
    cnt += (uint32_t)__popc(w[k]);
  }
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
  for (int q = 0; q < kExpThreads / 32; ++q) {
    const uint32_t t = s_warp[q];
    if (q < warp) warp_off += t;
    tile_sum += t;
  }
  if (warp == 0) {
    const uint32_t prefix = sel_lookback(status, tile, tile_sum, lane);
    if (lane == 0) {
      s_prefix = prefix;
      if (tile == (n_words - 1) / kExpTile) *n_cand_out = prefix + tile_sum;
    }
  }
  __syncthreads();
  uint32_t off = s_prefix + warp_off + (inc - cnt);
#pragma unroll
  for (int k = 0; k < kExpItems; ++k) {
    uint32_t bits = w[k];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      cand_ids[off++] = (w0 + k) * 32u + (uint32_t)b;
    }
  }
}

__global__ void __launch_bounds__(kPreBlock) band_project_kernel(GaussIn g, const float* __restrict__ c2w, RenderParams rp,
                                                                 FrameView f, const uint32_t* __restrict__ cand_ids,
                                                                 const uint32_t* __restrict__ n_cand,
                                                                 uint32_t* __restrict__ cand_key,
                                                                 uint32_t* __restrict__ depth_hist, DepthKeyPlan kp) {
  constexpr int kShStride = 49;                  // 3 + 45 coefficients, odd stride: conflict-free row reads
  __shared__ float s_c2w[16];
  __shared__ uint32_t s_tiles;
  __shared__ uint32_t s_dh[kSortMaxPasses * kSortMaxRadix];
  __shared__ float s_sh[kPreBlock / 32][32][kShStride];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 16) s_c2w[tid] = c2w[tid];
  if (tid == 0) s_tiles = 0;
  for (int i = tid; i < kSortMaxPasses * kSortMaxRadix; i += kPreBlock) s_dh[i] = 0;
  __syncthreads();
  const Pose ps = make_pose(s_c2w);
  const uint32_t total = *n_cand;
  uint32_t vis_count = 0, s7_count = 0, tiles_sum = 0;
  // whole warps iterate together (the SH rows of a warp's survivors are fetched cooperatively)
  for (uint32_t j0 = (blockIdx.x * (kPreBlock / 32) + warp) * 32; j0 < total; j0 += gridDim.x * kPreBlock) {
    const uint32_t j = j0 + lane;
    const bool have = j < total;
    const uint32_t i = have ? cand_ids[j] : 0u;
    bool want_sh = false;
    Projection o;
    float p[3] = {0.f, 0.f, 0.f};
    int tv0 = 0, tv1 = 0, tiles = 0;
    if (have) {
      p[0] = __ldg(g.pos + (size_t)i * 3); p[1] = __ldg(g.pos + (size_t)i * 3 + 1); p[2] = __ldg(g.pos + (size_t)i * 3 + 2);
      const float sr[3] = {__ldg(g.scale_raw + (size_t)i * 3), __ldg(g.scale_raw + (size_t)i * 3 + 1),
                           __ldg(g.scale_raw + (size_t)i * 3 + 2)};
      const float4 q4 = __ldg(reinterpret_cast<const float4*>(g.q_raw) + i);
      const float q[4] = {q4.x, q4.y, q4.z, q4.w};
      QuatScale qs;
      quat_scale_forward(sr, q, qs);
      float full[9];
      sigma_full(qs, full);
      const Cov3 S = sym_from_full(full);
      const bool vis = project_gaussian(p, S, __ldg(g.opacity_raw + i), ps, rp, o);
      s7_count += (vis || o.offscreen) ? 1u : 0u;
      vis_count += vis ? 1u : 0u;
      if (vis) {
        tv0 = max(o.tv0, rp.row_begin); tv1 = min(o.tv1, rp.row_end - 1);
        tiles = (tv1 >= tv0) ? (o.tu1 - o.tu0 + 1) * (tv1 - tv0 + 1) : 0;
      }
      want_sh = tiles > 0;
      if (!want_sh) cand_key[j] = kCulledKey;
    }
    // SH rows of the warp's survivors: one row per step, the 48 coefficients read by 32 lanes side by side (a lane
    // gathering its own row touches 32 different sectors per load instruction and thrashes L1)
    unsigned need = __ballot_sync(0xffffffffu, want_sh);
    __syncwarp();
    // eight rows per step: sixteen to twenty-four loads in flight per lane before the first store.  (LDGSTS copies of all
    // rows at once - no registers, 7 resident CTAs - measured slower: 158-177 us against 143-159 us per eighth of the 4K
    // frame; what bounds this kernel is the scattered 32-byte sector traffic, ~275 MB per band, not the latency chain.)
    constexpr int kRowsPerStep = 8;
    while (need) {
      int r[kRowsPerStep];
      float a[kRowsPerStep], b[kRowsPerStep], c[kRowsPerStep];
#pragma unroll
      for (int u = 0; u < kRowsPerStep; ++u) {
        r[u] = need ? (__ffs(need) - 1) : -1;
        if (need) need &= need - 1;
      }
#pragma unroll
      for (int u = 0; u < kRowsPerStep; ++u) {
        a[u] = b[u] = c[u] = 0.f;
        if (r[u] >= 0) {                                              // warp-uniform
          const uint32_t ir = __shfl_sync(0xffffffffu, i, r[u]);
          const float* rest = g.f_rest + (size_t)ir * 45;
          a[u] = __ldg(rest + lane);
          if (lane < 13) b[u] = __ldg(rest + 32 + lane);
          if (lane < 3) c[u] = __ldg(g.f_dc + (size_t)ir * 3 + lane);
        }
      }
#pragma unroll
      for (int u = 0; u < kRowsPerStep; ++u)
        if (r[u] >= 0) {
          s_sh[warp][r[u]][3 + lane] = a[u];
          if (lane < 13) s_sh[warp][r[u]][3 + 32 + lane] = b[u];
          if (lane < 3) s_sh[warp][r[u]][lane] = c[u];
        }
    }
    __syncwarp();
    if (!want_sh) continue;
    {
    const ViewDir vd = view_dir(p, ps.cam);
    float Y[16], rgb[3], acc[3];
    sh_basis(vd.d, Y);
    sh_color(&s_sh[warp][lane][0], &s_sh[warp][lane][3], Y, rgb, acc);
    float eu, ev;
    conic_extent(o.A11, o.A12, o.A22, rp.chi2, o.op, rp.alpha_cutoff, eu, ev);
    write_splat_record(f, (int)i, o, rgb, eu, ev, rp);
    const uint32_t dk = __float_as_uint(o.z);
    cand_key[j] = dk;
    f.depth_key[i] = dk;
    f.rect[i] = make_uint2((uint32_t)o.tu0 | ((uint32_t)o.tu1 << 16), (uint32_t)tv0 | ((uint32_t)tv1 << 16));
    f.super_touched[i] = (uint32_t)((o.tu1 / kSuperX - o.tu0 / kSuperX + 1) * (tv1 / kSuperY - tv0 / kSuperY + 1));
    tiles_sum += (uint32_t)tiles;
    if (depth_hist) depth_hist_add(s_dh, kp, dk);
    }
  }
  const uint32_t wv = __reduce_add_sync(0xffffffffu, vis_count), w7 = __reduce_add_sync(0xffffffffu, s7_count);
  const uint32_t wt = __reduce_add_sync(0xffffffffu, tiles_sum);
  if ((tid & 31) == 0) {
    if (wv) atomicAdd(&f.stats->n_visible, wv);
    if (w7) atomicAdd(&f.stats->n_in_frustum, w7);
    if (wt) atomicAdd(&s_tiles, wt);
  }
  __syncthreads();
  if (tid == 0 && s_tiles) atomicAdd(&f.stats->n_isect, s_tiles);
  if (depth_hist)
    for (int i = tid; i < kSortMaxPasses * kSortMaxRadix; i += kPreBlock) {
      const uint32_t c = s_dh[i];
      if (c) atomicAdd(&depth_hist[i], c);
    }
}

// (full frames only: a band of tile rows goes through band_select_kernel + band_project_kernel below)
__global__ void __launch_bounds__(kPreBlock) preprocess_fwd_tma_kernel(GaussIn g, const float* __restrict__ c2w,
                                                                       RenderParams rp, FrameView f, int n_chunks,
                                                                       uint32_t* __restrict__ depth_hist, DepthKeyPlan kp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PreStage* stage = reinterpret_cast<PreStage*>(smem_raw);       // [2]
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ float s_c2w[16];
  __shared__ uint32_t s_tiles;
  __shared__ uint32_t s_dh[kSortMaxPasses * kSortMaxRadix];      // digit histograms of the depth keys (all passes of the depth sort)
  const int tid = threadIdx.x;
  if (tid < 16) s_c2w[tid] = c2w[tid];
  for (int i = tid; i < kSortMaxPasses * kSortMaxRadix; i += kPreBlock) s_dh[i] = 0;
  if (tid == 0) {
    s_tiles = 0;
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const Pose ps = make_pose(s_c2w);

  // full chunks go through the bulk-copy engine; a ragged last chunk is staged by the threads
  auto issue = [&](int chunk, int buf) {
    const size_t n0 = (size_t)chunk * kPreBlock;
    PreStage& st = stage[buf];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this buffer are done
    mbar_expect_tx(&s_bar[buf], kPreStageBytes);
    bulk_g2s(st.rest, g.f_rest + n0 * 45, kPreBlock * 45 * 4, &s_bar[buf]);
    bulk_g2s(st.quat, g.q_raw + n0 * 4, kPreBlock * 4 * 4, &s_bar[buf]);
    bulk_g2s(st.pos, g.pos + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    bulk_g2s(st.scale, g.scale_raw + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    bulk_g2s(st.dc, g.f_dc + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    bulk_g2s(st.opac, g.opacity_raw + n0, kPreBlock * 4, &s_bar[buf]);
  };
  auto is_full = [&](int chunk) { return (chunk + 1) * kPreBlock <= g.n; };

  int chunk = blockIdx.x;
  if (chunk < n_chunks && is_full(chunk) && tid == 0) issue(chunk, 0);
  uint32_t vis_count = 0, s7_count = 0, tiles_sum = 0;   // per-thread tallies, reduced once at the end
  for (int it = 0; chunk < n_chunks; ++it, chunk += gridDim.x) {
    const int buf = it & 1;
    const int next = chunk + gridDim.x;
    if (next < n_chunks && is_full(next) && tid == 0) issue(next, buf ^ 1);
    PreStage& st = stage[buf];
    const int n0 = chunk * kPreBlock;
    const int count = min(kPreBlock, g.n - n0);
    if (is_full(chunk)) {
      mbar_wait(&s_bar[buf], (uint32_t)((it >> 1) & 1));
    } else {
      stage_rows<45>(g.f_rest, st.rest, n0, count);
      stage_rows<4>(g.q_raw, st.quat, n0, count);
      stage_rows<3>(g.pos, st.pos, n0, count);
      stage_rows<3>(g.scale_raw, st.scale, n0, count);
      stage_rows<3>(g.f_dc, st.dc, n0, count);
      stage_rows<1>(g.opacity_raw, st.opac, n0, count);
      __syncthreads();
    }
    if (tid < count) {
      const int i = n0 + tid;
      const float p[3] = {st.pos[3 * tid], st.pos[3 * tid + 1], st.pos[3 * tid + 2]};
      const float sr[3] = {st.scale[3 * tid], st.scale[3 * tid + 1], st.scale[3 * tid + 2]};
      const float4 q4 = reinterpret_cast<const float4*>(st.quat)[tid];
      const float q[4] = {q4.x, q4.y, q4.z, q4.w};
      QuatScale qs;
      quat_scale_forward(sr, q, qs);
      float full[9];
      sigma_full(qs, full);
      const Cov3 S = sym_from_full(full);
      Projection o;
      const bool vis = project_gaussian(p, S, st.opac[tid], ps, rp, o);
      s7_count += (vis || o.offscreen) ? 1u : 0u;
      vis_count += vis ? 1u : 0u;
      if (depth_hist) {            // (full frames only: a band's sort builds its own histograms after the compaction)
        const uint32_t dk = vis ? __float_as_uint(o.z) : kCulledKey;
        depth_hist_add(s_dh, kp, dk);
      }
      if (!vis) {
        f.depth_key[i] = kCulledKey;
        f.super_touched[i] = 0;
      } else {
        const ViewDir vd = view_dir(p, ps.cam);
        float Y[16], acc[3], rgb[3];
        sh_basis(vd.d, Y);
        sh_color(&st.dc[3 * tid], &st.rest[45 * tid], Y, rgb, acc);
        float eu, ev;
        conic_extent(o.A11, o.A12, o.A22, rp.chi2, o.op, rp.alpha_cutoff, eu, ev);
        // tile-row sharding (large bands): this rank only bins tile rows [row_begin, row_end)
        const int tv0 = max(o.tv0, rp.row_begin), tv1 = min(o.tv1, rp.row_end - 1);
        if (tv1 >= tv0) {
          write_splat_record(f, i, o, rgb, eu, ev, rp);
          f.depth_key[i] = __float_as_uint(o.z);
          f.rect[i] = make_uint2((uint32_t)o.tu0 | ((uint32_t)o.tu1 << 16), (uint32_t)tv0 | ((uint32_t)tv1 << 16));
          tiles_sum += (uint32_t)((o.tu1 - o.tu0 + 1) * (tv1 - tv0 + 1));
          f.super_touched[i] = (uint32_t)((o.tu1 / kSuperX - o.tu0 / kSuperX + 1) * (tv1 / kSuperY - tv0 / kSuperY + 1));
        } else {
          f.depth_key[i] = kCulledKey;
          f.super_touched[i] = 0;
        }
      }
    }
    __syncthreads();   // everyone is done with stage[buf] before it is refilled (two iterations from now)
  }
  // frame counters: one atomic per warp / block for the whole persistent loop
  const uint32_t wv = __reduce_add_sync(0xffffffffu, vis_count), w7 = __reduce_add_sync(0xffffffffu, s7_count);
  const uint32_t wt = __reduce_add_sync(0xffffffffu, tiles_sum);
  if ((tid & 31) == 0) {
    if (wv) atomicAdd(&f.stats->n_visible, wv);
    if (w7) atomicAdd(&f.stats->n_in_frustum, w7);
    if (wt) atomicAdd(&s_tiles, wt);
  }
  __syncthreads();
  if (tid == 0 && s_tiles) atomicAdd(&f.stats->n_isect, s_tiles);
  if (depth_hist)
    for (int i = tid; i < kSortMaxPasses * kSortMaxRadix; i += kPreBlock) {
      const uint32_t c = s_dh[i];
      if (c) atomicAdd(&depth_hist[i], c);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward: recompute the forward from the inputs, chain the 9 per-Gaussian gradients that the blend
// backward accumulated (grad_acc[n][12]) down to the leaves.  Dense outputs (zeros when culled).
// ------------------------------------------------------------------------------------------------
template <bool RAW_COV, bool RAW_SH>
__global__ void __launch_bounds__(kPreBlock) preprocess_bwd_kernel(GaussIn g, GaussGrad gg,
                                                                   const float* __restrict__ c2w, RenderParams rp,
                                                                   FrameView f) {
  __shared__ __align__(16) float s_pos[kPreBlock * 3];
  __shared__ __align__(16) float s_cov[kPreBlock * (RAW_COV ? 3 : 9)];   // in: scale/sigma, out: grads
  __shared__ __align__(16) float s_col[kPreBlock * 3];
  __shared__ __align__(16) float s_rest[RAW_SH ? kPreBlock * 45 : 4];
  __shared__ float s_c2w[16];

  const int n0 = blockIdx.x * kPreBlock;
  const int count = min(kPreBlock, g.n - n0);
  const int tid = threadIdx.x;
  if (tid < 16) s_c2w[tid] = c2w[tid];
  stage_rows<3>(g.pos, s_pos, n0, count);
  if (RAW_COV) stage_rows<3>(g.scale_raw, s_cov, n0, count); else stage_rows<9>(g.sigma, s_cov, n0, count);
  if (RAW_SH) { stage_rows<3>(g.f_dc, s_col, n0, count); stage_rows<45>(g.f_rest, s_rest, n0, count); }
  __syncthreads();
  const int i = n0 + tid;
  const bool live = tid < count;
  float gp[3] = {0.f, 0.f, 0.f}, g_op = 0.f;
  float g_cov[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // RAW_COV: [0..2] scale grads; else sigma grads
  float g_q[4] = {0.f, 0.f, 0.f, 0.f};
  float g_dc[3] = {0.f, 0.f, 0.f};
  bool vis = false;
  Pose ps;
  float Y[16];
  float g_acc[3] = {0.f, 0.f, 0.f};
  if (live) {
    ps = make_pose(s_c2w);
    const float p[3] = {s_pos[3 * tid], s_pos[3 * tid + 1], s_pos[3 * tid + 2]};
    Cov3 S;
    QuatScale qs;
    float q[4] = {0.f, 0.f, 0.f, 1.f};
    if (RAW_COV) {
      const float sr[3] = {s_cov[3 * tid], s_cov[3 * tid + 1], s_cov[3 * tid + 2]};
      const float4 q4 = ld_stream_f4(reinterpret_cast<const float4*>(g.q_raw) + i);
      q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
      quat_scale_forward(sr, q, qs);
      float full[9];
      sigma_full(qs, full);
      S = sym_from_full(full);
    } else {
      S = sym_from_full(&s_cov[9 * tid]);
    }
    Projection o;
    vis = project_gaussian(p, S, ld_stream_f(g.opacity_raw + i), ps, rp, o);
    if (vis) {
      const float4 a0 = reinterpret_cast<const float4*>(f.grad_acc)[3 * (size_t)i];
      const float4 a1 = reinterpret_cast<const float4*>(f.grad_acc)[3 * (size_t)i + 1];
      const float4 a2 = reinterpret_cast<const float4*>(f.grad_acc)[3 * (size_t)i + 2];
      const SplatGrad sg = splat_grad_from_moments(o, a0.x, a0.y, a0.z, a0.w, a1.x, a1.y);
      const float g_rgb[3] = {a1.z, a1.w, a2.x};
      float G[9];
      project_backward(p, S, ps, rp, o, sg, gp, G, g_op);
      if (RAW_COV) {
        quat_scale_backward(qs, q, G, g_cov, g_q);
      } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) g_cov[k] = G[k];
      }
      if (RAW_SH) {
        const ViewDir vd = view_dir(p, ps.cam);
        sh_basis(vd.d, Y);
        float rgb[3], acc[3];
        sh_color(&s_col[3 * tid], &s_rest[45 * tid], Y, rgb, acc);
        float gY[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) gY[k] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          g_acc[c] = g_rgb[c] * rgb[c] * (1.f - rgb[c]);
          g_dc[c] = g_acc[c] * Y[0];
#pragma unroll
          for (int k = 1; k < 16; ++k) gY[k] = fmaf(g_acc[c], s_rest[45 * tid + 15 * c + k - 1], gY[k]);
        }
        float gd[3], gpv[3];
        sh_basis_backward(vd.d, gY, gd);
        view_dir_backward(vd, gd, gpv);
        gp[0] += gpv[0]; gp[1] += gpv[1]; gp[2] += gpv[2];
      } else {
        g_dc[0] = g_rgb[0]; g_dc[1] = g_rgb[1]; g_dc[2] = g_rgb[2];   // dL/dcolor
      }
    }
  }
  __syncthreads();   // everyone is done reading the staged inputs; reuse the buffers for the outputs
  if (live) {
    s_pos[3 * tid] = gp[0]; s_pos[3 * tid + 1] = gp[1]; s_pos[3 * tid + 2] = gp[2];
    s_col[3 * tid] = g_dc[0]; s_col[3 * tid + 1] = g_dc[1]; s_col[3 * tid + 2] = g_dc[2];
    if (RAW_COV) {
      s_cov[3 * tid] = g_cov[0]; s_cov[3 * tid + 1] = g_cov[1]; s_cov[3 * tid + 2] = g_cov[2];
      reinterpret_cast<float4*>(gg.q_raw)[i] = make_float4(g_q[0], g_q[1], g_q[2], g_q[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) s_cov[9 * tid + k] = g_cov[k];
    }
    if (RAW_SH) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 1; k < 16; ++k) s_rest[45 * tid + 15 * c + k - 1] = vis ? g_acc[c] * Y[k] : 0.f;
    }
    gg.opacity_raw[i] = g_op;
  }
  __syncthreads();
  unstage_rows<3>(gg.pos, s_pos, n0, count);
  if (RAW_COV) unstage_rows<3>(gg.scale_raw, s_cov, n0, count); else unstage_rows<9>(gg.sigma, s_cov, n0, count);
  if (RAW_SH) { unstage_rows<3>(gg.f_dc, s_col, n0, count); unstage_rows<45>(gg.f_rest, s_rest, n0, count); }
  else unstage_rows<3>(gg.color, s_col, n0, count);
}

// ------------------------------------------------------------------------------------------------
// Blackwell path of the fully-raw backward: same persistent double-buffered bulk-copy pipeline as the
// forward, inputs + grad_acc in, and the six gradient row-blocks written back from shared memory with bulk
// async stores (cp.async.bulk shared -> global), so both directions run on the copy engine.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}

struct __align__(128) PreStageBwd {
  PreStage in;                       // inputs, overwritten in place by the gradients of the same shape
  float acc[kPreBlock * 12];         // grad_acc rows (6144 B)
};
constexpr uint32_t kPreStageBwdBytes = kPreStageBytes + kPreBlock * 12 * 4;

__global__ void __launch_bounds__(kPreBlock) preprocess_bwd_tma_kernel(GaussIn g, GaussGrad gg,
                                                                       const float* __restrict__ c2w,
                                                                       RenderParams rp, FrameView f, int n_chunks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PreStageBwd* stage = reinterpret_cast<PreStageBwd*>(smem_raw);       // [2]
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ float s_c2w[16];
  const int tid = threadIdx.x;
  if (tid < 16) s_c2w[tid] = c2w[tid];
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const Pose ps = make_pose(s_c2w);

  auto issue = [&](int chunk, int buf) {
    const size_t n0 = (size_t)chunk * kPreBlock;
    PreStageBwd& st = stage[buf];
    // the bulk stores that drained this buffer two iterations ago must have finished READING it
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&s_bar[buf], kPreStageBwdBytes);
    bulk_g2s(st.in.rest, g.f_rest + n0 * 45, kPreBlock * 45 * 4, &s_bar[buf]);
    bulk_g2s(st.in.quat, g.q_raw + n0 * 4, kPreBlock * 4 * 4, &s_bar[buf]);
    bulk_g2s(st.in.pos, g.pos + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    bulk_g2s(st.in.scale, g.scale_raw + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    bulk_g2s(st.in.dc, g.f_dc + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    bulk_g2s(st.in.opac, g.opacity_raw + n0, kPreBlock * 4, &s_bar[buf]);
    bulk_g2s(st.acc, f.grad_acc + n0 * 12, kPreBlock * 12 * 4, &s_bar[buf]);
  };
  auto is_full = [&](int chunk) { return (chunk + 1) * kPreBlock <= g.n; };

  int chunk = blockIdx.x;
  if (chunk < n_chunks && is_full(chunk) && tid == 0) issue(chunk, 0);
  for (int it = 0; chunk < n_chunks; ++it, chunk += gridDim.x) {
    const int buf = it & 1;
    const int next = chunk + gridDim.x;
    if (next < n_chunks && is_full(next) && tid == 0) issue(next, buf ^ 1);
    PreStageBwd& st = stage[buf];
    const int n0 = chunk * kPreBlock;
    const int count = min(kPreBlock, g.n - n0);
    const bool full = is_full(chunk);
    if (full) {
      mbar_wait(&s_bar[buf], (uint32_t)((it >> 1) & 1));
    } else {
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
      stage_rows<45>(g.f_rest, st.in.rest, n0, count);
      stage_rows<4>(g.q_raw, st.in.quat, n0, count);
      stage_rows<3>(g.pos, st.in.pos, n0, count);
      stage_rows<3>(g.scale_raw, st.in.scale, n0, count);
      stage_rows<3>(g.f_dc, st.in.dc, n0, count);
      stage_rows<1>(g.opacity_raw, st.in.opac, n0, count);
      stage_rows<12>(f.grad_acc, st.acc, n0, count);
      __syncthreads();
    }
    float gp[3] = {0.f, 0.f, 0.f}, g_op = 0.f, g_sr[3] = {0.f, 0.f, 0.f}, g_q[4] = {0.f, 0.f, 0.f, 0.f};
    float g_dc[3] = {0.f, 0.f, 0.f}, g_acc[3] = {0.f, 0.f, 0.f};
    float Y[16];
    bool vis = false;
    if (tid < count) {
      const float p[3] = {st.in.pos[3 * tid], st.in.pos[3 * tid + 1], st.in.pos[3 * tid + 2]};
      const float sr[3] = {st.in.scale[3 * tid], st.in.scale[3 * tid + 1], st.in.scale[3 * tid + 2]};
      const float4 q4 = reinterpret_cast<const float4*>(st.in.quat)[tid];
      const float q[4] = {q4.x, q4.y, q4.z, q4.w};
      QuatScale qs;
      quat_scale_forward(sr, q, qs);
      float fullm[9];
      sigma_full(qs, fullm);
      const Cov3 S = sym_from_full(fullm);
      Projection o;
      vis = project_gaussian(p, S, st.in.opac[tid], ps, rp, o);
      if (vis) {
        const float4 a0 = reinterpret_cast<const float4*>(st.acc)[3 * tid];
        const float4 a1 = reinterpret_cast<const float4*>(st.acc)[3 * tid + 1];
        const float4 a2 = reinterpret_cast<const float4*>(st.acc)[3 * tid + 2];
        const SplatGrad sg = splat_grad_from_moments(o, a0.x, a0.y, a0.z, a0.w, a1.x, a1.y);
        const float g_rgb[3] = {a1.z, a1.w, a2.x};
        float G[9];
        project_backward(p, S, ps, rp, o, sg, gp, G, g_op);
        quat_scale_backward(qs, q, G, g_sr, g_q);
        const ViewDir vd = view_dir(p, ps.cam);
        sh_basis(vd.d, Y);
        float rgb[3], acc[3], gY[16];
        sh_color(&st.in.dc[3 * tid], &st.in.rest[45 * tid], Y, rgb, acc);
#pragma unroll
        for (int k = 0; k < 16; ++k) gY[k] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          g_acc[c] = g_rgb[c] * rgb[c] * (1.f - rgb[c]);
          g_dc[c] = g_acc[c] * Y[0];
#pragma unroll
          for (int k = 1; k < 16; ++k) gY[k] = fmaf(g_acc[c], st.in.rest[45 * tid + 15 * c + k - 1], gY[k]);
        }
        float gd[3], gpv[3];
        sh_basis_backward(vd.d, gY, gd);
        view_dir_backward(vd, gd, gpv);
        gp[0] += gpv[0]; gp[1] += gpv[1]; gp[2] += gpv[2];
      }
    }
    __syncthreads();   // all inputs consumed: the stage now becomes the output staging area
    if (tid < count) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        st.in.pos[3 * tid + c] = gp[c];
        st.in.scale[3 * tid + c] = g_sr[c];
        st.in.dc[3 * tid + c] = g_dc[c];
#pragma unroll
        for (int k = 1; k < 16; ++k) st.in.rest[45 * tid + 15 * c + k - 1] = vis ? g_acc[c] * Y[k] : 0.f;
      }
      reinterpret_cast<float4*>(st.in.quat)[tid] = make_float4(g_q[0], g_q[1], g_q[2], g_q[3]);
      st.in.opac[tid] = g_op;
    }
    if (full) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // make the generic writes visible to the copy engine
      __syncthreads();
      if (tid == 0) {
        const size_t b = (size_t)n0;
        bulk_s2g(gg.f_rest + b * 45, st.in.rest, kPreBlock * 45 * 4);
        bulk_s2g(gg.q_raw + b * 4, st.in.quat, kPreBlock * 4 * 4);
        bulk_s2g(gg.pos + b * 3, st.in.pos, kPreBlock * 3 * 4);
        bulk_s2g(gg.scale_raw + b * 3, st.in.scale, kPreBlock * 3 * 4);
        bulk_s2g(gg.f_dc + b * 3, st.in.dc, kPreBlock * 3 * 4);
        bulk_s2g(gg.opacity_raw + b, st.in.opac, kPreBlock * 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      __syncthreads();
      unstage_rows<45>(gg.f_rest, st.in.rest, n0, count);
      unstage_rows<4>(gg.q_raw, st.in.quat, n0, count);
      unstage_rows<3>(gg.pos, st.in.pos, n0, count);
      unstage_rows<3>(gg.scale_raw, st.in.scale, n0, count);
      unstage_rows<3>(gg.f_dc, st.in.dc, n0, count);
      unstage_rows<1>(gg.opacity_raw, st.in.opac, n0, count);
      __syncthreads();
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
}

static FrameView make_view(void* ws, const FrameLayout& L) {
  FrameView f;
  f.rec0 = ws_ptr<float4>(ws, L.rec0); f.rec1 = ws_ptr<float4>(ws, L.rec1); f.rec2 = ws_ptr<float4>(ws, L.rec2);
  f.depth_key = ws_ptr<uint32_t>(ws, L.depth_key);
  f.rect = ws_ptr<uint2>(ws, L.rect);
  f.radius = ws_ptr<uint32_t>(ws, L.radius);
  f.super_touched = ws_ptr<uint32_t>(ws, L.super_touched);
  f.grad_acc = ws_ptr<float>(ws, L.grad_acc);
  f.stats = ws_ptr<b200gs_frame_stats>(ws, L.header);
  return f;
}

cudaError_t launch_preprocess_fwd(const GaussIn& g, const float* c2w, const RenderParams& rp, void* ws,
                                  const FrameLayout& L, cudaStream_t s, uint32_t* depth_hist, bool* hist_done) {
  if (hist_done) *hist_done = false;
  if (g.n <= 0) return cudaSuccess;
  const FrameView f = make_view(ws, L);
  const int grid = ceil_div(g.n, kPreBlock);
  const bool rc = g.scale_raw != nullptr, rs = g.f_dc != nullptr;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  static const bool no_tma = getenv("B200GS_NO_TMA") != nullptr;
  const bool band = rp.row_begin > 0 || rp.row_end < rp.tiles_y;
  if (band) depth_hist = nullptr;
  if (rc && rs && !no_tma && aligned16(g.pos) && aligned16(g.scale_raw) && aligned16(g.q_raw) && aligned16(g.f_dc) &&
      aligned16(g.f_rest) && aligned16(g.opacity_raw)) {
    static int sm_count = 0;
    static std::atomic<uint64_t> attr_set{0};
    once_per_device(attr_set, [] {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaFuncSetAttribute(preprocess_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(PreStage));
    });
    const int pgrid = grid < 3 * sm_count ? grid : 3 * sm_count;     // 3 resident CTAs per SM (2 x 30 KB stages each)
    preprocess_fwd_tma_kernel<<<pgrid, kPreBlock, 2 * sizeof(PreStage), s>>>(g, c2w, rp, f, grid, depth_hist, depth_key_plan(rp));
    if (hist_done) *hist_done = depth_hist != nullptr;
    return cudaGetLastError();
  }
  if (rc && rs) preprocess_fwd_kernel<true, true><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
  else if (rc && !rs) preprocess_fwd_kernel<true, false><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
  else if (!rc && rs) preprocess_fwd_kernel<false, true><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
  else preprocess_fwd_kernel<false, false><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
  return cudaGetLastError();
}

// Band frames from raw parameters: select + project (see above).  cand_ids / cand_key: u32[n] each; select_scratch:
// band_select_scratch_bytes(n) bytes, zeroed here.  *used = false when this route does not apply (precomputed
// sigma / color, unaligned arrays): the caller then runs launch_preprocess_fwd, whose kernels clip to the band too.
cudaError_t launch_band_select(const GaussIn& g, const float* c2w, const RenderParams& rp, void* ws, const FrameLayout& L,
                               uint32_t* flag_words, uint32_t* cand_ids, uint32_t* n_cand, void* select_scratch,
                               size_t select_scratch_bytes, bool* used, cudaStream_t s) {
  *used = false;
  if (g.n <= 0) return cudaSuccess;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  static const bool off = getenv("B200GS_NO_BAND_SELECT") != nullptr;
  // select + project pays when the band is a small part of the frame (measured on the 6M / 4K frame: an eighth of the
  // rows 150 us against 290 us for the fused kernel; half of the rows: the fused kernel wins)
  static const int max_pct = getenv("B200GS_BAND_SELECT_MAX_PCT") ? atoi(getenv("B200GS_BAND_SELECT_MAX_PCT")) : 35;
  if (off || !g.scale_raw || !g.f_dc || !aligned16(g.pos) || !aligned16(g.scale_raw) || !aligned16(g.q_raw) ||
      !aligned16(g.opacity_raw) || (rp.row_end - rp.row_begin) * 100 > max_pct * rp.tiles_y)
    return cudaSuccess;
  if (select_scratch_bytes < band_select_scratch_bytes(g.n)) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(select_scratch, 0, band_select_scratch_bytes(g.n), s);
  if (e != cudaSuccess) return e;
  const FrameView f = make_view(ws, L);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_chunks = ceil_div(g.n, 32 * kSelItems), want = ceil_div(n_chunks, kSelThreads / 32);
  const int grid = want < 5 * sm_count ? want : 5 * sm_count;
  band_select_kernel<<<grid, kSelThreads, 0, s>>>(g, c2w, rp, f.depth_key, flag_words);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const uint32_t n_words = (uint32_t)n_chunks * kSelItems;
  uint32_t* ticket = reinterpret_cast<uint32_t*>(select_scratch);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(select_scratch) + 256);
  band_expand_kernel<<<(n_words + kExpTile - 1) / kExpTile, kExpThreads, 0, s>>>(flag_words, n_words, cand_ids, n_cand, ticket,
                                                                              status);
  *used = true;
  return cudaGetLastError();
}

cudaError_t launch_band_project(const GaussIn& g, const float* c2w, const RenderParams& rp, void* ws, const FrameLayout& L,
                                const uint32_t* cand_ids, const uint32_t* n_cand, uint32_t* cand_key, uint32_t* depth_hist,
                                cudaStream_t s) {
  if (g.n <= 0) return cudaSuccess;
  const FrameView f = make_view(ws, L);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const int want = ceil_div(g.n, kPreBlock);
  const int grid = want < 5 * sm_count ? want : 5 * sm_count;     // 92 registers x 128 threads: 5 resident CTAs per SM
  band_project_kernel<<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f, cand_ids, n_cand, cand_key, depth_hist, depth_key_plan(rp));
  return cudaGetLastError();
}

cudaError_t launch_preprocess_bwd(const GaussIn& g, const GaussGrad& gg, const float* c2w, const RenderParams& rp,
                                  void* ws, const FrameLayout& L, cudaStream_t s) {
  if (g.n <= 0) return cudaSuccess;
  const FrameView f = make_view(ws, L);
  const int grid = ceil_div(g.n, kPreBlock);
  const bool rc = g.scale_raw != nullptr, rs = g.f_dc != nullptr;
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  static const bool no_tma = getenv("B200GS_NO_TMA") != nullptr;
  if (rc && rs && !no_tma && aligned16(g.pos) && aligned16(g.scale_raw) && aligned16(g.q_raw) && aligned16(g.f_dc) &&
      aligned16(g.f_rest) && aligned16(g.opacity_raw) && aligned16(gg.pos) && aligned16(gg.scale_raw) &&
      aligned16(gg.q_raw) && aligned16(gg.f_dc) && aligned16(gg.f_rest) && aligned16(gg.opacity_raw)) {
    static int sm_count = 0;
    static std::atomic<uint64_t> attr_set{0};
    once_per_device(attr_set, [] {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaFuncSetAttribute(preprocess_bwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(PreStageBwd));
    });
    const int pgrid = grid < 3 * sm_count ? grid : 3 * sm_count;
    preprocess_bwd_tma_kernel<<<pgrid, kPreBlock, 2 * sizeof(PreStageBwd), s>>>(g, gg, c2w, rp, f, grid);
    return cudaGetLastError();
  }
  if (rc && rs) preprocess_bwd_kernel<true, true><<<grid, kPreBlock, 0, s>>>(g, gg, c2w, rp, f);
  else if (rc && !rs) preprocess_bwd_kernel<true, false><<<grid, kPreBlock, 0, s>>>(g, gg, c2w, rp, f);
  else if (!rc && rs) preprocess_bwd_kernel<false, true><<<grid, kPreBlock, 0, s>>>(g, gg, c2w, rp, f);
  else preprocess_bwd_kernel<false, false><<<grid, kPreBlock, 0, s>>>(g, gg, c2w, rp, f);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stand-alone Sigma (gaussian.py:71-127) and its backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPreBlock) build_sigma_kernel(int n, const float* __restrict__ scale_raw,
                                                                const float* __restrict__ q_raw,
                                                                float* __restrict__ sigma) {
  __shared__ __align__(16) float s_in[kPreBlock * 3];
  __shared__ __align__(16) float s_out[kPreBlock * 9];
  const int n0 = blockIdx.x * kPreBlock, count = min(kPreBlock, n - n0), tid = threadIdx.x;
  stage_rows<3>(scale_raw, s_in, n0, count);
  __syncthreads();
  if (tid < count) {
    const float sr[3] = {s_in[3 * tid], s_in[3 * tid + 1], s_in[3 * tid + 2]};
    const float4 q4 = ld_stream_f4(reinterpret_cast<const float4*>(q_raw) + n0 + tid);
    const float q[4] = {q4.x, q4.y, q4.z, q4.w};
    QuatScale qs;
    quat_scale_forward(sr, q, qs);
    float full[9];
    sigma_full(qs, full);
#pragma unroll
    for (int k = 0; k < 9; ++k) s_out[9 * tid + k] = full[k];
  }
  __syncthreads();
  unstage_rows<9>(sigma, s_out, n0, count);
}

__global__ void __launch_bounds__(kPreBlock) build_sigma_bwd_kernel(int n, const float* __restrict__ scale_raw,
                                                                    const float* __restrict__ q_raw,
                                                                    const float* __restrict__ g_sigma,
                                                                    float* __restrict__ g_scale,
                                                                    float* __restrict__ g_q) {
  __shared__ __align__(16) float s_in[kPreBlock * 3];
  __shared__ __align__(16) float s_g[kPreBlock * 9];
  const int n0 = blockIdx.x * kPreBlock, count = min(kPreBlock, n - n0), tid = threadIdx.x;
  stage_rows<3>(scale_raw, s_in, n0, count);
  stage_rows<9>(g_sigma, s_g, n0, count);
  __syncthreads();
  float gs[3] = {0.f, 0.f, 0.f};
  if (tid < count) {
    const float sr[3] = {s_in[3 * tid], s_in[3 * tid + 1], s_in[3 * tid + 2]};
    const float4 q4 = ld_stream_f4(reinterpret_cast<const float4*>(q_raw) + n0 + tid);
    const float q[4] = {q4.x, q4.y, q4.z, q4.w};
    QuatScale qs;
    quat_scale_forward(sr, q, qs);
    // the reference's Sigma is a function of the full (not symmetrised) gradient: dL = <G, dSigma>, and
    // dSigma is symmetric, so only the symmetric part of G matters
    float G[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) G[3 * a + b] = 0.5f * (s_g[9 * tid + 3 * a + b] + s_g[9 * tid + 3 * b + a]);
    float gq[4];
    quat_scale_backward(qs, q, G, gs, gq);
    reinterpret_cast<float4*>(g_q)[n0 + tid] = make_float4(gq[0], gq[1], gq[2], gq[3]);
  }
  __syncthreads();
  if (tid < count) { s_in[3 * tid] = gs[0]; s_in[3 * tid + 1] = gs[1]; s_in[3 * tid + 2] = gs[2]; }
  __syncthreads();
  unstage_rows<3>(g_scale, s_in, n0, count);
}

cudaError_t launch_build_sigma(int n, const float* scale_raw, const float* q_raw, float* sigma, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  build_sigma_kernel<<<ceil_div(n, kPreBlock), kPreBlock, 0, s>>>(n, scale_raw, q_raw, sigma);
  return cudaGetLastError();
}
cudaError_t launch_build_sigma_bwd(int n, const float* scale_raw, const float* q_raw, const float* g_sigma,
                                   float* g_scale, float* g_q, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  build_sigma_bwd_kernel<<<ceil_div(n, kPreBlock), kPreBlock, 0, s>>>(n, scale_raw, q_raw, g_sigma, g_scale, g_q);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stand-alone SH evaluation (spherical_harmonics.py:70-166) and its backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPreBlock) eval_sh_kernel(int n, const float* __restrict__ f_dc,
                                                            const float* __restrict__ f_rest,
                                                            const float* __restrict__ pts,
                                                            const float* __restrict__ c2w,
                                                            float* __restrict__ color) {
  __shared__ __align__(16) float s_pos[kPreBlock * 3];
  __shared__ __align__(16) float s_dc[kPreBlock * 3];
  __shared__ __align__(16) float s_rest[kPreBlock * 45];
  const int n0 = blockIdx.x * kPreBlock, count = min(kPreBlock, n - n0), tid = threadIdx.x;
  stage_rows<3>(pts, s_pos, n0, count);
  stage_rows<3>(f_dc, s_dc, n0, count);
  stage_rows<45>(f_rest, s_rest, n0, count);
  __syncthreads();
  float rgb[3] = {0.f, 0.f, 0.f};
  if (tid < count) {
    const float cam[3] = {c2w[3], c2w[7], c2w[11]};
    const float p[3] = {s_pos[3 * tid], s_pos[3 * tid + 1], s_pos[3 * tid + 2]};
    const ViewDir vd = view_dir(p, cam);
    float Y[16], acc[3];
    sh_basis(vd.d, Y);
    sh_color(&s_dc[3 * tid], &s_rest[45 * tid], Y, rgb, acc);
  }
  __syncthreads();
  if (tid < count) { s_pos[3 * tid] = rgb[0]; s_pos[3 * tid + 1] = rgb[1]; s_pos[3 * tid + 2] = rgb[2]; }
  __syncthreads();
  unstage_rows<3>(color, s_pos, n0, count);
}

__global__ void __launch_bounds__(kPreBlock) eval_sh_bwd_kernel(int n, const float* __restrict__ f_dc,
                                                                const float* __restrict__ f_rest,
                                                                const float* __restrict__ pts,
                                                                const float* __restrict__ c2w,
                                                                const float* __restrict__ g_color,
                                                                float* __restrict__ g_dc_out,
                                                                float* __restrict__ g_rest_out,
                                                                float* __restrict__ g_pts_out) {
  __shared__ __align__(16) float s_pos[kPreBlock * 3];
  __shared__ __align__(16) float s_dc[kPreBlock * 3];
  __shared__ __align__(16) float s_gc[kPreBlock * 3];
  __shared__ __align__(16) float s_rest[kPreBlock * 45];
  const int n0 = blockIdx.x * kPreBlock, count = min(kPreBlock, n - n0), tid = threadIdx.x;
  stage_rows<3>(pts, s_pos, n0, count);
  stage_rows<3>(f_dc, s_dc, n0, count);
  stage_rows<3>(g_color, s_gc, n0, count);
  stage_rows<45>(f_rest, s_rest, n0, count);
  __syncthreads();
  float gp[3] = {0.f, 0.f, 0.f}, gdc[3] = {0.f, 0.f, 0.f}, g_acc[3] = {0.f, 0.f, 0.f};
  float Y[16];
  if (tid < count) {
    const float cam[3] = {c2w[3], c2w[7], c2w[11]};
    const float p[3] = {s_pos[3 * tid], s_pos[3 * tid + 1], s_pos[3 * tid + 2]};
    const ViewDir vd = view_dir(p, cam);
    float rgb[3], acc[3], gY[16];
    sh_basis(vd.d, Y);
    sh_color(&s_dc[3 * tid], &s_rest[45 * tid], Y, rgb, acc);
#pragma unroll
    for (int k = 0; k < 16; ++k) gY[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g_acc[c] = s_gc[3 * tid + c] * rgb[c] * (1.f - rgb[c]);
      gdc[c] = g_acc[c] * Y[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) gY[k] = fmaf(g_acc[c], s_rest[45 * tid + 15 * c + k - 1], gY[k]);
    }
    float gd[3];
    sh_basis_backward(vd.d, gY, gd);
    view_dir_backward(vd, gd, gp);
  }
  __syncthreads();
  if (tid < count) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s_pos[3 * tid + c] = gp[c];
      s_dc[3 * tid + c] = gdc[c];
#pragma unroll
      for (int k = 1; k < 16; ++k) s_rest[45 * tid + 15 * c + k - 1] = g_acc[c] * Y[k];
    }
  }
  __syncthreads();
  unstage_rows<3>(g_pts_out, s_pos, n0, count);
  unstage_rows<3>(g_dc_out, s_dc, n0, count);
  unstage_rows<45>(g_rest_out, s_rest, n0, count);
}

cudaError_t launch_eval_sh(int n, const float* f_dc, const float* f_rest, const float* pts, const float* c2w,
                           float* color, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  eval_sh_kernel<<<ceil_div(n, kPreBlock), kPreBlock, 0, s>>>(n, f_dc, f_rest, pts, c2w, color);
  return cudaGetLastError();
}
cudaError_t launch_eval_sh_bwd(int n, const float* f_dc, const float* f_rest, const float* pts, const float* c2w,
                               const float* g_color, float* g_dc, float* g_rest, float* g_pts, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  eval_sh_bwd_kernel<<<ceil_div(n, kPreBlock), kPreBlock, 0, s>>>(n, f_dc, f_rest, pts, c2w, g_color, g_dc, g_rest,
                                                                  g_pts);
  return cudaGetLastError();
}

}  // namespace gs
