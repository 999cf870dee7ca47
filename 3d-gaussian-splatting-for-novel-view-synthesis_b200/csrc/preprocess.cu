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

// Tile-row bands (a large frame sharded over GPUs): most Gaussians miss the band, so only the 44 B the projection
// needs go through the copy pipeline; the 192 B of SH coefficients are fetched - straight from global memory, one
// row per thread - for the band's survivors only.
struct __align__(128) PreStageBand {
  float quat[kPreBlock * 4];
  float pos[kPreBlock * 3];
  float scale[kPreBlock * 3];
  float opac[kPreBlock];
};
constexpr uint32_t kPreStageBandBytes = kPreBlock * (4 + 3 + 3 + 1) * 4;
template <bool BAND> struct PreStageOf { using type = PreStage; };
template <> struct PreStageOf<true> { using type = PreStageBand; };

__device__ __forceinline__ void sh_color_global(const float* __restrict__ dc, const float* __restrict__ rest,
                                                const float Y[16], float rgb[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = __ldg(dc + c) * Y[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) acc = fmaf(__ldg(rest + 15 * c + k - 1), Y[k], acc);
    rgb[c] = sigmoidf_(acc);
  }
}

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

template <bool BAND>
__global__ void __launch_bounds__(kPreBlock) preprocess_fwd_tma_kernel(GaussIn g, const float* __restrict__ c2w,
                                                                       RenderParams rp, FrameView f, int n_chunks,
                                                                       uint32_t* __restrict__ depth_hist) {
  using Stage = typename PreStageOf<BAND>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage* stage = reinterpret_cast<Stage*>(smem_raw);             // [2]
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ float s_c2w[16];
  __shared__ uint32_t s_tiles;
  __shared__ uint32_t s_dh[4][256];      // digit histograms of the depth keys (the depth sort's 4 passes)
  __shared__ uint32_t s_cand_cnt[kPreBlock / 32];
  __shared__ uint16_t s_cand[kPreBlock];  // BAND: stage rows of the Gaussians that may touch the band
  const int tid = threadIdx.x;
  if (tid < 16) s_c2w[tid] = c2w[tid];
  for (int i = tid; i < 4 * 256; i += kPreBlock) (&s_dh[0][0])[i] = 0;
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
    Stage& st = stage[buf];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this buffer are done
    mbar_expect_tx(&s_bar[buf], BAND ? kPreStageBandBytes : kPreStageBytes);
    if constexpr (!BAND) bulk_g2s(st.rest, g.f_rest + n0 * 45, kPreBlock * 45 * 4, &s_bar[buf]);
    bulk_g2s(st.quat, g.q_raw + n0 * 4, kPreBlock * 4 * 4, &s_bar[buf]);
    bulk_g2s(st.pos, g.pos + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    bulk_g2s(st.scale, g.scale_raw + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
    if constexpr (!BAND) bulk_g2s(st.dc, g.f_dc + n0 * 3, kPreBlock * 3 * 4, &s_bar[buf]);
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
    Stage& st = stage[buf];
    const int n0 = chunk * kPreBlock;
    const int count = min(kPreBlock, g.n - n0);
    if (is_full(chunk)) {
      mbar_wait(&s_bar[buf], (uint32_t)((it >> 1) & 1));
    } else {
      if constexpr (!BAND) stage_rows<45>(g.f_rest, st.rest, n0, count);
      stage_rows<4>(g.q_raw, st.quat, n0, count);
      stage_rows<3>(g.pos, st.pos, n0, count);
      stage_rows<3>(g.scale_raw, st.scale, n0, count);
      if constexpr (!BAND) stage_rows<3>(g.f_dc, st.dc, n0, count);
      stage_rows<1>(g.opacity_raw, st.opac, n0, count);
      __syncthreads();
    }
    // BAND: every thread runs the cheap test on its Gaussian; the candidates that remain (about one in eight, in random
    // positions) are compacted to the front of the CTA so that whole warps skip the expensive path - with the
    // candidates left where they are every warp would still execute it for its four or so live lanes.
    bool live = tid < count;
    int j = tid;                                   // row of the stage this thread projects
    if constexpr (BAND) {
      bool cand = false;
      if (live) {
        const float p[3] = {st.pos[3 * tid], st.pos[3 * tid + 1], st.pos[3 * tid + 2]};
        const float sr[3] = {st.scale[3 * tid], st.scale[3 * tid + 1], st.scale[3 * tid + 2]};
        cand = !band_cannot_touch(p, sr, st.opac[tid], ps, rp);
        if (!cand) {                               // (the frame counters of a band describe the band)
          f.depth_key[n0 + tid] = kCulledKey;
          f.super_touched[n0 + tid] = 0;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, cand);
      if ((tid & 31) == 0) s_cand_cnt[tid >> 5] = (uint32_t)__popc(m);
      __syncthreads();
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kPreBlock / 32; ++w) {
        const uint32_t c = s_cand_cnt[w];
        if (w < (tid >> 5)) before += c;
        total += c;
      }
      if (cand) s_cand[before + __popc(m & ((1u << (tid & 31)) - 1u))] = (uint16_t)tid;
      __syncthreads();
      live = (uint32_t)tid < total;
      if (live) j = s_cand[tid];
    }
    if (live) {
      const int i = n0 + j;
      const float p[3] = {st.pos[3 * j], st.pos[3 * j + 1], st.pos[3 * j + 2]};
      const float sr[3] = {st.scale[3 * j], st.scale[3 * j + 1], st.scale[3 * j + 2]};
      const float4 q4 = reinterpret_cast<const float4*>(st.quat)[j];
      const float q[4] = {q4.x, q4.y, q4.z, q4.w};
      QuatScale qs;
      quat_scale_forward(sr, q, qs);
      float full[9];
      sigma_full(qs, full);
      const Cov3 S = sym_from_full(full);
      Projection o;
      const bool vis = project_gaussian(p, S, st.opac[j], ps, rp, o);
      s7_count += (vis || o.offscreen) ? 1u : 0u;
      vis_count += vis ? 1u : 0u;
      // a survivor whose tile rect misses this rank's band [row_begin, row_end) is dropped from the frame
      int tv0 = 0, tv1 = 0, tiles = 0;
      if (vis) {
        tv0 = max(o.tv0, rp.row_begin); tv1 = min(o.tv1, rp.row_end - 1);
        tiles = (tv1 >= tv0) ? (o.tu1 - o.tu0 + 1) * (tv1 - tv0 + 1) : 0;
      }
      const bool in_frame = tiles > 0;
      // BAND: the keys are compacted before the sort, so the histograms cover the band's survivors only
      if (depth_hist && (!BAND || in_frame)) {
        const uint32_t dk = in_frame ? __float_as_uint(o.z) : kCulledKey;
        atomicAdd(&s_dh[0][dk & 255u], 1u); atomicAdd(&s_dh[1][(dk >> 8) & 255u], 1u);
        atomicAdd(&s_dh[2][(dk >> 16) & 255u], 1u); atomicAdd(&s_dh[3][dk >> 24], 1u);
      }
      if (!in_frame) {
        f.depth_key[i] = kCulledKey;
        f.super_touched[i] = 0;
      } else {
        const ViewDir vd = view_dir(p, ps.cam);
        float Y[16], rgb[3];
        sh_basis(vd.d, Y);
        if constexpr (BAND) {
          sh_color_global(g.f_dc + (size_t)i * 3, g.f_rest + (size_t)i * 45, Y, rgb);
        } else {
          float acc[3];
          sh_color(&st.dc[3 * j], &st.rest[45 * j], Y, rgb, acc);
        }
        float eu, ev;
        conic_extent(o.A11, o.A12, o.A22, rp.chi2, o.op, rp.alpha_cutoff, eu, ev);
        write_splat_record(f, i, o, rgb, eu, ev, rp);
        f.depth_key[i] = __float_as_uint(o.z);
        f.rect[i] = make_uint2((uint32_t)o.tu0 | ((uint32_t)o.tu1 << 16), (uint32_t)tv0 | ((uint32_t)tv1 << 16));
        tiles_sum += (uint32_t)tiles;
        f.super_touched[i] = (uint32_t)((o.tu1 / kSuperX - o.tu0 / kSuperX + 1) * (tv1 / kSuperY - tv0 / kSuperY + 1));
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
    for (int i = tid; i < 4 * 256; i += kPreBlock) {
      const uint32_t c = (&s_dh[0][0])[i];
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
  if (rc && rs && !no_tma && aligned16(g.pos) && aligned16(g.scale_raw) && aligned16(g.q_raw) && aligned16(g.f_dc) &&
      aligned16(g.f_rest) && aligned16(g.opacity_raw)) {
    static int sm_count = 0;
    if (!sm_count) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaFuncSetAttribute(preprocess_fwd_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(PreStage));
    }
    if (rp.row_begin > 0 || rp.row_end < rp.tiles_y) {
      // a band: 11 KB of stages per CTA, latency-bound on the per-survivor SH gathers; 90 registers x 128 threads
      // -> 5 resident CTAs per SM
      const int bgrid = grid < 5 * sm_count ? grid : 5 * sm_count;
      preprocess_fwd_tma_kernel<true><<<bgrid, kPreBlock, 2 * sizeof(PreStageBand), s>>>(g, c2w, rp, f, grid, depth_hist);
    } else {
      const int pgrid = grid < 3 * sm_count ? grid : 3 * sm_count;     // 3 resident CTAs per SM (2 x 30 KB stages each)
      preprocess_fwd_tma_kernel<false><<<pgrid, kPreBlock, 2 * sizeof(PreStage), s>>>(g, c2w, rp, f, grid, depth_hist);
    }
    if (hist_done) *hist_done = depth_hist != nullptr;
    return cudaGetLastError();
  }
  if (rc && rs) preprocess_fwd_kernel<true, true><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
  else if (rc && !rs) preprocess_fwd_kernel<true, false><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
  else if (!rc && rs) preprocess_fwd_kernel<false, true><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
  else preprocess_fwd_kernel<false, false><<<grid, kPreBlock, 0, s>>>(g, c2w, rp, f);
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
    if (!sm_count) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaFuncSetAttribute(preprocess_bwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(PreStageBwd));
    }
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
