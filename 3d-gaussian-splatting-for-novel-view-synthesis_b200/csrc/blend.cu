// Per-tile front-to-back alpha blend (forward) and its replay-in-reverse backward.
//
// Replaces (reference): render.py:317-410 (S16-S17, the Python loop over tiles) and its autograd.
// One 256-thread CTA per 16x16 tile; each warp owns an 8x4 pixel block.  Splat records of the tile's
// depth-sorted list are staged 256 at a time into shared memory (3 x float4 per splat, read back as
// warp-wide broadcasts), every pixel walks the batch front to back, and the CTA stops as soon as every
// pixel's transmittance is <= 5e-5 (render.py:387: a splat contributes iff T *before* it is > 5e-5).
//
// Roofline: FP32 issue + MUFU.EX2; HBM traffic is the 36-B record gather per
// intersection plus 12 B/pixel of output.
#include "common.cuh"

namespace gs {

constexpr int kBlendThreads = 256;
constexpr uint32_t kCountMask = 0x1FFFFFFFu;

struct PixelCoord { int px, py; bool inside; };

__device__ __forceinline__ PixelCoord pixel_of_thread(const RenderParams& rp, int tile_x, int tile_y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PixelCoord p;
  p.px = tile_x * kTile + ((warp & 1) << 3) + (lane & 7);
  p.py = tile_y * kTile + ((warp >> 1) << 2) + (lane >> 3);
  p.inside = (p.px < rp.W) && (p.py < rp.H);
  return p;
}

// alpha of one splat at one pixel; identical instruction sequence in forward and backward.
// Returns 0 when the splat does not contribute (render.py:362-374).
// Shared-memory reads with an explicit 32-bit shared address: keeps the per-visit address arithmetic to one
// IMAD (the generic-pointer form re-derives the shared window base inside the loop).
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2_ftz(float x) {   // q <= chi2 keeps the argument far from the denormal range
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float splat_alpha(const float4& r0, const float4& r1, float pxf, float pyf,
                                             const RenderParams& rp, float& du, float& dv, float& gval,
                                             float& araw) {
  du = pxf - r0.x;
  dv = pyf - r0.y;
  const float q = fmaf(r1.x * dv, dv, fmaf(r0.w * du, dv, (r0.z * du) * du));   // r0.w holds 2*A12
  if (!(q <= rp.chi2)) return 0.f;
  gval = ex2_ftz(q * -0.72134752044448170368f);   // exp(-q/2)
  araw = r1.y * gval;
  const float a = fminf(araw, rp.alpha_max);
  return (a >= rp.alpha_cutoff) ? a : 0.f;
}

// Can this splat contribute to any pixel of the warp's 8x4 block (centre wcx, wcy)?  Bounding box of the
// effective ellipse { q <= min(chi2, 2 ln(opacity / alpha_cutoff)) } (extents precomputed per splat,
// conservative) against the block.  (An exact ellipse-vs-rectangle test was measured: it costs more than
// the few extra visits it removes.)  One lane evaluates one splat.
__device__ __forceinline__ bool splat_touches_block(const float4& r0, float eu, float ev, float wcx, float wcy) {
  return (fabsf(r0.x - wcx) <= eu + 3.5f) && (fabsf(r0.y - wcy) <= ev + 1.5f);
}

__global__ void __launch_bounds__(kBlendThreads) blend_fwd_kernel(RenderParams rp, const uint2* __restrict__ ranges,
                                                                  const uint32_t* __restrict__ vals,
                                                                  const float4* __restrict__ rec0,
                                                                  const float4* __restrict__ rec1,
                                                                  const float4* __restrict__ rec2,
                                                                  float* __restrict__ image,
                                                                  float* __restrict__ final_T,
                                                                  uint32_t* __restrict__ n_contrib) {
  __shared__ float4 s_rec[3][kBlendThreads];      // [0] = rec0, [1] = rec1, [2] = rec2 of the staged batch
  float4* const s0 = s_rec[0];
  float4* const s1 = s_rec[1];
  float4* const s2 = s_rec[2];
  uint32_t sa0;   // opaque copy of the shared address: stops the compiler re-deriving it inside the visit loop
  asm volatile("mov.u32 %0, %1;" : "=r"(sa0) : "r"((uint32_t)__cvta_generic_to_shared(s_rec)));
  constexpr uint32_t kRecStride = kBlendThreads * 16;
  const int tile_x = blockIdx.x, tile_y = rp.row_begin + blockIdx.y;
  const uint2 range = ranges[tile_y * rp.tiles_x + tile_x];
  const PixelCoord pc = pixel_of_thread(rp, tile_x, tile_y);
  const float pxf = (float)pc.px, pyf = (float)pc.py;
  // T is the true transmittance (stored for the backward); Tl is its "live" copy that drops to 0 once
  // T <= 5e-5 (render.py:387: a splat contributes iff the transmittance BEFORE it is > 5e-5), so a finished
  // pixel needs no branch in the visit loop: its weights are simply zero.
  float T = 1.f, Tl = pc.inside ? 1.f : 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last = 0;
  // centre of this warp's 8x4 pixel block: a splat can only touch the block if its centre lies within
  // (ext_u + 3.5, ext_v + 1.5) of it, where ext_* are the conservative half-extents of its footprint
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float wcx = (float)(tile_x * kTile + ((warp & 1) << 3)) + 3.5f;
  const float wcy = (float)(tile_y * kTile + ((warp >> 1) << 2)) + 1.5f;
  for (uint32_t base = range.x; base < range.y; base += kBlendThreads) {
    if (__syncthreads_count(Tl == 0.f) == kBlendThreads) break;
    const uint32_t idx = base + threadIdx.x;
    if (idx < range.y) {
      const uint32_t id = vals[idx];
      s0[threadIdx.x] = rec0[id];
      s1[threadIdx.x] = rec1[id];
      s2[threadIdx.x] = rec2[id];
    }
    __syncthreads();
    const int cnt = (int)min((uint32_t)kBlendThreads, range.y - base);
    const uint32_t list_off = base - range.x + 1u;
    for (int chunk = 0; chunk < cnt; chunk += 32) {
      if (__all_sync(0xffffffffu, Tl == 0.f)) break;
      // each lane tests one splat of the chunk against the warp's pixel block
      const int jt = chunk + lane;
      bool touch = false;
      if (jt < cnt) {
        const float4 t2 = s2[jt];
        touch = splat_touches_block(s0[jt], t2.y, t2.z, wcx, wcy);
      }
      unsigned m = __ballot_sync(0xffffffffu, touch);
      while (m) {
        const int j = chunk + __ffs(m) - 1;
        m &= m - 1;
        const uint32_t aj = sa0 + (uint32_t)j * 16u;
        const float4 r0 = lds_f4(aj), r1 = lds_f4(aj + kRecStride);
        float du, dv, gval, araw;
        const float a = splat_alpha(r0, r1, pxf, pyf, rp, du, dv, gval, araw);
        if (a > 0.f && Tl > 0.f) {
          const float w = a * Tl;
          C0 = fmaf(w, r1.z, C0);
          C1 = fmaf(w, r1.w, C1);
          C2 = fmaf(w, lds_f(aj + 2 * kRecStride), C2);
          T = Tl * (1.f - a);
          last = list_off + (uint32_t)j;
          Tl = (T > 5e-5f) ? T : 0.f;
        }
      }
    }
  }
  if (pc.inside) {
    const size_t pix = (size_t)pc.py * rp.W + pc.px;
    uint32_t flags = last;
    if (C0 > 1.f) flags |= 1u << 29;
    if (C1 > 1.f) flags |= 1u << 30;
    if (C2 > 1.f) flags |= 1u << 31;
    image[3 * pix + 0] = fminf(fmaxf(C0, 0.f), 1.f);
    image[3 * pix + 1] = fminf(fmaxf(C1, 0.f), 1.f);
    image[3 * pix + 2] = fminf(fmaxf(C2, 0.f), 1.f);
    final_T[pix] = T;
    n_contrib[pix] = flags;
  }
}

cudaError_t launch_blend_fwd(const RenderParams& rp, const void* ws, const FrameLayout& L, const uint32_t* vals,
                             float* image, cudaStream_t s) {
  const int rows = rp.row_end - rp.row_begin;
  if (rows <= 0 || rp.tiles_x <= 0) return cudaSuccess;
  dim3 grid(rp.tiles_x, rows);
  blend_fwd_kernel<<<grid, kBlendThreads, 0, s>>>(
      rp, ws_ptr<uint2>(ws, L.ranges), vals, ws_ptr<float4>(ws, L.rec0), ws_ptr<float4>(ws, L.rec1),
      ws_ptr<float4>(ws, L.rec2), image, const_cast<float*>(ws_ptr<float>(ws, L.final_T)),
      const_cast<uint32_t*>(ws_ptr<uint32_t>(ws, L.n_contrib)));
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Backward.  Walks each tile's list back to front from the last contributor, recomputing alpha and
// recovering T_i = T_{i+1} / (1 - alpha_i).  Per-splat gradients (u, v, A11, A12, A22, opacity, r, g, b)
// are reduced over the warp's 32 pixels with shuffles, accumulated per batch in shared memory, and
// flushed with one global atomic per splat and component per tile.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kBlendThreads) blend_bwd_kernel(RenderParams rp, const uint2* __restrict__ ranges,
                                                                  const uint32_t* __restrict__ vals,
                                                                  const float4* __restrict__ rec0,
                                                                  const float4* __restrict__ rec1,
                                                                  const float4* __restrict__ rec2,
                                                                  const float* __restrict__ image_grad,
                                                                  const float* __restrict__ final_T,
                                                                  const uint32_t* __restrict__ n_contrib,
                                                                  float* __restrict__ grad_acc) {
  __shared__ float4 s_rec[2][kBlendThreads];
  float4* const s0 = s_rec[0];
  float4* const s1 = s_rec[1];
  uint32_t sa0;   // opaque copy of the shared address: stops the compiler re-deriving it inside the visit loop
  asm volatile("mov.u32 %0, %1;" : "=r"(sa0) : "r"((uint32_t)__cvta_generic_to_shared(s_rec)));
  constexpr uint32_t kRecStride = kBlendThreads * 16;
  __shared__ float s_cb[kBlendThreads];
  __shared__ float2 s_ext[kBlendThreads];
  __shared__ uint32_t s_id[kBlendThreads];
  __shared__ float s_grad[kBlendThreads][9];
  __shared__ uint32_t s_max;
  const int tile_x = blockIdx.x, tile_y = rp.row_begin + blockIdx.y;
  const uint2 range = ranges[tile_y * rp.tiles_x + tile_x];
  const PixelCoord pc = pixel_of_thread(rp, tile_x, tile_y);
  const float pxf = (float)pc.px, pyf = (float)pc.py;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float wcx = (float)(tile_x * kTile + ((warp & 1) << 3)) + 3.5f;
  const float wcy = (float)(tile_y * kTile + ((warp >> 1) << 2)) + 1.5f;
  // which of the nine totals this lane ends up holding after the halving reduction (-1: none):
  // index within the half-warp's five values = 3*h8 + (h8 ? h4 : 2*h4 + h2 ...) - see the steps below
  int slot = -1;
  {
    const int h16 = (lane >> 4) & 1, h8 = (lane >> 3) & 1, h4 = (lane >> 2) & 1, h2 = (lane >> 1) & 1;
    // step 8: h8=0 keeps x0,x1,x2 ; h8=1 keeps x3,x4,-.  step 4: h4=0 keeps y0,y1 ; h4=1 keeps y2,-.
    // step 2: h2=0 keeps z0 ; h2=1 keeps z1.
    int y = h4 ? 2 : h2;                 // index among (y0,y1,y2) ; (y2 only valid when h2 == 0)
    bool ok = !(h4 && h2);
    int x = h8 ? 3 + y : y;              // index among x0..x4 ; h8=1 has only x3,x4
    if (h8 && y >= 2) ok = false;
    int v = h16 ? 5 + x : x;             // 0..4 = u,v,A11,A12,A22 ; 5..8 = op,r,g,b ; 9 = padding
    if (v >= 9) ok = false;
    if (ok && (lane & 1) == 0) slot = v;
  }
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, T = 1.f;
  uint32_t last = 0;
  if (pc.inside) {
    const size_t pix = (size_t)pc.py * rp.W + pc.px;
    const uint32_t flags = n_contrib[pix];
    last = flags & kCountMask;
    // final clamp(0,1): gradient passes where 0 <= C <= 1 (C >= 0 always)
    g0 = (flags & (1u << 29)) ? 0.f : image_grad[3 * pix + 0];
    g1 = (flags & (1u << 30)) ? 0.f : image_grad[3 * pix + 1];
    g2 = (flags & (1u << 31)) ? 0.f : image_grad[3 * pix + 2];
    T = final_T[pix];
  }
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const uint32_t wlast = __reduce_max_sync(0xffffffffu, last);
  if (lane == 0 && wlast) atomicMax(&s_max, wlast);
  __syncthreads();
  const uint32_t max_last = s_max;
  if (max_last == 0) return;
  float rc0 = 0.f, rc1 = 0.f, rc2 = 0.f;   // colour accumulated behind the current splat, normalised by T_{i+1}
  const int nb = (int)((max_last + kBlendThreads - 1) / kBlendThreads);
  for (int b = nb - 1; b >= 0; --b) {
    const uint32_t boff = (uint32_t)b * kBlendThreads;
    const int cnt = (int)min((uint32_t)kBlendThreads, max_last - boff);
    if ((int)threadIdx.x < cnt) {
      const uint32_t id = vals[range.x + boff + threadIdx.x];
      s_id[threadIdx.x] = id;
      s0[threadIdx.x] = rec0[id];
      s1[threadIdx.x] = rec1[id];
      const float4 t2 = rec2[id];
      s_cb[threadIdx.x] = t2.x;
      s_ext[threadIdx.x] = make_float2(t2.y, t2.z);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) s_grad[threadIdx.x][k] = 0.f;
    __syncthreads();
    for (int chunk = ((cnt - 1) >> 5) << 5; chunk >= 0; chunk -= 32) {
      if (boff + (uint32_t)chunk >= wlast) continue;        // nobody in this warp consumed these splats
      const int jt = chunk + lane;
      bool touch = false;
      if (jt < cnt) {
        const float2 te = s_ext[jt];
        touch = splat_touches_block(s0[jt], te.x, te.y, wcx, wcy);
      }
      unsigned m = __ballot_sync(0xffffffffu, touch);
      while (m) {
        const int bit = 31 - __clz(m);
        m &= ~(1u << bit);
        const int j = chunk + bit;
        const bool active = (boff + (uint32_t)j) < last;
        float a = 0.f, du = 0.f, dv = 0.f, gval = 0.f, araw = 0.f;
        const uint32_t aj = sa0 + (uint32_t)j * 16u;
        const float4 r0 = lds_f4(aj), r1 = lds_f4(aj + kRecStride);
        if (active) a = splat_alpha(r0, r1, pxf, pyf, rp, du, dv, gval, araw);
        const bool hit = a > 0.f;
        if (!__any_sync(0xffffffffu, hit)) continue;
        float v_u = 0.f, v_v = 0.f, v_a11 = 0.f, v_a12 = 0.f, v_a22 = 0.f, v_op = 0.f, v_r = 0.f, v_g = 0.f, v_b = 0.f;
        if (hit) {
          const float cb = s_cb[j];
          const float Ti = T / (1.f - a);
          T = Ti;
          const float w = a * Ti;
          v_r = g0 * w; v_g = g1 * w; v_b = g2 * w;
          const float dalpha = Ti * (g0 * (r1.z - rc0) + g1 * (r1.w - rc1) + g2 * (cb - rc2));
          rc0 = fmaf(a, r1.z - rc0, rc0);
          rc1 = fmaf(a, r1.w - rc1, rc1);
          rc2 = fmaf(a, cb - rc2, rc2);
          const float draw = (araw <= rp.alpha_max) ? dalpha : 0.f;   // clamp_max passes on <=
          v_op = draw * gval;
          const float dq = -0.5f * araw * draw;                       // d/dq of op*exp(-q/2)
          const float B2 = r0.w;            // 2*A12
          v_u = -dq * (2.f * r0.z * du + B2 * dv);
          v_v = -dq * (2.f * r1.x * dv + B2 * du);
          v_a11 = dq * du * du;
          v_a12 = dq * 2.f * du * dv;
          v_a22 = dq * dv * dv;
        }
        // Reduce the 9 per-pixel values over the warp by recursive halving: at every step a lane keeps
        // one half of its values and trades the other half with its partner, so 12 shuffles (instead of
        // 45) leave each total in one lane pair; those lanes then add into shared memory in parallel.
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
        float a0, a1, a2, a3, a4;
        {
          const float k0 = h16 ? v_op : v_u,   t0 = h16 ? v_u : v_op;
          const float k1 = h16 ? v_r : v_v,    t1 = h16 ? v_v : v_r;
          const float k2 = h16 ? v_g : v_a11,  t2 = h16 ? v_a11 : v_g;
          const float k3 = h16 ? v_b : v_a12,  t3 = h16 ? v_a12 : v_b;
          const float k4 = h16 ? 0.f : v_a22,  t4 = h16 ? v_a22 : 0.f;
          a0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 16);
          a1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 16);
          a2 = k2 + __shfl_xor_sync(0xffffffffu, t2, 16);
          a3 = k3 + __shfl_xor_sync(0xffffffffu, t3, 16);
          a4 = k4 + __shfl_xor_sync(0xffffffffu, t4, 16);
        }
        // low half-warp: (u, v, A11, A12, A22)   high half-warp: (op, r, g, b, 0)
        float b0, b1, b2;
        {
          const float k0 = h8 ? a3 : a0, t0 = h8 ? a0 : a3;
          const float k1 = h8 ? a4 : a1, t1 = h8 ? a1 : a4;
          const float k2 = h8 ? 0.f : a2, t2 = h8 ? a2 : 0.f;
          b0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 8);
          b1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 8);
          b2 = k2 + __shfl_xor_sync(0xffffffffu, t2, 8);
        }
        // h8 = 0: (x0, x1, x2)   h8 = 1: (x3, x4, 0)   of the half-warp's five values
        float c0, c1;
        {
          const float k0 = h4 ? b2 : b0, t0 = h4 ? b0 : b2;
          const float k1 = h4 ? 0.f : b1, t1 = h4 ? b1 : 0.f;
          c0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 4);
          c1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 4);
        }
        // h4 = 0: (y0, y1)   h4 = 1: (y2, 0)
        float d0;
        {
          const float k0 = h2 ? c1 : c0, t0 = h2 ? c0 : c1;
          d0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 2);
        }
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
        if (slot >= 0) atomicAdd(&s_grad[j][slot], d0);
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      float* dst = grad_acc + (size_t)s_id[threadIdx.x] * 12;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float v = s_grad[threadIdx.x][k];
        if (v != 0.f) atomicAdd(dst + k, v);
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_blend_bwd(const RenderParams& rp, void* ws, const FrameLayout& L, const uint32_t* vals,
                             const float* image_grad, int n, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(ws_ptr<float>(ws, L.grad_acc), 0, (size_t)(n > 0 ? n : 1) * 12 * sizeof(float), s);
  if (e != cudaSuccess) return e;
  const int rows = rp.row_end - rp.row_begin;
  if (rows <= 0 || rp.tiles_x <= 0) return cudaSuccess;
  dim3 grid(rp.tiles_x, rows);
  blend_bwd_kernel<<<grid, kBlendThreads, 0, s>>>(
      rp, ws_ptr<uint2>(ws, L.ranges), vals, ws_ptr<float4>(ws, L.rec0), ws_ptr<float4>(ws, L.rec1),
      ws_ptr<float4>(ws, L.rec2), image_grad, ws_ptr<float>(ws, L.final_T), ws_ptr<uint32_t>(ws, L.n_contrib),
      ws_ptr<float>(ws, L.grad_acc));
  return cudaGetLastError();
}

}  // namespace gs
