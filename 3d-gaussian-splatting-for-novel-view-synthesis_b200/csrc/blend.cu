// Per-tile front-to-back alpha blend (forward) and its replay-in-reverse backward.
//
// Replaces (reference): render.py:317-410 (S16-S17, the Python loop over tiles) and its autograd.
// One 256-thread CTA per 16x16 tile; each warp owns an 8x4 pixel block.  Splat records of the tile's
// depth-sorted list are staged 256 at a time into shared memory, every warp first compacts the batch to
// the splats whose footprint can touch its pixel block (one splat per lane, ballot + prefix popcount into
// a per-warp list of shared addresses), then every pixel walks that list front to back, and the CTA stops
// as soon as every pixel's transmittance is <= 5e-5 (render.py:387: a splat contributes iff T *before* it
// is > 5e-5).
//
// The splat records arrive from preprocess already in the form the per-(pixel, splat) visit wants
// (write_splat_record, preprocess.cu): with c = -log2(e)/2 the exponent of  alpha_raw = op * exp(-q/2) = 2^e  is
//     e = c*A11*du^2 + c*2*A12*du*dv + c*A22*dv^2 + log2(op)
// and both gates of render.py:362-374 collapse into one compare:
//     q <= chi2  and  min(alpha_raw, alpha_max) >= alpha_cutoff   <=>   e >= max(c*chi2 + log2(op), log2(alpha_cutoff)).
// The list entries (Gaussian ids) of the next batch are fetched while the current one is blended, so a batch
// costs one global round trip, not two.  (A cp.async double-buffered staging was measured and dropped: the
// extra shared memory shrinks L1 and the kernel gets slower - 262 us -> 360 us on the headline frame.)
//
// Roofline: FP32 issue + MUFU.EX2; HBM traffic is the 36-B record gather per
// intersection plus 12 B/pixel of output.
#include <stdlib.h>

#include "common.cuh"

namespace gs {

constexpr int kBlendThreads = 256;
constexpr int kBlendWarps = kBlendThreads / 32;
constexpr uint32_t kCountMask = 0x1FFFFFFFu;
// (the records arrive pre-scaled by c = -log2(e)/2 = -0.7213475: exp(-q/2) = 2^(c q); see write_splat_record in preprocess.cu)
constexpr uint32_t kRecStride = kBlendThreads * 16;      // bytes between the staged float4 planes

struct PixelCoord { int px, py; bool inside; };

__device__ __forceinline__ PixelCoord pixel_of_thread(const RenderParams& rp, int tile_x, int tile_y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PixelCoord p;
  p.px = tile_x * kTile + ((warp & 1) << 3) + (lane & 7);
  p.py = tile_y * kTile + ((warp >> 1) << 2) + (lane >> 3);
  p.inside = (p.px < rp.W) && (p.py < rp.H);
  return p;
}

// Shared-memory reads with an explicit 32-bit shared address (the per-warp lists hold such addresses).
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float ex2_ftz(float x) {   // arguments are >= log2(alpha_cutoff): far from the denormal range
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Shared address of a __shared__ object as an opaque register value: stops the compiler from re-deriving
// the shared window base (S2UR/ULEA chains) inside the visit loops.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  uint32_t a;
  asm volatile("mov.u32 %0, %1;" : "=r"(a) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return a;
}

// exponent e of alpha_raw = 2^e at pixel offset (du, dv) from the splat centre (record layout: preprocess.cu)
__device__ __forceinline__ float splat_exponent(const float4& s0, const float4& s1, float du, float dv) {
  const float t1 = fmaf(s0.z, du, s0.w * dv);
  const float t2 = fmaf(s1.x * dv, dv, s1.y);
  return fmaf(du, t1, t2);
}

// Compacts the staged splats [0, lim) of the buffer at `buf` that can touch this warp's 8x4 pixel block (centre
// wcx, wcy) into the warp's list (shared addresses of their rec0 entries, in list order); one lane tests one
// splat: bounding box of the effective ellipse { q <= min(chi2, 2 ln(opacity / alpha_cutoff)) } (half-extents
// precomputed per splat, conservative) against the block.  (An exact ellipse-vs-rectangle test was measured:
// it costs more than the few extra visits it removes.)  Returns the number of entries.
template <uint32_t kStride = kRecStride>
__device__ __forceinline__ uint32_t compact_touching(uint32_t buf, uint32_t sa_list, int lim, int lane, float wcx,
                                                     float wcy, float hx = 3.5f, float hy = 1.5f) {
  uint32_t nw = 0;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t a = buf + (uint32_t)lane * 16u;
  for (int jt = lane; jt - lane < lim; jt += 32, a += 512u) {
    const float2 c = lds_f2(a);                                   // (u, v); slots >= lim hold stale data: masked below
    const float2 e = lds_f2(a + 2 * kStride + 8u);                // (ext_u, ext_v)
    const bool touch = (jt < lim) & (fabsf(c.x - wcx) <= e.x + hx) & (fabsf(c.y - wcy) <= e.y + hy);
    const unsigned m = __ballot_sync(0xffffffffu, touch);
    if (touch) sts_u32(sa_list + (nw + __popc(m & lt)) * 4u, a);
    nw += __popc(m);
  }
  __syncwarp();
  return nw;
}

// OVERLAPPED only selects a second instance of the same code: the frame pipeline launches it beside the next frame's
// binning kernels and gives it a larger shared-memory carve-out (launch_blend_fwd), a per-function attribute.
template <bool OVERLAPPED>
__global__ void __launch_bounds__(kBlendThreads) blend_fwd_kernel(RenderParams rp, const uint2* __restrict__ ranges,
                                                                  const uint32_t* __restrict__ vals,
                                                                  const float4* __restrict__ rec0,
                                                                  const float4* __restrict__ rec1,
                                                                  const float4* __restrict__ rec2,
                                                                  float* __restrict__ image,
                                                                  float* __restrict__ final_T,
                                                                  uint32_t* __restrict__ n_contrib, int row_stores) {
  __shared__ float4 s_rec[3][kBlendThreads];
  __shared__ uint32_t s_list[kBlendWarps][kBlendThreads];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sa0 = smem_addr(s_rec);
  const uint32_t sa_list = smem_addr(&s_list[warp][0]);
  const int tile_x = blockIdx.x, tile_y = rp.row_begin + blockIdx.y;
  const uint2 range = ranges[tile_y * rp.tiles_x + tile_x];
  const PixelCoord pc = pixel_of_thread(rp, tile_x, tile_y);
  const float pxf = (float)pc.px, pyf = (float)pc.py;
  const float amax = rp.alpha_max;
  float T = 1.f, Tl = pc.inside ? 1.f : 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last16 = 0;
  const float wcx = (float)(tile_x * kTile + ((warp & 1) << 3)) + 3.5f;
  const float wcy = (float)(tile_y * kTile + ((warp >> 1) << 2)) + 1.5f;
  uint32_t idx = range.x + threadIdx.x;
  uint32_t id_cur = (idx < range.y) ? vals[idx] : 0u;
  for (uint32_t base = range.x; base < range.y; base += kBlendThreads) {
    if (__syncthreads_count(Tl == 0.f) == kBlendThreads) break;
    if (idx < range.y) {
      const float4 a0 = rec0[id_cur], a1 = rec1[id_cur], a2 = rec2[id_cur];
      s_rec[0][threadIdx.x] = a0;
      s_rec[1][threadIdx.x] = a1;
      s_rec[2][threadIdx.x] = a2;
    }
    idx += kBlendThreads;
    id_cur = (idx < range.y) ? vals[idx] : 0u;
    __syncthreads();
    const int cnt = (int)min((uint32_t)kBlendThreads, range.y - base);
    if (__all_sync(0xffffffffu, Tl == 0.f)) continue;
    const uint32_t nw = compact_touching(sa0, sa_list, cnt, lane, wcx, wcy);
    const uint32_t lastbase = (base - range.x + 1u) * 16u - sa0;
    for (uint32_t i = 0; i < nw;) {
      if (__all_sync(0xffffffffu, Tl == 0.f)) break;
      const uint32_t iend = min(i + 32u, nw);
#pragma unroll 4
      for (; i < iend; ++i) {
        const uint32_t aj = lds_u32(sa_list + i * 4u);
        const float4 r0 = lds_f4(aj), r1 = lds_f4(aj + kRecStride);
        const float du = pxf - r0.x, dv = pyf - r0.y;
        const float e = splat_exponent(r0, r1, du, dv);
        if (e >= r1.z && Tl > 0.f) {
          const float a = fminf(ex2_ftz(e), amax);
          const float w = a * Tl;
          const float2 gb = lds_f2(aj + 2 * kRecStride);
          C0 = fmaf(w, r1.w, C0);
          C1 = fmaf(w, gb.x, C1);
          C2 = fmaf(w, gb.y, C2);
          T = fmaf(-a, Tl, Tl);
          last16 = aj + lastbase;
          Tl = (T > 5e-5f) ? T : 0.f;
        }
      }
    }
  }
  if (pc.inside) {
    const size_t pix = (size_t)pc.py * rp.W + pc.px;
    uint32_t flags = last16 >> 4;
    if (C0 > 1.f) flags |= 1u << 29;
    if (C1 > 1.f) flags |= 1u << 30;
    if (C2 > 1.f) flags |= 1u << 31;
    final_T[pix] = T;
    n_contrib[pix] = flags;
    if (!row_stores) {
      image[3 * pix + 0] = fminf(fmaxf(C0, 0.f), 1.f);
      image[3 * pix + 1] = fminf(fmaxf(C1, 0.f), 1.f);
      image[3 * pix + 2] = fminf(fmaxf(C2, 0.f), 1.f);
    }
  }
  if (!row_stores) return;
  // Frame buffer on ANOTHER GPU (tile-row bands: image = a peer-mapped address): the tile's 16 x 16 x 3 colours leave
  // through shared memory, so that every store instruction of a warp writes 128 consecutive bytes of an image row (a
  // tile row is 192 B) instead of every third float of four rows - whole 32-byte sectors, each written once, a third of
  // the write requests on NVLink.  (Into local memory the L2 merges the three partial stores and this costs 2 %.)
  float* s_out = reinterpret_cast<float*>(&s_rec[0][0]);
  __syncthreads();                        // every warp is done with the staged records
  {
    const int o = 3 * ((((warp >> 1) << 2) + (lane >> 3)) * kTile + ((warp & 1) << 3) + (lane & 7));
    s_out[o + 0] = fminf(fmaxf(C0, 0.f), 1.f);
    s_out[o + 1] = fminf(fmaxf(C1, 0.f), 1.f);
    s_out[o + 2] = fminf(fmaxf(C2, 0.f), 1.f);
  }
  __syncthreads();
  const int x0 = tile_x * kTile, y0 = tile_y * kTile;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int f = k * kBlendThreads + (int)threadIdx.x;
    const int r = f / (3 * kTile), j = f - r * (3 * kTile);
    if (y0 + r < rp.H && x0 + j / 3 < rp.W) image[3 * ((size_t)(y0 + r) * rp.W + x0) + j] = s_out[f];
  }
}

cudaError_t launch_blend_fwd(const RenderParams& rp, const void* ws, const FrameLayout& L, const uint32_t* vals,
                             float* image, cudaStream_t s, bool overlapped, bool row_stores) {
  const int rows = rp.row_end - rp.row_begin;
  if (rows <= 0 || rp.tiles_x <= 0) return cudaSuccess;
  dim3 grid(rp.tiles_x, rows);
  if (overlapped) {
    // Six resident CTAs use 120 KB of shared memory; by default the SM is then configured with just enough of it
    // and the (high-priority) binning CTAs of the next frame - 46 KB for a radix pass - have to wait until blend
    // CTAs retire.  65 % leaves them room at the price of some L1: 2 175 -> 2 230 pipelined frames/s on the headline
    // workload (50 %: 2 142, 80 %: 2 147, 100 %: 1 832; one frame at a time it costs 0.8 %, hence the second instance).
    static std::atomic<uint64_t> once{0};
    once_per_device(once, [] { cudaFuncSetAttribute(blend_fwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 65); });
    blend_fwd_kernel<true><<<grid, kBlendThreads, 0, s>>>(
        rp, ws_ptr<uint2>(ws, L.ranges), vals, ws_ptr<float4>(ws, L.rec0), ws_ptr<float4>(ws, L.rec1),
        ws_ptr<float4>(ws, L.rec2), image, const_cast<float*>(ws_ptr<float>(ws, L.final_T)),
        const_cast<uint32_t*>(ws_ptr<uint32_t>(ws, L.n_contrib)), row_stores ? 1 : 0);
    return cudaGetLastError();
  }
  blend_fwd_kernel<false><<<grid, kBlendThreads, 0, s>>>(
      rp, ws_ptr<uint2>(ws, L.ranges), vals, ws_ptr<float4>(ws, L.rec0), ws_ptr<float4>(ws, L.rec1),
      ws_ptr<float4>(ws, L.rec2), image, const_cast<float*>(ws_ptr<float>(ws, L.final_T)),
      const_cast<uint32_t*>(ws_ptr<uint32_t>(ws, L.n_contrib)), row_stores ? 1 : 0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Backward.  Walks each tile's list back to front from the last contributor, recomputing alpha and
// recovering T_i = T_{i+1} / (1 - alpha_i).  Per (pixel, splat) it forms dL/dq (q = the Mahalanobis form)
// and its first and second moments in the pixel offset, plus dL/dcolour:
//     grad_acc row = (Mx, My, Mxx, Mxy, Myy, M0, r, g, b),  M0 = sum dL/dq, Mx = sum dL/dq*du, Mxy = sum dL/dq*du*dv, ...
// from which the per-Gaussian pass (preprocess backward, splat_grad_from_moments) derives dL/d(u, v, conic,
// opacity) - that algebra is per Gaussian, not per pixel.  The nine values are reduced over the warp's 32
// pixels by recursive halving (12 shuffles) and added to grad_acc with one coalesced RED per (warp, splat).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlendThreads) blend_bwd_kernel(RenderParams rp, const uint2* __restrict__ ranges,
                                                                  const uint32_t* __restrict__ vals,
                                                                  const float4* __restrict__ rec0,
                                                                  const float4* __restrict__ rec1,
                                                                  const float4* __restrict__ rec2,
                                                                  const float* __restrict__ image_grad,
                                                                  const float* __restrict__ final_T,
                                                                  const uint32_t* __restrict__ n_contrib,
                                                                  float* __restrict__ grad_acc) {
  __shared__ float4 s_rec[3][kBlendThreads];         // staged records
  __shared__ uint32_t s_id[kBlendThreads];           // Gaussian ids of the staged splats
  __shared__ uint32_t s_list[kBlendWarps][kBlendThreads];
  __shared__ uint32_t s_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sa0 = smem_addr(s_rec);
  const uint32_t sa_id = smem_addr(s_id);
  const uint32_t sa_list = smem_addr(&s_list[warp][0]);
  const int tile_x = blockIdx.x, tile_y = rp.row_begin + blockIdx.y;
  const uint2 range = ranges[tile_y * rp.tiles_x + tile_x];
  const PixelCoord pc = pixel_of_thread(rp, tile_x, tile_y);
  const float pxf = (float)pc.px, pyf = (float)pc.py;
  const float amax = rp.alpha_max;
  const float wcx = (float)(tile_x * kTile + ((warp & 1) << 3)) + 3.5f;
  const float wcy = (float)(tile_y * kTile + ((warp >> 1) << 2)) + 1.5f;
  // which of the nine totals this lane ends up holding after the halving reduction (-1: none)
  int slot = -1;
  {
    const int h16 = (lane >> 4) & 1, h8 = (lane >> 3) & 1, h4 = (lane >> 2) & 1, h2 = (lane >> 1) & 1;
    // step 8: h8=0 keeps x0,x1,x2 ; h8=1 keeps x3,x4,-.  step 4: h4=0 keeps y0,y1 ; h4=1 keeps y2,-.
    // step 2: h2=0 keeps z0 ; h2=1 keeps z1.
    int y = h4 ? 2 : h2;                 // index among (y0,y1,y2) ; (y2 only valid when h2 == 0)
    bool ok = !(h4 && h2);
    int x = h8 ? 3 + y : y;              // index among x0..x4 ; h8=1 has only x3,x4
    if (h8 && y >= 2) ok = false;
    int v = h16 ? 5 + x : x;             // 0..4 = Mx,My,Mxx,Mxy,Myy ; 5..8 = M0,r,g,b ; 9 = padding
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
  float* const my_acc = grad_acc + (slot >= 0 ? slot : 0);
  const int nb = (int)((max_last + kBlendThreads - 1) / kBlendThreads);
  // Batches back to front; the list entries (Gaussian ids) are fetched one batch ahead.
  const uint32_t* const list = vals + range.x;
  uint32_t pos = (uint32_t)(nb - 1) * kBlendThreads + threadIdx.x;       // list position this thread stages
  uint32_t id_cur = (pos < max_last) ? list[pos] : 0u;
  const uint32_t buf = sa0;
  for (int b = nb - 1; b >= 0; --b) {
    const uint32_t boff = (uint32_t)b * kBlendThreads;
    const int cnt = (int)min((uint32_t)kBlendThreads, max_last - boff);
    if (b != nb - 1) __syncthreads();      // every warp is done with the previous batch
    if ((int)threadIdx.x < cnt) {
      const float4 a0 = rec0[id_cur], a1 = rec1[id_cur], a2 = rec2[id_cur];
      s_rec[0][threadIdx.x] = a0;
      s_rec[1][threadIdx.x] = a1;
      s_rec[2][threadIdx.x] = a2;
      s_id[threadIdx.x] = id_cur;
    }
    if (b >= 1) id_cur = list[boff - kBlendThreads + threadIdx.x];
    __syncthreads();
    if (wlast > boff) {                                                // else nobody in this warp consumed these splats
      const int lim = (int)min((uint32_t)cnt, wlast - boff);
      const uint32_t nw = compact_touching(buf, sa_list, lim, lane, wcx, wcy);
      // splat j of this batch was consumed by this pixel iff boff + j < last  <=>  its rec0 address < last_addr
      const int last_addr = (int)buf + ((int)last - (int)boff) * 16;
      for (uint32_t i = nw; i-- > 0;) {
        const uint32_t aj = lds_u32(sa_list + i * 4u);
        const float4 r0 = lds_f4(aj), r1 = lds_f4(aj + kRecStride);
        const float du = pxf - r0.x, dv = pyf - r0.y;
        const float e = splat_exponent(r0, r1, du, dv);
        const bool hit = ((int)aj < last_addr) && (e >= r1.z);
        if (!__any_sync(0xffffffffu, hit)) continue;
        const float2 gb = lds_f2(aj + 2 * kRecStride);
        float mx = 0.f, my = 0.f, mxx = 0.f, mxy = 0.f, myy = 0.f, m0 = 0.f, v_r = 0.f, v_g = 0.f, v_b = 0.f;
        if (hit) {
          const float araw = ex2_ftz(e);
          const float a = fminf(araw, amax);
          const float Ti = __fdividef(T, 1.f - a);
          T = Ti;
          const float w = a * Ti;
          v_r = g0 * w; v_g = g1 * w; v_b = g2 * w;
          const float d0c = r1.w - rc0, d1c = gb.x - rc1, d2c = gb.y - rc2;
          const float dalpha = Ti * fmaf(g2, d2c, fmaf(g1, d1c, g0 * d0c));
          rc0 = fmaf(a, d0c, rc0);
          rc1 = fmaf(a, d1c, rc1);
          rc2 = fmaf(a, d2c, rc2);
          const float draw = (araw <= amax) ? dalpha : 0.f;            // clamp_max passes on <=
          m0 = -0.5f * araw * draw;                                    // dL/dq: d/dq of op*exp(-q/2)
          mx = m0 * du; my = m0 * dv;
          mxx = mx * du; mxy = mx * dv; myy = my * dv;
        }
        // Reduce the 9 per-pixel values over the warp by recursive halving: at every step a lane keeps
        // one half of its values and trades the other half with its partner, so 12 shuffles (instead of
        // 45) leave each total in one lane pair; those lanes then issue one RED.
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
        float a0, a1, a2, a3, a4;
        {
          const float k0 = h16 ? m0 : mx,    t0 = h16 ? mx : m0;
          const float k1 = h16 ? v_r : my,   t1 = h16 ? my : v_r;
          const float k2 = h16 ? v_g : mxx,  t2 = h16 ? mxx : v_g;
          const float k3 = h16 ? v_b : mxy,  t3 = h16 ? mxy : v_b;
          const float k4 = h16 ? 0.f : myy,  t4 = h16 ? myy : 0.f;
          a0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 16);
          a1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 16);
          a2 = k2 + __shfl_xor_sync(0xffffffffu, t2, 16);
          a3 = k3 + __shfl_xor_sync(0xffffffffu, t3, 16);
          a4 = k4 + __shfl_xor_sync(0xffffffffu, t4, 16);
        }
        // low half-warp: (Mx, My, Mxx, Mxy, Myy)   high half-warp: (M0, r, g, b, 0)
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
        if (slot >= 0) {
          const uint32_t id = lds_u32(sa_id + ((aj - buf) >> 2));
          atomicAdd(my_acc + (size_t)id * 12, d0);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward, PPL pixels per lane.  Half of a hit visit of the kernel above is the nine-value warp reduction and the
// RED behind it, and both are per (warp, splat), not per pixel: a warp that covers an 8 x (4 PPL) pixel block - lane
// (x, y) owns the pixels (x, y + 4 k), k < PPL - sums its pixels' contributions in registers first and pays the
// reduction, the RED and the record loads once for PPL times as many pixels.  PPL = 2: four warps per tile, 8x8 blocks.
// ------------------------------------------------------------------------------------------------
template <int PPL>
__global__ void __launch_bounds__(kBlendThreads / PPL) blend_bwd_ppl_kernel(
    RenderParams rp, const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals, const float4* __restrict__ rec0,
    const float4* __restrict__ rec1, const float4* __restrict__ rec2, const float* __restrict__ image_grad,
    const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib, float* __restrict__ grad_acc) {
  constexpr int kThreads = kBlendThreads / PPL;
  constexpr int kWarps = kThreads / 32;
  __shared__ float4 s_rec[3][kBlendThreads];         // staged records (256 splats per batch, PPL per thread)
  __shared__ uint32_t s_id[kBlendThreads];
  __shared__ uint32_t s_list[kWarps][kBlendThreads];
  __shared__ uint32_t s_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sa0 = smem_addr(s_rec);
  const uint32_t sa_id = smem_addr(s_id);
  const uint32_t sa_list = smem_addr(&s_list[warp][0]);
  const int tile_x = blockIdx.x, tile_y = rp.row_begin + blockIdx.y;
  const uint2 range = ranges[tile_y * rp.tiles_x + tile_x];
  // warp block: 8 wide, 4 PPL high; warps tile the 16x16 tile two across
  const int bx = tile_x * kTile + ((warp & 1) << 3), by = tile_y * kTile + (warp >> 1) * (4 * PPL);
  const int px = bx + (lane & 7), py0 = by + (lane >> 3);
  const float pxf = (float)px;
  const float amax = rp.alpha_max;
  const float wcx = (float)bx + 3.5f, wcy = (float)by + (2.f * PPL - 0.5f);
  int slot = -1;       // which of the nine totals this lane holds after the halving reduction (see blend_bwd_kernel)
  {
    const int h16 = (lane >> 4) & 1, h8 = (lane >> 3) & 1, h4 = (lane >> 2) & 1, h2 = (lane >> 1) & 1;
    int y = h4 ? 2 : h2;
    bool ok = !(h4 && h2);
    int x = h8 ? 3 + y : y;
    if (h8 && y >= 2) ok = false;
    int v = h16 ? 5 + x : x;
    if (v >= 9) ok = false;
    if (ok && (lane & 1) == 0) slot = v;
  }
  float g0[PPL], g1[PPL], g2[PPL], T[PPL], pyf[PPL], rc0[PPL], rc1[PPL], rc2[PPL];
  uint32_t last[PPL];
  uint32_t lmax = 0;
#pragma unroll
  for (int k = 0; k < PPL; ++k) {
    const int py = py0 + 4 * k;
    pyf[k] = (float)py;
    g0[k] = g1[k] = g2[k] = 0.f; T[k] = 1.f; last[k] = 0; rc0[k] = rc1[k] = rc2[k] = 0.f;
    if (px < rp.W && py < rp.H) {
      const size_t pix = (size_t)py * rp.W + px;
      const uint32_t flags = n_contrib[pix];
      last[k] = flags & kCountMask;
      g0[k] = (flags & (1u << 29)) ? 0.f : image_grad[3 * pix + 0];
      g1[k] = (flags & (1u << 30)) ? 0.f : image_grad[3 * pix + 1];
      g2[k] = (flags & (1u << 31)) ? 0.f : image_grad[3 * pix + 2];
      T[k] = final_T[pix];
    }
    lmax = max(lmax, last[k]);
  }
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  const uint32_t wlast = __reduce_max_sync(0xffffffffu, lmax);
  if (lane == 0 && wlast) atomicMax(&s_max, wlast);
  __syncthreads();
  const uint32_t max_last = s_max;
  if (max_last == 0) return;
  float* const my_acc = grad_acc + (slot >= 0 ? slot : 0);
  const int nb = (int)((max_last + kBlendThreads - 1) / kBlendThreads);
  const uint32_t* const list = vals + range.x;
  uint32_t id_cur[PPL];
#pragma unroll
  for (int k = 0; k < PPL; ++k) {
    const uint32_t pos = (uint32_t)(nb - 1) * kBlendThreads + k * kThreads + threadIdx.x;
    id_cur[k] = (pos < max_last) ? list[pos] : 0u;
  }
  const uint32_t buf = sa0;
  for (int b = nb - 1; b >= 0; --b) {
    const uint32_t boff = (uint32_t)b * kBlendThreads;
    const int cnt = (int)min((uint32_t)kBlendThreads, max_last - boff);
    if (b != nb - 1) __syncthreads();      // every warp is done with the previous batch
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
      const int j = k * kThreads + (int)threadIdx.x;
      if (j < cnt) {
        const float4 a0 = rec0[id_cur[k]], a1 = rec1[id_cur[k]], a2 = rec2[id_cur[k]];
        s_rec[0][j] = a0;
        s_rec[1][j] = a1;
        s_rec[2][j] = a2;
        s_id[j] = id_cur[k];
      }
    }
    if (b >= 1) {
#pragma unroll
      for (int k = 0; k < PPL; ++k) id_cur[k] = list[boff - kBlendThreads + k * kThreads + threadIdx.x];
    }
    __syncthreads();
    if (wlast > boff) {                                                // else nobody in this warp consumed these splats
      const int lim = (int)min((uint32_t)cnt, wlast - boff);
      const uint32_t nw = compact_touching(buf, sa_list, lim, lane, wcx, wcy, 3.5f, 2.f * PPL - 0.5f);
      int last_addr[PPL];
#pragma unroll
      for (int k = 0; k < PPL; ++k) last_addr[k] = (int)buf + ((int)last[k] - (int)boff) * 16;
      for (uint32_t i = nw; i-- > 0;) {
        const uint32_t aj = lds_u32(sa_list + i * 4u);
        const float4 r0 = lds_f4(aj), r1 = lds_f4(aj + kRecStride);
        const float du = pxf - r0.x;
        float dv[PPL], e[PPL];
        bool hit[PPL], any = false;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          dv[k] = pyf[k] - r0.y;
          e[k] = splat_exponent(r0, r1, du, dv[k]);
          hit[k] = ((int)aj < last_addr[k]) && (e[k] >= r1.z);
          any |= hit[k];
        }
        if (!__any_sync(0xffffffffu, any)) continue;
        const float2 gb = lds_f2(aj + 2 * kRecStride);
        float mx = 0.f, my = 0.f, mxx = 0.f, mxy = 0.f, myy = 0.f, m0 = 0.f, v_r = 0.f, v_g = 0.f, v_b = 0.f;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          if (hit[k]) {
            const float araw = ex2_ftz(e[k]);
            const float a = fminf(araw, amax);
            const float Ti = __fdividef(T[k], 1.f - a);
            T[k] = Ti;
            const float w = a * Ti;
            v_r = fmaf(g0[k], w, v_r); v_g = fmaf(g1[k], w, v_g); v_b = fmaf(g2[k], w, v_b);
            const float d0c = r1.w - rc0[k], d1c = gb.x - rc1[k], d2c = gb.y - rc2[k];
            const float dalpha = Ti * fmaf(g2[k], d2c, fmaf(g1[k], d1c, g0[k] * d0c));
            rc0[k] = fmaf(a, d0c, rc0[k]);
            rc1[k] = fmaf(a, d1c, rc1[k]);
            rc2[k] = fmaf(a, d2c, rc2[k]);
            const float draw = (araw <= amax) ? dalpha : 0.f;            // clamp_max passes on <=
            const float q = -0.5f * araw * draw;                          // dL/dq of this pixel
            const float qv = q * dv[k];
            m0 += q;
            my += qv;
            myy = fmaf(qv, dv[k], myy);
          }
        }
        mx = m0 * du; mxx = mx * du; mxy = my * du;                       // du is shared by the lane's pixels
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
        float a0, a1, a2, a3, a4;
        {
          const float k0 = h16 ? m0 : mx,    t0 = h16 ? mx : m0;
          const float k1 = h16 ? v_r : my,   t1 = h16 ? my : v_r;
          const float k2 = h16 ? v_g : mxx,  t2 = h16 ? mxx : v_g;
          const float k3 = h16 ? v_b : mxy,  t3 = h16 ? mxy : v_b;
          const float k4 = h16 ? 0.f : myy,  t4 = h16 ? myy : 0.f;
          a0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 16);
          a1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 16);
          a2 = k2 + __shfl_xor_sync(0xffffffffu, t2, 16);
          a3 = k3 + __shfl_xor_sync(0xffffffffu, t3, 16);
          a4 = k4 + __shfl_xor_sync(0xffffffffu, t4, 16);
        }
        float b0, b1, b2;
        {
          const float k0 = h8 ? a3 : a0, t0 = h8 ? a0 : a3;
          const float k1 = h8 ? a4 : a1, t1 = h8 ? a1 : a4;
          const float k2 = h8 ? 0.f : a2, t2 = h8 ? a2 : 0.f;
          b0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 8);
          b1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 8);
          b2 = k2 + __shfl_xor_sync(0xffffffffu, t2, 8);
        }
        float c0, c1;
        {
          const float k0 = h4 ? b2 : b0, t0 = h4 ? b0 : b2;
          const float k1 = h4 ? 0.f : b1, t1 = h4 ? b1 : 0.f;
          c0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 4);
          c1 = k1 + __shfl_xor_sync(0xffffffffu, t1, 4);
        }
        float d0;
        {
          const float k0 = h2 ? c1 : c0, t0 = h2 ? c0 : c1;
          d0 = k0 + __shfl_xor_sync(0xffffffffu, t0, 2);
        }
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
        if (slot >= 0) {
          const uint32_t id = lds_u32(sa_id + ((aj - buf) >> 2));
          atomicAdd(my_acc + (size_t)id * 12, d0);
        }
      }
    }
  }
}

cudaError_t launch_blend_bwd(const RenderParams& rp, void* ws, const FrameLayout& L, const uint32_t* vals,
                             const float* image_grad, int n, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(ws_ptr<float>(ws, L.grad_acc), 0, (size_t)(n > 0 ? n : 1) * 12 * sizeof(float), s);
  if (e != cudaSuccess) return e;
  const int rows = rp.row_end - rp.row_begin;
  if (rows <= 0 || rp.tiles_x <= 0) return cudaSuccess;
  dim3 grid(rp.tiles_x, rows);
  // measured on the headline frame: 1 pixel per lane 686 us, 2 pixels 573 us, 4 pixels 668 us (96 registers, coarse culling)
  static const int ppl = getenv("B200GS_BWD_PPL") ? atoi(getenv("B200GS_BWD_PPL")) : 2;
#define GS_BWD_ARGS rp, ws_ptr<uint2>(ws, L.ranges), vals, ws_ptr<float4>(ws, L.rec0), ws_ptr<float4>(ws, L.rec1), \
      ws_ptr<float4>(ws, L.rec2), image_grad, ws_ptr<float>(ws, L.final_T), ws_ptr<uint32_t>(ws, L.n_contrib),      \
      ws_ptr<float>(ws, L.grad_acc)
  if (ppl == 2) blend_bwd_ppl_kernel<2><<<grid, kBlendThreads / 2, 0, s>>>(GS_BWD_ARGS);
  else if (ppl == 4) blend_bwd_ppl_kernel<4><<<grid, kBlendThreads / 4, 0, s>>>(GS_BWD_ARGS);
  else blend_bwd_kernel<<<grid, kBlendThreads, 0, s>>>(GS_BWD_ARGS);
#undef GS_BWD_ARGS
  return cudaGetLastError();
}

}  // namespace gs
