// Fused L1 + SSIM training loss, forward and backward.
//
// Replaces (reference): gaussian_splatting/losses.py:27-185 - F.l1_loss, fifteen 11x11 conv2d calls
// (five per channel) plus ~40 elementwise kernels and three .item() syncs per view - and its autograd.
// The 2-D Gaussian window is the outer product of a 1-D one (losses.py:148-154), so the five windowed
// means are evaluated separably (11 + 11 taps instead of 121) on a 32x32 output tile with a 5-pixel halo
// staged in shared memory; zero padding as conv2d(padding=5) does.  Images are [B,H,W,3] (the render
// output layout, no permute).
//
// Forward, per pixel and channel (x = pred, y = target; E[.] = windowed mean):
//   mu1 = E[x], mu2 = E[y], s1 = E[xx]-mu1^2, s2 = E[yy]-mu2^2, s12 = E[xy]-mu1 mu2
//   S = (2 mu1 mu2 + C1)(2 s12 + C2) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2))
//   l1 = mean|x-y|, ssim = 1 - mean S, total = lambda_l1 l1 + lambda_ssim ssim
// and, when a backward will follow, the three partials of S the backward convolves again:
//   A = dS/dE[x], B = dS/dE[xx], C = dS/dE[xy]       (planar [B,3,H,W] maps in the workspace)
// Backward:  dtotal/dx(q) = -lambda_ssim/(3P) * (G*A + 2 x G*B + y G*C)(q) + lambda_l1/(3P) * sign(x-y)(q).
//
// Roofline: HBM (reads pred/target once, writes/reads three maps), ~130 FMA per pixel-channel per pass.
#include "common.cuh"

namespace gs {

constexpr int kLossTile = 32;
constexpr int kLossR = 5;
constexpr int kLossIn = kLossTile + 2 * kLossR;   // 42
constexpr int kLossThreads = 256;
constexpr float kSsimC1 = 0.01f * 0.01f, kSsimC2 = 0.03f * 0.03f;

// g = exp(-c^2 / (2 * 1.5^2)) / sum for c = -5..5, the fp32 values torch computes (losses.py:148-151;
// printed by oracle/make_golden_loss.py)
__constant__ float c_win[11] = {0.0010283803567290306f, 0.0075987582094967365f, 0.036000773310661316f,
                                0.10936068743467331f,   0.21300552785396576f,   0.26601171493530273f,
                                0.21300552785396576f,   0.10936068743467331f,   0.036000773310661316f,
                                0.0075987582094967365f, 0.0010283803567290306f};

struct LossLayout {
  size_t ticket, out_acc, partials, maps, total;
  size_t n_blocks, map_floats;
};

static LossLayout loss_layout(int n_img, int H, int W, bool with_grad) {
  LossLayout L;
  L.n_blocks = (size_t)ceil_div(W, kLossTile) * ceil_div(H, kLossTile) * (size_t)(n_img > 0 ? n_img : 1);
  L.map_floats = (size_t)(n_img > 0 ? n_img : 1) * 3 * (size_t)H * W;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.ticket = take(256);
  L.out_acc = take(256);
  L.partials = take(L.n_blocks * sizeof(float2));
  L.maps = take(with_grad ? 3 * L.map_floats * sizeof(float) : 0);
  L.total = off;
  return L;
}

size_t loss_workspace_bytes(int n_img, int H, int W, bool with_grad) { return loss_layout(n_img, H, W, with_grad).total; }

// Packed FP32 (Blackwell FFMA2 / FMUL2): one instruction, two IEEE operations on a register pair - the same results as
// two scalar fmaf / multiplies, half the issue slots.  The separable filter applies one tap weight to five quantities
// at once, so the pairs (x, y) and (xx, yy) share their (w, w) operand.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Loads the (kLossIn x kLossIn) halo tile of one channel of an interleaved [H,W,3] image, zero outside.
// (One element per thread and step with a division by 42: a row-per-warp mapping without the division was measured and
// is 11 % slower - 10 of its 32 lanes load the second half of a 42-wide row; the kernel is bound by load / shared-memory
// instructions, not by arithmetic.)
__device__ __forceinline__ void load_tile_hwc(const float* __restrict__ img, int H, int W, int c, int x0, int y0,
                                              float (*s)[kLossIn + 1]) {
  for (int i = threadIdx.x; i < kLossIn * kLossIn; i += kLossThreads) {
    const int r = i / kLossIn, q = i - r * kLossIn;
    const int y = y0 + r - kLossR, x = x0 + q - kLossR;
    s[r][q] = (y >= 0 && y < H && x >= 0 && x < W) ? img[((size_t)y * W + x) * 3 + c] : 0.f;
  }
}
__device__ __forceinline__ void load_tile_planar(const float* __restrict__ map, int H, int W, int x0, int y0,
                                                 float (*s)[kLossIn + 1]) {
  for (int i = threadIdx.x; i < kLossIn * kLossIn; i += kLossThreads) {
    const int r = i / kLossIn, q = i - r * kLossIn;
    const int y = y0 + r - kLossR, x = x0 + q - kLossR;
    s[r][q] = (y >= 0 && y < H && x >= 0 && x < W) ? map[(size_t)y * W + x] : 0.f;
  }
}

__global__ void __launch_bounds__(kLossThreads) l1_ssim_fwd_kernel(const float* __restrict__ pred,
                                                                   const float* __restrict__ target, int H, int W,
                                                                   float* __restrict__ maps, size_t map_floats,
                                                                   float2* __restrict__ partials, uint32_t* ticket,
                                                                   float* __restrict__ out, float lambda_l1,
                                                                   float lambda_ssim, double inv_count) {
  __shared__ float sx[kLossIn][kLossIn + 1];
  __shared__ float sy[kLossIn][kLossIn + 1];
  __shared__ float sh[5][kLossIn][kLossTile + 1];  // horizontally filtered x, y, xx, yy, xy (row stride 33 words:
                                                   // conflict-free for row-per-lane writes and column-per-lane reads)
  __shared__ float s_red[2][kLossThreads / 32];
  __shared__ bool s_last;
  const int x0 = blockIdx.x * kLossTile, y0 = blockIdx.y * kLossTile, b = blockIdx.z;
  const size_t img_off = (size_t)b * H * W * 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float w[11];
  f32x2 w2[11];
#pragma unroll
  for (int t = 0; t < 11; ++t) { w[t] = c_win[t]; w2[t] = pack2(w[t], w[t]); }
  float sum_l1 = 0.f, sum_s = 0.f;
  for (int c = 0; c < 3; ++c) {
    __syncthreads();     // previous channel's vertical pass is done with sh, sx, sy
    load_tile_hwc(pred + img_off, H, W, c, x0, y0, sx);
    load_tile_hwc(target + img_off, H, W, c, x0, y0, sy);
    __syncthreads();
    // horizontal pass: one work item = 4 consecutive outputs of one row, computed from a sliding window of 14
    // inputs (each input is loaded once and feeds the up to 4 outputs it belongs to); lanes take consecutive rows
    for (int item = tid; item < kLossIn * (kLossTile / 4); item += kLossThreads) {
      const int r = item % kLossIn, q0 = (item / kLossIn) * 4;
      f32x2 a01[4], a23[4];       // (E[x], E[y]) and (E[xx], E[yy]) as register pairs
      float a4[4];                // E[xy]
#pragma unroll
      for (int j = 0; j < 4; ++j) { a01[j] = 0ull; a23[j] = 0ull; a4[j] = 0.f; }
#pragma unroll
      for (int t = 0; t < 14; ++t) {
        const float x = sx[r][q0 + t], y = sy[r][q0 + t];
        const f32x2 xy2 = pack2(x, y), sq2 = mul2(xy2, xy2);
        const float xy = x * y;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = t - j;
          if (k >= 0 && k < 11) {
            a01[j] = fma2(w2[k], xy2, a01[j]);
            a23[j] = fma2(w2[k], sq2, a23[j]);
            a4[j] = fmaf(w[k], xy, a4[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e0, e1, e2, e3;
        unpack2(a01[j], e0, e1);
        unpack2(a23[j], e2, e3);
        sh[0][r][q0 + j] = e0; sh[1][r][q0 + j] = e1; sh[2][r][q0 + j] = e2; sh[3][r][q0 + j] = e3;
        sh[4][r][q0 + j] = a4[j];
      }
    }
    __syncthreads();
    // vertical pass: a thread owns 4 vertically adjacent pixels of one column, same sliding window
    f32x2 v01[4], v23[4];
    float v4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { v01[j] = 0ull; v23[j] = 0ull; v4[j] = 0.f; }
#pragma unroll
    for (int t = 0; t < 14; ++t) {
      const f32x2 p01 = pack2(sh[0][4 * warp + t][lane], sh[1][4 * warp + t][lane]);
      const f32x2 p23 = pack2(sh[2][4 * warp + t][lane], sh[3][4 * warp + t][lane]);
      const float p4 = sh[4][4 * warp + t][lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = t - j;
        if (k >= 0 && k < 11) {
          v01[j] = fma2(w2[k], p01, v01[j]);
          v23[j] = fma2(w2[k], p23, v23[j]);
          v4[j] = fmaf(w[k], p4, v4[j]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int py = 4 * warp + k, px = lane;
      const int y = y0 + py, x = x0 + px;
      float mu1, mu2, exx, eyy;
      unpack2(v01[k], mu1, mu2);
      unpack2(v23[k], exx, eyy);
      const float exy = v4[k];
      if (y < H && x < W) {
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = exx - mu1_sq, s2 = eyy - mu2_sq, s12 = exy - mu12;
        const float n1 = 2.f * mu12 + kSsimC1, n2 = 2.f * s12 + kSsimC2;
        const float d1 = mu1_sq + mu2_sq + kSsimC1, d2 = s1 + s2 + kSsimC2;
        const float inv = 1.f / (d1 * d2);
        const float S = n1 * n2 * inv;
        sum_s += S;
        sum_l1 += fabsf(sx[py + kLossR][px + kLossR] - sy[py + kLossR][px + kLossR]);
        if (maps) {
          const size_t o = ((size_t)(b * 3 + c) * H + y) * W + x;
          maps[o] = 2.f * mu2 * (n2 - n1) * inv - 2.f * mu1 * S / d1 + 2.f * mu1 * S / d2;   // dS/dE[x]
          maps[map_floats + o] = -S / d2;                                                     // dS/dE[xx]
          maps[2 * map_floats + o] = 2.f * n1 * inv;                                          // dS/dE[xy]
        }
      }
    }
  }
  // block sums -> partials; the last block to finish adds all partials up in double and writes the result
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum_l1 += __shfl_xor_sync(0xffffffffu, sum_l1, o);
    sum_s += __shfl_xor_sync(0xffffffffu, sum_s, o);
  }
  if (lane == 0) { s_red[0][warp] = sum_l1; s_red[1][warp] = sum_s; }
  __syncthreads();
  const uint32_t n_blocks = gridDim.x * gridDim.y * gridDim.z;
  const uint32_t bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (tid == 0) {
    float a = 0.f, s = 0.f;
#pragma unroll
    for (int k = 0; k < kLossThreads / 32; ++k) { a += s_red[0][k]; s += s_red[1][k]; }
    partials[bid] = make_float2(a, s);
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == n_blocks - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double a = 0.0, s = 0.0;
  for (uint32_t i = tid; i < n_blocks; i += kLossThreads) {
    const float2 p = __ldcg(&partials[i]);      // written by other blocks: read through L2
    a += p.x; s += p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  __shared__ double s_d[2][kLossThreads / 32];
  if (lane == 0) { s_d[0][warp] = a; s_d[1][warp] = s; }
  __syncthreads();
  if (tid == 0) {
    double ta = 0.0, ts = 0.0;
    for (int k = 0; k < kLossThreads / 32; ++k) { ta += s_d[0][k]; ts += s_d[1][k]; }
    const float l1 = (float)(ta * inv_count);
    const float ssim = 1.f - (float)(ts * inv_count);
    out[0] = l1;
    out[1] = ssim;
    out[2] = lambda_l1 * l1 + lambda_ssim * ssim;     // losses.py:177
    *ticket = 0;
  }
}

__global__ void __launch_bounds__(kLossThreads) l1_ssim_bwd_kernel(const float* __restrict__ pred,
                                                                   const float* __restrict__ target, int H, int W,
                                                                   const float* __restrict__ maps, size_t map_floats,
                                                                   float k_l1, float k_ssim,
                                                                   const float* __restrict__ grad_total,
                                                                   float* __restrict__ grad_pred) {
  __shared__ float sa[3][kLossIn][kLossIn + 1];
  __shared__ float sh[3][kLossIn][kLossTile + 1];
  const int x0 = blockIdx.x * kLossTile, y0 = blockIdx.y * kLossTile, b = blockIdx.z;
  const size_t img_off = (size_t)b * H * W * 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float g = grad_total ? *grad_total : 1.f;
  const float ks = k_ssim * g, k1 = k_l1 * g;
  float w[11];
  f32x2 w2[11];
#pragma unroll
  for (int t = 0; t < 11; ++t) { w[t] = c_win[t]; w2[t] = pack2(w[t], w[t]); }
  for (int c = 0; c < 3; ++c) {
    __syncthreads();
    const size_t plane = (size_t)(b * 3 + c) * H * W;
#pragma unroll
    for (int m = 0; m < 3; ++m) load_tile_planar(maps + m * map_floats + plane, H, W, x0, y0, sa[m]);
    __syncthreads();
    for (int item = tid; item < kLossIn * (kLossTile / 4); item += kLossThreads) {
      const int r = item % kLossIn, q0 = (item / kLossIn) * 4;
      f32x2 a01[4];
      float a2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { a01[j] = 0ull; a2[j] = 0.f; }
#pragma unroll
      for (int t = 0; t < 14; ++t) {
        const f32x2 p01 = pack2(sa[0][r][q0 + t], sa[1][r][q0 + t]);
        const float p2 = sa[2][r][q0 + t];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = t - j;
          if (k >= 0 && k < 11) {
            a01[j] = fma2(w2[k], p01, a01[j]);
            a2[j] = fmaf(w[k], p2, a2[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e0, e1;
        unpack2(a01[j], e0, e1);
        sh[0][r][q0 + j] = e0; sh[1][r][q0 + j] = e1; sh[2][r][q0 + j] = a2[j];
      }
    }
    __syncthreads();
    f32x2 v01[4];
    float v2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { v01[j] = 0ull; v2[j] = 0.f; }
#pragma unroll
    for (int t = 0; t < 14; ++t) {
      const f32x2 p01 = pack2(sh[0][4 * warp + t][lane], sh[1][4 * warp + t][lane]);
      const float p2 = sh[2][4 * warp + t][lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = t - j;
        if (k >= 0 && k < 11) {
          v01[j] = fma2(w2[k], p01, v01[j]);
          v2[j] = fmaf(w[k], p2, v2[j]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int py = 4 * warp + k, px = lane;
      const int y = y0 + py, x = x0 + px;
      float gA, gB;
      unpack2(v01[k], gA, gB);
      const float gC = v2[k];
      if (y < H && x < W) {
        const size_t o = img_off + ((size_t)y * W + x) * 3 + c;
        const float xv = pred[o], yv = target[o];
        const float d = xv - yv;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        grad_pred[o] = ks * (gA + 2.f * xv * gB + yv * gC) + k1 * sgn;
      }
    }
  }
}

cudaError_t launch_l1_ssim_fwd(const float* pred, const float* target, int n_img, int H, int W, float lambda_l1,
                               float lambda_ssim, void* ws, bool with_grad, float* out3, cudaStream_t s) {
  const LossLayout L = loss_layout(n_img, H, W, with_grad);
  uint32_t* ticket = ws_ptr<uint32_t>(ws, L.ticket);
  cudaError_t e = cudaMemsetAsync(ticket, 0, 4, s);
  if (e != cudaSuccess) return e;
  dim3 grid(ceil_div(W, kLossTile), ceil_div(H, kLossTile), n_img);
  const double inv_count = 1.0 / ((double)n_img * H * W * 3.0);
  l1_ssim_fwd_kernel<<<grid, kLossThreads, 0, s>>>(pred, target, H, W, with_grad ? ws_ptr<float>(ws, L.maps) : nullptr,
                                                  L.map_floats, ws_ptr<float2>(ws, L.partials), ticket, out3,
                                                  lambda_l1, lambda_ssim, inv_count);
  return cudaGetLastError();
}

cudaError_t launch_l1_ssim_bwd(const float* pred, const float* target, int n_img, int H, int W, float lambda_l1,
                               float lambda_ssim, const void* ws, const float* grad_total, float* grad_pred,
                               cudaStream_t s) {
  const LossLayout L = loss_layout(n_img, H, W, true);
  dim3 grid(ceil_div(W, kLossTile), ceil_div(H, kLossTile), n_img);
  const double count = (double)n_img * H * W * 3.0;
  l1_ssim_bwd_kernel<<<grid, kLossThreads, 0, s>>>(pred, target, H, W, ws_ptr<float>(ws, L.maps), L.map_floats,
                                                  (float)(lambda_l1 / count), (float)(-lambda_ssim / count), grad_total,
                                                  grad_pred);
  return cudaGetLastError();
}

}  // namespace gs
