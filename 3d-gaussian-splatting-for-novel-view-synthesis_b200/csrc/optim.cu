// Fused multi-tensor Adam step and gradient-norm clipping.
//
// Replaces (reference): scripts/train.py:394-401 + :538 - torch.optim.Adam over six parameter groups
// (pos, opacity_raw, f_dc, f_rest, scale_raw, q_raw; per-group learning rates, eps = 1e-15), whose
// multi-tensor implementation makes ~8 passes over the 59 floats per Gaussian - and scripts/train.py:536,
// torch.nn.utils.clip_grad_norm_(model.pos, 1.0).  Here the step of ALL tensors is one launch that reads
// (param, grad, exp_avg, exp_avg_sq) once and writes (param, exp_avg, exp_avg_sq) once: 28 B per element,
// the HBM floor of the update.  Arithmetic follows torch/optim/adam.py (_single_tensor_adam /
// _multi_tensor_adam, weight_decay = 0, amsgrad = False, maximize = False), in fp32:
//   m  <- m + (g - m) * (1 - beta1)                              (lerp_)
//   v  <- v * beta2 + (1 - beta2) * g * g                        (mul_, addcmul_)
//   p  <- p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// with the two bias corrections evaluated on the host in double, as torch does with Python floats.
//
// Roofline: HBM.
#include "common.cuh"

namespace gs {

constexpr int kAdamThreads = 256;
constexpr int kAdamVec = 4;                                   // floats per thread and iteration
constexpr int kAdamChunk = kAdamThreads * kAdamVec * 4;       // 4096 elements per block
constexpr int kAdamMaxTensors = 8;

struct AdamTensor {
  float* p; const float* g; float* m; float* v;
  long long n;
  float step_size, bc2_sqrt;       // lr / (1 - beta1^t), sqrt(1 - beta2^t)
  int chunk_begin;                 // first block of this tensor
};
struct AdamTable {
  AdamTensor t[kAdamMaxTensors];
  int n_tensors;
  float beta1, beta2, eps, one_minus_beta1, one_minus_beta2;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamTable& tb, const AdamTensor& t) {
  adam_update_f32(p, g, m, v, tb.one_minus_beta1, tb.beta2, tb.one_minus_beta2, tb.eps, t.step_size, t.bc2_sqrt);
}

__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(const __grid_constant__ AdamTable tb) {
  int ti = 0;
#pragma unroll
  for (int k = 1; k < kAdamMaxTensors; ++k)
    if (k < tb.n_tensors && (int)blockIdx.x >= tb.t[k].chunk_begin) ti = k;
  const AdamTensor& t = tb.t[ti];
  const long long base = (long long)((int)blockIdx.x - t.chunk_begin) * kAdamChunk;
  const long long end = min(base + (long long)kAdamChunk, t.n);
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                        reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15u) == 0;
  if (vec_ok && end - base == kAdamChunk) {
    float4* p4 = reinterpret_cast<float4*>(t.p + base);
    const float4* g4 = reinterpret_cast<const float4*>(t.g + base);
    float4* m4 = reinterpret_cast<float4*>(t.m + base);
    float4* v4 = reinterpret_cast<float4*>(t.v + base);
    float4 p[4], g[4], m[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = k * kAdamThreads + threadIdx.x;
      p[k] = __ldcs(p4 + i); g[k] = __ldcs(g4 + i); m[k] = __ldcs(m4 + i); v[k] = __ldcs(v4 + i);   // touched once per step
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      adam_update(p[k].x, g[k].x, m[k].x, v[k].x, tb, t);
      adam_update(p[k].y, g[k].y, m[k].y, v[k].y, tb, t);
      adam_update(p[k].z, g[k].z, m[k].z, v[k].z, tb, t);
      adam_update(p[k].w, g[k].w, m[k].w, v[k].w, tb, t);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = k * kAdamThreads + threadIdx.x;
      __stcs(p4 + i, p[k]); __stcs(m4 + i, m[k]); __stcs(v4 + i, v[k]);
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += kAdamThreads) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      adam_update(p, t.g[i], m, v, tb, t);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
    }
  }
}

cudaError_t launch_adam_step(const b200gs_adam_tensor* tensors, int n_tensors, double beta1, double beta2, double eps,
                             cudaStream_t s) {
  for (int first = 0; first < n_tensors; first += kAdamMaxTensors) {
    AdamTable tb;
    const int cnt = (n_tensors - first) < kAdamMaxTensors ? (n_tensors - first) : kAdamMaxTensors;
    tb.n_tensors = 0;
    tb.beta1 = (float)beta1; tb.beta2 = (float)beta2; tb.eps = (float)eps;
    tb.one_minus_beta1 = (float)(1.0 - beta1); tb.one_minus_beta2 = (float)(1.0 - beta2);
    long long blocks = 0;
    for (int k = 0; k < cnt; ++k) {
      const b200gs_adam_tensor& a = tensors[first + k];
      if (a.numel <= 0) continue;
      AdamTensor& t = tb.t[tb.n_tensors++];
      t.p = a.param; t.g = a.grad; t.m = a.exp_avg; t.v = a.exp_avg_sq; t.n = a.numel;
      const double bc1 = 1.0 - pow(beta1, (double)a.step), bc2 = 1.0 - pow(beta2, (double)a.step);
      t.step_size = (float)(a.lr / bc1);
      t.bc2_sqrt = (float)sqrt(bc2);
      t.chunk_begin = (int)blocks;
      blocks += (a.numel + kAdamChunk - 1) / kAdamChunk;
    }
    if (blocks == 0) continue;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    adam_step_kernel<<<(unsigned)blocks, kAdamThreads, 0, s>>>(tb);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ---- clip_grad_norm_ (torch/nn/utils/clip_grad.py): total_norm = ||(g_1, ..., g_k)||_2 over ALL the tensors of the call,
//      every g_i *= min(1, max_norm / (total_norm + 1e-6)).  The reference clips one tensor (scripts/train.py:536:
//      model.pos); several tensors go through the same two kernels with a table of pointers, like adam_step. ----
constexpr int kClipThreads = 256;
constexpr int kClipChunk = kClipThreads * 16;
constexpr int kClipMaxTensors = B200GS_CLIP_MAX_TENSORS;

struct ClipTensor {
  float* g;
  long long n;
  int block_begin;                 // first block of this tensor
};
struct ClipTable {
  ClipTensor t[kClipMaxTensors];
  int n_tensors;
};

static size_t clip_blocks(long long numel) { return (size_t)((numel + kClipChunk - 1) / kClipChunk); }

size_t clip_workspace_bytes(long long numel) { return 256 + align_up((clip_blocks(numel) + 1) * sizeof(float), 256); }

size_t clip_workspace_bytes_multi(const int64_t* numel, int n_tensors) {
  size_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) blocks += clip_blocks(numel[i] > 0 ? numel[i] : 0);
  return 256 + align_up((blocks + 1) * sizeof(float), 256);
}

__device__ __forceinline__ const ClipTensor& clip_tensor_of_block(const ClipTable& tb) {
  int ti = 0;
#pragma unroll
  for (int k = 1; k < kClipMaxTensors; ++k)
    if (k < tb.n_tensors && (int)blockIdx.x >= tb.t[k].block_begin) ti = k;
  return tb.t[ti];
}

__global__ void __launch_bounds__(kClipThreads) grad_sqnorm_kernel(const __grid_constant__ ClipTable tb,
                                                                   float* __restrict__ partials, uint32_t* ticket,
                                                                   float* __restrict__ coef, float* __restrict__ total_norm,
                                                                   float max_norm) {
  __shared__ float s_w[kClipThreads / 32];
  __shared__ double s_d[kClipThreads / 32];
  __shared__ bool s_last;
  const ClipTensor& t = clip_tensor_of_block(tb);
  const float* __restrict__ g = t.g;
  const long long n = t.n;
  const long long base = (long long)((int)blockIdx.x - t.block_begin) * kClipChunk;
  float acc = 0.f;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const long long i = base + k * kClipThreads + threadIdx.x;
    if (i < n) { const float x = g[i]; acc = fmaf(x, x, acc); }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_w[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kClipThreads / 32; ++k) t += s_w[k];
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double d = 0.0;
  for (uint32_t i = threadIdx.x; i < gridDim.x; i += kClipThreads) d += (double)__ldcg(&partials[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) s_d[warp] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < kClipThreads / 32; ++k) t += s_d[k];
    const float norm = (float)sqrt(t);
    const float c = max_norm / (norm + 1e-6f);
    *coef = c < 1.f ? c : 1.f;                         // clip_coef_clamped
    if (total_norm) *total_norm = norm;
    *ticket = 0;
  }
}

__global__ void __launch_bounds__(kClipThreads) grad_scale_kernel(const __grid_constant__ ClipTable tb,
                                                                  const float* __restrict__ coef) {
  const float c = *coef;
  if (c == 1.f) return;                                // nothing to clip: leave the gradients bit-for-bit alone
  const ClipTensor& t = clip_tensor_of_block(tb);
  float* __restrict__ g = t.g;
  const long long n = t.n;
  const long long base = (long long)((int)blockIdx.x - t.block_begin) * kClipChunk;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const long long i = base + k * kClipThreads + threadIdx.x;
    if (i < n) g[i] *= c;
  }
}

cudaError_t launch_clip_grad_norm_multi(float* const* grads, const int64_t* numel, int n_tensors, double max_norm, void* ws,
                                        float* total_norm_out, cudaStream_t s) {
  ClipTable tb;
  tb.n_tensors = 0;
  size_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (numel[i] <= 0) continue;                       // an empty tensor adds nothing to the norm
    if (tb.n_tensors == kClipMaxTensors) return cudaErrorInvalidValue;
    ClipTensor& t = tb.t[tb.n_tensors++];
    t.g = grads[i];
    t.n = numel[i];
    t.block_begin = (int)blocks;
    blocks += clip_blocks(numel[i]);
  }
  for (int k = tb.n_tensors; k < kClipMaxTensors; ++k) tb.t[k] = ClipTensor{nullptr, 0, 0x7fffffff};
  if (blocks == 0) {
    if (total_norm_out) return cudaMemsetAsync(total_norm_out, 0, sizeof(float), s);
    return cudaSuccess;
  }
  uint32_t* ticket = ws_ptr<uint32_t>(ws, 0);
  float* coef = ws_ptr<float>(ws, 64);
  float* partials = ws_ptr<float>(ws, 256);
  cudaError_t e = cudaMemsetAsync(ticket, 0, 4, s);
  if (e != cudaSuccess) return e;
  grad_sqnorm_kernel<<<(unsigned)blocks, kClipThreads, 0, s>>>(tb, partials, ticket, coef, total_norm_out, (float)max_norm);
  grad_scale_kernel<<<(unsigned)blocks, kClipThreads, 0, s>>>(tb, coef);
  return cudaGetLastError();
}

cudaError_t launch_clip_grad_norm(float* grad, long long numel, double max_norm, void* ws, float* total_norm_out,
                                  cudaStream_t s) {
  const int64_t n = numel;
  return launch_clip_grad_norm_multi(&grad, &n, 1, max_norm, ws, total_norm_out, s);
}

}  // namespace gs
