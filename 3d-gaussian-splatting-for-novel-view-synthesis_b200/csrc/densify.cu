// Adaptive density control: prune / split / clone as stream compaction.
//
// Replaces (reference): scripts/train.py:89-195, GaussianModel.densify_and_prune / _prune_points / _split_points /
// _clone_points - about thirty boolean-mask gathers and six torch.cat over the parameter tensors, with several
// .any() host syncs - called every 100 iterations (train.py:544-557).  Semantics kept exactly:
//   keep  = not (sigmoid(opacity_raw) < opacity_threshold)                       (prune)
//   on the kept rows, with g = ||pos.grad||_2 and smax = max(exp(scale_raw)):
//   split = smax >  scale_threshold and g > max_grad     -> the row STAYS and a copy is appended with
//                                                           pos + (noise * exp(scale_raw)) * 0.1,  scale_raw - 0.5
//   clone = smax <= scale_threshold and g > max_grad     -> the row stays and an identical copy is appended
//   output rows: kept rows in their order, then the split copies in order, then the clones in order.
// (max_screen_size is accepted and unused, as in the reference.)
//
// Plan: one flag kernel, three exclusive scans (the single-pass scan of scan_sort.cu) -> offsets and the three counts;
// apply: one kernel per tensor that reads every source element once and writes it to its (up to three) destinations.
//
// Roofline: HBM (reads 59 N floats once, writes the new arrays once).
#include "common.cuh"

namespace gs {

constexpr int kDensifyThreads = 256;

__global__ void __launch_bounds__(kDensifyThreads) densify_flags_kernel(int n, const float* __restrict__ opacity_raw,
                                                                        const float* __restrict__ scale_raw,
                                                                        const float* __restrict__ pos_grad, float thr_op,
                                                                        float max_grad, float thr_scale,
                                                                        uint32_t* __restrict__ keep,
                                                                        uint32_t* __restrict__ split,
                                                                        uint32_t* __restrict__ clone) {
  const int i = blockIdx.x * kDensifyThreads + threadIdx.x;
  if (i >= n) return;
  const float op = 1.0f / (1.0f + expf(-opacity_raw[i]));                      // torch.sigmoid
  const bool k = !(op < thr_op);
  bool sp = false, cl = false;
  if (k && pos_grad) {
    const float gx = pos_grad[3 * i], gy = pos_grad[3 * i + 1], gz = pos_grad[3 * i + 2];
    const float g = sqrtf(gx * gx + gy * gy + gz * gz);                        // grads['pos'].norm(dim=-1)
    const float smax = fmaxf(expf(scale_raw[3 * i]), fmaxf(expf(scale_raw[3 * i + 1]), expf(scale_raw[3 * i + 2])));
    const bool hot = g > max_grad;
    sp = hot && (smax > thr_scale);
    cl = hot && (smax <= thr_scale);
  }
  keep[i] = k ? 1u : 0u;
  split[i] = sp ? 1u : 0u;
  clone[i] = cl ? 1u : 0u;
}

// MODE 0: plain copy, 1: pos (the split copy is displaced), 2: scale_raw (the split copy shrinks by 0.5)
template <int MODE>
__global__ void __launch_bounds__(kDensifyThreads) densify_apply_kernel(long long total, int w, const float* __restrict__ in,
                                                                        float* __restrict__ out,
                                                                        const uint32_t* __restrict__ keep,
                                                                        const uint32_t* __restrict__ split,
                                                                        const uint32_t* __restrict__ clone,
                                                                        const uint32_t* __restrict__ keep_off,
                                                                        const uint32_t* __restrict__ split_off,
                                                                        const uint32_t* __restrict__ clone_off,
                                                                        const uint32_t* __restrict__ counts,
                                                                        const float* __restrict__ scale_raw,
                                                                        const float* __restrict__ noise) {
  const long long e = (long long)blockIdx.x * kDensifyThreads + threadIdx.x;
  if (e >= total) return;
  const int i = (int)(e / w), c = (int)(e - (long long)i * w);
  if (!keep[i]) return;
  const float v = in[e];
  const uint32_t n_keep = counts[0], n_split = counts[1];
  out[(long long)keep_off[i] * w + c] = v;
  if (split[i]) {
    float nv = v;
    const uint32_t k = split_off[i];
    if (MODE == 1) nv = v + (noise[3 * (long long)k + c] * expf(scale_raw[3 * (long long)i + c])) * 0.1f;
    if (MODE == 2) nv = v - 0.5f;
    out[((long long)n_keep + k) * w + c] = nv;
  }
  if (clone[i]) out[((long long)n_keep + n_split + clone_off[i]) * w + c] = v;
}

size_t densify_workspace_bytes(int n) {
  const size_t a = align_up((size_t)(n > 0 ? n : 1) * sizeof(uint32_t), 256);
  return 256 + 6 * a + align_up(scan_scratch_bytes((uint32_t)(n > 0 ? n : 1)), 256);
}

struct DensifyLayout { uint32_t *counts, *keep, *split, *clone, *keep_off, *split_off, *clone_off; void* scratch; size_t scratch_bytes; };

static DensifyLayout densify_layout(void* ws, int n) {
  const size_t a = align_up((size_t)(n > 0 ? n : 1) * sizeof(uint32_t), 256);
  char* p = static_cast<char*>(ws);
  DensifyLayout L;
  L.counts = reinterpret_cast<uint32_t*>(p); p += 256;
  L.keep = reinterpret_cast<uint32_t*>(p); p += a;
  L.split = reinterpret_cast<uint32_t*>(p); p += a;
  L.clone = reinterpret_cast<uint32_t*>(p); p += a;
  L.keep_off = reinterpret_cast<uint32_t*>(p); p += a;
  L.split_off = reinterpret_cast<uint32_t*>(p); p += a;
  L.clone_off = reinterpret_cast<uint32_t*>(p); p += a;
  L.scratch = p;
  L.scratch_bytes = align_up(scan_scratch_bytes((uint32_t)(n > 0 ? n : 1)), 256);
  return L;
}

cudaError_t launch_densify_plan(int n, const float* opacity_raw, const float* scale_raw, const float* pos_grad,
                                double thr_op, double max_grad, double thr_scale, void* ws, uint32_t* counts_host,
                                cudaStream_t s) {
  const DensifyLayout L = densify_layout(ws, n);
  cudaError_t e = cudaMemsetAsync(L.counts, 0, 256, s);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    densify_flags_kernel<<<ceil_div(n, kDensifyThreads), kDensifyThreads, 0, s>>>(
        n, opacity_raw, scale_raw, pos_grad, (float)thr_op, (float)max_grad, (float)thr_scale, L.keep, L.split, L.clone);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const uint32_t* flags[3] = {L.keep, L.split, L.clone};
    uint32_t* offs[3] = {L.keep_off, L.split_off, L.clone_off};
    for (int k = 0; k < 3; ++k)
      if ((e = launch_exclusive_scan(flags[k], nullptr, offs[k], (uint32_t)n, L.counts + k, L.scratch, L.scratch_bytes, s)) !=
          cudaSuccess)
        return e;
  }
  if (counts_host) e = cudaMemcpyAsync(counts_host, L.counts, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
  return e;
}

cudaError_t launch_densify_apply(int n, const void* ws, const float* const in[6], float* const out[6], const float* noise,
                                 cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const DensifyLayout L = densify_layout(const_cast<void*>(ws), n);
  // order of the six tensors: pos[3] opacity_raw[1] f_dc[3] f_rest[45] scale_raw[3] q_raw[4]
  static const int widths[6] = {3, 1, 3, 45, 3, 4};
  const float* scale_raw = in[4];
  for (int t = 0; t < 6; ++t) {
    const long long total = (long long)n * widths[t];
    const unsigned blocks = (unsigned)((total + kDensifyThreads - 1) / kDensifyThreads);
    if (t == 0)
      densify_apply_kernel<1><<<blocks, kDensifyThreads, 0, s>>>(total, widths[t], in[t], out[t], L.keep, L.split, L.clone,
                                                                 L.keep_off, L.split_off, L.clone_off, L.counts, scale_raw, noise);
    else if (t == 4)
      densify_apply_kernel<2><<<blocks, kDensifyThreads, 0, s>>>(total, widths[t], in[t], out[t], L.keep, L.split, L.clone,
                                                                 L.keep_off, L.split_off, L.clone_off, L.counts, scale_raw, noise);
    else
      densify_apply_kernel<0><<<blocks, kDensifyThreads, 0, s>>>(total, widths[t], in[t], out[t], L.keep, L.split, L.clone,
                                                                 L.keep_off, L.split_off, L.clone_off, L.counts, scale_raw, noise);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace gs
