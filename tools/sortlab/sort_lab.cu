// Stand-alone timing + correctness harness for the radix sort of scan_sort.cu (no Python, no torch: a GPU call spends
// its seconds on the kernels).  Build: tools/sortlab/build.sh; run: ./sort_lab_mb3 [n] [reps]
//   keys: depth keys of a frame (float bits of z in (near, far), ~10 % culled = 0xFFFFFFFF), or `bits`-bit random keys
//   env: B200GS_SORT_V1=1 (first form of the pass kernel), B200GS_SORT_DIGIT_BITS=8|9
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "../../3d-gaussian-splatting-for-novel-view-synthesis_b200/csrc/common.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static uint32_t fbits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }

struct Case { const char* name; uint32_t n; int bits; bool depth; uint32_t sub, maxk; bool with_vals; };

static void run(const Case& c, int reps) {
  std::mt19937 rng(1234);
  std::vector<uint32_t> keys(c.n), vals(c.n);
  if (c.depth) {
    std::lognormal_distribution<float> dz(1.6f, 0.45f);
    std::uniform_real_distribution<float> u01(0.f, 1.f);
    for (uint32_t i = 0; i < c.n; ++i) {
      float z = std::min(std::max(dz(rng), 0.1001f), 99.9f);
      keys[i] = u01(rng) < 0.10f ? 0xFFFFFFFFu : fbits(z);
    }
  } else {
    for (uint32_t i = 0; i < c.n; ++i) keys[i] = c.bits >= 32 ? rng() : (rng() & ((1u << c.bits) - 1u));
  }
  for (uint32_t i = 0; i < c.n; ++i) vals[i] = c.with_vals ? (rng() % c.n) : i;
  // reference: stable sort by transformed key
  auto tk = [&](uint32_t k) { uint32_t t = k - c.sub; return t < c.maxk ? t : c.maxk; };
  std::vector<uint32_t> perm(c.n);
  std::iota(perm.begin(), perm.end(), 0u);
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return tk(keys[a]) < tk(keys[b]); });

  uint32_t *d_src, *d_vsrc, *ka, *va, *kb, *vb;
  CK(cudaMalloc(&d_src, c.n * 4)); CK(cudaMalloc(&d_vsrc, c.n * 4));
  CK(cudaMalloc(&ka, c.n * 4)); CK(cudaMalloc(&va, c.n * 4)); CK(cudaMalloc(&kb, c.n * 4)); CK(cudaMalloc(&vb, c.n * 4));
  CK(cudaMemcpy(d_src, keys.data(), c.n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_vsrc, vals.data(), c.n * 4, cudaMemcpyHostToDevice));
  const size_t sb = gs::sort_scratch_bytes(c.n);
  void* scratch; CK(cudaMalloc(&scratch, sb));
  void* flush; const size_t fl = 256u << 20; CK(cudaMalloc(&flush, fl));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> hot, cold;
  int in_a = 0;
  for (int r = 0; r < reps + 3; ++r) {
    const bool flushed = (r & 1);
    if (flushed) CK(cudaMemsetAsync(flush, r, fl, s));
    else CK(cudaMemcpyAsync(ka, d_src, c.n * 4, cudaMemcpyDeviceToDevice, s));   // touches the keys: L2-warm, as after preprocess
    CK(cudaEventRecord(e0, s));
    CK(gs::launch_radix_sort(d_src, c.with_vals ? d_vsrc : nullptr, ka, va, kb, vb, c.n, nullptr, 0, c.bits, scratch, sb, &in_a, s,
                             false, c.sub, c.maxk));
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 3) (flushed ? cold : hot).push_back(ms * 1e3f);
  }
  std::vector<uint32_t> ok(c.n), ov(c.n);
  CK(cudaMemcpy(ok.data(), in_a ? ka : kb, c.n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ov.data(), in_a ? va : vb, c.n * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (uint32_t i = 0; i < c.n; ++i) {
    const uint32_t want_v = c.with_vals ? vals[perm[i]] : perm[i];
    if (ov[i] != want_v || ok[i] != tk(keys[perm[i]])) { if (bad < 3) printf("  mismatch at %u: got (%u,%u) want (%u,%u)\n", i, ok[i], ov[i], tk(keys[perm[i]]), want_v); ++bad; }
  }
  std::sort(hot.begin(), hot.end()); std::sort(cold.begin(), cold.end());
  const gs::SortPasses sp = gs::sort_passes(0, c.bits);
  printf("%-28s n=%8u bits=%2d passes=%d x %d-bit  hot median %7.1f us (min %7.1f)  cold median %7.1f us  %s\n", c.name, c.n, c.bits,
         sp.num, sp.digit_bits, hot[hot.size() / 2], hot[0], cold[cold.size() / 2], bad ? "WRONG" : "ok");
  if (bad) printf("  %zu mismatches\n", bad);
  cudaFree(d_src); cudaFree(d_vsrc); cudaFree(ka); cudaFree(va); cudaFree(kb); cudaFree(vb); cudaFree(scratch); cudaFree(flush);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 20;
  const uint32_t base = fbits(0.1f), span = fbits(100.f) - base;
  int kb = 1; while (((1u << kb) - 1u) < span) ++kb;
  const uint32_t maxk = (1u << kb) - 1u;
  const Case cases[] = {
      {"depth 1M raw 32-bit", 1000000, 32, true, 0, 0xFFFFFFFFu, false},
      {"depth 1M offset keys", 1000000, kb, true, base, maxk, false},
      {"depth 850k offset keys", 850000, kb, true, base, maxk, true},
      {"depth 3M offset keys", 3000000, kb, true, base, maxk, false},
      {"depth 6M offset keys", 6000000, kb, true, base, maxk, false},
      {"supertile 1.2M 8-bit", 1200000, 8, false, 0, 0xFFFFFFFFu, true},
      {"supertile 8M 10-bit", 8000000, 10, false, 0, 0xFFFFFFFFu, true},
      {"random 1M 32-bit", 1000000, 32, false, 0, 0xFFFFFFFFu, true},
      {"tail 4097 13-bit", 4097, 13, false, 0, 0xFFFFFFFFu, true},
      {"tail 100001 27-bit", 100001, 27, false, 0, 0xFFFFFFFFu, true},
  };
  for (const Case& c : cases) run(c, reps);
  return 0;
}
