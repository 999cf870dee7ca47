// Per-block phase timeline of one onesweep pass (SORT_TRACE build of scan_sort.cu, included as one translation unit).
#define SORT_TRACE 1
#include "../../3d-gaussian-splatting-for-novel-view-synthesis_b200/csrc/scan_sort.cu"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const uint32_t n = argc > 1 ? atoi(argv[1]) : 1000000;
  const int bits = argc > 2 ? atoi(argv[2]) : 8;
  const uint32_t cap = argc > 4 ? atoi(argv[4]) : n;     // host-side bound (grid size); the real count sits on the device
  const bool skew = argc > 5 && atoi(argv[5]);
  std::mt19937 rng(7);
  std::vector<uint32_t> keys(cap);
  std::exponential_distribution<float> ex(1.f / 40.f);
  for (auto& k : keys) k = skew ? std::min(255u, (uint32_t)ex(rng)) : rng();
  uint32_t *src, *ka, *va, *kb, *vb;
  CK(cudaMalloc(&src, cap * 4)); CK(cudaMalloc(&ka, cap * 4)); CK(cudaMalloc(&va, cap * 4)); CK(cudaMalloc(&kb, cap * 4)); CK(cudaMalloc(&vb, cap * 4));
  CK(cudaMemcpy(src, keys.data(), cap * 4, cudaMemcpyHostToDevice));
  uint32_t* n_dev; CK(cudaMalloc(&n_dev, 4)); CK(cudaMemcpy(n_dev, &n, 4, cudaMemcpyHostToDevice));
  const size_t sb = gs::sort_scratch_bytes(cap);
  void* scratch; CK(cudaMalloc(&scratch, sb));
  int in_a;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int r = 0; r < 5; ++r) {
    CK(cudaMemcpy(ka, src, n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    CK(gs::launch_radix_sort(src, src, ka, va, kb, vb, cap, cap != n ? n_dev : nullptr, 0, bits, scratch, sb, &in_a, 0, false));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  // mirror of launch_radix_sort's block plan: argv[3] = threads per block (256 narrow / 512 wide), argv[6] = blocks per SM
  const int threads = argc > 3 ? atoi(argv[3]) : 256, bps = argc > 6 ? atoi(argv[6]) : (threads == 512 ? 1 : 2);
  uint32_t grid = 148 * bps; if (grid > 480) grid = 480;
  { const uint32_t mt = threads * 4; if (grid > (n + mt - 1) / mt) grid = (n + mt - 1) / mt; }
  uint32_t items = ((n + grid - 1) / grid + threads - 1) / threads;
  items = items < 4 ? 4 : (items > (threads == 512 ? 14u : 16u) ? (threads == 512 ? 14u : 16u) : items);
  const int tile = items * threads;
  const int nblk = (n + tile - 1) / tile;
  printf("plan: %d threads, %d blocks per SM -> grid %u, tile %d, %d tiles\n", threads, bps, grid, tile, nblk);
  std::vector<unsigned long long> tr(4096 * 12);
  CK(cudaMemcpyFromSymbol(tr.data(), gs::g_sort_trace, tr.size() * 8));
  int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double ghz = 1.965;   // SM clock under load on this pool (nvidia-smi); the attribute reports the boost ceiling
  printf("n=%u bits=%d blocks=%d   memset+hist+pass(es): %.1f us   (clock attr %d kHz, using %.3f GHz)\n", n, bits, nblk, ms * 1e3, clk_khz, ghz);
  const char* names[] = {"ticket->keys arrived", "ranking", "digit bookkeeping", "scatter", "look-back (thread 0)", "wait for block", "write-out"};
  unsigned long long gt0 = ~0ull;
  for (int b = 0; b < nblk && b < 4096; ++b) gt0 = std::min(gt0, tr[b * 12 + 10]);
  for (int ph = 0; ph < 7; ++ph) {
    double sum = 0, mn = 1e30, mx = 0;
    for (int b = 0; b < nblk && b < 4096; ++b) {
      const double d = (double)(tr[b * 12 + ph + 1] - tr[b * 12 + ph]) / ghz * 1e-3;
      sum += d; mn = std::min(mn, d); mx = std::max(mx, d);
    }
    printf("  %-24s avg %6.2f us  min %6.2f  max %6.2f\n", names[ph], sum / nblk, mn, mx);
  }
  double s_mx = 0, e_mx = 0;
  for (int b = 0; b < nblk && b < 4096; ++b) {
    const double st = (double)(tr[b * 12 + 10] - gt0) * 1e-3;
    const double en = st + (double)(tr[b * 12 + 7] - tr[b * 12 + 0]) / ghz * 1e-3;
    s_mx = std::max(s_mx, st); e_mx = std::max(e_mx, en);
  }
  printf("  last block start %.2f us after the first; last block end %.2f us after the first start\n", s_mx, e_mx);
  {   // the slowest blocks: which SM, how many working blocks shared it
    std::vector<int> per_sm(256, 0);
    for (int b = 0; b < nblk && b < 4096; ++b) per_sm[tr[b * 12 + 9] & 255]++;
    std::vector<std::pair<double, int>> ends;
    for (int b = 0; b < nblk && b < 4096; ++b) ends.push_back({(double)(tr[b * 12 + 10] - gt0) * 1e-3 + (double)(tr[b * 12 + 7] - tr[b * 12 + 0]) / ghz * 1e-3, b});
    std::sort(ends.rbegin(), ends.rend());
    for (int i = 0; i < 6 && i < (int)ends.size(); ++i) {
      const int b = ends[i].second;
      printf("  slow vbid %4d (cta %4llu, sm %3llu shared by %d) start +%6.2f us end %6.2f:", b, tr[b * 12 + 11], tr[b * 12 + 9], per_sm[tr[b * 12 + 9] & 255],
             (double)(tr[b * 12 + 10] - gt0) * 1e-3, ends[i].first);
      for (int ph = 0; ph < 7; ++ph) printf(" %5.2f", (double)(tr[b * 12 + ph + 1] - tr[b * 12 + ph]) / ghz * 1e-3);
      printf("\n");
    }
  }
  for (int b : {0, 1, nblk / 4, nblk / 2, nblk - 2, nblk - 1}) {
    if (b < 0 || b >= nblk) continue;
    printf("  vbid %4d (cta %4llu) start +%6.2f us:", b, tr[b * 12 + 11], (double)(tr[b * 12 + 10] - gt0) * 1e-3);
    for (int ph = 0; ph < 7; ++ph) printf(" %5.2f", (double)(tr[b * 12 + ph + 1] - tr[b * 12 + ph]) / ghz * 1e-3);
    printf("\n");
  }
  return 0;
}
