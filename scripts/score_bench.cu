// Stand-alone timing + self-check of the two scoring kernels of a rank event (merge_score_rows_kernel, merge_score_kernel)
// on controlled mixes of child pairs -- what scripts/hbm_kernels.py is for the streaming kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/score_bench scripts/score_bench.cu \
//        -Lphylo_b200 -lvcsmc_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../phylo_b200'
//   scripts/score_bench [reps [scenario]]
// Per scenario: `count` particles whose children are (leaf, internal node) [rows kernel] or two internal nodes [generic
// kernel], drawn uniformly from n_a x n_b child pairs and listed in grouped order, as the event kernel leaves them.
// Reference values: the generic kernel in identity order (order == null: its per-lane leaf path), which the GPU tests pin
// against the oracle.  Prints time per launch, FP64-pipe operations / s and the fraction of 18.2 T op/s.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../phylo_b200/csrc/launch.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 20;
  const int only = argc > 2 ? atoi(argv[2]) : -1;   // run one scenario only (for ncu)
  const int N = 64, S = 10000, n_slots = 24;
  const int64_t K = 65536;
  std::mt19937_64 rng(7);
  std::vector<uint8_t> codes((size_t)N * S);
  for (auto& c : codes) c = (uint8_t)(1u << (rng() & 3));
  for (int i = 0; i < 40; ++i) codes[rng() % codes.size()] = 15;                // a few gaps
  for (int i = 0; i < 6; ++i) codes[(size_t)3 * S + rng() % S] = 5;            // ambiguity codes on one leaf
  std::vector<double> pool((size_t)n_slots * S * 4);
  std::uniform_real_distribution<double> U(0.05, 1.0);
  for (auto& x : pool) x = U(rng);
  std::vector<double> P((size_t)K * 32), pi = {0.1, 0.2, 0.3, 0.4};
  for (int64_t k = 0; k < K * 8; ++k) {
    double r[4], s = 0;
    for (int j = 0; j < 4; ++j) { r[j] = U(rng); s += r[j]; }
    for (int j = 0; j < 4; ++j) P[k * 4 + j] = r[j] / s;
  }
  uint8_t* d_codes; double *d_pool, *d_P, *d_pi, *d_ell, *d_ref;
  int32_t *d_l, *d_r, *d_order, *d_count, *d_perm; uint8_t* d_tstate;
  const int Sp = vcsmc::leaf_sort_stride(S);
  CK(cudaMalloc(&d_codes, codes.size())); CK(cudaMalloc(&d_pool, pool.size() * 8)); CK(cudaMalloc(&d_P, P.size() * 8));
  CK(cudaMalloc(&d_pi, 32)); CK(cudaMalloc(&d_ell, K * 16 * 8)); CK(cudaMalloc(&d_ref, K * 16 * 8));
  CK(cudaMalloc(&d_l, K * 4)); CK(cudaMalloc(&d_r, K * 4)); CK(cudaMalloc(&d_order, K * 4)); CK(cudaMalloc(&d_count, 16));
  CK(cudaMalloc(&d_perm, (size_t)N * Sp * 4)); CK(cudaMalloc(&d_tstate, (size_t)N * (Sp / 128) + 16));
  CK(cudaMemcpy(d_codes, codes.data(), codes.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pool, pool.data(), pool.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_P, P.data(), P.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pi, pi.data(), 32, cudaMemcpyHostToDevice));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  if (vcsmc::launch_leaf_sort(d_codes, S, N, S, d_perm, d_tstate, st)) { printf("leaf_sort failed\n"); return 1; }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  struct Scn { const char* name; int leaf; int n_a, n_b; int64_t count; };
  const Scn scn[] = {
    {"rows late-heavy  6x6 pairs", 1, 6, 6, 32768},   {"rows mid       30x15 pairs", 1, 30, 15, 16384},
    {"rows early      56x4 pairs", 1, 56, 4, 8192},   {"rows light      3x8 pairs ", 1, 3, 8, 6144},
    {"rows all        8x8 pairs ", 1, 8, 8, 65536},   {"rows tiny      20x10 pairs", 1, 20, 10, 1024},
    {"generic  10 nodes (45 prs)", 0, 10, 10, 8192},  {"generic  20 nodes (190 p) ", 0, 20, 20, 32768},
    {"generic   6 nodes (15 prs)", 0, 6, 6, 2048},
  };
  double worst = 0;
  int scn_idx = -1;
  for (const Scn& sc : scn) {
    if (++scn_idx != only && only >= 0) continue;
    // children of every particle; the scored ones first / last in `order`, grouped by pair
    std::vector<int32_t> l(K), r(K), order(K);
    std::vector<std::pair<int64_t, int32_t>> key(K);
    for (int64_t k = 0; k < K; ++k) {
      int a, b;
      if (sc.leaf) { a = -(int)(rng() % sc.n_a) - 1; b = (int)(rng() % sc.n_b); }
      else { a = (int)(rng() % sc.n_a); do b = (int)(rng() % sc.n_b); while (b == a); }
      if (rng() & 1) std::swap(a, b);   // either side may hold the leaf
      l[k] = a; r[k] = b;
      const int lo = std::min(a, b), hi = std::max(a, b);
      key[k] = {((int64_t)(lo + 1000) << 20) | (hi + 1000), (int32_t)k};
    }
    // the first `count` particles (by index) are the scored list; sort them by pair
    std::sort(key.begin(), key.begin() + sc.count);
    for (int64_t k = 0; k < K; ++k) order[k] = key[k].second;
    if (!sc.leaf) std::reverse(order.begin(), order.end());   // the generic list sits at the END of order
    int32_t cnt[2] = {sc.leaf ? (int32_t)sc.count : 0, sc.leaf ? 0 : (int32_t)sc.count};
    CK(cudaMemcpy(d_l, l.data(), K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_r, r.data(), K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_order, order.data(), K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_count, cnt, 8, cudaMemcpyHostToDevice));
    int parts_ref = 0, parts = 0;
    CK(cudaMemsetAsync(d_ref, 0, K * 16 * 8, st)); CK(cudaMemsetAsync(d_ell, 0, K * 16 * 8, st));
    if (vcsmc::launch_merge_score(d_codes, S, d_pool, S, d_l, d_r, nullptr, d_P, d_pi, K, nullptr, S, 0, 0, nullptr, nullptr, d_ref, &parts_ref, st)) return 1;
    auto run = [&] { return vcsmc::launch_merge_score(d_codes, S, d_pool, S, d_l, d_r, d_order, d_P, d_pi, K, d_count, S, 0, 1, d_perm, d_tstate, d_ell, &parts, st); };
    for (int i = 0; i < 3; ++i) if (run()) return 1;
    // the timed launches replay as one CUDA graph (no host launch cost between them)
    cudaGraph_t graph; cudaGraphExec_t gexec;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < reps; ++i) run();
    CK(cudaStreamEndCapture(st, &graph)); CK(cudaGraphInstantiate(&gexec, graph, 0));
    CK(cudaGraphLaunch(gexec, st)); CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    CK(cudaGraphLaunch(gexec, st));
    CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    CK(cudaGraphExecDestroy(gexec)); CK(cudaGraphDestroy(graph));
    std::vector<double> a(K * parts), b(K * parts_ref);
    CK(cudaMemcpy(a.data(), d_ell, a.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), d_ref, b.size() * 8, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int64_t i = 0; i < sc.count; ++i) {
      const int64_t k = key[i].second;
      double x = 0, y = 0;
      for (int p = 0; p < parts; ++p) x += a[k * parts + p];
      for (int p = 0; p < parts_ref; ++p) y += b[k * parts_ref + p];
      err = std::max(err, std::fabs(x - y) / std::fabs(y));
    }
    worst = std::max(worst, err);
    const double ops = (double)sc.count * S * (sc.leaf ? 5.0 : 18.0);
    printf("%s count %6lld: %8.1f us  %6.2f T op/s  frac %.3f  (both launches; max rel err vs identity-order generic kernel %.2e)\n",
           sc.name, (long long)sc.count, ms * 1e3, ops / ms / 1e9, ops / ms / 1e9 / 18.2, err);
  }
  printf(worst < 1e-12 ? "CHECK OK\n" : "CHECK FAILED (%.3e)\n", worst);
  return worst < 1e-12 ? 0 : 1;
}
