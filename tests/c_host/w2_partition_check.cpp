// CPU check of the pair-wgrad work split (sfm-nerf_b200/csrc/mlp_layout.cuh: w2_partition): every (item, tile) is covered exactly
// once, no pair gets more than two units, and the heaviest pair is within a few per cent of the mean.
//   g++ -std=c++17 -I sfm-nerf_b200/csrc tests/c_host/w2_partition_check.cpp -o w2_check && ./w2_check
#include <cstdio>
#include <cstring>
#include <vector>
#include "mlp_layout.cuh"
using namespace nerf;
int main() {
  int bad = 0;
  const int layouts[] = {0, 1, 2, 3, make_layout_id(6, 2, 1, 0, 256), make_layout_id(10, 10, 1, 0, 256), make_layout_id(10, 4, 1, 0, 128),
                         make_layout_id(3, 8, 1, 1, 128), make_layout_id(0, 0, 0, 0, 128)};
  for (int layout : layouts) {
    ModelLayout m;
    if (!build_model_layout(layout, m)) return 2;
    for (int n_pairs : {74, 66, 9, 96, 120}) {
      for (int n_tiles : {1, 3, 15, 16, 17, 64, 100, 2048, 2049, 16384, 960000}) {
        static int32_t n_units[kW2MaxPairs];
        static W2Unit unit[kW2MaxPairs][kW2MaxUnits];
        std::memset(n_units, 0, sizeof(n_units));
        const int used = w2_partition(m, n_tiles, n_pairs, n_units, unit);
        if (used < 1) { std::printf("layout %d pairs %d tiles %d: partition failed\n", layout, n_pairs, n_tiles); ++bad; continue; }
        std::vector<std::vector<int>> seen(m.n_w2, std::vector<int>(n_tiles < 100000 ? n_tiles : 0, 0));
        std::vector<long long> covered(m.n_w2, 0);
        double worst = 0.0, total = 0.0;
        for (int p = 0; p < used; ++p) {
          double bytes = 0.0;
          for (int u = 0; u < n_units[p]; ++u) {
            const W2Unit& q = unit[p][u];
            if (q.item < 0 || q.item >= m.n_w2 || q.t0 < 0 || q.t1 > n_tiles || q.t0 >= q.t1) { ++bad; continue; }
            covered[q.item] += q.t1 - q.t0;
            if (n_tiles < 100000) for (int t = q.t0; t < q.t1; ++t) ++seen[q.item][t];
            bytes += (double)m.w2[q.item].cost * (q.t1 - q.t0);
            if (u > 0 && unit[p][u - 1].item >= q.item) ++bad;
          }
          if (bytes > worst) worst = bytes;
          total += bytes;
        }
        for (int i = 0; i < m.n_w2; ++i) {
          if (covered[i] != n_tiles) { std::printf("layout %d pairs %d tiles %d item %d covered %lld\n", layout, n_pairs, n_tiles, i, covered[i]); ++bad; }
          if (n_tiles < 100000) for (int t = 0; t < n_tiles; ++t) if (seen[i][t] != 1) { ++bad; break; }
        }
        const double ratio = worst / (total / used);
        if (n_tiles >= 2048 && used == (n_pairs > kW2MaxPairs ? kW2MaxPairs : n_pairs) && ratio > 1.01) { std::printf("imbalance %.4f (layout %d pairs %d tiles %d)\n", ratio, layout, n_pairs, n_tiles); ++bad; }
        if (layout == 0 && n_pairs == 74 && n_tiles == 2048) std::printf("C2 shape: %d pairs, heaviest / mean = %.4f\n", used, ratio);
      }
    }
    // ring geometry of every item fits
    for (int i = 0; i < m.n_w2; ++i) {
      const W2Item& it = m.w2[i];
      if (it.depth < 2 || it.depth * it.stage_bytes > kW2RingBytes || it.depth * it.n_loads > 9) { std::printf("item %d geometry\n", i); ++bad; }
      int cols = 0;
      for (int q = 0; q < it.n_mmas; ++q) if (it.mma[q].d_col + it.mma[q].n > cols) cols = it.mma[q].d_col + it.mma[q].n;
      if (cols > 512) { std::printf("item %d needs %d TMEM columns\n", i, cols); ++bad; }
    }
  }
  std::printf(bad ? "FAILED (%d)\n" : "ok\n", bad);
  return bad ? 1 : 0;
}
