// Host-side gradient schedule: orders the terms of a table into atom-disjoint WAVES of at most 32 (include/b200mol.h,
// b200mol_term_table). The kernels (ff.cuh forTerms<true>) give one wave to one warp, lane = term, and add the
// gradient contributions with plain shared-memory read-modify-writes - no atomics, bit-reproducible sums.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace b200 {
namespace {
constexpr int kWave = 32;

// Dense pair table: rounds of the circle-method tournament. With M odd, the pairs {i, j} with (i + j) mod M == r form a
// matching (atom i meets exactly one j per round; the atom with 2 i == r mod M sits out), so a round is conflict-free
// and an all-pairs table gives rounds of (M - 1) / 2 terms: ceil(that / 32) waves each.
void scheduleRounds(const int16_t* idx, int beg, int end, int nAtoms, int32_t* perm, std::vector<int32_t>& waves) {
  const int M = (nAtoms & 1) ? nAtoms : nAtoms + 1;
  std::iota(perm + beg, perm + end, beg);
  auto key = [&](int t) {
    const int i = idx[2 * t], j = idx[2 * t + 1];
    return std::make_pair((i + j) % M, std::min(i, j));
  };
  std::sort(perm + beg, perm + end, [&](int a, int b) { return key(a) < key(b); });
  int inWave = 0, lastRound = -1;
  for (int p = beg; p < end; ++p) {
    const int r = key(perm[p]).first;
    if (r != lastRound || inWave == kWave) {
      waves.push_back(p);
      inWave = 0;
    }
    lastRound = r;
    ++inWave;
  }
}

// Anything else: first-fit colouring. A term goes to the EARLIEST wave that has room and holds none of its atoms (one
// bitset of atoms per wave). The number of waves is bounded below by the busiest atom's term count (a ring carbon sits in
// ~30 torsions), so the sparse bonded tables end up with short waves - they are ~10 % of a molecule's terms. A term that
// names an atom twice (degenerate input) is fine: its adds to that atom come from ONE lane, in program order.
void scheduleFirstFit(const int16_t* idx, int K, int beg, int end, int nAtoms, int32_t* perm, std::vector<int32_t>& waves) {
  const int                          words = (nAtoms + 63) / 64;
  std::vector<uint64_t>              used;  // [wave][words]
  std::vector<int>                   fill, waveOf(end - beg);
  for (int t = beg; t < end; ++t) {
    const int16_t* a = idx + static_cast<size_t>(K) * t;
    int            w = 0;
    for (;; ++w) {
      if (w == static_cast<int>(fill.size())) {
        fill.push_back(0);
        used.resize(used.size() + words, 0);
        break;
      }
      if (fill[w] == kWave) continue;
      bool clash = false;
      for (int k = 0; k < K && !clash; ++k) clash = (used[static_cast<size_t>(w) * words + (a[k] >> 6)] >> (a[k] & 63)) & 1u;
      if (!clash) break;
    }
    ++fill[w];
    waveOf[t - beg] = w;
    for (int k = 0; k < K; ++k) used[static_cast<size_t>(w) * words + (a[k] >> 6)] |= 1ull << (a[k] & 63);
  }
  std::iota(perm + beg, perm + end, beg);
  std::stable_sort(perm + beg, perm + end, [&](int x, int y) { return waveOf[x - beg] < waveOf[y - beg]; });
  int p = beg;
  for (int f : fill) {
    waves.push_back(p);
    p += f;
  }
}
}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200mol_schedule_waves(int32_t nMols, const int32_t* h_starts, const int16_t* h_idx, int K, int32_t* h_perm,
                                      int32_t* h_mol_waves, int32_t* h_waves, int64_t* n_waves) {
  return guarded([&] {
    B200_REQUIRE(nMols >= 0 && h_starts && h_perm && h_mol_waves && h_waves && n_waves, "null pointer");
    B200_REQUIRE(K >= 1 && K <= 8, "K must be in 1..8");
    std::vector<int32_t> waves;
    for (int m = 0; m < nMols; ++m) {
      const int beg = h_starts[m], end = h_starts[m + 1];
      B200_REQUIRE(beg <= end, "term starts must be non-decreasing");
      h_mol_waves[m] = static_cast<int32_t>(waves.size());
      if (beg == end) continue;
      B200_REQUIRE(h_idx, "null index table");
      int nAtoms = 0;
      for (int t = K * beg; t < K * end; ++t) {
        B200_REQUIRE(h_idx[t] >= 0, "negative atom index in molecule %d", m);
        nAtoms = std::max(nAtoms, h_idx[t] + 1);
      }
      if (K == 2 && end - beg >= 2 * nAtoms) scheduleRounds(h_idx, beg, end, nAtoms, h_perm, waves);
      else scheduleFirstFit(h_idx, K, beg, end, nAtoms, h_perm, waves);
    }
    h_mol_waves[nMols] = static_cast<int32_t>(waves.size());
    waves.push_back(h_starts[nMols]);
    std::copy(waves.begin(), waves.end(), h_waves);
    *n_waves = static_cast<int64_t>(waves.size()) - 1;
  });
}
