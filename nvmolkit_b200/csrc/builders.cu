// Host-side term construction: smoothed bounds matrix (+ RDKit's CrystalFFDetails as plain arrays) -> the DG and ETK term
// tables of include/b200mol.h. Pure arithmetic, no RDKit. Replaces the reference's flatteners
//   constructForceFieldContribs / construct3DForceFieldContribs (rdkit_extensions/dist_geom_flattened_builder.cpp:472-541)
// with these differences of LAYOUT only: records are {int16 local indices, fp64 parameters} rows instead of SoA columns
// with global int32 indices, and the 1-2 / free 1-3 windows carry their half-width (0.01 A) and a `fixed` flag instead of
// being centred on build-time coordinates - the kernels re-centre them on the geometry the ETK stage starts from
// (what the reference's refresh does, src/etkdg_stage_etk_minimization.cu:32-64).
#include <cmath>
#include <vector>

#include "common.cuh"

namespace b200 {
namespace {

// RDKit BoundsMatrix: upper bound at [min][max], lower bound at [max][min].
struct Bounds {
  const double* m;
  int           n;
  double        ub(int i, int j) const { return i < j ? m[static_cast<size_t>(i) * n + j] : m[static_cast<size_t>(j) * n + i]; }
  double        lb(int i, int j) const { return i < j ? m[static_cast<size_t>(j) * n + i] : m[static_cast<size_t>(i) * n + j]; }
};

// Inversion coefficients and force constant of an improper centre: sp2 C / N / O, else the group-15 formula
// (dist_geom_flattened_builder.cpp:178-235); the result already carries the / 3.
void inversionCoefficients(int z, bool cBoundToO, double& k, double& c0, double& c1, double& c2) {
  if (z == 6 || z == 7 || z == 8) {
    c0 = 1.0;
    c1 = -1.0;
    c2 = 0.0;
    k  = cBoundToO ? 50.0 : 6.0;
  } else {
    double w = M_PI / 180.0;
    switch (z) {
      case 15: w *= 84.4339; break;
      case 33: w *= 86.9735; break;
      case 51: w *= 87.7047; break;
      case 83: w *= 90.0; break;
      default: break;
    }
    c2 = 1.0;
    c1 = -4.0 * std::cos(w);
    c0 = -(c1 * std::cos(w) + c2 * std::cos(2.0 * w));
    k  = 22.0 / (c0 + c1 + c2);
  }
  k /= 3.0;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200mol_dg_terms_from_bounds(int32_t nAtoms, const double* h_bounds, int32_t nChiral, const int32_t* h_chiral_atoms,
                                            const double* h_chiral_bounds, int dim, double basinSizeTol, int16_t* h_dist_idx,
                                            double* h_dist_par, int16_t* h_chiral_idx, double* h_chiral_par,
                                            int16_t* h_fourth_idx, int32_t* h_counts3) {
  return guarded([&] {
    B200_REQUIRE(nAtoms >= 0 && nAtoms <= 32767, "atom count out of range");
    B200_REQUIRE(nAtoms == 0 || h_bounds, "null bounds matrix");
    B200_REQUIRE(h_dist_idx && h_dist_par && h_counts3, "null output");
    B200_REQUIRE(dim == 3 || dim == 4, "dim must be 3 or 4");
    const Bounds b{h_bounds, nAtoms};
    int          nd = 0;
    for (int i = 1; i < nAtoms; ++i)
      for (int j = 0; j < i; ++j) {
        const double l = b.lb(i, j), u = b.ub(i, j);
        if (u - l <= basinSizeTol) {  // basinSizeTol = 1e8 in the ETKDG pipeline: every pair (src/etkdg.cpp, embedder_utils.cpp:54-66)
          h_dist_idx[2 * nd]     = static_cast<int16_t>(i);
          h_dist_idx[2 * nd + 1] = static_cast<int16_t>(j);
          h_dist_par[3 * nd]     = l * l;
          h_dist_par[3 * nd + 1] = u * u;
          h_dist_par[3 * nd + 2] = 1.0;
          ++nd;
        }
      }
    B200_REQUIRE(nChiral >= 0 && (nChiral == 0 || (h_chiral_atoms && h_chiral_bounds && h_chiral_idx && h_chiral_par)), "null chiral arrays");
    for (int c = 0; c < nChiral; ++c) {
      for (int k = 0; k < 4; ++k) {
        const int a = h_chiral_atoms[4 * c + k];
        B200_REQUIRE(a >= 0 && a < nAtoms, "chiral set %d names atom %d of %d", c, a, nAtoms);
        h_chiral_idx[4 * c + k] = static_cast<int16_t>(a);
      }
      h_chiral_par[2 * c]     = h_chiral_bounds[2 * c + 1];  // table order {volUpper, volLower}; input {lower, upper}
      h_chiral_par[2 * c + 1] = h_chiral_bounds[2 * c];
    }
    int nf = 0;
    if (dim == 4) {
      B200_REQUIRE(nAtoms == 0 || h_fourth_idx, "null fourth-dimension output");
      for (; nf < nAtoms; ++nf) h_fourth_idx[nf] = static_cast<int16_t>(nf);
    }
    h_counts3[0] = nd;
    h_counts3[1] = nChiral;
    h_counts3[2] = nf;
  });
}

extern "C" int b200mol_etk_terms_from_details(int32_t nAtoms, const double* h_bounds, const b200mol_crystalff_details* d,
                                              int useBasicKnowledge, b200mol_etk_term_buffers* out, int32_t* h_counts6,
                                              int32_t* h_num_impropers) {
  return guarded([&] {
    B200_REQUIRE(nAtoms >= 0 && nAtoms <= 32767, "atom count out of range");
    B200_REQUIRE(d && out && h_counts6 && h_num_impropers, "null pointer");
    B200_REQUIRE(nAtoms == 0 || h_bounds, "null bounds matrix");
    const Bounds      b{h_bounds, nAtoms};
    std::vector<char> paired(static_cast<size_t>(nAtoms) * nAtoms, 0), improperCentre(nAtoms, 0);
    auto              mark = [&](int i, int j) { paired[static_cast<size_t>(i < j ? i : j) * nAtoms + (i < j ? j : i)] = 1; };
    auto              atom = [&](int a, const char* what) {
      B200_REQUIRE(a >= 0 && a < nAtoms, "%s names atom %d of %d", what, a, nAtoms);
      return static_cast<int16_t>(a);
    };
    // 1. experimental torsions: six force constants and six signs each (:124-176)
    for (int t = 0; t < d->nTorsions; ++t) {
      const int32_t* a = d->torsionAtoms + 4 * t;
      B200_REQUIRE(a[0] != a[1] && a[0] != a[2] && a[0] != a[3] && a[1] != a[2] && a[1] != a[3] && a[2] != a[3], "degenerate torsion %d", t);
      for (int k = 0; k < 4; ++k) out->torsion_idx[4 * t + k] = atom(a[k], "torsion");
      mark(a[0], a[3]);
      for (int k = 0; k < 6; ++k) {
        out->torsion_par[12 * t + k]     = d->torsionV[6 * t + k];
        out->torsion_par[12 * t + 6 + k] = static_cast<double>(d->torsionSigns[6 * t + k]);
      }
    }
    // 2. improper (inversion) terms: three permutations per centre, force scaling 10 (:237-305)
    int ni = 0;
    if (useBasicKnowledge) {
      static const int perm[3][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {2, 1, 3, 0}};
      for (int t = 0; t < d->nImpropers; ++t) {
        const int32_t* a = d->improperAtoms + 6 * t;
        double         k, c0, c1, c2;
        inversionCoefficients(a[4], a[5] != 0, k, c0, c1, c2);
        for (int p = 0; p < 3; ++p, ++ni) {
          for (int q = 0; q < 4; ++q) out->improper_idx[4 * ni + q] = atom(a[perm[p][q]], "improper");
          out->improper_par[4 * ni]     = c0;
          out->improper_par[4 * ni + 1] = c1;
          out->improper_par[4 * ni + 2] = c2;
          out->improper_par[4 * ni + 3] = k * 10.0;
        }
        improperCentre[a[1]] = 1;
      }
    }
    *h_num_impropers = useBasicKnowledge ? d->nImpropers : 0;
    // 3. 1-2 windows: current distance +- 0.01, k = 100 (:323-352); stored as a window of half-width 0.01 around the
    // middle of the bounds, fixed = 0 -> re-centred on the stage's starting geometry by the kernel
    constexpr double kTol = 0.01, kKnown = 100.0;
    for (int t = 0; t < d->nBonds; ++t) {
      const int i = d->bonds[2 * t], j = d->bonds[2 * t + 1];
      out->dist12_idx[2 * t]     = atom(i, "bond");
      out->dist12_idx[2 * t + 1] = atom(j, "bond");
      mark(i, j);
      const double mid = 0.5 * (b.lb(i, j) + b.ub(i, j));
      double*      q   = out->dist12_par + 4 * t;
      q[0] = mid - kTol, q[1] = mid + kTol, q[2] = kKnown, q[3] = 0.0;
    }
    // 4. 1-3 terms (:373-430): triple-bond angle 179..180 | improper-constrained centre: the bounds, fixed | else +- 0.01
    int n13 = 0, na = 0;
    for (int t = 0; t < d->nAngles; ++t) {
      const int32_t* a = d->angles + 4 * t;
      const int      i = a[0], c = a[1], j = a[2];
      atom(i, "angle"), atom(c, "angle"), atom(j, "angle");
      mark(i, j);
      if (useBasicKnowledge && a[3] != 0) {
        out->angle13_idx[3 * na] = static_cast<int16_t>(i), out->angle13_idx[3 * na + 1] = static_cast<int16_t>(c);
        out->angle13_idx[3 * na + 2] = static_cast<int16_t>(j);
        out->angle13_par[2 * na] = 179.0, out->angle13_par[2 * na + 1] = 180.0;
        ++na;
      } else {
        out->dist13_idx[2 * n13] = static_cast<int16_t>(i), out->dist13_idx[2 * n13 + 1] = static_cast<int16_t>(j);
        double* q = out->dist13_par + 4 * n13;
        if (improperCentre[c]) {
          q[0] = b.lb(i, j), q[1] = b.ub(i, j), q[2] = kKnown, q[3] = 1.0;
        } else {
          const double mid = 0.5 * (b.lb(i, j) + b.ub(i, j));
          q[0] = mid - kTol, q[1] = mid + kTol, q[2] = kKnown, q[3] = 0.0;
        }
        ++n13;
      }
    }
    // 5. long-range: every pair not named above, its bounds, k = 10 x boundsMatForceScaling (:432-470)
    int nl = 0;
    for (int i = 1; i < nAtoms; ++i)
      for (int j = 0; j < i; ++j) {
        if (paired[static_cast<size_t>(j) * nAtoms + i]) continue;
        out->longrange_idx[2 * nl] = static_cast<int16_t>(i), out->longrange_idx[2 * nl + 1] = static_cast<int16_t>(j);
        double* q = out->longrange_par + 3 * nl;
        q[0] = b.lb(i, j), q[1] = b.ub(i, j), q[2] = d->boundsMatForceScaling * 10.0;
        ++nl;
      }
    h_counts6[0] = d->nTorsions, h_counts6[1] = ni, h_counts6[2] = d->nBonds, h_counts6[3] = n13, h_counts6[4] = na, h_counts6[5] = nl;
  });
}
