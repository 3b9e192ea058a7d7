/*
 * oracle_build.c — TEST INFRASTRUCTURE ONLY (see oracle_fp.c). CPU restatement of the reference's term construction from a
 * smoothed bounds matrix, written independently of the product's builders (nvmolkit_b200/csrc/builders.cu) and compared
 * with them bit for bit in tests/test_builders.py.
 *
 * Follows rdkit_extensions/dist_geom_flattened_builder.cpp (nvMolKit v0.5.0):
 *   :56-86    addDistViolationContribs   (i > j, lb^2 / ub^2 / weight 1, kept when ub - lb <= basinSizeTol)
 *   :88-109   addChiralViolationContribs :111-122 addFourthDimContribs
 *   :124-176  addExperimentalTorsionTerms (marks the 1-4 pair)      :178-235 calcInversionCoefficientsAndForceConstant
 *   :237-305  addImproperTorsionTerms (3 permutations, x10)          :323-352 add12Terms   :373-430 add13Terms
 *   :432-470  addLongRangeDistanceConstraints (k = 10 x boundsMatForceScaling)
 * Output = SoA columns in the reference's own order (the tests convert both sides to a canonical form).
 * Pinning: no RDKit here; pinned by hand-computed cases in tests/test_builders.py (inversion coefficients of C / N / P,
 * a 4-atom chain). "parity unpinned" against a live RDKit for the CrystalFFDetails contents themselves.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static double ub_(const double* m, int n, int i, int j) { return i < j ? m[(size_t)i * n + j] : m[(size_t)j * n + i]; }
static double lb_(const double* m, int n, int i, int j) { return i < j ? m[(size_t)j * n + i] : m[(size_t)i * n + j]; }

/* returns the number of distance terms; idx [n][2] int32, par [n][3] = lb2, ub2, weight */
int oracle_dg_dist_terms(int n, const double* bounds, double basin, int32_t* idx, double* par) {
  int k = 0;
  for (int i = 1; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      const double l = lb_(bounds, n, i, j), u = ub_(bounds, n, i, j);
      if (u - l <= basin) {
        idx[2 * k] = i;
        idx[2 * k + 1] = j;
        par[3 * k] = l * l;
        par[3 * k + 1] = u * u;
        par[3 * k + 2] = 1.0;
        ++k;
      }
    }
  return k;
}

/* k (already / 3), C0, C1, C2 */
void oracle_inversion_coefficients(int z, int c_bound_to_o, double* out4) {
  double res, c0, c1, c2;
  if (z == 6 || z == 7 || z == 8) {
    c0 = 1.0;
    c1 = -1.0;
    c2 = 0.0;
    res = c_bound_to_o ? 50.0 : 6.0;
  } else {
    double w = M_PI / 180.0;
    if (z == 15) w *= 84.4339;
    else if (z == 33) w *= 86.9735;
    else if (z == 51) w *= 87.7047;
    else if (z == 83) w *= 90.0;
    c2 = 1.0;
    c1 = -4.0 * cos(w);
    c0 = -(c1 * cos(w) + c2 * cos(2.0 * w));
    res = 22.0 / (c0 + c1 + c2);
  }
  out4[0] = res / 3.0;
  out4[1] = c0;
  out4[2] = c1;
  out4[3] = c2;
}

/* ETK terms. Inputs as in b200mol_crystalff_details. Outputs (caller-sized):
 *   imp_idx [3 nImp][4], imp_par [3 nImp][4] = C0, C1, C2, k      d12 [nBonds]: idx[2], par[4] = half-width window centre
 *   ... the 1-2 / free 1-3 windows are written as {mid - 0.01, mid + 0.01, 100, 0} with mid = (lb + ub) / 2: the product's
 *   position-free encoding of "current distance +- 0.01" (the kernel re-centres it), see builders.cu.
 * counts[5] = nImproperTerms, n13, nAngle13, nLongRange, numImpropers */
void oracle_etk_terms(int n, const double* bounds, int nTor, const int32_t* torAtoms, int nImp, const int32_t* impAtoms, int nBonds,
                      const int32_t* bondAtoms, int nAng, const int32_t* angAtoms, double scaling, int basic, int32_t* imp_idx,
                      double* imp_par, double* d12_par, int32_t* d13_idx, double* d13_par, int32_t* a13_idx, double* a13_par,
                      int32_t* lr_idx, double* lr_par, int32_t* counts) {
  char* pair = calloc((size_t)n * n + 1, 1);
  char* constrained = calloc((size_t)n + 1, 1);
#define MARK(a, b) pair[(size_t)((a) < (b) ? (a) : (b)) * n + ((a) < (b) ? (b) : (a))] = 1
  for (int t = 0; t < nTor; ++t) MARK(torAtoms[4 * t], torAtoms[4 * t + 3]);
  int ni = 0;
  if (basic) {
    for (int t = 0; t < nImp; ++t) {
      const int32_t* a = impAtoms + 6 * t;
      double         c[4];
      oracle_inversion_coefficients(a[4], a[5], c);
      for (int p = 0; p < 3; ++p, ++ni) {
        int o0, o2, o3;
        if (p == 0) { o0 = 0; o2 = 2; o3 = 3; }
        else if (p == 1) { o0 = 0; o2 = 3; o3 = 2; }
        else { o0 = 2; o2 = 3; o3 = 0; }
        imp_idx[4 * ni] = a[o0];
        imp_idx[4 * ni + 1] = a[1];
        imp_idx[4 * ni + 2] = a[o2];
        imp_idx[4 * ni + 3] = a[o3];
        imp_par[4 * ni] = c[1];
        imp_par[4 * ni + 1] = c[2];
        imp_par[4 * ni + 2] = c[3];
        imp_par[4 * ni + 3] = c[0] * 10.0;
      }
      constrained[a[1]] = 1;
    }
  }
  for (int t = 0; t < nBonds; ++t) {
    const int i = bondAtoms[2 * t], j = bondAtoms[2 * t + 1];
    MARK(i, j);
    const double mid = 0.5 * (lb_(bounds, n, i, j) + ub_(bounds, n, i, j));
    d12_par[4 * t] = mid - 0.01;
    d12_par[4 * t + 1] = mid + 0.01;
    d12_par[4 * t + 2] = 100.0;
    d12_par[4 * t + 3] = 0.0;
  }
  int n13 = 0, na = 0;
  for (int t = 0; t < nAng; ++t) {
    const int i = angAtoms[4 * t], c = angAtoms[4 * t + 1], j = angAtoms[4 * t + 2];
    MARK(i, j);
    if (basic && angAtoms[4 * t + 3] != 0) {
      a13_idx[3 * na] = i;
      a13_idx[3 * na + 1] = c;
      a13_idx[3 * na + 2] = j;
      a13_par[2 * na] = 179.0;
      a13_par[2 * na + 1] = 180.0;
      ++na;
    } else {
      d13_idx[2 * n13] = i;
      d13_idx[2 * n13 + 1] = j;
      if (constrained[c]) {
        d13_par[4 * n13] = lb_(bounds, n, i, j);
        d13_par[4 * n13 + 1] = ub_(bounds, n, i, j);
        d13_par[4 * n13 + 3] = 1.0;
      } else {
        const double mid = 0.5 * (lb_(bounds, n, i, j) + ub_(bounds, n, i, j));
        d13_par[4 * n13] = mid - 0.01;
        d13_par[4 * n13 + 1] = mid + 0.01;
        d13_par[4 * n13 + 3] = 0.0;
      }
      d13_par[4 * n13 + 2] = 100.0;
      ++n13;
    }
  }
  int nl = 0;
  for (int i = 1; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (!pair[(size_t)j * n + i]) {
        lr_idx[2 * nl] = i;
        lr_idx[2 * nl + 1] = j;
        lr_par[3 * nl] = lb_(bounds, n, i, j);
        lr_par[3 * nl + 1] = ub_(bounds, n, i, j);
        lr_par[3 * nl + 2] = scaling * 10.0;
        ++nl;
      }
#undef MARK
  counts[0] = ni;
  counts[1] = n13;
  counts[2] = na;
  counts[3] = nl;
  counts[4] = basic ? nImp : 0;
  free(pair);
  free(constrained);
}
