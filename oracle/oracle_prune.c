/*
 * oracle_prune.c — TEST INFRASTRUCTURE ONLY (see oracle_fp.c). RMS pruning of conformers on the CPU, following
 * rdkit_extensions/conformer_pruning.cpp:96-137 of the reference (RDKit's _isConfFarFromRest): a conformer is kept iff its
 * best-alignment sum of squared deviations to every conformer kept before it is >= nSel * thresh^2, for every self match.
 * The optimal SSD is computed the way RDKit's AlignPoints does - largest eigenvalue of Horn's 4x4 quaternion matrix, here
 * by Jacobi rotations - i.e. by a different route than the product's closed-form 3x3 singular values (csrc/prune.cu).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static double max_eig4(double A[4][4]) {
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
    if (off < 1e-30) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double th = 0.5 * atan2(2.0 * A[p][q], A[q][q] - A[p][p]);
        const double c = cos(th), s = sin(th);
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
      }
  }
  double m = A[0][0];
  for (int k = 1; k < 4; ++k)
    if (A[k][k] > m) m = A[k][k];
  return m;
}

/* best-alignment SSD of point lists a[ia[i]] and b[ib[i]], i < n */
double oracle_best_ssd(const double* a, const int16_t* ia, const double* b, const int16_t* ib, int n) {
  double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      ca[k] += a[3 * (ia ? ia[i] : i) + k];
      cb[k] += b[3 * (ib ? ib[i] : i) + k];
    }
  for (int k = 0; k < 3; ++k) ca[k] /= n, cb[k] /= n;
  double S[3][3] = {{0}}, G = 0.0;
  for (int i = 0; i < n; ++i) {
    double x[3], y[3];
    for (int k = 0; k < 3; ++k) {
      x[k] = a[3 * (ia ? ia[i] : i) + k] - ca[k];
      y[k] = b[3 * (ib ? ib[i] : i) + k] - cb[k];
      G += x[k] * x[k] + y[k] * y[k];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) S[r][c] += x[r] * y[c];
  }
  double N[4][4] = {
    {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
    {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
    {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
    {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
  const double ssd = G - 2.0 * max_eig4(N);
  return ssd > 0.0 ? ssd : 0.0;
}

/* conformers [c0, c1) of one molecule; K matches of L atoms (match_atoms NULL: identity over L = all atoms) */
void oracle_rms_prune_mol(int c0, int c1, const int32_t* conf_atom_start, const double* xyz, int K, int L, const int16_t* match_atoms,
                          double thresh, const uint8_t* valid, uint8_t* keep) {
  const double limit = L * thresh * thresh;
  for (int c = c0; c < c1; ++c) {
    int ok = !valid || valid[c];
    for (int k = c0; k < c && ok; ++k) {
      if (!keep[k]) continue;
      for (int m = 0; m < K && ok; ++m)
        if (oracle_best_ssd(xyz + 3 * (size_t)conf_atom_start[c], match_atoms, xyz + 3 * (size_t)conf_atom_start[k],
                            match_atoms ? match_atoms + (size_t)m * L : NULL, L) < limit)
          ok = 0;
    }
    keep[c] = (uint8_t)ok;
  }
}
