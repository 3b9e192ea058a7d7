/*
 * oracle_ff.c — TEST INFRASTRUCTURE ONLY (see oracle_fp.c). Path B on the CPU, pure fp64:
 *   MMFF94, distance-geometry (DG) and experimental-torsion (ETK) energy + gradient over flattened term tables,
 *   and the BFGS minimiser of RDKit (BFGSOpt.h + ForceField::minimize gradient scaling).
 *
 * What each function follows (nvMolKit v0.5.0 checkout; the reference itself declares these "1:1 ports" of RDKit):
 *   MMFF term math   src/forcefields/mmff_kernels_device.cuh:28-661 (without its fp32 shortcuts)
 *   DG / ETK math    src/forcefields/dist_geom_kernels_device.cuh:37-830
 *   BFGS             src/minimizer/bfgs_minimize.cu:80-162 (line-search setup), :202-311 (perturb / backtrack),
 *                    :327-356 (restore on failure), :610-632 (max step), :732-776 (direction, TOLX),
 *                    :797-851 (gradient scaling, RDKit >= 2025.09 rule), :873-918 (gradient convergence),
 *                    src/minimizer/bfgs_hessian.cu:37-239 (inverse-Hessian update), :978-1053 (driver loop).
 * Pinning: the BFGS driver is pinned by the reference's analytic systems (tests/test_bfgs_minimizer.cu:822-930,
 * 1159-1240) in tests/test_oracle_golden.py; force-field terms are pinned by finite differences of their own energies
 * and by hand-computed single-term values. RDKit-generated fixtures for real molecules do not exist in this container
 * ("parity unpinned" against RDKit for the MMFF/DG parametrisation — DESIGN.md).
 *
 * Term-table layout = include/b200mol.h (b200mol_term_table): CSR starts per molecule, int16 molecule-local atom
 * indices [n][K], fp64 parameters [n][P].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const int32_t* starts;
  const int16_t* idx;
  const double*  par;
} TermTable;

typedef struct {
  int32_t        nMols;
  const int32_t* atomCounts;
  TermTable      bond, angle, strbend, oop, torsion, vdw, ele;
  TermTable      distc, posc, anglec, torsc; /* restraints */
} MmffSystem;

typedef struct {
  int32_t        nMols;
  const int32_t* atomCounts;
  TermTable      dist, chiral, fourth;
} DgSystem;

typedef struct {
  int32_t        nMols;
  const int32_t* atomCounts;
  TermTable      torsion, improper, dist12, dist13, angle13, longrange;
} EtkSystem;

#define DEG2RAD (M_PI / 180.0)
#define RAD2DEG (180.0 / M_PI)
static inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int    is_zero(double v) { return v < 1.0e-10 && v > -1.0e-10; }
static inline void   cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* =========================================================================================== MMFF */
static double mmff_bond(const double* p, int i, int j, double r0, double kb, double* g) {
  double d[3] = {p[3 * i] - p[3 * j], p[3 * i + 1] - p[3 * j + 1], p[3 * i + 2] - p[3 * j + 2]};
  const double dist = sqrt(dot3(d, d)), dr = dist - r0, cs = -2.0;
  if (g) {
    const double de = 143.9325 * kb * dr * (1.0 + 1.5 * cs * dr + 2.0 * 7.0 / 12.0 * cs * cs * dr * dr);
    for (int k = 0; k < 3; ++k) {
      const double v = dist > 0.0 ? de * d[k] / dist : kb * 0.01;
      g[3 * i + k] += v;
      g[3 * j + k] -= v;
    }
  }
  return 143.9325 / 2.0 * kb * dr * dr * (1.0 + cs * dr + 7.0 / 12.0 * cs * cs * dr * dr);
}

static double mmff_angle(const double* p, int i, int j, int k, double theta0, double ka, int linear, double* g) {
  double       d1[3], d2[3];
  for (int c = 0; c < 3; ++c) {
    d1[c] = p[3 * i + c] - p[3 * j + c];
    d2[c] = p[3 * k + c] - p[3 * j + c];
  }
  const double l1sq = dot3(d1, d1), l2sq = dot3(d2, d2);
  const double l1 = sqrt(l1sq), l2 = sqrt(l2sq);
  const double cosT = clampd(dot3(d1, d2) / (l1 * l2), -1.0, 1.0);
  const double theta = RAD2DEG * acos(cosT), dT = theta - theta0;
  if (g) {
    const double sinSq = 1.0 - cosT * cosT;
    if (!(is_zero(sinSq) || is_zero(l1sq) || is_zero(l2sq))) {
      const double de = linear ? -143.9325 * ka * sqrt(sinSq)
                               : 143.9325 * DEG2RAD * ka * dT * (1.0 + (-0.006981317 * 1.5) * dT);
      const double cf = de * (-1.0 / sqrt(sinSq));
      for (int c = 0; c < 3; ++c) {
        const double n1 = d1[c] / l1, n2 = d2[c] / l2;
        const double a = (n2 - cosT * n1) / l1, b = (n1 - cosT * n2) / l2;
        g[3 * i + c] += cf * a;
        g[3 * j + c] += cf * (-a - b);
        g[3 * k + c] += cf * b;
      }
    }
  }
  if (linear) return 143.9325 * ka * (1.0 + cosT);
  return 0.5 * 143.9325 * DEG2RAD * DEG2RAD * ka * dT * dT * (1.0 + (-0.4 * DEG2RAD) * dT);
}

static double mmff_strbend(const double* p, int i, int j, int k, double theta0, double r01, double r02, double f1,
                           double f2, double* g) {
  double d1[3], d2[3];
  for (int c = 0; c < 3; ++c) {
    d1[c] = p[3 * i + c] - p[3 * j + c];
    d2[c] = p[3 * k + c] - p[3 * j + c];
  }
  const double l1 = sqrt(dot3(d1, d1)), l2 = sqrt(dot3(d2, d2));
  const double cosT = clampd(dot3(d1, d2) / (l1 * l2), -1.0, 1.0);
  const double theta = RAD2DEG * acos(cosT), dT = theta - theta0, dr1 = l1 - r01, dr2 = l2 - r02;
  if (g) {
    const double pre = 143.9325 * DEG2RAD;
    double       invSin = 1.0 / sqrt(1.0 - cosT * cosT);
    if (!(invSin < 1.0e8)) invSin = 1.0e8;
    const double bt = RAD2DEG * (f1 * dr1 + f2 * dr2) * invSin;
    for (int c = 0; c < 3; ++c) {
      const double n1 = d1[c] / l1, n2 = d2[c] / l2;
      const double a = (n2 - cosT * n1) / l1, b = (n1 - cosT * n2) / l2;
      g[3 * i + c] += pre * (dT * n1 * f1 - a * bt);
      g[3 * j + c] += pre * (-dT * (n1 * f1 + n2 * f2) + (a + b) * bt);
      g[3 * k + c] += pre * (dT * n2 * f2 - b * bt);
    }
  }
  return 2.51210 * dT * (dr1 * f1 + dr2 * f2);
}

static double mmff_oop(const double* p, int i, int j, int k, int l, double koop, double* g) {
  double ji[3], jk[3], jl[3];
  for (int c = 0; c < 3; ++c) {
    ji[c] = p[3 * i + c] - p[3 * j + c];
    jk[c] = p[3 * k + c] - p[3 * j + c];
    jl[c] = p[3 * l + c] - p[3 * j + c];
  }
  const double li = sqrt(dot3(ji, ji)), lk = sqrt(dot3(jk, jk)), ll = sqrt(dot3(jl, jl));
  for (int c = 0; c < 3; ++c) {
    ji[c] /= li;
    jk[c] /= lk;
    jl[c] /= ll;
  }
  double nji[3] = {-ji[0], -ji[1], -ji[2]}, n[3];
  cross(nji, jk, n);
  const double nl = sqrt(dot3(n, n));
  for (int c = 0; c < 3; ++c) n[c] /= nl;
  const double sinChi = clampd(dot3(jl, n), -1.0, 1.0);
  const double chi    = RAD2DEG * asin(sinChi);
  if (g) {
    const double cosChiSq = 1.0 - sinChi * sinChi;
    const double invCosChi = cosChiSq > 0.0 ? 1.0 / sqrt(cosChiSq) : 1.0e8;
    const double cosT = clampd(dot3(ji, jk), -1.0, 1.0);
    const double invSinT = 1.0 / sqrt(fmax(1.0 - cosT * cosT, 1.0e-8));
    const double de = 143.9325 * DEG2RAD * koop * chi;
    double t1[3], t2[3], t3[3];
    cross(jl, jk, t1);
    cross(ji, jl, t2);
    cross(jk, ji, t3);
    const double term1 = invCosChi * invSinT, term2 = sinChi * invCosChi * invSinT * invSinT;
    for (int c = 0; c < 3; ++c) {
      const double g1 = (t1[c] * term1 - (ji[c] - jk[c] * cosT) * term2) / li;
      const double g3 = (t2[c] * term1 - (jk[c] - ji[c] * cosT) * term2) / lk;
      const double g4 = (t3[c] * term1 - jl[c] * sinChi * invCosChi) / ll;
      g[3 * i + c] += de * g1;
      g[3 * j + c] += -de * (g1 + g3 + g4);
      g[3 * k + c] += de * g3;
      g[3 * l + c] += de * g4;
    }
  }
  return 0.5 * 143.9325 * DEG2RAD * DEG2RAD * koop * chi * chi;
}

static double mmff_torsion(const double* p, int i, int j, int k, int l, double V1, double V2, double V3, double* g) {
  double d1[3], d2[3], d4[3], nd2[3];
  for (int c = 0; c < 3; ++c) {
    d1[c]  = p[3 * i + c] - p[3 * j + c];
    d2[c]  = p[3 * k + c] - p[3 * j + c];
    d4[c]  = p[3 * l + c] - p[3 * k + c];
    nd2[c] = -d2[c];
  }
  double c1[3], c2[3];
  cross(d1, d2, c1);
  cross(nd2, d4, c2);
  double inv1 = 1.0 / sqrt(dot3(c1, c1)), inv2 = 1.0 / sqrt(dot3(c2, c2));
  if (g) {
    if (!(inv1 < 1.0e5)) inv1 = 1.0e5;
    if (!(inv2 < 1.0e5)) inv2 = 1.0e5;
  }
  const double cosPhiE = clampd(dot3(c1, c2) * (1.0 / sqrt(dot3(c1, c1))) * (1.0 / sqrt(dot3(c2, c2))), -1.0, 1.0);
  if (g) {
    double u1[3], u2[3];
    for (int c = 0; c < 3; ++c) {
      u1[c] = c1[c] * inv1;
      u2[c] = c2[c] * inv2;
    }
    const double cosPhi = clampd(dot3(u1, u2), -1.0, 1.0);
    const double sinSq  = 1.0 - cosPhi * cosPhi;
    double       sinTerm = 0.0;
    if (sinSq > 0.0) sinTerm = 0.5 * (V1 - 2.0 * V2 * (2.0 * cosPhi) + 3.0 * V3 * (3.0 - 4.0 * sinSq));
    double dT[6];
    for (int c = 0; c < 3; ++c) {
      dT[c]     = inv1 * (u2[c] - cosPhi * u1[c]);
      dT[3 + c] = inv2 * (u1[c] - cosPhi * u2[c]);
    }
    const double dx1 = d1[0], dy1 = d1[1], dz1 = d1[2], dx2 = d2[0], dy2 = d2[1], dz2 = d2[2], dx4 = d4[0],
                 dy4 = d4[1], dz4 = d4[2];
    g[3 * i + 0] += sinTerm * (dT[2] * dy2 - dT[1] * dz2);
    g[3 * i + 1] += sinTerm * (dT[0] * dz2 - dT[2] * dx2);
    g[3 * i + 2] += sinTerm * (dT[1] * dx2 - dT[0] * dy2);
    g[3 * j + 0] += sinTerm * (dT[1] * (dz2 - dz1) + dT[2] * (dy1 - dy2) + dT[4] * (-dz4) + dT[5] * (dy4));
    g[3 * j + 1] += sinTerm * (dT[0] * (dz1 - dz2) + dT[2] * (dx2 - dx1) + dT[3] * (dz4) + dT[5] * (-dx4));
    g[3 * j + 2] += sinTerm * (dT[0] * (dy2 - dy1) + dT[1] * (dx1 - dx2) + dT[3] * (-dy4) + dT[4] * (dx4));
    g[3 * k + 0] += sinTerm * (dT[1] * (dz1) + dT[2] * (-dy1) + dT[4] * (dz4 + dz2) + dT[5] * (-dy4 - dy2));
    g[3 * k + 1] += sinTerm * (dT[0] * (-dz1) + dT[2] * (dx1) + dT[3] * (-dz4 - dz2) + dT[5] * (dx4 + dx2));
    g[3 * k + 2] += sinTerm * (dT[0] * (dy1) + dT[1] * (-dx1) + dT[3] * (dy4 + dy2) + dT[4] * (-dx4 - dx2));
    g[3 * l + 0] += sinTerm * (dT[4] * (-dz2) - dT[5] * (-dy2));
    g[3 * l + 1] += sinTerm * (dT[5] * (-dx2) - dT[3] * (-dz2));
    g[3 * l + 2] += sinTerm * (dT[3] * (-dy2) - dT[4] * (-dx2));
  }
  const double phi = acos(cosPhiE);
  return 0.5 * (V1 * (1.0 + cosPhiE) + V2 * (1.0 - cos(2.0 * phi)) + V3 * (1.0 + cos(3.0 * phi)));
}

static double mmff_vdw(const double* p, int i, int j, double R, double eps, double* g) {
  double d[3] = {p[3 * i] - p[3 * j], p[3 * i + 1] - p[3 * j + 1], p[3 * i + 2] - p[3 * j + 2]};
  const double d2 = dot3(d, d), dist = sqrt(d2);
  if (g) {
    const double q = dist / R, q2 = q * q, q6 = q2 * q2 * q2, q7 = q6 * q, q7p = q7 + 0.12;
    const double t = 1.07 / (q + 0.07), t2 = t * t, t7 = t2 * t2 * t2 * t;
    const double de = eps / R * t7 * (-1.12 * 7.0 * q6 / (q7p * q7p) + ((-1.12 * 7.0 / q7p + 14.0) / (q + 0.07)));
    for (int c = 0; c < 3; ++c) {
      const double v = dist <= 0.0 ? R * 0.01 : de * d[c] / dist;
      g[3 * i + c] += v;
      g[3 * j + c] -= v;
    }
  }
  const double R2 = R * R, R7 = R2 * R2 * R2 * R, dist7 = d2 * d2 * d2 * dist;
  const double t1 = 1.07 * R / (dist + 0.07 * R), t1sq = t1 * t1, t17 = t1sq * t1sq * t1sq * t1;
  return eps * t17 * (1.12 * R7 / (dist7 + 0.12 * R7) - 2.0);
}

static double mmff_ele(const double* p, int i, int j, double chargeTerm, int dielModel, int is14, double* g) {
  double d[3] = {p[3 * i] - p[3 * j], p[3 * i + 1] - p[3 * j + 1], p[3 * i + 2] - p[3 * j + 2]};
  const double dist = sqrt(dot3(d, d)), rb = dist + 0.05;
  if (g) {
    double num = -332.0716 * chargeTerm, den = rb * rb;
    if (dielModel == 2) {
      num *= 2.0;
      den *= rb;
    }
    double de = num / den;
    if (is14) de *= 0.75;
    for (int c = 0; c < 3; ++c) {
      const double v = de * d[c] / dist;
      g[3 * i + c] += v;
      g[3 * j + c] -= v;
    }
  }
  double e = 332.0716 * chargeTerm / (dielModel == 2 ? rb * rb : rb);
  if (is14) e *= 0.75;
  return e;
}

/* Energy of molecule `mol` at pos[nAtoms*3]; if grad != NULL the gradient is ACCUMULATED into it.
 * perType[7] (optional) receives the per-term-type energies (bond, angle, strbend, oop, torsion, vdw, ele). */
/* =========================================================================================== restraints
 * The four "constraint" contribs of RDKit's MMFF / UFF force fields as the reference evaluates them
 * (src/forcefields/mmff_kernels_device.cuh:673-1036): flat-bottomed distance (1/2 k), position (1/2 k beyond maxDispl), angle
 * and signed-dihedral windows in degrees (k, no 1/2; the dihedral window is periodic). */
static double norm_deg(double a) {
  a = fmod(a, 360.0);
  if (a < -180.0) a += 360.0;
  else if (a > 180.0) a -= 360.0;
  return a;
}
static double restraint_terms(const TermTable* D, const TermTable* P, const TermTable* A, const TermTable* T, int mol, const double* p,
                              double* g) {
  double e = 0.0;
  for (int t = D->starts[mol]; t < D->starts[mol + 1]; ++t) {
    const int    i = D->idx[2 * t], j = D->idx[2 * t + 1];
    const double mn = D->par[3 * t], mx = D->par[3 * t + 1], k = D->par[3 * t + 2];
    double       d[3] = {p[3 * i] - p[3 * j], p[3 * i + 1] - p[3 * j + 1], p[3 * i + 2] - p[3 * j + 2]};
    const double d2 = dot3(d, d);
    double       bound;
    if (d2 < mn * mn) bound = mn;
    else if (d2 > mx * mx) bound = mx;
    else continue;
    const double dist = sqrt(d2);
    e += 0.5 * k * (dist - bound) * (dist - bound);
    if (g)
      for (int c = 0; c < 3; ++c) {
        const double v = k * (dist - bound) / fmax(1.0e-8, dist) * d[c];
        g[3 * i + c] += v;
        g[3 * j + c] -= v;
      }
  }
  for (int t = P->starts[mol]; t < P->starts[mol + 1]; ++t) {
    const int     i = P->idx[t];
    const double* q = P->par + 5 * t;
    double        d[3] = {p[3 * i] - q[0], p[3 * i + 1] - q[1], p[3 * i + 2] - q[2]};
    const double  dist = sqrt(dot3(d, d)), over = dist - q[3];
    if (over <= 0.0) continue;
    e += 0.5 * q[4] * over * over;
    if (g)
      for (int c = 0; c < 3; ++c) g[3 * i + c] += over * q[4] / fmax(dist, 1.0e-8) * d[c];
  }
  for (int t = A->starts[mol]; t < A->starts[mol + 1]; ++t) {
    const int     i = A->idx[3 * t], j = A->idx[3 * t + 1], k = A->idx[3 * t + 2];
    const double* q = A->par + 3 * t;
    double        r1[3], r2[3], rp[3], c0[3], c1[3];
    for (int c = 0; c < 3; ++c) {
      r1[c] = p[3 * i + c] - p[3 * j + c];
      r2[c] = p[3 * k + c] - p[3 * j + c];
    }
    const double l1 = fmax(1.0e-5, dot3(r1, r1)), l2 = fmax(1.0e-5, dot3(r2, r2));
    const double ang = RAD2DEG * acos(clampd(dot3(r1, r2) / sqrt(l1 * l2), -1.0, 1.0));
    const double at = ang < q[0] ? ang - q[0] : (ang > q[1] ? ang - q[1] : 0.0);
    e += q[2] * at * at;
    if (!g || is_zero(at)) continue;
    cross(r2, r1, rp);
    const double pre = 2.0 * RAD2DEG * q[2] * at / fmax(1.0e-5, sqrt(dot3(rp, rp)));
    cross(r1, rp, c0);
    cross(r2, rp, c1);
    for (int c = 0; c < 3; ++c) {
      const double a = c0[c] * (-pre / l1), b = c1[c] * (pre / l2);
      g[3 * i + c] += a;
      g[3 * j + c] -= a + b;
      g[3 * k + c] += b;
    }
  }
  for (int t = T->starts[mol]; t < T->starts[mol + 1]; ++t) {
    const int16_t* ix = T->idx + 4 * t;
    const double*  q = T->par + 3 * t;
    const double * p1 = p + 3 * ix[0], *p2 = p + 3 * ix[1], *p3 = p + 3 * ix[2], *p4 = p + 3 * ix[3];
    double         r0[3], r1[3], r2[3], r3[3], tt0[3], tt1[3], t0[3], t1[3], m[3];
    for (int c = 0; c < 3; ++c) {
      r0[c] = p1[c] - p2[c];
      r1[c] = p3[c] - p2[c];
      r2[c] = -r1[c];
      r3[c] = p4[c] - p3[c];
    }
    cross(r0, r1, tt0);
    cross(r2, r3, tt1);
    const double d0 = fmax(sqrt(dot3(tt0, tt0)), 1.0e-5), d1 = fmax(sqrt(dot3(tt1, tt1)), 1.0e-5);
    for (int c = 0; c < 3; ++c) {
      t0[c] = tt0[c] / d0;
      t1[c] = tt1[c] / d1;
    }
    const double cosPhi = clampd(dot3(t0, t1), -1.0, 1.0);
    cross(t0, r1, m);
    const double phi = RAD2DEG * -atan2(dot3(m, t1) / fmax(sqrt(dot3(m, m)), 1.0e-5), cosPhi);
    double       target = phi;
    if (!(phi > q[0] && phi < q[1]) && !(phi > q[0] && q[0] > q[1]) && !(phi < q[1] && q[0] > q[1]))
      target = fabs(norm_deg(phi - q[0])) < fabs(norm_deg(phi - q[1])) ? q[0] : q[1];
    const double term = norm_deg(phi - target);
    e += q[2] * term * term;
    if (!g || is_zero(term)) continue;
    double d23[3] = {p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2]};
    const double pre = 2.0 * RAD2DEG * q[2] * term / fmax(sqrt(dot3(d23, d23)), 1.0e-8);
    double       tmp0[3], tmp1[3], dedt0[3], dedt1[3], r31[3], r42[3], a[3], b[3];
    cross(tt0, r2, tmp0);
    cross(tt1, r1, tmp1);
    const double n0 = fmax(dot3(tt0, tt0), 1.0e-8), n1 = fmax(dot3(tt1, tt1), 1.0e-8);
    for (int c = 0; c < 3; ++c) {
      dedt0[c] = tmp0[c] / n0 * pre;
      dedt1[c] = tmp1[c] / n1 * pre;
      r31[c] = p3[c] - p1[c];
      r42[c] = p4[c] - p2[c];
    }
    cross(r2, dedt0, a);
    for (int c = 0; c < 3; ++c) g[3 * ix[0] + c] += a[c];
    cross(r31, dedt0, a);
    cross(r3, dedt1, b);
    for (int c = 0; c < 3; ++c) g[3 * ix[1] + c] += a[c] - b[c];
    cross(r0, dedt0, a);
    cross(r42, dedt1, b);
    for (int c = 0; c < 3; ++c) g[3 * ix[2] + c] += a[c] + b[c];
    cross(r2, dedt1, a);
    for (int c = 0; c < 3; ++c) g[3 * ix[3] + c] += a[c];
  }
  return e;
}

double oracle_mmff_energy_grad(const MmffSystem* s, int mol, const double* pos, double* grad, double* perType) {
  double e[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int t = s->bond.starts[mol]; t < s->bond.starts[mol + 1]; ++t)
    e[0] += mmff_bond(pos, s->bond.idx[2 * t], s->bond.idx[2 * t + 1], s->bond.par[2 * t], s->bond.par[2 * t + 1], grad);
  for (int t = s->angle.starts[mol]; t < s->angle.starts[mol + 1]; ++t)
    e[1] += mmff_angle(pos, s->angle.idx[3 * t], s->angle.idx[3 * t + 1], s->angle.idx[3 * t + 2], s->angle.par[3 * t],
                       s->angle.par[3 * t + 1], s->angle.par[3 * t + 2] != 0.0, grad);
  for (int t = s->strbend.starts[mol]; t < s->strbend.starts[mol + 1]; ++t) {
    const double* q = s->strbend.par + 5 * t;
    e[2] += mmff_strbend(pos, s->strbend.idx[3 * t], s->strbend.idx[3 * t + 1], s->strbend.idx[3 * t + 2], q[0], q[1],
                         q[2], q[3], q[4], grad);
  }
  for (int t = s->oop.starts[mol]; t < s->oop.starts[mol + 1]; ++t)
    e[3] += mmff_oop(pos, s->oop.idx[4 * t], s->oop.idx[4 * t + 1], s->oop.idx[4 * t + 2], s->oop.idx[4 * t + 3],
                     s->oop.par[t], grad);
  for (int t = s->torsion.starts[mol]; t < s->torsion.starts[mol + 1]; ++t) {
    const double* q = s->torsion.par + 3 * t;
    e[4] += mmff_torsion(pos, s->torsion.idx[4 * t], s->torsion.idx[4 * t + 1], s->torsion.idx[4 * t + 2],
                         s->torsion.idx[4 * t + 3], q[0], q[1], q[2], grad);
  }
  for (int t = s->vdw.starts[mol]; t < s->vdw.starts[mol + 1]; ++t)
    e[5] += mmff_vdw(pos, s->vdw.idx[2 * t], s->vdw.idx[2 * t + 1], s->vdw.par[2 * t], s->vdw.par[2 * t + 1], grad);
  for (int t = s->ele.starts[mol]; t < s->ele.starts[mol + 1]; ++t)
    e[6] += mmff_ele(pos, s->ele.idx[2 * t], s->ele.idx[2 * t + 1], s->ele.par[3 * t], (int)s->ele.par[3 * t + 1],
                     s->ele.par[3 * t + 2] != 0.0, grad);
  if (perType) memcpy(perType, e, sizeof(e));
  return e[0] + e[1] + e[2] + e[3] + e[4] + e[5] + e[6] + restraint_terms(&s->distc, &s->posc, &s->anglec, &s->torsc, mol, pos, grad);
}

/* =========================================================================================== DG (dim 3 or 4) */
double oracle_dg_energy_grad(const DgSystem* s, int mol, int dim, double chiralWeight, double fourthWeight,
                             const double* pos, double* grad) {
  double e = 0.0;
  for (int t = s->dist.starts[mol]; t < s->dist.starts[mol + 1]; ++t) {
    const int    i = s->dist.idx[2 * t], j = s->dist.idx[2 * t + 1];
    const double lb2 = s->dist.par[3 * t], ub2 = s->dist.par[3 * t + 1], w = s->dist.par[3 * t + 2];
    double       d2 = 0.0;
    for (int c = 0; c < dim; ++c) d2 += (pos[i * dim + c] - pos[j * dim + c]) * (pos[i * dim + c] - pos[j * dim + c]);
    double val = 0.0, pre = 0.0;
    int    active = 0;
    if (d2 > ub2) {
      val    = d2 / ub2 - 1.0;
      pre    = 4.0 * (d2 / ub2 - 1.0) / ub2;
      active = 1;
    } else if (d2 < lb2) {
      val    = 2.0 * lb2 / (lb2 + d2) - 1.0;
      pre    = 8.0 * lb2 * (1.0 - 2.0 * lb2 / (d2 + lb2)) / ((d2 + lb2) * (d2 + lb2));
      active = 1;
    }
    if (val > 0.0) e += w * val * val;
    if (grad && active)
      for (int c = 0; c < dim; ++c) {
        const double v = w * pre * (pos[i * dim + c] - pos[j * dim + c]);
        grad[i * dim + c] += v;
        grad[j * dim + c] -= v;
      }
  }
  for (int t = s->chiral.starts[mol]; t < s->chiral.starts[mol + 1]; ++t) {
    const int16_t* ix = s->chiral.idx + 4 * t;
    const double   ub = s->chiral.par[2 * t], lb = s->chiral.par[2 * t + 1];
    const double * p1 = pos + ix[0] * dim, *p2 = pos + ix[1] * dim, *p3 = pos + ix[2] * dim, *p4 = pos + ix[3] * dim;
    double         v1[3], v2[3], v3[3], c23[3];
    for (int c = 0; c < 3; ++c) {
      v1[c] = p1[c] - p4[c];
      v2[c] = p2[c] - p4[c];
      v3[c] = p3[c] - p4[c];
    }
    cross(v2, v3, c23);
    const double vol = dot3(v1, c23);
    double       pre = 0.0;
    int          active = 0;
    if (vol < lb) {
      e += chiralWeight * (vol - lb) * (vol - lb);
      pre    = chiralWeight * (vol - lb);
      active = 1;
    } else if (vol > ub) {
      e += chiralWeight * (vol - ub) * (vol - ub);
      pre    = chiralWeight * (vol - ub);
      active = 1;
    }
    if (grad && active) { /* RDKit quirk kept: prefactor has no factor 2 */
      double* g1 = grad + ix[0] * dim, *g2 = grad + ix[1] * dim, *g3 = grad + ix[2] * dim, *g4 = grad + ix[3] * dim;
      g1[0] += pre * (v2[1] * v3[2] - v2[2] * v3[1]);
      g1[1] += pre * (v2[2] * v3[0] - v2[0] * v3[2]);
      g1[2] += pre * (v2[0] * v3[1] - v2[1] * v3[0]);
      g2[0] += pre * (v3[1] * v1[2] - v3[2] * v1[1]);
      g2[1] += pre * (v3[2] * v1[0] - v3[0] * v1[2]);
      g2[2] += pre * (v3[0] * v1[1] - v3[1] * v1[0]);
      g3[0] += pre * (v2[2] * v1[1] - v2[1] * v1[2]);
      g3[1] += pre * (v2[0] * v1[2] - v2[2] * v1[0]);
      g3[2] += pre * (v2[1] * v1[0] - v2[0] * v1[1]);
      g4[0] += pre * (p1[2] * (p2[1] - p3[1]) + p2[2] * (p3[1] - p1[1]) + p3[2] * (p1[1] - p2[1]));
      g4[1] += pre * (p1[0] * (p2[2] - p3[2]) + p2[0] * (p3[2] - p1[2]) + p3[0] * (p1[2] - p2[2]));
      g4[2] += pre * (p1[1] * (p2[0] - p3[0]) + p2[1] * (p3[0] - p1[0]) + p3[1] * (p1[0] - p2[0]));
    }
  }
  if (dim == 4) {
    for (int t = s->fourth.starts[mol]; t < s->fourth.starts[mol + 1]; ++t) {
      const int    a = s->fourth.idx[t];
      const double w4 = pos[a * 4 + 3];
      e += fourthWeight * w4 * w4;
      if (grad) grad[a * 4 + 3] += fourthWeight * w4; /* RDKit quirk kept: no factor 2 */
    }
  }
  return e;
}

/* =========================================================================================== ETK (4-D storage, xyz used) */
static double etk_cos_phi(const double* p1, const double* p2, const double* p3, const double* p4) {
  double r1[3], r2[3], r3[3], r4[3], t1[3], t2[3];
  for (int c = 0; c < 3; ++c) {
    r1[c] = p1[c] - p2[c];
    r2[c] = p3[c] - p2[c];
    r3[c] = p2[c] - p3[c];
    r4[c] = p4[c] - p3[c];
  }
  cross(r1, r2, t1);
  cross(r3, r4, t2);
  const double comb = dot3(t1, t1) * dot3(t2, t2);
  if (is_zero(comb)) return 0.0;
  return clampd(dot3(t1, t2) / sqrt(comb), -1.0, 1.0);
}

/* refPos (optional): reference geometry; the 1-2 / 1-3 windows whose `fixed` flag is 0 are re-centred on the reference
 * distance keeping their half-width (src/etkdg_stage_etk_minimization.cu:32-64). */
double oracle_etk_energy_grad_ref(const EtkSystem* s, int mol, const double* pos, double* grad, int plain,
                                  const double* refPos) {
  double e = 0.0;
  for (int t = s->torsion.starts[mol]; t < s->torsion.starts[mol + 1]; ++t) {
    const int16_t* ix = s->torsion.idx + 4 * t;
    const double*  fc = s->torsion.par + 12 * t; /* 6 force constants then 6 signs */
    const double*  sg = fc + 6;
    const double * p1 = pos + 4 * ix[0], *p2 = pos + 4 * ix[1], *p3 = pos + 4 * ix[2], *p4 = pos + 4 * ix[3];
    const double   c = etk_cos_phi(p1, p2, p3, p4);
    const double   c2 = c * c, c3 = c * c2, c4 = c * c3, c5 = c * c4, c6 = c * c5;
    const double   cosm[6] = {c, 2 * c2 - 1, 4 * c3 - 3 * c, 8 * c4 - 8 * c2 + 1, 16 * c5 - 20 * c3 + 5 * c,
                              32 * c6 - 48 * c4 + 18 * c2 - 1};
    for (int m = 0; m < 6; ++m) e += fc[m] * (1.0 + sg[m] * cosm[m]);
    if (grad) {
      double r1[3], r2[3], r3[3], r4[3], t0[3], t1[3];
      for (int k = 0; k < 3; ++k) {
        r1[k] = p1[k] - p2[k];
        r2[k] = p3[k] - p2[k];
        r3[k] = -r2[k];
        r4[k] = p4[k] - p3[k];
      }
      cross(r1, r2, t0);
      cross(r3, r4, t1);
      const double d02 = dot3(t0, t0), d12 = dot3(t1, t1);
      if (is_zero(d02) || is_zero(d12)) continue;
      const double i0 = 1.0 / sqrt(d02), i1 = 1.0 / sqrt(d12);
      for (int k = 0; k < 3; ++k) {
        t0[k] *= i0;
        t1[k] *= i1;
      }
      const double cp = clampd(dot3(t0, t1), -1.0, 1.0);
      const double sSq = 1.0 - cp * cp, sp = sSq > 0.0 ? sqrt(sSq) : 0.0;
      const double q2 = cp * cp, q3 = cp * q2, q4 = cp * q3, q5 = cp * q4;
      /* RDKit quirk kept: the 6-fold term uses fc[4]*sign[4] (dist_geom_kernels_device.cuh:519-525) */
      const double dE = (-fc[0] * sg[0] * sp - 2.0 * fc[1] * sg[1] * (2.0 * cp * sp) -
                         3.0 * fc[2] * sg[2] * (4.0 * q2 * sp - sp) - 4.0 * fc[3] * sg[3] * (8.0 * q3 * sp - 4.0 * cp * sp) -
                         5.0 * fc[4] * sg[4] * (16.0 * q4 * sp - 12.0 * q2 * sp + sp) -
                         6.0 * fc[4] * sg[4] * (32.0 * q5 * sp - 32.0 * q3 * sp + 6.0 * sp));
      const double sinTerm = -dE * (is_zero(sp) ? 1.0 / cp : 1.0 / sp);
      double       a[3], b[3];
      for (int k = 0; k < 3; ++k) {
        a[k] = i0 * (t1[k] - cp * t0[k]);
        b[k] = i1 * (t0[k] - cp * t1[k]);
      }
      double* g1 = grad + 4 * ix[0], *g2 = grad + 4 * ix[1], *g3 = grad + 4 * ix[2], *g4 = grad + 4 * ix[3];
      g1[0] += sinTerm * (a[2] * r2[1] - a[1] * r2[2]);
      g1[1] += sinTerm * (a[0] * r2[2] - a[2] * r2[0]);
      g1[2] += sinTerm * (a[1] * r2[0] - a[0] * r2[1]);
      g4[0] += sinTerm * (b[1] * r3[2] - b[2] * r3[1]);
      g4[1] += sinTerm * (b[2] * r3[0] - b[0] * r3[2]);
      g4[2] += sinTerm * (b[0] * r3[1] - b[1] * r3[0]);
      g2[0] += sinTerm * (a[1] * (r2[2] - r1[2]) + a[2] * (r1[1] - r2[1]) + b[1] * (-r4[2]) + b[2] * (r4[1]));
      g2[1] += sinTerm * (a[0] * (r1[2] - r2[2]) + a[2] * (r2[0] - r1[0]) + b[0] * (r4[2]) + b[2] * (-r4[0]));
      g2[2] += sinTerm * (a[0] * (r2[1] - r1[1]) + a[1] * (r1[0] - r2[0]) + b[0] * (-r4[1]) + b[1] * (r4[0]));
      g3[0] += sinTerm * (a[1] * r1[2] + a[2] * (-r1[1]) + b[1] * (r4[2] - r3[2]) + b[2] * (r3[1] - r4[1]));
      g3[1] += sinTerm * (a[0] * (-r1[2]) + a[2] * r1[0] + b[0] * (r3[2] - r4[2]) + b[2] * (r4[0] - r3[0]));
      g3[2] += sinTerm * (a[0] * r1[1] + a[1] * (-r1[0]) + b[0] * (r4[1] - r3[1]) + b[1] * (r3[0] - r4[0]));
    }
  }
  if (!plain) {
    for (int t = s->improper.starts[mol]; t < s->improper.starts[mol + 1]; ++t) {
      const int16_t* ix = s->improper.idx + 4 * t;
      const double   C0 = s->improper.par[4 * t], C1 = s->improper.par[4 * t + 1], C2 = s->improper.par[4 * t + 2],
                   fk = s->improper.par[4 * t + 3];
      const double * p1 = pos + 4 * ix[0], *p2 = pos + 4 * ix[1], *p3 = pos + 4 * ix[2], *p4 = pos + 4 * ix[3];
      double         ji[3], jk[3], jl[3];
      for (int k = 0; k < 3; ++k) {
        ji[k] = p1[k] - p2[k];
        jk[k] = p3[k] - p2[k];
        jl[k] = p4[k] - p2[k];
      }
      const double l2i = dot3(ji, ji), l2k = dot3(jk, jk), l2l = dot3(jl, jl);
      double       cosY = 0.0;
      if (!(l2i < 1.0e-16 || l2k < 1.0e-16 || l2l < 1.0e-16)) {
        double n[3];
        cross(ji, jk, n);
        const double nf = 1.0 / sqrt(l2i * l2k);
        for (int k = 0; k < 3; ++k) n[k] *= nf;
        const double l2n = dot3(n, n);
        if (!(l2n < 1.0e-16)) cosY = dot3(n, jl) / sqrt(l2l) / sqrt(l2n);
      }
      const double sSq = 1.0 - cosY * cosY, sinY = sSq > 0.0 ? sqrt(sSq) : 0.0;
      e += fk * (C0 + C1 * sinY + C2 * (2.0 * sinY * sinY - 1.0));
      if (grad) {
        if (is_zero(l2i) || is_zero(l2k) || is_zero(l2l)) continue;
        const double ii = 1.0 / sqrt(l2i), ik = 1.0 / sqrt(l2k), il = 1.0 / sqrt(l2l);
        double       a[3], b[3], c[3], na[3], n[3];
        for (int k = 0; k < 3; ++k) {
          a[k]  = ji[k] * ii;
          b[k]  = jk[k] * ik;
          c[k]  = jl[k] * il;
          na[k] = -a[k];
        }
        cross(na, b, n);
        const double inl = 1.0 / sqrt(dot3(n, n));
        for (int k = 0; k < 3; ++k) n[k] *= inl;
        const double cY = clampd(dot3(n, c), -1.0, 1.0);
        const double sY = fmax(sqrt(1.0 - cY * cY), 1.0e-8);
        const double cT = clampd(dot3(a, b), -1.0, 1.0);
        const double sTsq = 1.0 - cT * cT, sT = fmax(sqrt(sTsq), 1.0e-8);
        const double dE = -fk * (C1 * cY - 4.0 * C2 * cY * sY);
        double       t1[3], t2[3], t3[3];
        cross(c, b, t1);
        cross(a, c, t2);
        cross(b, a, t3);
        const double inv1 = 1.0 / (sY * sT), term2 = cY / (sY * sTsq), cOs = cY / sY;
        for (int k = 0; k < 3; ++k) {
          const double g1 = (t1[k] * inv1 - (a[k] - b[k] * cT) * term2) * ii;
          const double g3 = (t2[k] * inv1 - (b[k] - a[k] * cT) * term2) * ik;
          const double g4 = (t3[k] * inv1 - c[k] * cOs) * il;
          grad[4 * ix[0] + k] += dE * g1;
          grad[4 * ix[1] + k] += -dE * (g1 + g3 + g4);
          grad[4 * ix[2] + k] += dE * g3;
          grad[4 * ix[3] + k] += dE * g4;
        }
      }
    }
  }
  const TermTable* dts[3] = {&s->dist12, &s->dist13, &s->longrange};
  for (int q = 0; q < 3; ++q) {
    const TermTable* T = dts[q];
    const int        P = q < 2 ? 4 : 3;
    for (int t = T->starts[mol]; t < T->starts[mol + 1]; ++t) {
      const int    i = T->idx[2 * t], j = T->idx[2 * t + 1];
      double       mn = T->par[P * t], mx = T->par[P * t + 1];
      const double fk = T->par[P * t + 2];
      if (P == 4 && refPos && T->par[P * t + 3] == 0.0) {
        double r2 = 0.0;
        for (int c = 0; c < 3; ++c) r2 += (refPos[4 * i + c] - refPos[4 * j + c]) * (refPos[4 * i + c] - refPos[4 * j + c]);
        const double dref = sqrt(r2), half = (mx - mn) / 2.0;
        mn = dref - half;
        mx = dref + half;
      }
      double       d2 = 0.0;
      for (int c = 0; c < 3; ++c) d2 += (pos[4 * i + c] - pos[4 * j + c]) * (pos[4 * i + c] - pos[4 * j + c]);
      double diff, pre;
      if (d2 < mn * mn) {
        const double d = sqrt(d2);
        diff           = mn - d;
        pre            = fk * (d - mn) / fmax(1.0e-8, d);
      } else if (d2 > mx * mx) {
        const double d = sqrt(d2);
        diff           = d - mx;
        pre            = fk * (d - mx) / fmax(1.0e-8, d);
      } else {
        continue;
      }
      e += 0.5 * fk * diff * diff;
      if (grad)
        for (int c = 0; c < 3; ++c) {
          const double v = pre * (pos[4 * i + c] - pos[4 * j + c]);
          grad[4 * i + c] += v;
          grad[4 * j + c] -= v;
        }
    }
  }
  for (int t = s->angle13.starts[mol]; t < s->angle13.starts[mol + 1]; ++t) {
    const int16_t* ix = s->angle13.idx + 3 * t;
    const double   mn = s->angle13.par[2 * t], mx = s->angle13.par[2 * t + 1], fk = 1.0;
    const double * p1 = pos + 4 * ix[0], *p2 = pos + 4 * ix[1], *p3 = pos + 4 * ix[2];
    double         r1[3], r2[3];
    for (int c = 0; c < 3; ++c) {
      r1[c] = p1[c] - p2[c];
      r2[c] = p3[c] - p2[c];
    }
    const double l1 = dot3(r1, r1), l2 = dot3(r2, r2);
    if (!is_zero(l1 * l2)) {
      const double ang = RAD2DEG * acos(clampd(dot3(r1, r2) / sqrt(l1 * l2), -1.0, 1.0));
      const double at  = ang < mn ? ang - mn : (ang > mx ? ang - mx : 0.0);
      e += fk * at * at;
    }
    if (grad) {
      const double m1 = fmax(1.0e-5, l1), m2 = fmax(1.0e-5, l2);
      const double cT = clampd(dot3(r1, r2) / sqrt(m1 * m2), -1.0, 1.0);
      const double ang = RAD2DEG * acos(cT);
      const double at  = ang < mn ? ang - mn : (ang > mx ? ang - mx : 0.0);
      const double dE  = 2.0 * RAD2DEG * fk * at;
      double       rp[3], e1[3], e3[3];
      cross(r2, r1, rp);
      const double pre = dE / sqrt(fmax(dot3(rp, rp), 1.0e-10));
      cross(r1, rp, e1);
      cross(r2, rp, e3);
      for (int c = 0; c < 3; ++c) {
        const double a = e1[c] * (-pre / m1), b = e3[c] * (pre / m2);
        grad[4 * ix[0] + c] += a;
        grad[4 * ix[1] + c] += -(a + b);
        grad[4 * ix[2] + c] += b;
      }
    }
  }
  return e;
}

double oracle_etk_energy_grad(const EtkSystem* s, int mol, const double* pos, double* grad, int plain) {
  return oracle_etk_energy_grad_ref(s, mol, pos, grad, plain, NULL);
}

/* =========================================================================================== test potential
 * E = sum_i w_i (x_i - c_i)^p  (tests/test_bfgs_minimizer.cu:822-930 uses p = 4, c_i = i; harmonic systems p = 2) */
typedef struct {
  int           n, power;
  const double *w, *c;
} PolySystem;
static double poly_energy_grad(const PolySystem* s, const double* x, double* g) {
  double e = 0.0;
  for (int i = 0; i < s->n; ++i) {
    const double d = x[i] - s->c[i];
    if (s->power == 2) {
      e += s->w[i] * d * d;
      if (g) g[i] += 2.0 * s->w[i] * d;
    } else {
      e += s->w[i] * d * d * d * d;
      if (g) g[i] += 4.0 * s->w[i] * d * d * d;
    }
  }
  return e;
}

/* =========================================================================================== BFGS */
typedef double (*EnergyGradFn)(const void* ctx, const double* x, double* gradOrNull);

static double scale_grad(int n, double* g, int scaleGrads) {
  double gradScale = scaleGrads ? 0.1 : 1.0, maxGrad = 0.0;
  for (int i = 0; i < n; ++i) {
    if (scaleGrads) g[i] *= gradScale;
    if (fabs(g[i]) > maxGrad) maxGrad = fabs(g[i]);
  }
  if (scaleGrads && maxGrad > 10.0) {
    while (maxGrad * gradScale > 10.0) gradScale *= 0.5;
    for (int i = 0; i < n; ++i) g[i] *= gradScale;
  }
  return gradScale;
}

/* Returns 0 when converged, 1 when maxIters ran out. *energyOut = energy re-evaluated at the final point. */
static int bfgs_minimize(int n, double* pos, EnergyGradFn fn, const void* ctx, int maxIters, double gradTol,
                         int scaleGrads, double* energyOut, int* itersOut) {
  const double FUNCTOL = 1e-4, MOVETOL = 1e-7, TOLX = 4. * 3e-8, EPS = 3e-8;
  double *     grad = calloc(n, 8), *dir = calloc(n, 8), *newPos = calloc(n, 8), *dGrad = calloc(n, 8),
         *hdg = calloc(n, 8), *H = calloc((size_t)n * n, 8);
  for (int i = 0; i < n; ++i) H[(size_t)i * n + i] = 1.0;
  double fp = fn(ctx, pos, NULL);
  memset(grad, 0, 8 * n);
  fn(ctx, pos, grad);
  double gradScale = scale_grad(n, grad, scaleGrads);
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    dir[i] = -grad[i];
    sum += pos[i] * pos[i];
  }
  const double maxStep = 100.0 * fmax(sqrt(sum), (double)n);
  int          status = 1, iter = 0;
  for (iter = 0; iter < maxIters; ++iter) {
    /* ---- line search ---- */
    double dsum = 0.0;
    for (int i = 0; i < n; ++i) dsum += dir[i] * dir[i];
    dsum = sqrt(dsum);
    if (dsum > maxStep)
      for (int i = 0; i < n; ++i) dir[i] *= maxStep / dsum;
    double slope = 0.0, test = 0.0;
    for (int i = 0; i < n; ++i) {
      slope += dir[i] * grad[i];
      const double t = fabs(dir[i]) / fmax(fabs(pos[i]), 1.0);
      if (t > test) test = t;
    }
    const double lambdaMin = MOVETOL / test;
    double       lambda = 1.0, lambda2 = 0.0, val2 = 0.0, newVal = fp;
    int          accepted = 0;
    for (int it = 0; it < 1000; ++it) {
      if (lambda < lambdaMin) break;
      for (int i = 0; i < n; ++i) newPos[i] = pos[i] + lambda * dir[i];
      newVal = fn(ctx, newPos, NULL);
      if (newVal - fp <= FUNCTOL * lambda * slope) {
        accepted = 1;
        break;
      }
      double tmp;
      if (it == 0) {
        tmp = -slope / (2.0 * (newVal - fp - slope));
      } else {
        const double rhs1 = newVal - fp - lambda * slope, rhs2 = val2 - fp - lambda2 * slope;
        const double a = (rhs1 / (lambda * lambda) - rhs2 / (lambda2 * lambda2)) / (lambda - lambda2);
        const double b = (-lambda2 * rhs1 / (lambda * lambda) + lambda * rhs2 / (lambda2 * lambda2)) / (lambda - lambda2);
        if (a == 0.0) {
          tmp = -slope / (2.0 * b);
        } else {
          const double disc = b * b - 3 * a * slope;
          if (disc < 0.0) tmp = 0.5 * lambda;
          else if (b <= 0.0) tmp = (-b + sqrt(disc)) / (3.0 * a);
          else tmp = -slope / (b + sqrt(disc));
        }
        if (tmp > 0.5 * lambda) tmp = 0.5 * lambda;
      }
      lambda2 = lambda;
      val2    = newVal;
      lambda  = fmax(tmp, 0.1 * lambda);
    }
    if (!accepted) memcpy(newPos, pos, 8 * n); /* nothing was done */
    fp = newVal;
    /* ---- direction + TOLX ---- */
    test = 0.0;
    for (int i = 0; i < n; ++i) {
      dir[i] = newPos[i] - pos[i];
      pos[i] = newPos[i];
      const double t = fabs(dir[i]) / fmax(fabs(pos[i]), 1.0);
      if (t > test) test = t;
      dGrad[i] = grad[i];
    }
    if (test < TOLX) {
      status = 0;
      break;
    }
    memset(grad, 0, 8 * n);
    fn(ctx, pos, grad);
    gradScale = scale_grad(n, grad, scaleGrads);
    test      = 0.0;
    const double term = fmax(fp * gradScale, 1.0);
    for (int i = 0; i < n; ++i) {
      const double t = fabs(grad[i]) * fmax(fabs(pos[i]), 1.0);
      if (t > test) test = t;
      dGrad[i] = grad[i] - dGrad[i];
    }
    if (test / term < gradTol) {
      status = 0;
      break;
    }
    /* ---- inverse Hessian ---- */
    double fac = 0, fae = 0, sumDGrad = 0, sumXi = 0;
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      for (int j = 0; j < n; ++j) a += H[(size_t)i * n + j] * dGrad[j];
      hdg[i] = a;
    }
    for (int i = 0; i < n; ++i) {
      fac += dGrad[i] * dir[i];
      fae += dGrad[i] * hdg[i];
      sumDGrad += dGrad[i] * dGrad[i];
      sumXi += dir[i] * dir[i];
    }
    if (fac > sqrt(EPS * sumDGrad * sumXi)) {
      fac              = 1.0 / fac;
      const double fad = 1.0 / fae;
      for (int i = 0; i < n; ++i) dGrad[i] = fac * dir[i] - fad * hdg[i];
      for (int i = 0; i < n; ++i) {
        const double pxi = fac * dir[i], hdgi = fad * hdg[i], dgi = fae * dGrad[i];
        for (int j = 0; j < n; ++j) H[(size_t)i * n + j] += pxi * dir[j] - hdgi * hdg[j] + dgi * dGrad[j];
      }
    }
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      for (int j = 0; j < n; ++j) a += H[(size_t)i * n + j] * grad[j];
      newPos[i] = -a;
    }
    memcpy(dir, newPos, 8 * n);
  }
  if (energyOut) *energyOut = fn(ctx, pos, NULL);
  if (itersOut) *itersOut = iter;
  free(grad);
  free(dir);
  free(newPos);
  free(dGrad);
  free(hdg);
  free(H);
  return status;
}

typedef struct {
  const MmffSystem* s;
  int               mol;
} MmffCtx;
static double mmff_fn(const void* c, const double* x, double* g) {
  const MmffCtx* m = (const MmffCtx*)c;
  return oracle_mmff_energy_grad(m->s, m->mol, x, g, NULL);
}
typedef struct {
  const DgSystem* s;
  int             mol, dim;
  double          cw, fw;
} DgCtx;
static double dg_fn(const void* c, const double* x, double* g) {
  const DgCtx* m = (const DgCtx*)c;
  return oracle_dg_energy_grad(m->s, m->mol, m->dim, m->cw, m->fw, x, g);
}
typedef struct {
  const EtkSystem* s;
  int              mol, plain;
  const double*    ref;
} EtkCtx;
static double etk_fn(const void* c, const double* x, double* g) {
  const EtkCtx* m = (const EtkCtx*)c;
  return oracle_etk_energy_grad_ref(m->s, m->mol, x, g, m->plain, m->ref);
}
static double poly_fn(const void* c, const double* x, double* g) { return poly_energy_grad((const PolySystem*)c, x, g); }

/* Batch drivers: conformer c uses molecule confMol[c], coordinates pos[confAtomStart[c]*dim ...]. OpenMP over conformers. */
void oracle_mmff_minimize(const MmffSystem* s, int nConf, const int32_t* confMol, const int32_t* confAtomStart,
                          double* pos, int maxIters, double gradTol, double* energies, int8_t* converged,
                          int32_t* iters) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < nConf; ++c) {
    MmffCtx ctx = {s, confMol[c]};
    int     it  = 0;
    const int st = bfgs_minimize(3 * s->atomCounts[confMol[c]], pos + 3 * (size_t)confAtomStart[c], mmff_fn, &ctx,
                                 maxIters, gradTol, 1, &energies[c], &it);
    if (converged) converged[c] = st == 0;
    if (iters) iters[c] = it;
  }
}
void oracle_dg_minimize(const DgSystem* s, int dim, double chiralWeight, double fourthWeight, int nConf,
                        const int32_t* confMol, const int32_t* confAtomStart, double* pos, int maxIters,
                        double gradTol, double* energies, int8_t* converged, int32_t* iters) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < nConf; ++c) {
    DgCtx ctx = {s, confMol[c], dim, chiralWeight, fourthWeight};
    int   it  = 0;
    const int st = bfgs_minimize(dim * s->atomCounts[confMol[c]], pos + (size_t)dim * confAtomStart[c], dg_fn, &ctx,
                                 maxIters, gradTol, 1, &energies[c], &it);
    if (converged) converged[c] = st == 0;
    if (iters) iters[c] = it;
  }
}
void oracle_etk_minimize(const EtkSystem* s, int plain, int recentre, int nConf, const int32_t* confMol,
                         const int32_t* confAtomStart, double* pos, int maxIters, double gradTol, double* energies,
                         int8_t* converged, int32_t* iters) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < nConf; ++c) {
    const int nn  = 4 * s->atomCounts[confMol[c]];
    double*   ref = NULL;
    if (recentre) {
      ref = (double*)malloc(sizeof(double) * nn);
      memcpy(ref, pos + 4 * (size_t)confAtomStart[c], sizeof(double) * nn);
    }
    EtkCtx ctx = {s, confMol[c], plain, ref};
    int    it  = 0;
    const int st = bfgs_minimize(nn, pos + 4 * (size_t)confAtomStart[c], etk_fn, &ctx, maxIters, gradTol, 1, &energies[c], &it);
    free(ref);
    if (converged) converged[c] = st == 0;
    if (iters) iters[c] = it;
  }
}
int oracle_poly_minimize(int n, int power, const double* w, const double* c, double* x, int maxIters, double gradTol,
                         int scaleGrads, double* energy, int32_t* iters) {
  PolySystem s = {n, power, w, c};
  int        it = 0;
  const int  st = bfgs_minimize(n, x, poly_fn, &s, maxIters, gradTol, scaleGrads, energy, &it);
  if (iters) *iters = it;
  return st;
}
double oracle_poly_energy_grad(int n, int power, const double* w, const double* c, const double* x, double* grad) {
  PolySystem s = {n, power, w, c};
  return poly_energy_grad(&s, x, grad);
}

/* Single-conformer entry points used by oracle_etkdg.c. maxRestarts > 0: re-run while unconverged (repeatUntilConverged). */
int oracle_dg_minimize_one(const DgSystem* s, int mol, int dim, double cw, double fw, double* pos, int maxIters,
                           double gradTol, int maxRestarts, double* energy) {
  DgCtx ctx = {s, mol, dim, cw, fw};
  int   st  = 1;
  for (int r = 0;; ++r) {
    st = bfgs_minimize(dim * s->atomCounts[mol], pos, dg_fn, &ctx, maxIters, gradTol, 1, energy, NULL);
    if (st == 0 || r >= maxRestarts) break;
  }
  return st;
}
int oracle_etk_minimize_one(const EtkSystem* s, int mol, int plain, const double* ref, double* pos, int maxIters,
                            double gradTol, double* energy) {
  EtkCtx ctx = {s, mol, plain, ref};
  return bfgs_minimize(4 * s->atomCounts[mol], pos, etk_fn, &ctx, maxIters, gradTol, 1, energy, NULL);
}

/* =========================================================================================== UFF
 * src/forcefields/uff_kernels_device.cuh:37-590 (RDKit ForceFields::UFF contribs), fp64. */
typedef struct {
  int32_t        nMols;
  const int32_t* atomCounts;
  TermTable      bond, angle, torsion, inversion, vdw;
  TermTable      distc, posc, anglec, torsc; /* restraints */
} UffSystem;

double oracle_uff_energy_grad(const UffSystem* s, int mol, const double* p, double* g) {
  double e = 0.0;
  for (int t = s->bond.starts[mol]; t < s->bond.starts[mol + 1]; ++t) {
    const int    i = s->bond.idx[2 * t], j = s->bond.idx[2 * t + 1];
    const double r0 = s->bond.par[2 * t], k = s->bond.par[2 * t + 1];
    double       d[3] = {p[3 * i] - p[3 * j], p[3 * i + 1] - p[3 * j + 1], p[3 * i + 2] - p[3 * j + 2]};
    const double dist = sqrt(dot3(d, d));
    e += 0.5 * k * (dist - r0) * (dist - r0);
    if (g)
      for (int c = 0; c < 3; ++c) {
        const double v = dist > 0.0 ? k * (dist - r0) * d[c] / dist : k * 0.01;
        g[3 * i + c] += v;
        g[3 * j + c] -= v;
      }
  }
  for (int t = s->angle.starts[mol]; t < s->angle.starts[mol + 1]; ++t) {
    const int     i = s->angle.idx[3 * t], j = s->angle.idx[3 * t + 1], k = s->angle.idx[3 * t + 2];
    const double* q = s->angle.par + 6 * t;
    const int     order = (int)q[2];
    double        d1[3], d2[3];
    for (int c = 0; c < 3; ++c) {
      d1[c] = p[3 * i + c] - p[3 * j + c];
      d2[c] = p[3 * k + c] - p[3 * j + c];
    }
    const double l1sq = dot3(d1, d1), l2sq = dot3(d2, d2);
    if (l1sq <= 0.0 || l2sq <= 0.0) continue;
    const double l1 = sqrt(l1sq), l2 = sqrt(l2sq);
    const double c = clampd(dot3(d1, d2) / (l1 * l2), -1.0, 1.0), sSq = 1.0 - c * c, c2t = c * c - sSq;
    const int    corr = order > 0 && order < 5 && c > 0.8660;
    double       term;
    if (order == 0) {
      term = q[3] + q[4] * c + q[5] * c2t;
    } else {
      double r = 0.0;
      if (order == 1) r = -c;
      else if (order == 2) r = c2t;
      else if (order == 3) r = c * (c * c - 3.0 * sSq);
      else if (order == 4) r = c * c * c * c - 6.0 * c * c * sSq + sSq * sSq;
      term = (1.0 - r) / (double)(order * order);
    }
    e += q[1] * term;
    if (corr) e += exp(-20.0 * (acos(c) - q[0] + 0.25));
    if (g && !is_zero(sSq)) {
      const double sn = fmax(sqrt(sSq), 1.0e-8), s2t = 2.0 * sn * c;
      double       dE;
      if (order == 0) {
        dE = -q[1] * (q[4] * sn + 2.0 * q[5] * s2t);
      } else {
        double r = 0.0;
        if (order == 1) r = -sn;
        else if (order == 2) r = s2t;
        else if (order == 3) r = sn * (3.0 - 4.0 * sn * sn);
        else if (order == 4) r = c * sn * (4.0 - 8.0 * sn * sn);
        dE = (order >= 1 && order <= 4) ? r * q[1] / (double)order : 0.0;
      }
      if (corr) dE += -20.0 * exp(-20.0 * (acos(c) - q[0] + 0.25));
      const double cf = dE / (-sn);
      for (int x = 0; x < 3; ++x) {
        const double n1 = d1[x] / l1, n2 = d2[x] / l2;
        const double a = (n2 - c * n1) / l1, b = (n1 - c * n2) / l2;
        g[3 * i + x] += cf * a;
        g[3 * j + x] += cf * (-a - b);
        g[3 * k + x] += cf * b;
      }
    }
  }
  for (int t = s->torsion.starts[mol]; t < s->torsion.starts[mol + 1]; ++t) {
    const int16_t* ix = s->torsion.idx + 4 * t;
    const double   fk = s->torsion.par[3 * t], cosTerm = s->torsion.par[3 * t + 2];
    const int      order = (int)s->torsion.par[3 * t + 1];
    double         r0[3], r1[3], r2[3], r3[3], t0[3], t1[3];
    for (int c = 0; c < 3; ++c) {
      r0[c] = p[3 * ix[0] + c] - p[3 * ix[1] + c];
      r1[c] = p[3 * ix[2] + c] - p[3 * ix[1] + c];
      r2[c] = -r1[c];
      r3[c] = p[3 * ix[3] + c] - p[3 * ix[2] + c];
    }
    cross(r0, r1, t0);
    cross(r2, r3, t1);
    const double d0 = sqrt(dot3(t0, t0)), d1 = sqrt(dot3(t1, t1));
    if (order != 2 && order != 3 && order != 6) continue;
    {
      const double c = (is_zero(d0) || is_zero(d1)) ? 0.0 : clampd(dot3(t0, t1) / (d0 * d1), -1.0, 1.0);
      const double sSq = 1.0 - c * c;
      double       cn;
      if (order == 2) cn = 1.0 - 2.0 * sSq;
      else if (order == 3) cn = c * (c * c - 3.0 * sSq);
      else cn = 1.0 + sSq * (-32.0 * sSq * sSq + 48.0 * sSq - 18.0);
      e += fk / 2.0 * (1.0 - cosTerm * cn);
    }
    if (g && !(is_zero(d0) || is_zero(d1))) {
      for (int c = 0; c < 3; ++c) {
        t0[c] /= d0;
        t1[c] /= d1;
      }
      const double c = clampd(dot3(t0, t1), -1.0, 1.0), sSq = 1.0 - c * c, sn = sSq > 0.0 ? sqrt(sSq) : 0.0;
      double       r;
      if (order == 2) r = 2.0 * sn * c;
      else if (order == 3) r = sn * (3.0 - 4.0 * sSq);
      else r = c * sn * (32.0 * sSq * (sSq - 1.0) + 6.0);
      const double dE = r * fk / 2.0 * cosTerm * -1.0 * (double)order;
      const double sinTerm = dE * (is_zero(sn) ? (1.0 / fmax(fabs(c), 1.0e-8)) : (1.0 / sn));
      double       a[3], b[3];
      for (int x = 0; x < 3; ++x) {
        a[x] = (t1[x] - c * t0[x]) / d0;
        b[x] = (t0[x] - c * t1[x]) / d1;
      }
      double *g1 = g + 3 * ix[0], *g2 = g + 3 * ix[1], *g3 = g + 3 * ix[2], *g4 = g + 3 * ix[3];
      g1[0] += sinTerm * (a[2] * r1[1] - a[1] * r1[2]);
      g1[1] += sinTerm * (a[0] * r1[2] - a[2] * r1[0]);
      g1[2] += sinTerm * (a[1] * r1[0] - a[0] * r1[1]);
      g2[0] += sinTerm * (a[1] * (r1[2] - r0[2]) + a[2] * (r0[1] - r1[1]) + b[1] * (-r3[2]) + b[2] * (r3[1]));
      g2[1] += sinTerm * (a[0] * (r0[2] - r1[2]) + a[2] * (r1[0] - r0[0]) + b[0] * (r3[2]) + b[2] * (-r3[0]));
      g2[2] += sinTerm * (a[0] * (r1[1] - r0[1]) + a[1] * (r0[0] - r1[0]) + b[0] * (-r3[1]) + b[1] * (r3[0]));
      g3[0] += sinTerm * (a[1] * r0[2] + a[2] * (-r0[1]) + b[1] * (r3[2] - r2[2]) + b[2] * (r2[1] - r3[1]));
      g3[1] += sinTerm * (a[0] * (-r0[2]) + a[2] * r0[0] + b[0] * (r2[2] - r3[2]) + b[2] * (r3[0] - r2[0]));
      g3[2] += sinTerm * (a[0] * r0[1] + a[1] * (-r0[0]) + b[0] * (r3[1] - r2[1]) + b[1] * (r2[0] - r3[0]));
      g4[0] += sinTerm * (b[1] * r2[2] - b[2] * r2[1]);
      g4[1] += sinTerm * (b[2] * r2[0] - b[0] * r2[2]);
      g4[2] += sinTerm * (b[0] * r2[1] - b[1] * r2[0]);
    }
  }
  for (int t = s->inversion.starts[mol]; t < s->inversion.starts[mol + 1]; ++t) {
    const int16_t* ix = s->inversion.idx + 4 * t;
    const double   fk = s->inversion.par[4 * t], C0 = s->inversion.par[4 * t + 1], C1 = s->inversion.par[4 * t + 2],
                 C2 = s->inversion.par[4 * t + 3];
    double ji[3], jk[3], jl[3];
    for (int c = 0; c < 3; ++c) {
      ji[c] = p[3 * ix[0] + c] - p[3 * ix[1] + c];
      jk[c] = p[3 * ix[2] + c] - p[3 * ix[1] + c];
      jl[c] = p[3 * ix[3] + c] - p[3 * ix[1] + c];
    }
    const double l2i = dot3(ji, ji), l2k = dot3(jk, jk), l2l = dot3(jl, jl);
    double       cosY = 0.0;
    if (!(l2i < 1.0e-16 || l2k < 1.0e-16 || l2l < 1.0e-16)) {
      double n[3];
      cross(ji, jk, n);
      const double sc = sqrt(l2i) * sqrt(l2k);
      for (int c = 0; c < 3; ++c) n[c] /= sc;
      const double l2n = dot3(n, n);
      if (!(l2n < 1.0e-16)) cosY = dot3(n, jl) / (sqrt(l2l) * sqrt(l2n));
    }
    const double sSq = 1.0 - cosY * cosY, sinY = sSq > 0.0 ? sqrt(sSq) : 0.0;
    e += fk * (C0 + C1 * sinY + C2 * (2.0 * sinY * sinY - 1.0));
    if (g) {
      const double dI = sqrt(l2i), dK = sqrt(l2k), dL = sqrt(l2l);
      if (is_zero(dI) || is_zero(dK) || is_zero(dL)) continue;
      double a[3], b[3], c[3], na[3], n[3];
      for (int x = 0; x < 3; ++x) {
        a[x]  = ji[x] / dI;
        b[x]  = jk[x] / dK;
        c[x]  = jl[x] / dL;
        na[x] = -a[x];
      }
      cross(na, b, n);
      const double nn = sqrt(dot3(n, n));
      if (nn <= 0.0) continue;
      for (int x = 0; x < 3; ++x) n[x] /= nn;
      const double cY = clampd(dot3(n, c), -1.0, 1.0), sY = fmax(sqrt(1.0 - cY * cY), 1.0e-8);
      const double cT = clampd(dot3(a, b), -1.0, 1.0), sTsq = 1.0 - cT * cT, sT = fmax(sqrt(sTsq), 1.0e-8);
      const double dE = -fk * (C1 * cY - 4.0 * C2 * cY * sY);
      double       t1[3], t2[3], t3[3];
      cross(c, b, t1);
      cross(a, c, t2);
      cross(b, a, t3);
      const double term1 = sY * sT, term2 = cY / (sY * sTsq);
      for (int x = 0; x < 3; ++x) {
        const double g1 = (t1[x] / term1 - (a[x] - b[x] * cT) * term2) / dI;
        const double g3 = (t2[x] / term1 - (b[x] - a[x] * cT) * term2) / dK;
        const double g4 = (t3[x] / term1 - c[x] * cY / sY) / dL;
        g[3 * ix[0] + x] += dE * g1;
        g[3 * ix[1] + x] += -dE * (g1 + g3 + g4);
        g[3 * ix[2] + x] += dE * g3;
        g[3 * ix[3] + x] += dE * g4;
      }
    }
  }
  for (int t = s->vdw.starts[mol]; t < s->vdw.starts[mol + 1]; ++t) {
    const int    i = s->vdw.idx[2 * t], j = s->vdw.idx[2 * t + 1];
    const double x = s->vdw.par[3 * t], eps = s->vdw.par[3 * t + 1], thr = s->vdw.par[3 * t + 2];
    double       d[3] = {p[3 * i] - p[3 * j], p[3 * i + 1] - p[3 * j + 1], p[3 * i + 2] - p[3 * j + 2]};
    const double dist = sqrt(dot3(d, d));
    if (dist > thr) continue;
    if (dist > 0.0) {
      const double r = x / dist, r2 = r * r, r6 = r2 * r2 * r2;
      e += eps * (r6 * r6 - 2.0 * r6);
    }
    if (g) {
      for (int c = 0; c < 3; ++c) {
        double v = 100.0;
        if (dist > 0.0) {
          const double r = x / dist, r2 = r * r, r7 = r * r2 * r2 * r2, r13 = r7 * r2 * r2 * r2;
          v = 12.0 * eps / x * (r7 - r13) * d[c] / dist;
        }
        g[3 * i + c] += v;
        g[3 * j + c] -= v;
      }
    }
  }
  e += restraint_terms(&s->distc, &s->posc, &s->anglec, &s->torsc, mol, p, g);
  return e;
}

typedef struct {
  const UffSystem* s;
  int              mol;
} UffCtx;
static double uff_fn(const void* c, const double* x, double* g) {
  const UffCtx* m = (const UffCtx*)c;
  return oracle_uff_energy_grad(m->s, m->mol, x, g);
}
void oracle_uff_minimize(const UffSystem* s, int nConf, const int32_t* confMol, const int32_t* confAtomStart,
                         double* pos, int maxIters, double gradTol, double* energies, int8_t* converged,
                         int32_t* iters) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < nConf; ++c) {
    UffCtx ctx = {s, confMol[c]};
    int    it  = 0;
    const int st = bfgs_minimize(3 * s->atomCounts[confMol[c]], pos + 3 * (size_t)confAtomStart[c], uff_fn, &ctx,
                                 maxIters, gradTol, 1, &energies[c], &it);
    if (converged) converged[c] = st == 0;
    if (iters) iters[c] = it;
  }
}
