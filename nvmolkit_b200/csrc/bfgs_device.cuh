// Device-side BFGS over a flattened force field for ONE conformer held in shared memory (see bfgs.cu for the design).
#pragma once
#include "ff.cuh"

namespace b200 {

constexpr int kT     = 256;  // threads per CTA
constexpr int kWarps = kT / 32;

__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warpMaxD(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide reductions; every thread receives the same value. `red` is kWarps doubles of shared memory.
__device__ __forceinline__ double blockSum(double v, double* red) {
  v = warpSum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ double blockMax(double v, double* red) {
  v = warpMaxD(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = red[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) t = fmax(t, red[w]);
  return t;
}

// RDKit ForceField::minimize gradient cap (>= 2025.09: |g|), bfgs_minimize.cu:797-851.
__device__ inline double scaleGrad(int n, double* grad, bool scaleGrads, double* red) {
  const int tid       = threadIdx.x;
  double    gradScale = scaleGrads ? 0.1 : 1.0, mx = 0.0;
  for (int i = tid; i < n; i += kT) {
    if (scaleGrads) grad[i] *= gradScale;
    mx = fmax(mx, fabs(grad[i]));
  }
  mx = blockMax(mx, red);
  if (scaleGrads && mx > 10.0) {
    while (mx * gradScale > 10.0) gradScale *= 0.5;
    for (int i = tid; i < n; i += kT) grad[i] *= gradScale;
  }
  __syncthreads();
  return gradScale;
}

template <class FF>
__device__ double energyOf(const typename FF::View& v, const double* x, double* red) {
  return blockSum(FF::template eval<false>(v, x, nullptr, threadIdx.x, kT), red);
}
template <class FF>
__device__ void gradOf(const typename FF::View& v, const double* x, double* grad, int n) {
  for (int i = threadIdx.x; i < n; i += kT) grad[i] = 0.0;
  __syncthreads();
  FF::template eval<true>(v, x, grad, threadIdx.x, kT);
  __syncthreads();
}

// Shared-memory working set of one CTA: six vectors of maxN doubles + the per-CTA inverse-Hessian slab (global/L2).
// HT = storage type of the inverse Hessian: double (default; bit-for-bit the RDKit recurrence) or float (half the slab
// traffic, products and sums still accumulate in fp64) for the embedding stages whose trajectories are chaotic anyway.
template <class HT = double>
struct BfgsWorkT {
  double *pos, *grad, *dir, *newPos, *dGrad, *hdg;  // shared memory, maxN each
  HT*     H;                                        // [n*n] global slab of this CTA
  double* red;                                      // kWarps doubles of shared memory
  double* scratch;                                  // shared memory, 4 * maxN doubles (scaled vectors of the Hessian passes)
};
constexpr int kBfgsVectors = 10;
template <class HT>
__host__ __device__ inline int bfgsLd(int n) {
  constexpr int per = 128 / static_cast<int>(sizeof(HT));
  return (n + per - 1) / per * per;
}  // six working vectors + four scratch vectors of maxN doubles
using BfgsWork = BfgsWorkT<double>;
template <class HT = double>
__device__ __forceinline__ BfgsWorkT<HT> carveWork(double* sm, int maxN, HT* H, double* red) {
  return {sm, sm + maxN, sm + 2 * maxN, sm + 3 * maxN, sm + 4 * maxN, sm + 5 * maxN, H, red, sm + 6 * maxN};
}

struct BfgsOutcome {
  int    status;  // 0 converged, 1 not
  int    iters;   // BFGS iterations of the last (re)start
  double energy;  // energy at w.pos (re-evaluated)
};

// Minimises w.pos[0..n) in place. maxRestarts > 0 re-runs (H = I, fresh gradient) while the run ends unconverged:
// RDKit's `while (needMore) needMore = field->minimize(...)` (src/etkdg_stage_distgeom_minimize.cu repeatUntilConverged).
#ifdef B200_BFGS_TIMING
static __device__ unsigned long long g_bfgsClk[8];  // per translation unit
static inline void readBfgsClocks(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_bfgsClk, sizeof(g_bfgsClk));
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_bfgsClk, z, sizeof(z));
}
#define B200_T0() const long long t0_ = clock64()
#define B200_T1(slot) tim[slot] += clock64() - t0_
#else
#define B200_T0()
#define B200_T1(slot)
#endif

template <class FF, class HT = double>
__device__ BfgsOutcome bfgsMinimize(const typename FF::View& view, const BfgsWorkT<HT>& w, int n, int maxIters,
                                    double gradTol, bool scaleGrads, int maxRestarts) {
  constexpr double FUNCTOL = 1e-4, MOVETOL = 1e-7, TOLX = 4. * 3e-8, EPS = 3e-8;
  double *pos = w.pos, *grad = w.grad, *dir = w.dir, *newPos = w.newPos, *dGrad = w.dGrad, *hdg = w.hdg, *red = w.red;
  HT*     H   = w.H;
  const int tid = threadIdx.x;
  // leading dimension of the slab: rows padded to 128 bytes so that every warp access is whole, aligned cache lines
  const int ld = bfgsLd<HT>(n);
  int       status = 1, iter = 0;
#ifdef B200_BFGS_TIMING
  long long tim[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tAll = clock64();
#endif
  for (int restart = 0;; ++restart) {
    __syncthreads();
    for (size_t i = tid; i < static_cast<size_t>(n) * ld; i += kT) H[i] = HT(0);
    __syncthreads();
    for (int i = tid; i < n; i += kT) H[static_cast<size_t>(i) * ld + i] = HT(1);

    double fp = energyOf<FF>(view, pos, red);
    gradOf<FF>(view, pos, grad, n);
    double gradScale = scaleGrad(n, grad, scaleGrads, red);
    double s2        = 0.0;
    for (int i = tid; i < n; i += kT) {
      dir[i] = -grad[i];
      s2 += pos[i] * pos[i];
    }
    const double maxStep = 100.0 * fmax(sqrt(blockSum(s2, red)), static_cast<double>(n));
    status               = 1;
    for (iter = 0; iter < maxIters; ++iter) {
      // ---------------- line search (bfgs_minimize.cu:80-162, 202-356) ----------------
      double t = 0.0;
      for (int i = tid; i < n; i += kT) t += dir[i] * dir[i];
      const double dsum = sqrt(blockSum(t, red));
      if (dsum > maxStep) {
        const double sc = maxStep / dsum;
        for (int i = tid; i < n; i += kT) dir[i] *= sc;
      }
      double sl = 0.0, tst = 0.0;
      for (int i = tid; i < n; i += kT) {
        sl += dir[i] * grad[i];
        tst = fmax(tst, fabs(dir[i]) / fmax(fabs(pos[i]), 1.0));
      }
      const double slope     = blockSum(sl, red);
      const double lambdaMin = MOVETOL / blockMax(tst, red);
      double       lambda = 1.0, lambda2 = 0.0, val2 = 0.0, newVal = fp;
      bool         accepted = false;
      for (int it = 0; it < 1000; ++it) {
        if (lambda < lambdaMin) break;
        for (int i = tid; i < n; i += kT) newPos[i] = pos[i] + lambda * dir[i];
        __syncthreads();
        {
          B200_T0();
          newVal = energyOf<FF>(view, newPos, red);
          B200_T1(0);
        }
        if (newVal - fp <= FUNCTOL * lambda * slope) {
          accepted = true;
          break;
        }
        double tmp;
        if (it == 0) {
          tmp = -slope / (2.0 * (newVal - fp - slope));
        } else {
          const double rhs1 = newVal - fp - lambda * slope, rhs2 = val2 - fp - lambda2 * slope;
          const double a    = (rhs1 / (lambda * lambda) - rhs2 / (lambda2 * lambda2)) / (lambda - lambda2);
          const double bq   = (-lambda2 * rhs1 / (lambda * lambda) + lambda * rhs2 / (lambda2 * lambda2)) / (lambda - lambda2);
          if (a == 0.0) {
            tmp = -slope / (2.0 * bq);
          } else {
            const double disc = bq * bq - 3 * a * slope;
            if (disc < 0.0) tmp = 0.5 * lambda;
            else if (bq <= 0.0) tmp = (-bq + sqrt(disc)) / (3.0 * a);
            else tmp = -slope / (bq + sqrt(disc));
          }
          if (tmp > 0.5 * lambda) tmp = 0.5 * lambda;
        }
        lambda2 = lambda;
        val2    = newVal;
        lambda  = fmax(tmp, 0.1 * lambda);
      }
      __syncthreads();
      if (!accepted)
        for (int i = tid; i < n; i += kT) newPos[i] = pos[i];  // "nothing was done"
      fp = newVal;
      // ---------------- direction, TOLX (bfgs_minimize.cu:732-776) ----------------
      tst = 0.0;
      for (int i = tid; i < n; i += kT) {
        const double xi = newPos[i] - pos[i];
        dir[i]          = xi;
        pos[i]          = newPos[i];
        tst             = fmax(tst, fabs(xi) / fmax(fabs(pos[i]), 1.0));
        dGrad[i]        = grad[i];
      }
      if (blockMax(tst, red) < TOLX) {
        status = 0;
        break;
      }
      {
        B200_T0();
        gradOf<FF>(view, pos, grad, n);
        B200_T1(1);
      }
      gradScale = scaleGrad(n, grad, scaleGrads, red);
      tst       = 0.0;
      for (int i = tid; i < n; i += kT) {
        tst      = fmax(tst, fabs(grad[i]) * fmax(fabs(pos[i]), 1.0));
        dGrad[i] = grad[i] - dGrad[i];
      }
      if (blockMax(tst, red) / fmax(fp * gradScale, 1.0) < gradTol) {
        status = 0;
        break;
      }
      // ---------------- inverse Hessian (bfgs_hessian.cu:37-239) ----------------
      B200_T0();
      // hdg = H * dGrad, thread per COLUMN (H is symmetric): every load is coalesced along a row, consecutive rows are
      // independent loads (deep memory-level parallelism) and no cross-lane reduction is needed. The warp-per-row form
      // exposed one L2 round trip + a shuffle tree per row and was 78 % of the embedder's time (profiles/).
      // Work item = (column j, row segment): items are spread evenly over the CTA whatever n is; partial sums meet in
      // shared memory (a few hundred fp64 atomics per pass). With fp32 slabs the whole pass runs in fp32 (no
      // conversions: F2F and fp64 are the slow pipes of this part, profiles/r01_path_b_summary.md).
      using AT         = HT;  // arithmetic type of the Hessian passes = storage type
      constexpr int kSlots = sizeof(HT) == 4 ? 8 : 4;  // column slots a lane keeps in registers per sweep
      AT*       vD     = reinterpret_cast<AT*>(w.scratch);  // dGrad as AT
      for (int i = tid; i < n; i += kT) {
        hdg[i] = 0.0;
        vD[i]  = static_cast<AT>(dGrad[i]);
      }
      __syncthreads();
      // Row-streaming, column-accumulating: a warp walks whole rows (contiguous, line-aligned loads), each lane keeps
      // the partial sums of ITS columns (H is symmetric: sum_i H[i][j] v_i = (H v)_j), rows are independent so several
      // are in flight; the eight warps' partial column sums meet in shared memory. No shuffles, no per-row round trip.
      {
        const int lane = tid & 31, warp = tid >> 5;
        for (int c0 = 0; c0 < n; c0 += 32 * kSlots) {  // kSlots column slots per lane per sweep
          AT acc[kSlots];
#pragma unroll
          for (int k = 0; k < kSlots; ++k) acc[k] = AT(0);
#pragma unroll 2
          for (int i = warp; i < n; i += kWarps) {
            const HT* hr = H + static_cast<size_t>(i) * ld + c0 + lane;
            const AT  vi = vD[i];
#pragma unroll
            for (int k = 0; k < kSlots; ++k)
              if (c0 + lane + 32 * k < n) acc[k] += hr[32 * k] * vi;
          }
#pragma unroll
          for (int k = 0; k < kSlots; ++k)
            if (c0 + lane + 32 * k < n) atomicAdd(&hdg[c0 + lane + 32 * k], static_cast<double>(acc[k]));
        }
      }
      __syncthreads();
      B200_T1(2);
      double f1 = 0, f2 = 0, f3 = 0, f4 = 0;
      for (int i = tid; i < n; i += kT) {
        f1 += dGrad[i] * dir[i];
        f2 += dGrad[i] * hdg[i];
        f3 += dGrad[i] * dGrad[i];
        f4 += dir[i] * dir[i];
      }
      double       fac      = blockSum(f1, red);
      const double fae      = blockSum(f2, red);
      const double sumDGrad = blockSum(f3, red);
      const double sumXi    = blockSum(f4, red);
      const bool   update   = fac > sqrt(EPS * sumDGrad * sumXi);
      double       fad      = 0.0;
      if (update) {
        fac = 1.0 / fac;
        fad = 1.0 / fae;
        for (int i = tid; i < n; i += kT) dGrad[i] = fac * dir[i] - fad * hdg[i];
      }
      __syncthreads();
      // fused: rank-2 update of row + dot with the new gradient -> next direction (into newPos, free here)
      const long long tU_ = clock64();
      (void)tU_;
      // per-row scalars of the rank-2 update, staged in newPos (free here) as fac*xi_i | hdg holds fad*hdg_i after scaling
      // thread per column: H[i][j] += (fac xi_i) xi_j - (fad hdg_i) hdg_j + (fae u_i) u_j ; a_j += H[i][j] g_i
      // scaled row vectors (index i) and plain column vectors (index j), in the arithmetic type:
      //   H[i][j] += sx_i x_j - sh_i h_j + su_i u_j ,  a_j += H[i][j] g_i        (3 + 1 FMAs per element)
      AT* sx = reinterpret_cast<AT*>(w.scratch);
      AT* sh = sx + n;
      AT* su = sh + n;
      AT* vg = su + n;
      AT* vx = vg + n;
      AT* vh = vx + n;
      AT* vu = vh + n;  // 7 n elements of AT <= 4 n doubles when AT = float; AT = double keeps x/h/u in place (below)
      if constexpr (sizeof(AT) == 4) {
        for (int i = tid; i < n; i += kT) {
          sx[i] = static_cast<AT>(fac * dir[i]);
          sh[i] = static_cast<AT>(fad * hdg[i]);
          su[i] = static_cast<AT>(fae * dGrad[i]);
          vg[i] = static_cast<AT>(grad[i]);
          vx[i] = static_cast<AT>(dir[i]);
          vh[i] = static_cast<AT>(hdg[i]);
          vu[i] = static_cast<AT>(dGrad[i]);
          newPos[i] = 0.0;
        }
      } else {
        for (int i = tid; i < n; i += kT) {
          sx[i] = static_cast<AT>(fac * dir[i]);
          sh[i] = static_cast<AT>(fad * hdg[i]);
          su[i] = static_cast<AT>(fae * dGrad[i]);
          newPos[i] = 0.0;
        }
      }
      __syncthreads();
      {
        const int lane = tid & 31, warp = tid >> 5;
        for (int c0 = 0; c0 < n; c0 += 32 * kSlots) {
          AT acc[kSlots], xj[kSlots], hj[kSlots], uj[kSlots];
#pragma unroll
          for (int k = 0; k < kSlots; ++k) {
            const int j = c0 + lane + 32 * k;
            acc[k]      = AT(0);
            if constexpr (sizeof(AT) == 4) {
              xj[k] = j < n ? vx[j] : AT(0);
              hj[k] = j < n ? vh[j] : AT(0);
              uj[k] = j < n ? vu[j] : AT(0);
            } else {
              xj[k] = j < n ? dir[j] : 0.0;
              hj[k] = j < n ? hdg[j] : 0.0;
              uj[k] = j < n ? dGrad[j] : 0.0;
            }
          }
#pragma unroll 2
          for (int i = warp; i < n; i += kWarps) {
            HT*      hr = H + static_cast<size_t>(i) * ld + c0 + lane;
            const AT si = sx[i], ti = sh[i], wi2 = su[i];
            AT       gi;
            if constexpr (sizeof(AT) == 4) gi = vg[i];
            else gi = grad[i];
            if (update) {
#pragma unroll
              for (int k = 0; k < kSlots; ++k)
                if (c0 + lane + 32 * k < n) {
                  const AT h = hr[32 * k] + (si * xj[k] - ti * hj[k] + wi2 * uj[k]);
                  hr[32 * k] = h;
                  acc[k] += h * gi;
                }
            } else {
#pragma unroll
              for (int k = 0; k < kSlots; ++k)
                if (c0 + lane + 32 * k < n) acc[k] += hr[32 * k] * gi;
            }
          }
#pragma unroll
          for (int k = 0; k < kSlots; ++k)
            if (c0 + lane + 32 * k < n) atomicAdd(&newPos[c0 + lane + 32 * k], -static_cast<double>(acc[k]));
        }
      }
      __syncthreads();
#ifdef B200_BFGS_TIMING
      tim[3] += clock64() - tU_;
      tim[4] += 1;
#endif
      for (int i = tid; i < n; i += kT) dir[i] = newPos[i];
      __syncthreads();
    }
    if (status == 0 || restart >= maxRestarts) break;
  }
  __syncthreads();
  BfgsOutcome out;
  out.status = status;
  out.iters  = iter;
  out.energy = energyOf<FF>(view, pos, red);
#ifdef B200_BFGS_TIMING
  if (tid == 0) {
    tim[5] = clock64() - tAll;
    for (int k = 0; k < 6; ++k) atomicAdd(&g_bfgsClk[k], static_cast<unsigned long long>(tim[k]));
  }
#endif
  return out;
}

}  // namespace b200
