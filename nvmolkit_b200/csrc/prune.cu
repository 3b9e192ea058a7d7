// RMS pruning of embedded conformers on the device. RDKit semantics
// (EmbedParameters::pruneRmsThresh, restated by the reference in rdkit_extensions/conformer_pruning.cpp:96-137): walk a
// molecule's conformers in order and keep one only if, after the best rigid alignment, its sum of squared deviations to
// EVERY conformer kept before it is at least nSel * thresh^2 - for every symmetry-equivalent atom mapping ("self match")
// when the caller provides them. The reference does this on the host after the GPU work and refuses it for DEVICE output
// (src/etkdg.cpp:106-110); here one CTA per molecule does the greedy walk, one warp per (kept conformer, match).
//
// Best-alignment SSD in closed form: with centred point sets, SSD = |a|^2 + |b|^2 - 2 (s1 + s2 + d s3), s_i the singular
// values of the 3x3 covariance H and d = sign(det H); s_i^2 are the eigenvalues of H^T H, which a symmetric 3x3 matrix
// yields trigonometrically. (RDKit's AlignPoints reaches the same optimum through the quaternion eigenproblem.)
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kPT = 256;

__device__ double bestSsd(int n, double sa2, double sb2, const double* sa, const double* sb, const double* sab) {
  // centred second moments from raw sums
  const double inv = 1.0 / n;
  double       H[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) H[3 * r + c] = sab[3 * r + c] - sa[r] * sb[c] * inv;
  const double Ga = sa2 - (sa[0] * sa[0] + sa[1] * sa[1] + sa[2] * sa[2]) * inv;
  const double Gb = sb2 - (sb[0] * sb[0] + sb[1] * sb[1] + sb[2] * sb[2]) * inv;
  double       M[6];  // H^T H: xx xy xz yy yz zz
  M[0] = H[0] * H[0] + H[3] * H[3] + H[6] * H[6];
  M[1] = H[0] * H[1] + H[3] * H[4] + H[6] * H[7];
  M[2] = H[0] * H[2] + H[3] * H[5] + H[6] * H[8];
  M[3] = H[1] * H[1] + H[4] * H[4] + H[7] * H[7];
  M[4] = H[1] * H[2] + H[4] * H[5] + H[7] * H[8];
  M[5] = H[2] * H[2] + H[5] * H[5] + H[8] * H[8];
  double       e1, e2, e3;
  const double p1 = M[1] * M[1] + M[2] * M[2] + M[4] * M[4];
  const double q  = (M[0] + M[3] + M[5]) / 3.0;
  const double p2 = (M[0] - q) * (M[0] - q) + (M[3] - q) * (M[3] - q) + (M[5] - q) * (M[5] - q) + 2.0 * p1;
  if (p2 <= 0.0) {
    e1 = e2 = e3 = q;
  } else {
    const double p  = sqrt(p2 / 6.0), ip = 1.0 / p;
    const double b0 = (M[0] - q) * ip, b1 = M[1] * ip, b2 = M[2] * ip, b3 = (M[3] - q) * ip, b4 = M[4] * ip, b5 = (M[5] - q) * ip;
    double       r  = 0.5 * (b0 * (b3 * b5 - b4 * b4) - b1 * (b1 * b5 - b4 * b2) + b2 * (b1 * b4 - b3 * b2));
    r               = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) / 3.0;
    e1               = q + 2.0 * p * cos(phi);
    e3               = q + 2.0 * p * cos(phi + 2.0943951023931954923);
    e2               = 3.0 * q - e1 - e3;
  }
  const double s1 = sqrt(fmax(e1, 0.0)), s2 = sqrt(fmax(e2, 0.0)), s3 = sqrt(fmax(e3, 0.0));
  const double det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
  const double ssd = Ga + Gb - 2.0 * (s1 + s2 + (det < 0.0 ? -s3 : s3));
  return ssd > 0.0 ? ssd : 0.0;
}

// mol_conf_start[nMols+1]: conformers of molecule m are [start[m], start[m+1]) (contiguous, in the order they were embedded);
// conf_atom_start[nConf+1] into xyz; match tables per molecule (see the C-ABI comment); keep[nConf] out.
__global__ void __launch_bounds__(kPT) rmsPruneKernel(int nMols, const int32_t* molConfStart, const int32_t* confAtomStart, const double* xyz,
                                                    const int32_t* matchOffset, const int32_t* matchLen, const int16_t* matchAtoms,
                                                    double thresh, const uint8_t* confValid, uint8_t* keep) {
  __shared__ int near;
  const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nW = kPT / 32;
  for (int m = blockIdx.x; m < nMols; m += gridDim.x) {
    const int c0 = molConfStart[m], c1 = molConfStart[m + 1];
    const int L  = matchLen ? matchLen[m] : (c1 > c0 ? confAtomStart[c0 + 1] - confAtomStart[c0] : 0);
    const int K  = matchLen ? (L > 0 ? (matchOffset[m + 1] - matchOffset[m]) / L : 0) : 1;
    const int16_t* mt = matchAtoms ? matchAtoms + matchOffset[m] : nullptr;
    const double   limit = L * thresh * thresh;
    for (int c = c0; c < c1; ++c) {
      __syncthreads();
      if (threadIdx.x == 0) near = 0;
      __syncthreads();
      const bool valid = !confValid || confValid[c];
      if (valid && L > 0 && K > 0) {
        const double* a = xyz + static_cast<size_t>(confAtomStart[c]) * 3;
        // one warp per (earlier kept conformer k, match): reference points = match 0 on c, probe points = the match on k
        for (int job = warp; job < (c - c0) * K; job += nW) {
          const int k = c0 + job / K, mi = job % K;
          if (!keep[k]) continue;  // (written by thread 0 before the barrier of this iteration)
          const double* b = xyz + static_cast<size_t>(confAtomStart[k]) * 3;
          double        sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0}, sab[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, sa2 = 0.0, sb2 = 0.0;
          for (int i = lane; i < L; i += 32) {
            const int ia = mt ? mt[i] : i, ib = mt ? mt[mi * L + i] : i;
            const double ax = a[3 * ia], ay = a[3 * ia + 1], az = a[3 * ia + 2];
            const double bx = b[3 * ib], by = b[3 * ib + 1], bz = b[3 * ib + 2];
            sa[0] += ax, sa[1] += ay, sa[2] += az;
            sb[0] += bx, sb[1] += by, sb[2] += bz;
            sa2 += ax * ax + ay * ay + az * az;
            sb2 += bx * bx + by * by + bz * bz;
            sab[0] += ax * bx, sab[1] += ax * by, sab[2] += ax * bz;
            sab[3] += ay * bx, sab[4] += ay * by, sab[5] += ay * bz;
            sab[6] += az * bx, sab[7] += az * by, sab[8] += az * bz;
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              sa[q] += __shfl_xor_sync(0xffffffffu, sa[q], o);
              sb[q] += __shfl_xor_sync(0xffffffffu, sb[q], o);
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) sab[q] += __shfl_xor_sync(0xffffffffu, sab[q], o);
            sa2 += __shfl_xor_sync(0xffffffffu, sa2, o);
            sb2 += __shfl_xor_sync(0xffffffffu, sb2, o);
          }
          if (lane == 0 && bestSsd(L, sa2, sb2, sa, sb, sab) < limit) near = 1;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) keep[c] = (valid && !near) ? 1 : 0;
    }
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200mol_rms_prune(int32_t nMols, const int32_t* d_mol_conf_start, const int32_t* d_conf_atom_start, const double* d_xyz,
                                 const int32_t* d_match_offset, const int32_t* d_match_len, const int16_t* d_match_atoms,
                                 double rms_thresh, const uint8_t* d_conf_valid, uint8_t* d_keep, void* stream) {
  return guarded([&] {
    if (nMols <= 0) return;
    B200_REQUIRE(d_mol_conf_start && d_conf_atom_start && d_xyz && d_keep, "null pointer");
    B200_REQUIRE((d_match_offset == nullptr) == (d_match_len == nullptr) && (d_match_len == nullptr) == (d_match_atoms == nullptr),
                 "match tables come together or not at all");
    B200_REQUIRE(rms_thresh >= 0.0, "negative RMS threshold");
    int blocks = smCount() * 4;
    if (blocks > nMols) blocks = nMols;
    rmsPruneKernel<<<blocks, kPT, 0, asStream(stream)>>>(nMols, d_mol_conf_start, d_conf_atom_start, d_xyz, d_match_offset, d_match_len,
                                                         d_match_atoms, rms_thresh, d_conf_valid, d_keep);
    B200_LAUNCHED();
  });
}
