// Popcount similarity on the 5th-generation tensor cores (tcgen05), sm_100a: the thresholded neighbour pass of the fused
// Butina path and the materialised Tanimoto / cosine matrices.
//
// |A & B| of two bit vectors is the dot product of their 0/1 expansions, so the N x M intersection-count matrix is a
// GEMM with exact small-integer sums. The SIMT tile (tanimoto.cu) sits on the POPC issue roof (64 POPC per 2048-bit
// pair, 16 lanes/clk/SM -> 7e10 pairs/s, profiles/r01_path_a_summary.md); the reference reaches its tensor path through
// `mma.sync ... b1 ... and.popc`, which ptxas lowers on sm_100a to bit-slicing LOP3s plus eight IMMA.16832.U8.U8 per
// 256-bit step - the legacy warp-level path. Here the contraction is issued natively:
//
//   pre-pass   bits -> packed E2M1 (fp4) once per fingerprint set, 1 KB per 2048-bit row; for the neighbour pass a row
//              operand is the SUM of S = 4 fingerprints and a column operand the sum of C = 1, 2 or 4 (values <= 4 and
//              products <= 16 are exact), so one accumulator bounds S * C pair counts. Fingerprints that are no multiple
//              of 256 bits take the int8 tile (0/1 bytes, kind::i8, 128 x 256).
//   tile       128 x 224 accumulators per step of a persistent CTA (CTA pairs share the column operand: TMA multicast)
//   warp 0     TMA producer: [128 | 224 rows][128 B] K-chunks, SWIZZLE_128B, mbarrier ring (4 stages; 3 when materialising)
//   warp 1     one elected thread issues tcgen05.mma.cta_group::1.kind::mxf4.block_scale (M128 N224 K64, every scale
//              factor 1.0), accumulators in TMEM (2 x 224 columns: the next tile's MMAs overlap this tile's epilogue)
//   warps 2-9  count epilogue: the two warps of a TMEM lane quarter alternate tiles; tcgen05.ld 32x32b, one subtract and
//              one max tree per 32 columns decide "no pair of this group can reach its threshold"; survivors go to a
//              candidate list (staged per warp in shared memory) that verifyCandidatesKernel re-counts exactly with the
//              integer threshold table (bit-exact with the fp64 predicate, see tanimoto.cu). Unsuperposed (S = C = 1) the
//              same warps apply the exact test themselves: neighbour counts for both endpoints + warp-aggregated edges.
//   warps 2-17 materialise epilogue: fp64 Tanimoto / cosine values staged per warp with the 128-byte swizzle and stored
//              by TMA (cp.async.bulk.tensor store).
// A pilot over a prefix sample picks C for the data at hand; a candidate-list overflow reruns with fewer pairs per
// accumulator before anything has been counted. Measurements: profiles/r02_path_a_summary.md.
//
// Replaces crossSimilarityKernelTensorOp (src/similarity_kernels.cu:104-240) + the Triton count kernel
// (nvmolkit/_fusedButina.py:99-179) for the fused Butina pass.
#include "profile.cuh"
#include "similarity.cuh"
#include "tma.cuh"

namespace b200 {
namespace {

constexpr int kTM       = 128;
constexpr int kTN       = 256;
constexpr int kTK       = 128;  // bytes (= bits of the fingerprint) per K chunk
constexpr int kStagesCount = 4;  // smem ring depth of the count mode (int8 tile)
#ifndef B200_STAGES_FP4
#define B200_STAGES_FP4 4
#endif
constexpr int kStagesCountFp4 = B200_STAGES_FP4;  // ... of the fp4 count tile
// RN(1/u) of the materialise epilogue: from the table (one 8-byte gather per element through L1) or computed
// (__drcp_rn). Same value either way; measured at 32k x 32k: table 2.75 ms, computed 3.10 ms.
#ifndef B200_RECIP_TABLE
#define B200_RECIP_TABLE 1
#endif
constexpr bool kRecipTable = B200_RECIP_TABLE != 0;
constexpr int kStagesMat   = 3;  // the materialise modes are bound by the fp64 output (11.4k clocks of HBM write per tile and SM),
                                 // but with TWO stages the 16 operand chunks of a tile took 16 L2 round trips / 2 = 16k clocks:
                                 // three stages bring the operand stream under the write time. Shared memory also holds 16 x 4 KB
                                 // of staging for the TMA stores; the reciprocal table moved to global memory (L1) to make room
constexpr int kStagesPair  = 6;  // pair-MMA count mode: 30 KB per stage and CTA  // materialise modes: one stage less, the space holds the reciprocal table
constexpr int kEpiWarpsCount = 8;   // count mode: two warps per TMEM lane quarter share the column blocks
constexpr int kEpiWarpsMat   = 16;  // materialise modes: four per quarter (the fp64 epilogue is the long pole there)
constexpr int epiWarps(int mode) { return mode == 0 ? kEpiWarpsCount : kEpiWarpsMat; }
constexpr int threadsTC(int mode) { return 64 + 32 * epiWarps(mode); }  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int kABytes   = kTM * kTK;
#ifndef B200_TN_FP4
#define B200_TN_FP4 224
#endif
constexpr int kTNFp4    = B200_TN_FP4;  // fp4 count mode: 2 x 224 accumulator columns leave TMEM columns 448..511 for the scale factors
constexpr int kGroupRows = 8192;  // fingerprints per row group: the unit of L2 reuse of the column operand AND of the
                                  // multi-GPU row split (independent of the tile variant and of the superposition factor)
constexpr int kRunStat   = 16;  // tile columns per unit of the A-stationary tile (the row operand is loaded once per unit)
constexpr int kStagesStat = 3;  // its ring holds the column operand only (28 KB per stage, next to 128 KB of row operand)
constexpr int kMaxChunksStat = 8;  // row operand resident in shared memory: 8 K-chunks x 16 KB (fingerprints <= 2048 bits)

struct TcParams {
  uint32_t        n;  // X == Y (symmetric) or nX/nY
  uint32_t        nY;
  int             kChunks;
  uint32_t        tilesM, tilesN;
  int             symmetric;
  uint32_t        groupOffset, groupStride;  // multi-GPU: this rank owns tile-row groups with group % stride == offset
  const int32_t*  popX;
  const int32_t*  popY;
  const uint16_t* thresh;
  int             threshLen;
  int             sign;
  int32_t*        counts;
  int32_t*        countsY;  // == counts in symmetric mode; nullptr in X-vs-Y mode (rows only)
  int2*           edges;
  unsigned long long* edgeCursor;
  unsigned long long  edgeCap;
  double*         out;  // materialise modes: [n][nY] fp64
  int             recipLen;  // 2 * bits (materialise Tanimoto)
  const double*   recipG;    // RN(1/u), u = 0 .. recipLen (global memory, 32 KB: L1-resident)
  // Row superposition (count mode): a row of the X operand is the SUM of superS consecutive fingerprints (values 0..4,
  // exact in E2M1), so one accumulator bounds superS pair counts at once; n / tilesM then count SUPER rows, popX holds
  // the smallest popcount of each super row, and the epilogue only lists candidates (super row, column) for the exact
  // verification kernel. rowSpan = fingerprints per tile row = kTM * superS.
  int                 outTma;  // materialise modes: the epilogue leaves through TMA stores (out 16-byte aligned, nY even)
  int                 superS;
  int                 superC;   // the same for the columns of the Y operand: one accumulator then bounds superS * superC pair counts
  uint32_t            colSpan;  // fingerprints per tile column = TN * superC
  float               alpha;    // (1 - cutoff) / (2 - cutoff), rounded down: a pair can only pass with c >= alpha (|A| + |B|)
  uint32_t            rowSpan;
  uint32_t            groupTiles;  // tile rows per row group (a power of two): kGroupRows fingerprints whatever superS is
  int2*               cand;
  unsigned long long* candCursor;
  unsigned long long  candCap;
};

enum TcMode : int { kTcCount = 0, kTcTanimoto = 1, kTcCosine = 2 };
}  // namespace
int g_tensorCluster = 1;  // fp4 count tile in clusters of two CTAs with a multicast column operand (option "similarity_tensor_cluster")
int g_tensorFp4 = 1;  // count mode on block-scaled fp4 operands (option "similarity_tensor_fp4"; 0 = int8 tile)
namespace {
// -DB200_TC_MMAONLY (with B200_TC_TIMING; tools/pair_pass_timing.py): nothing is loaded, nothing waits, no epilogue - the
// MMA thread issues the same instruction stream on whatever shared memory holds. Results are garbage; the time per tile
// is the tensor pipe's own rate for this instruction mix (the floor the real kernel is compared with).
// -DB200_TC_MMAONLY=2: the same with the TMA producer and the operand waits back in (still no epilogue): what the operand
// stream alone costs the tensor pipe.
#ifdef B200_TC_MMAONLY
constexpr bool kMmaOnly = true;
constexpr bool kMmaFed  = (B200_TC_MMAONLY + 0) == 2;
#else
constexpr bool kMmaOnly = false;
constexpr bool kMmaFed  = false;
#endif
#ifdef B200_TC_TIMING
// clock64() attribution of the count tile: [0] MMA thread total, [1] its wait for operands (fullBar / aFull), [2] its wait for a
// free accumulator (tmemEmpty), [3] epilogue warp 0 total, [4] its wait for a finished accumulator (tmemFull), [5] its
// staging barrier, [6] tiles (MMA thread), [7] producer wait for a free stage
__device__ unsigned long long g_tcClk[8];
#define TC_T0() const long long tc0_ = clock64()
#define TC_T1(slot) tcAcc[slot] += clock64() - tc0_
#else
#define TC_T0()
#define TC_T1(slot)
#endif

__global__ void expandBitsKernel(const uint32_t* __restrict__ fp, size_t nWords, uint4* __restrict__ out) {
  const size_t w = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= nWords) return;
  const uint32_t x = fp[w];
  uint32_t       b[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {  // 4 bits -> 4 bytes of 0/1 (bit j of the nibble lands in byte j)
    const uint32_t nib = (x >> (4 * q)) & 0xFu;
    b[q]               = (nib * 0x00204081u) & 0x01010101u;
  }
  out[2 * w]     = make_uint4(b[0], b[1], b[2], b[3]);
  out[2 * w + 1] = make_uint4(b[4], b[5], b[6], b[7]);
}

// bits -> packed E2M1 (fp4): bit = 1 -> 0x2 (1.0), bit = 0 -> 0x0; two elements per byte, 1 KB per 2048-bit row. The
// contraction is invariant to the order of the K elements as long as both operands use the same one.
__global__ void expandBitsFp4Kernel(const uint32_t* __restrict__ fp, size_t nWords, uint4* __restrict__ out) {
  const size_t w = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= nWords) return;
  const uint32_t x = fp[w];
  uint32_t       b[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // 8 bits -> 8 nibbles
    uint32_t v = (x >> (8 * q)) & 0xFFu;
    v          = (v | (v << 12)) & 0x000F000Fu;
    v          = (v | (v << 6)) & 0x03030303u;
    v          = (v | (v << 3)) & 0x11111111u;
    b[q]       = v << 1;
  }
  out[w] = make_uint4(b[0], b[1], b[2], b[3]);
}

// Superposed row operand: fp4 row R = sum over s < S of the 0/1 expansions of fingerprints S R + s (0..4 -> E2M1 codes
// 0x0 0x2 0x4 0x5 0x6, all exact). One thread per 32 fingerprint bits.
__global__ void expandBitsFp4SuperKernel(const uint32_t* __restrict__ fp, size_t n, int words, int S, size_t nSuper,
                                         uint4* __restrict__ out) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nSuper * static_cast<size_t>(words)) return;
  const size_t R = t / words;
  const int    w = static_cast<int>(t % words);
  uint32_t     b[4] = {0, 0, 0, 0};  // 8 nibble counters each
  for (int s = 0; s < S; ++s) {
    const size_t i = R * S + s;
    if (i >= n) break;
    const uint32_t x = fp[i * words + w];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = (x >> (8 * q)) & 0xFFu;
      v          = (v | (v << 12)) & 0x000F000Fu;
      v          = (v | (v << 6)) & 0x03030303u;
      v          = (v | (v << 3)) & 0x11111111u;
      b[q] += v;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t code = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) code |= ((0x65420u >> (4 * ((b[q] >> (4 * k)) & 0xFu))) & 0xFu) << (4 * k);
    b[q] = code;
  }
  out[t] = make_uint4(b[0], b[1], b[2], b[3]);
}

// smallest popcount among the fingerprints of each super row (the epilogue's conservative pre-filter needs the lowest
// threshold any pair of the group can have)
__global__ void superMinPopKernel(const int32_t* __restrict__ pop, size_t n, int S, size_t nSuper, int32_t* __restrict__ out) {
  const size_t R = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (R >= nSuper) return;
  int m = 0x3fffffff;
  for (int s = 0; s < S; ++s)
    if (R * S + s < n) m = min(m, pop[R * S + s]);
  out[R] = m;
}

// Exact verification of the candidates of a superposed pass: one warp per (super row R, super column J), for each of
// the S * C pairs (i = S R + s, j = C J + c) of the group the exact count |X_i & Y_j| and the exact integer threshold
// test; counts for both endpoints and the (i < j) edge list exactly as the unsuperposed epilogue produces them.
// (Measured and rejected: reading a 512-bit prefix of both rows first and dropping pairs whose exact upper bound
// cq + min(|X| - xa, |Y| - yb) is below the threshold - 7 of 8 pairs leave after a quarter of the bytes, but the second,
// dependent round of loads costs more than the traffic saved: 9.7 ms against 8.2.)
// A block takes 64 consecutive candidates at a time (one coalesced load; the warp that listed them worked on one quarter
// of a tile row, so their row operands are L1 / L2 hits), a warp one candidate: all of its 128-bit row loads and the two
// popcount loads are in flight together; edges are parked in 64 shared-memory slots per warp that leave with ONE global
// atomic per flush.
__global__ void __launch_bounds__(256) verifyCandidatesKernel(const uint32_t* __restrict__ x, const uint32_t* __restrict__ y, int words,
                                                             const int2* __restrict__ cand, unsigned long long nCand, int S, int C,
                                                             uint32_t nRows, uint32_t nCols, int symmetric,
                                                             const int32_t* __restrict__ popX, const int32_t* __restrict__ popY,
                                                             const uint16_t* __restrict__ thresh, int sign, int32_t* counts,
                                                             int32_t* countsY, int2* edges, unsigned long long* edgeCursor,
                                                             unsigned long long edgeCap) {
  constexpr int kStage = 64, kFlight = 1, kChunk = 64;  // a block takes 64 consecutive candidates at a time: the epilogue
  // warp that listed them worked on ONE quarter of a tile row (32 super rows = 32 KB of fingerprints), so most of their
  // row operands are L1 hits for the block
  __shared__ int2 stage[8][kStage];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = 32 / (S * C), sub = lane / G, gl = lane % G;  // S, C are 1, 2 or 4: G lanes per pair
  const int chunks = words / 4;                                // uint4 per fingerprint
  int       nStaged = 0;
  auto flush = [&]() {
    if (nStaged == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(edgeCursor, static_cast<unsigned long long>(nStaged));
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    for (int k = lane; k < nStaged; k += 32)
      if (base + k < edgeCap) edges[base + k] = stage[warp][k];
    __syncwarp();
    nStaged = 0;
  };
  __shared__ int2 chunkCand[kChunk];
  for (unsigned long long chunk = static_cast<unsigned long long>(blockIdx.x) * kChunk; chunk < nCand; chunk += static_cast<unsigned long long>(gridDim.x) * kChunk) {
  // the block's 64 candidates in one coalesced load (a warp fetching its own entry would start every candidate with a
  // dependent L2 round trip)
  __syncthreads();
  if (threadIdx.x < kChunk) chunkCand[threadIdx.x] = chunk + threadIdx.x < nCand ? cand[chunk + threadIdx.x] : make_int2(-1, -1);
  __syncthreads();
  for (int it = 0; it < kChunk / (8 * kFlight); ++it) {
    uint32_t     iOf[kFlight], jOf[kFlight];
    bool         live[kFlight];
    int          popSum[kFlight];
    const uint4 *xi[kFlight], *yj[kFlight];
#pragma unroll
    for (int f = 0; f < kFlight; ++f) {
      const int2 rj   = chunkCand[it * 8 * kFlight + warp * kFlight + f];
      const bool have = rj.x >= 0;
      iOf[f]          = static_cast<uint32_t>(rj.x) * S + sub / C;
      jOf[f]          = static_cast<uint32_t>(rj.y) * C + sub % C;
      live[f]         = have && iOf[f] < nRows && jOf[f] < nCols && !(symmetric && iOf[f] >= jOf[f]);
      xi[f]           = reinterpret_cast<const uint4*>(x + static_cast<size_t>(live[f] ? iOf[f] : 0) * words);
      yj[f]           = reinterpret_cast<const uint4*>(y + static_cast<size_t>(live[f] ? jOf[f] : 0) * words);
      popSum[f]       = (gl == 0 && live[f]) ? popX[iOf[f]] + popY[jOf[f]] : 0;  // (in flight together with the rows)
    }
    int cnt[kFlight] = {};
    for (int q = gl; q < chunks; q += 4 * G) {  // (4 chunks per lane and candidate in flight)
      uint4 a[kFlight][4], b4[kFlight][4];
#pragma unroll
      for (int f = 0; f < kFlight; ++f)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (q + u * G < chunks) a[f][u] = xi[f][q + u * G], b4[f][u] = yj[f][q + u * G];
#pragma unroll
      for (int f = 0; f < kFlight; ++f)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (q + u * G < chunks)
            cnt[f] += __popc(a[f][u].x & b4[f][u].x) + __popc(a[f][u].y & b4[f][u].y) + __popc(a[f][u].z & b4[f][u].z) +
                      __popc(a[f][u].w & b4[f][u].w);
    }
#pragma unroll
    for (int f = 0; f < kFlight; ++f) {
      int v = cnt[f];
      for (int o = G >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const bool hit = gl == 0 && live[f] && v >= thresh[popSum[f]];
      if (hit) {
        atomicAdd(counts + iOf[f], sign);
        if (countsY) atomicAdd(countsY + jOf[f], sign);
      }
      if (edges) {
        const unsigned hits = __ballot_sync(0xffffffffu, hit);
        if (hits) {
          const int total = __popc(hits);  // <= 16
          if (nStaged + total > kStage) flush();
          if (hit) stage[warp][nStaged + __popc(hits & ((1u << lane) - 1u))] = make_int2(static_cast<int>(iOf[f]), static_cast<int>(jOf[f]));
          nStaged += total;
        }
      }
    }
  }
  }
  if (edges) flush();
}

__global__ void recipTableKernel(double* __restrict__ r, int len) {  // r[u] = RN(1 / u), r[0] = 0
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u <= len) r[u] = u ? __drcp_rn(static_cast<double>(u)) : 0.0;
}

// lo[S] = min(thresh[S..len-1]): the threshold table need not be monotone (cutoff 0 admits only even |A|+|B|), its
// lower envelope is, and that is what the epilogue's one-compare pre-filter may use. One block, any len <= 16384.
__global__ void __launch_bounds__(1024) threshSuffixMinKernel(const uint16_t* __restrict__ thresh, int len, uint16_t* __restrict__ lo) {
  __shared__ int part[1024];
  const int      per = (len + 1023) / 1024, beg = threadIdx.x * per, end = min(len, beg + per);
  int            m   = 0x7fffffff;
  for (int i = end - 1; i >= beg; --i) m = min(m, static_cast<int>(thresh[i]));
  part[threadIdx.x] = m;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // inclusive suffix-min over the chunk minima
    const int v = threadIdx.x + o < 1024 ? part[threadIdx.x + o] : 0x7fffffff;
    __syncthreads();
    part[threadIdx.x] = min(part[threadIdx.x], v);
    __syncthreads();
  }
  int run = threadIdx.x + 1 < 1024 ? part[threadIdx.x + 1] : 0x7fffffff;  // everything to the right of this chunk
  for (int i = end - 1; i >= beg; --i) {
    run   = min(run, static_cast<int>(thresh[i]));
    lo[i] = static_cast<uint16_t>(run);
  }
}

__device__ __forceinline__ uint64_t makeSmemDesc(uint32_t smemByteAddr) {
  // K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), LBO unused, descriptor version 1 (sm_100), layout type 2
  return static_cast<uint64_t>((smemByteAddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}
constexpr uint32_t kIdescI8 = (2u << 4)                 // D = S32
                              | (0u << 7) | (0u << 10)  // A, B = unsigned 8-bit
                              | (static_cast<uint32_t>(kTN >> 3) << 17) | (static_cast<uint32_t>(kTM >> 4) << 24);

__device__ __forceinline__ void ummaI8(uint32_t tmemD, uint64_t aDesc, uint64_t bDesc, uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD),
    "l"(aDesc), "l"(bDesc), "r"(kIdescI8), "r"(accumulate)
    : "memory");
}
// Block-scaled fp4 (kind::mxf4, K = 64 elements = 32 bytes per instruction, twice the int8 rate): operands are the
// packed E2M1 0/1 expansions, every UE8M0 scale factor is 1.0 (0x7F; TMEM columns 448..511 are filled with it, so the
// scale-factor layout is immaterial), accumulation in fp32 is exact (sums <= 4096).
// Descriptor bits (cute/arch/mma_sm100_desc.hpp, InstrDescriptorBlockScaled): a/b format E2M1 = 1 at [7,10) / [10,13),
// N >> 3 at [17,23), scale format UE8M0 = 1 at bit 23, M >> 4 at [24,29), scale-factor ids 0, K = 64 (bit 31 = 0).
constexpr uint32_t kIdescMxf4 = (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTNFp4 >> 3) << 17) | (1u << 23) |
                                (static_cast<uint32_t>(kTM >> 4) << 24);
__device__ __forceinline__ void ummaMxf4(uint32_t tmemD, uint64_t aDesc, uint64_t bDesc, uint32_t accumulate, uint32_t tmemSfa,
                                         uint32_t tmemSfb) {
  asm volatile(
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmemD),
    "l"(aDesc), "l"(bDesc), "r"(kIdescMxf4), "r"(accumulate), "r"(tmemSfa), "r"(tmemSfb)
    : "memory");
}
__device__ __forceinline__ void tmemStore32Const(uint32_t taddr, uint32_t v) {
  asm volatile(
    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
    "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr),
    "r"(v)
    : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ummaCommit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void ummaCommitMulticast(uint64_t* bar, uint16_t ctaMask) {  // same barrier offset in every CTA of the mask
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smemAddr(bar)),
               "h"(ctaMask)
               : "memory");
}
// CTA-pair MMA (cta_group::2): M = 256 (128 rows in each CTA's TMEM), the N x K operand split across the two CTAs'
// shared memory (N/2 rows each, same offsets), issued by the leader alone.
constexpr uint32_t kIdescMxf4Pair = (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTNFp4 >> 3) << 17) | (1u << 23) |
                                    (static_cast<uint32_t>((2 * kTM) >> 4) << 24);
__device__ __forceinline__ void ummaMxf4Pair(uint32_t tmemD, uint64_t aDesc, uint64_t bDesc, uint32_t accumulate, uint32_t tmemSfa,
                                             uint32_t tmemSfb) {
  asm volatile(
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmemD),
    "l"(aDesc), "l"(bDesc), "r"(kIdescMxf4Pair), "r"(accumulate), "r"(tmemSfa), "r"(tmemSfb)
    : "memory");
}
__device__ __forceinline__ void ummaCommitPairMulticast(uint64_t* bar, uint16_t ctaMask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smemAddr(bar)),
               "h"(ctaMask)
               : "memory");
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmemLoad32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
    "[%32];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
    : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Work units. A unit = one tile (CL = 0) or two vertically adjacent tiles of one tile column (CTA pairs: unit row r
// holds tile rows 2 r and 2 r + 1, the two CTAs share the column operand). Unit rows come in groups of G (p.groupTiles
// tile rows = kGroupRows fingerprints), a group sweeps the tile columns row-fastest (L2 reuse of the row operand). Only units that can hold a pair
// are enumerated: a rank walks the groups it OWNS, and a group starts at the first tile column that reaches past the
// diagonal for its top tile row (closed form), so that neither the other ranks' groups nor the lower triangle cost
// loop iterations (round 1 walked all of them: 215 ns per skipped unit, the 1 -> 8 GPU limiter).
// Ownership is serpentine over cycles of `groupStride` groups (even cycles: group = cycle * stride + offset, odd
// cycles mirrored), which balances the triangle's shrinking rows across ranks to < 0.1 %.
// A CTA walks units first, first + step, ... of the concatenated owned groups; (cycle, index in group) are kept
// incrementally, the only divisions are by compile-time constants and happen once per group.
// RUN > 1 (the A-stationary tile, CL = 3): a unit is a RUN of consecutive tile columns of one unit row - the row
// operand is loaded once per unit and stays in shared memory - and the units of a group go chunk-major (all rows of the
// group take column chunk q, then q + 1): the CTA pairs of a group stream the same column tiles at about the same time,
// so a column tile comes from HBM once per group and from L2 for the other rows.
template <int TN, bool PAIR, int RUN = 1>
struct UnitWalk {
  uint32_t cycle, inGroup, units, gRows, tn0, group, step, G, gShift;
  bool     done;
  __device__ UnitWalk(const TcParams& p, uint32_t first, uint32_t stepBy)
      : cycle(0), inGroup(first), units(0), gRows(0), tn0(0), group(0), step(stepBy), G(PAIR ? p.groupTiles / 2 : p.groupTiles),
        gShift(0), done(false) {
    while ((1u << gShift) < G) ++gShift;  // G is a power of two
    settle(p);
  }
  __device__ void settle(const TcParams& p) {  // make (cycle, inGroup) point at an existing unit, or set done
    const uint32_t unitRows = PAIR ? (p.tilesM + 1) / 2 : p.tilesM;
    for (;;) {
      if (static_cast<uint64_t>(cycle) * p.groupStride * G >= unitRows) {
        done = true;
        return;
      }
      group = cycle * p.groupStride + ((cycle & 1u) ? p.groupStride - 1 - p.groupOffset : p.groupOffset);
      units = 0;
      if (static_cast<uint64_t>(group) * G < unitRows) {
        gRows = min(G, unitRows - group * G);
        // first tile column holding a pair with row < col for the group's top tile row tm0 = group * groupTiles:
        // (tn + 1) * colSpan - 1 > tm0 * rowSpan   (rowSpan / colSpan = fingerprints per tile row / tile column)
        tn0   = p.symmetric ? (group * p.groupTiles * p.rowSpan + 1u) / p.colSpan : 0u;
        if (tn0 < p.tilesN) units = gRows * ((p.tilesN - tn0 + RUN - 1) / RUN);
      }
      if (inGroup < units) return;
      inGroup -= units;
      ++cycle;
    }
  }
  __device__ void next(const TcParams& p) {
    inGroup += step;
    if (inGroup >= units) {
      inGroup -= units;
      ++cycle;
      settle(p);
    }
  }
  // tiles [tnBeg, tnEnd) of tile row tm for CTA `rank` of the pair (0 when unpaired); false = nothing to do
  __device__ bool coords(const TcParams& p, uint32_t rank, uint32_t& tm, uint32_t& tnBeg, uint32_t& tnEnd) const {
    uint32_t row, col;
    if (gRows == G) {
      row = inGroup & (G - 1);  // G is a power of two
      col = inGroup >> gShift;
    } else {  // the last, partial group
      row = inGroup % gRows;
      col = inGroup / gRows;
    }
    tnBeg              = tn0 + col * RUN;
    tnEnd              = min(p.tilesN, tnBeg + RUN);
    const uint32_t tr  = group * G + row;  // unit row
    const uint32_t top = PAIR ? 2 * tr : tr;
    // the upper tile decides for both CTAs of a pair (if it has no pair with row < col, neither has the lower one); a
    // lower tile past the end or below the diagonal still runs - its loads are zero-filled / its predicates reject all.
    // A tile column is useful iff (tn + 1) * colSpan - 1 > top * rowSpan: monotone in tn, so a run is clipped from the left.
    if (p.symmetric) tnBeg = max(tnBeg, (top * p.rowSpan + 1u) / p.colSpan);
    if (tnBeg >= tnEnd) return false;
    tm = top + (PAIR ? rank : 0u);
    return true;
  }
};

// host twin of the walk's unit count (sizes the grid)
template <int TN, bool PAIR, int RUN = 1>
uint64_t countUnits(const TcParams& p) {
  const uint32_t G        = PAIR ? p.groupTiles / 2 : p.groupTiles;
  const uint32_t unitRows = PAIR ? (p.tilesM + 1) / 2 : p.tilesM;
  uint64_t       total    = 0;
  for (uint32_t cycle = 0; static_cast<uint64_t>(cycle) * p.groupStride * G < unitRows; ++cycle) {
    const uint32_t group = cycle * p.groupStride + ((cycle & 1u) ? p.groupStride - 1 - p.groupOffset : p.groupOffset);
    if (static_cast<uint64_t>(group) * G >= unitRows) continue;
    const uint32_t gRows = std::min(G, unitRows - group * G);
    const uint32_t tn0   = p.symmetric ? (group * p.groupTiles * p.rowSpan + 1u) / p.colSpan : 0u;
    if (tn0 < p.tilesN) total += static_cast<uint64_t>(gRows) * ((p.tilesN - tn0 + RUN - 1) / RUN);
  }
  return total;
}

// CL: 0 = one CTA per tile; 1 = CTA pair, column operand multicast; 2 = CTA pair with cta_group::2 MMAs (each CTA stages
// half of the column operand, the leader issues M = 256 instructions for both); 3 = CTA pair, column operand multicast,
// ROW OPERAND STATIONARY: a unit is a run of kRunStat tile columns of one tile row, the row tile's K chunks (128 KB)
// are loaded once per unit into their own shared-memory region and only the column operand streams through the ring.
// Per pair that is 4.3 B from L2 instead of 8.6 (the pass was bound by L2 -> SM delivery at 9.6 TB/s and by 3.7 TB/s of
// HBM re-reads, profiles/r02_path_a_summary.md); the K chunks of the next unit's row tile are requested as soon as the
// last tile of the current unit has consumed them, so the reload hides behind that tile's remaining MMAs.
template <int MODE, bool FP4, int CL>
__global__ void __launch_bounds__(threadsTC(MODE), 1)
  simTensorKernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smemRaw[];
  constexpr int kEpiWarps = epiWarps(MODE), kThreadsTC = threadsTC(MODE);
  // Count mode: the two warps of a TMEM lane quarter ALTERNATE tiles (warps 2..5 take the even accumulator buffer, warps
  // 6..9 the odd one) instead of splitting the column blocks of one tile. A tile's epilogue is a chain of latencies (column
  // popcounts from L2, a barrier, seven TMEM reads one after the other) of ~7.7k clocks against ~4.7k of MMA issue; two
  // tiles in flight hide it. The materialise modes keep all their warps on one tile (bound by the fp64 output stream).
  constexpr bool ALT   = MODE == kTcCount;
  constexpr int  kGrp  = ALT ? 4 : kEpiWarps;          // warps that share one tile
  constexpr int  kParts = ALT ? 1 : kEpiWarps / 4;     // ... and how many of them share one TMEM lane quarter
  constexpr int TN      = FP4 ? kTNFp4 : kTN;  // tile columns; the accumulator stages sit TN TMEM columns apart
  constexpr int kBBytes = TN * kTK;
  constexpr bool FP4C = FP4 && MODE == kTcCount;  // the count-mode extras of the fp4 tile (pre-filter, candidate list)
  static_assert(!CL || FP4, "the two-CTA cluster is wired for the fp4 count tile");
  const uint32_t rank      = CL ? clusterCtaRank() : 0u;
  const uint32_t firstUnit = CL ? blockIdx.x / 2 : blockIdx.x, unitStep = CL ? gridDim.x / 2 : gridDim.x;
  constexpr bool P2          = CL == 2;
  constexpr bool ST          = CL == 3;  // row operand stationary
  using Walk                 = UnitWalk<TN, CL != 0, ST ? kRunStat : 1>;
  constexpr int  kBStage     = P2 ? kBBytes / 2 : kBBytes;  // bytes of the column operand one CTA stages per K chunk
  constexpr int  kStageBytes = ST ? kBStage : kABytes + kBStage;
  constexpr int  kStagesTC   = ST ? kStagesStat : (P2 ? kStagesPair : (MODE == kTcCount ? (FP4 ? kStagesCountFp4 : kStagesCount) : kStagesMat));
  constexpr int  kAResident  = ST ? kMaxChunksStat * kABytes : 0;  // the stationary row tile, ahead of the ring
  __shared__ uint64_t fullBar[kStagesTC], emptyBar[kStagesTC], tmemFull[2], tmemEmpty[2];
  __shared__ uint64_t aFull[ST ? kMaxChunksStat : 1], aEmpty[ST ? kMaxChunksStat : 1];
  __shared__ uint32_t tmemBase;
  __shared__ int      popB[2][kTN];
  __shared__ int      popA[2][kTM];
  __shared__ int      colAcc[2][kTN];
  __shared__ int      popBMin[2][kEpiWarps];
  // candidates of a superposed pass wait here, per epilogue warp, and leave 64 at a time: one global atomic per flush
  // instead of one per 32 x 32 block that holds a candidate (a ~1k-clock round trip most blocks paid: with 8 pairs per
  // accumulator more than half of the blocks have a survivor; profiles/r02_path_a_summary.md)
  constexpr int kCandStage = kStagesCountFp4 > 4 ? 32 : 64;
  __shared__ int2 candStage[FP4C ? kEpiWarps : 1][FP4C ? kCandStage : 1];
  __shared__ __align__(16) float colAdj[2][FP4C ? kTN : 4];  // fp4 count tile: alpha * |B_j| (rounded down), +inf for columns past the end

  const uint32_t smemA    = (smemAddr(smemRaw) + 1023u) & ~1023u;  // (stationary tile: the row operand's K chunks)
  const uint32_t smemBase = smemA + kAResident;                    // the ring
  uint8_t*       smemGen  = smemRaw + (smemBase - smemAddr(smemRaw));
  uint16_t*      threshS  = reinterpret_cast<uint16_t*>(smemGen + kStagesTC * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tmaPrefetchDesc(&tmA);
    tmaPrefetchDesc(&tmB);
    for (int s = 0; s < kStagesTC; ++s) {
      mbarInit(&fullBar[s], 1);
      mbarInit(&emptyBar[s], (CL == 1 || ST) ? 2 : 1);  // multicast pair: the MMAs of both CTAs read what the pair's producers overwrite
    }
    if constexpr (ST)
      for (int s = 0; s < kMaxChunksStat; ++s) {
        mbarInit(&aFull[s], 1);
        mbarInit(&aEmpty[s], 1);
      }
    else if (kMmaOnly) mbarInit(&aFull[0], 1);
    for (int s = 0; s < 2; ++s) {
      mbarInit(&tmemFull[s], 1);
      mbarInit(&tmemEmpty[s], P2 ? 2 * kGrp : kGrp);  // one arrival per warp working on the tile (pair MMA: of both CTAs, at the leader)
    }
    fenceBarrierInit();
  }
  if (warp == 1) {  // TMEM: 512 columns = two 128 x 256 s32 accumulators
    if constexpr (P2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&tmemBase)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(&tmemBase)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  if constexpr (!ST)
    for (int i = threadIdx.x; i < 2 * p.threshLen; i += kThreadsTC) threshS[i] = p.thresh[i];  // table | its suffix-min
  // (stationary tile: shared memory is full of operands; the tables are read through L1 - once per row and tile plus once
  // per surviving pair)
  const uint16_t* threshT   = ST ? p.thresh : threshS;
  const uint16_t* threshLoS = threshT + p.threshLen;
  const double* __restrict__ recipS = p.recipG;  // materialise Tanimoto: RN(1/u), u = |A u B| <= 2 * bits
  // materialise modes: one [32 rows][128 B] box per epilogue warp, written with the 128-byte swizzle the output tensor map
  // expects and handed to cp.async.bulk.tensor (store)
  const uint32_t stagingAddr = (smemAddr(threshS) + 1023u) & ~1023u;
  tcFenceBefore();
  __syncthreads();
  tcFenceAfter();
  const uint32_t tmem = tmemBase;
  if constexpr (CL) clusterSync();  // the peer's barriers exist before anything is multicast at them
  if constexpr (FP4) {
    // every scale factor = 1.0: fill TMEM columns 448..511 of all 128 lanes (a warp reaches its own lane quarter)
    if (warp >= 2 && warp < 6) {
      const uint32_t q = static_cast<uint32_t>(warp & 3) * 32u;
      tmemStore32Const(tmem + 448 + (q << 16), 0x7F7F7F7Fu);
      tmemStore32Const(tmem + 480 + (q << 16), 0x7F7F7F7Fu);
    }
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && (!kMmaOnly || kMmaFed)) {
#ifdef B200_TC_TIMING
      long long tcAcc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
      int      stage = 0;
      uint32_t phase = 0;
      uint32_t aPhase = 0;
      for (Walk w(p, firstUnit, unitStep); !w.done; w.next(p)) {
        uint32_t tm, tnBeg, tnEnd;
        if (!w.coords(p, rank, tm, tnBeg, tnEnd)) continue;
        for (uint32_t tn = tnBeg; tn < tnEnd; ++tn) {
        for (int kc = 0; kc < p.kChunks; ++kc) {
          if constexpr (ST) {
            if (tn == tnBeg) {  // this unit's row tile, chunk kc: as soon as the previous unit's last tile is done with it
              mbarWait(&aEmpty[kc], aPhase ^ 1);
              mbarExpectTx(&aFull[kc], kABytes);
              tmaLoad2D(smemRaw + (smemA - smemAddr(smemRaw)) + kc * kABytes, &tmA, kc * kTK, tm * kTM, &aFull[kc]);
            }
          }
          {
            TC_T0();
            mbarWait(&emptyBar[stage], phase ^ 1);
            TC_T1(7);
          }
          uint8_t* dst = smemGen + stage * kStageBytes;
          if constexpr (P2) {
            // both CTAs' loads are counted on the LEADER's barrier (it alone issues the MMAs); each CTA stages its own
            // 128 rows and its half of the tile's columns
            constexpr int kHalfRows = TN / 2;
            if (rank == 0) mbarExpectTx(&fullBar[stage], 2 * kStageBytes);
            tmaLoad2DPair(dst, &tmA, kc * kTK, tm * kTM, &fullBar[stage]);
            tmaLoad2DPair(dst + kABytes, &tmB, kc * kTK, tn * TN + rank * kHalfRows, &fullBar[stage]);
          } else if constexpr (ST) {
            constexpr int kHalfRows = TN / 2;
            mbarExpectTx(&fullBar[stage], kBBytes);
            tmaLoad2DMulticast(dst + rank * (kHalfRows * kTK), &tmB, kc * kTK, tn * TN + rank * kHalfRows, &fullBar[stage],
                               static_cast<uint16_t>(3));
          } else {
            mbarExpectTx(&fullBar[stage], kABytes + kBBytes);
            tmaLoad2D(dst, &tmA, kc * kTK, tm * kTM, &fullBar[stage]);
            if constexpr (CL == 1) {
              // half of the shared column operand each, delivered to both CTAs (their barriers count the bytes)
              constexpr int kHalfRows = TN / 2;
              tmaLoad2DMulticast(dst + kABytes + rank * (kHalfRows * kTK), &tmB, kc * kTK, tn * TN + rank * kHalfRows,
                                 &fullBar[stage], static_cast<uint16_t>(3));
            } else {
              tmaLoad2D(dst + kABytes, &tmB, kc * kTK, tn * TN, &fullBar[stage]);
            }
          }
          if (++stage == kStagesTC) {
            stage = 0;
            phase ^= 1;
          }
        }
        }
        aPhase ^= 1;
      }
#ifdef B200_TC_TIMING
      atomicAdd(&g_tcClk[7], static_cast<unsigned long long>(tcAcc[7]));
#endif
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && !(P2 && rank != 0)) {  // (pair MMA: the leader issues for both CTAs)
#ifdef B200_TC_TIMING
      long long       tcAcc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long tcStart  = clock64();
#endif
      int      stage = 0;
      uint32_t phase = 0, local = 0;
      uint32_t aPhase = 0;
      for (Walk w(p, firstUnit, unitStep); !w.done; w.next(p)) {
        uint32_t tm, tnBeg, tnEnd;
        if (!w.coords(p, rank, tm, tnBeg, tnEnd)) continue;
        for (uint32_t tn = tnBeg; tn < tnEnd; ++tn) {
        const uint32_t as = local & 1, accPhase = (local >> 1) & 1;
        {
          TC_T0();
          if (!kMmaOnly) mbarWait(&tmemEmpty[as], accPhase ^ 1);
          TC_T1(2);
        }
        tcFenceAfter();
        const uint32_t dAddr = tmem + as * TN;
        for (int kc = 0; kc < p.kChunks; ++kc) {
          {
            TC_T0();
            if constexpr (ST) {
              if (tn == tnBeg) mbarWait(&aFull[kc], aPhase);
            }
            if (!kMmaOnly || kMmaFed) mbarWait(&fullBar[stage], phase);
            TC_T1(1);
          }
          tcFenceAfter();
          const uint32_t sAddr = smemBase + stage * kStageBytes;
          const uint64_t aDesc = makeSmemDesc(ST ? smemA + kc * kABytes : sAddr), bDesc = makeSmemDesc(ST ? sAddr : sAddr + kABytes);
#pragma unroll
          for (int k = 0; k < kTK / 32; ++k)  // K = 32 bytes per instruction: +32 B = +2 in the 16-byte address field
          {
            if constexpr (P2) ummaMxf4Pair(dAddr, aDesc + 2 * k, bDesc + 2 * k, (kc | k) != 0 ? 1u : 0u, tmem + 448, tmem + 480);
            else if constexpr (FP4) ummaMxf4(dAddr, aDesc + 2 * k, bDesc + 2 * k, (kc | k) != 0 ? 1u : 0u, tmem + 448, tmem + 480);
            else ummaI8(dAddr, aDesc + 2 * k, bDesc + 2 * k, (kc | k) != 0 ? 1u : 0u);
          }
          if constexpr (P2) ummaCommitPairMulticast(&emptyBar[stage], static_cast<uint16_t>(3));
          else if constexpr (CL == 1 || ST) ummaCommitMulticast(&emptyBar[stage], static_cast<uint16_t>(3));
          else ummaCommit(&emptyBar[stage]);  // frees the smem stage when these MMAs retire
          if constexpr (ST) {
            if (tn + 1 == tnEnd) ummaCommit(&aEmpty[kc]);  // the unit's last tile: the row tile's chunk may be replaced
          }
          if (++stage == kStagesTC) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (P2) ummaCommitPairMulticast(&tmemFull[as], static_cast<uint16_t>(3));  // both CTAs' epilogues
        else ummaCommit(&tmemFull[as]);
        ++local;
        }
        aPhase ^= 1;
      }
      if (kMmaOnly && !ST) {  // everything issued has retired before the clock is read
        ummaCommit(&aFull[0]);
        mbarWait(&aFull[0], 0);
      }
#ifdef B200_TC_TIMING
      atomicAdd(&g_tcClk[0], static_cast<unsigned long long>(clock64() - tcStart));
      atomicAdd(&g_tcClk[1], static_cast<unsigned long long>(tcAcc[1]));
      atomicAdd(&g_tcClk[2], static_cast<unsigned long long>(tcAcc[2]));
      atomicAdd(&g_tcClk[6], static_cast<unsigned long long>(local));
#endif
    }
  } else {
    // ===================== epilogue (warps 2 .. 2 + kEpiWarps) =====================
    const int      ew      = warp - 2;             // 0..kEpiWarps-1
    const int      quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int      part    = ew >> 2;              // count: which accumulator buffer this warp serves; else its share of the column blocks
    const int      gw      = ALT ? (ew & 3) : ew;  // warp index among the warps sharing the tile
    const int      et      = gw * 32 + lane;       // thread index among them
    const int      barId   = ALT ? 1 + part : 1;   // their named barrier
    uint32_t       local   = 0;
    uint32_t       tnOf[2] = {0, 0};  // tile column each accumulator-side buffer last served
    int            nStaged = 0;       // entries of candStage[ew] (the same in every lane)
    auto flushCandidates = [&]() {
      if constexpr (FP4C) {
        if (nStaged == 0) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.candCursor, static_cast<unsigned long long>(nStaged));
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int k = lane; k < nStaged; k += 32)
          if (base + k < p.candCap) p.cand[base + k] = candStage[ew][k];
        __syncwarp();
        nStaged = 0;
      }
    };
#ifdef B200_TC_TIMING
    long long       tcAcc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tcStart  = clock64();
#endif
    for (Walk w(p, firstUnit, unitStep); !w.done && !kMmaOnly; w.next(p)) {
      uint32_t tm, tnBeg, tnEnd;
      if (!w.coords(p, rank, tm, tnBeg, tnEnd)) continue;
      for (uint32_t tn = tnBeg; tn < tnEnd; ++tn) {
      const uint32_t as = local & 1, accPhase = (local >> 1) & 1;
      if (ALT && static_cast<int>(as) != part) {  // the other four warps' tile
        ++local;
        continue;
      }
      const uint32_t gr = tm * kTM + quarter * 32 + lane;
      const int      pa = gr < p.n ? __ldg(p.popX + gr) : 0;  // (issued ahead of the staging loads: one latency, not two)
      if (ALT && p.countsY && local >= 2)  // the four warps are done with this buffer's previous tile (its column counts)
        asm volatile("bar.sync %0, %1;" ::"r"(barId), "n"(32 * kGrp) : "memory");
      // stage this tile's column popcounts
      int minPb = 0x3fffffff;  // smallest |B| among this tile's valid columns (pre-filter of the threshold test)
      for (int c = et; c < TN; c += 32 * kGrp) {
        if (MODE == kTcCount && p.countsY && local >= 2) {
          // column counts this buffer collected two tiles ago
          const int v = colAcc[as][c];
          if (v) atomicAdd(p.countsY + tnOf[as] * TN + c, p.sign * v);
        }
        const uint32_t gc = tn * TN + c;
        const int      pb = gc < p.nY ? __ldg(p.popY + gc) : 0;
        popB[as][c]       = pb;
        colAcc[as][c]     = 0;
        if constexpr (FP4C) colAdj[as][c] = gc < p.nY ? __fmul_rd(p.alpha, static_cast<float>(pb)) : 3.0e38f;
        if (gc < p.nY) minPb = min(minPb, pb);
      }
      tnOf[as] = tn;
      if constexpr (MODE == kTcCount) {
#pragma unroll
        for (int o = 16; o; o >>= 1) minPb = min(minPb, __shfl_xor_sync(0xffffffffu, minPb, o));
        if (lane == 0) popBMin[as][gw] = minPb;
      }
      if (MODE != kTcCount && et < kTM) {
        const uint32_t ga = tm * kTM + et;
        popA[as][et]      = ga < p.n ? __ldg(p.popX + ga) : 0;
      }
      {
        TC_T0();
        asm volatile("bar.sync %0, %1;" ::"r"(barId), "n"(32 * kGrp) : "memory");
        TC_T1(5);
      }
      int thMin = 0;
      if constexpr (MODE == kTcCount) {
        int mpb = popBMin[as][0];
#pragma unroll
        for (int k = 1; k < kGrp; ++k) mpb = min(mpb, popBMin[as][k]);
        thMin = (gr < p.n && mpb < 0x3fffffff) ? static_cast<int>(threshLoS[pa + mpb]) : 0x3fffffff;  // no valid pair: all out
      }
      // fp4 count tile: a pair (or a superposed group of pairs) can only pass with c >= alpha (|A| + |B|) (the exact
      // threshold is the smallest integer the fp64 predicate accepts, never below alpha S - 1e-12), so the pre-filter is
      //   acc - alpha |B_j|  >=  alpha |A_i| - 1/2        (both products rounded down: conservative)
      // per column, instead of one bound from the smallest |B| of the whole tile.
      const float fRowTh = gr < p.n ? __fmul_rd(p.alpha, static_cast<float>(pa)) - 0.5f : 3.0e38f;
      (void)fRowTh;
      // every (row, column) of the tile is a pair with row fingerprints < column fingerprints: the tile's last row group
      // ends before its first column group starts
      const bool interior = !p.symmetric || (static_cast<uint64_t>(tm) * kTM + kTM) * p.superS <= static_cast<uint64_t>(tn) * TN * p.superC;
      (void)interior;
      {
        TC_T0();
        mbarWait(&tmemFull[as], accPhase);
        TC_T1(4);
      }
      tcFenceAfter();
      int rowHits = 0;
      constexpr int kCb = TN / 32, kCbPer = (kCb + kParts - 1) / kParts;  // column blocks of 32; the warps of a quarter split them
      for (int cb = ALT ? 0 : part * kCbPer; cb < (ALT ? kCb : min(kCb, (part + 1) * kCbPer)); ++cb) {
        uint32_t r[32];
        tmemLoad32(tmem + as * TN + cb * 32 + (static_cast<uint32_t>(quarter * 32) << 16), r);
        if constexpr (MODE != kTcCount) {
          if (p.outTma) {
            // Lane = output row: 32 consecutive fp64 values of ITS row, 16 at a time into the warp's staging box (128-bit
            // stores, conflict-free under the 128-byte swizzle: chunk c of row r sits at chunk c ^ (r & 7)), then one TMA
            // store of the [32 rows][16 columns] box; the tensor map clips rows >= n and columns >= nY. This replaces
            // 160 SHFL (the register transpose) + 32 scattered 256-byte stores per block: the epilogue, not HBM, bounded
            // the materialised matrix at 0.43 of the copy bandwidth.
            const uint32_t stg = stagingAddr + static_cast<uint32_t>(ew) * 4096u;
            const int      pak = popA[as][quarter * 32 + lane];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              // the 16 values of this lane's row first, in registers: they do not need the staging box, so the TMA
              // engine reads the previous box while they are computed (the wait used to come first and serialised the two)
              double v2[8][2];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j   = 16 * h + 2 * c + e;
                  const int cnt = FP4 ? __float2int_rn(__uint_as_float(r[j])) : static_cast<int>(r[j]);
                  const int pb  = popB[as][cb * 32 + j];
                  double    v   = 0.0;
                  if (cnt != 0) {
                    if constexpr (MODE == kTcTanimoto) {
                      const int    u  = pak + pb - cnt;
                      const double dc = __hiloint2double(0x43300000, cnt) - 4503599627370496.0;
                      const double du = __hiloint2double(0x43300000, u) - 4503599627370496.0;
                      const double rc = kRecipTable ? __ldg(recipS + u) : __drcp_rn(du), q0 = __dmul_rn(dc, rc);
                      v               = __fma_rn(__fma_rn(-q0, du, dc), rc, q0);
                    } else {
                      v = __ddiv_rn(static_cast<double>(cnt), __dsqrt_rn(__dmul_rn(static_cast<double>(pak), static_cast<double>(pb))));
                    }
                  }
                  v2[c][e] = v;
                }
              }
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous box has been read
              __syncwarp();
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint32_t at = stg + static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(c ^ (lane & 7)) << 4);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(at), "d"(v2[c][0]), "d"(v2[c][1]) : "memory");
              }
              fenceProxyAsync();
              __syncwarp();
              if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                               reinterpret_cast<uint64_t>(&tmOut)),
                             "r"(static_cast<int>(tn * TN + cb * 32 + 16 * h)), "r"(static_cast<int>(tm * kTM + quarter * 32)), "r"(stg)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
            continue;
          }
          // Transpose the 32 x 32 block in registers (5 butterfly rounds of SHFL) so that lane = column and k = row:
          // every store instruction then writes 32 consecutive doubles of one output row (8 full sectors) instead of
          // 32 scattered 8-byte pieces.
#pragma unroll
          for (int sft = 16; sft >= 1; sft >>= 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (!(j & sft)) {
                const bool     hi   = (lane & sft) != 0;
                const uint32_t send = hi ? r[j] : r[j + sft];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, sft);
                if (hi) r[j] = recv;
                else r[j + sft] = recv;
              }
            }
          }
          const uint32_t gc = tn * TN + cb * 32 + lane;
          const int      pb = popB[as][cb * 32 + lane];
          if (gc < p.nY) {
            const uint32_t row0 = tm * kTM + quarter * 32;
            double*        ocol = p.out + static_cast<size_t>(row0) * p.nY + gc;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              if (row0 + k >= p.n) break;
              const int c   = FP4 ? __float2int_rn(__uint_as_float(r[k])) : static_cast<int>(r[k]);
              const int pak = popA[as][quarter * 32 + k];
              double    v   = 0.0;
              if (c != 0) {
                if constexpr (MODE == kTcTanimoto) {
                  // c / u through one reciprocal + one Newton step: exhaustively verified on the CPU to equal the
                  // correctly rounded quotient for every 1 <= c <= u <= 8192 (tests/test_oracle_golden.py)
                  // int -> double through the 2^52 trick (1 DADD) and RN(1/u) from the shared-memory table
                  const int    u  = pak + pb - c;
                  const double dc = __hiloint2double(0x43300000, c) - 4503599627370496.0;
                  const double du = __hiloint2double(0x43300000, u) - 4503599627370496.0;
                  const double rc = kRecipTable ? __ldg(recipS + u) : __drcp_rn(du), q0 = __dmul_rn(dc, rc);
                  v               = __fma_rn(__fma_rn(-q0, du, dc), rc, q0);
                } else {
                  v = __ddiv_rn(static_cast<double>(c), __dsqrt_rn(__dmul_rn(static_cast<double>(pak), static_cast<double>(pb))));
                }
              }
              __stcs(ocol + static_cast<size_t>(k) * p.nY, v);
            }
          }
          continue;
        }
        // Pre-filter: the threshold grows with |A| + |B|, so c < thresh[|A| + min |B| of the tile] rules a pair out with
        // one compare; the exact table test (and the bounds / upper-triangle predicates) runs for the survivors only —
        // a handful per million pairs on fingerprint data.
        // The common case is "no survivor in these 32 columns": a max tree (31 independent-ish min/max instructions, depth
        // 5) and ONE compare decide it. Building the bit mask directly was a chain of 32 dependent compare-select-or
        // triples per block; with two epilogue warps per scheduler nothing hides that latency, and ~1,100 instructions
        // per warp and tile at one issue every ~7 clocks made the epilogue, not the MMA, pace the tile
        // (profiles/r02_path_a_summary.md).
        uint32_t maybe = 0;
        bool     hot;
        if constexpr (FP4C) {
          const float4* adj4 = reinterpret_cast<const float4*>(&colAdj[as][cb * 32]);
          float         vv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = adj4[j];  // (the same address for every lane: a broadcast read)
            vv[4 * j]     = __uint_as_float(r[4 * j]) - a.x;
            vv[4 * j + 1] = __uint_as_float(r[4 * j + 1]) - a.y;
            vv[4 * j + 2] = __uint_as_float(r[4 * j + 2]) - a.z;
            vv[4 * j + 3] = __uint_as_float(r[4 * j + 3]) - a.w;
          }
          float m[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) m[j] = fmaxf(vv[j], vv[j + 16]);
#pragma unroll
          for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
            for (int j = 0; j < w; ++j) m[j] = fmaxf(m[j], m[j + w]);
          hot = m[0] >= fRowTh;
          if (hot) {
            // 32 independent compare-selects and an OR tree: a chain of 32 dependent ORs is ~150 clocks of latency for
            // the one or two warps a scheduler has here
            uint32_t b[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) b[j] = vv[j] >= fRowTh ? (1u << j) : 0u;
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1)
#pragma unroll
              for (int j = 0; j < w; ++j) b[j] |= b[j + w];
            maybe = b[0];
          }
        } else {
          int m[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) m[j] = max(static_cast<int>(r[j]), static_cast<int>(r[j + 16]));
#pragma unroll
          for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
            for (int j = 0; j < w; ++j) m[j] = max(m[j], m[j + w]);
          hot = m[0] >= thMin;
          if (hot) {
#pragma unroll
            for (int j = 0; j < 32; ++j) maybe |= (static_cast<int>(r[j]) >= thMin ? 1u : 0u) << j;
          }
        }
        uint32_t mask = 0;
        if (p.superS * p.superC > 1) {
          // superposed operands: the accumulator is the SUM of superS * superC pair counts, so "sum below the smallest
          // threshold any pair of the group can have" rejected all of them above; what is left goes to the exact
          // verification kernel as (super row, super column)
          // (rows and columns past the end never get here: their thresholds / adjustments are +inf; only a tile the
          // diagonal crosses has to look at each survivor)
          if (FP4 && interior) mask = maybe;
          else
            while (maybe) {
              const int j = __ffs(maybe) - 1;
              maybe &= maybe - 1;
              const uint32_t gc = tn * TN + cb * 32 + j;
              if (gr < p.n && gc < p.nY &&
                  (!p.symmetric || gr * static_cast<uint32_t>(p.superS) + 1u < (gc + 1u) * static_cast<uint32_t>(p.superC)))
                mask |= 1u << j;
            }
          const unsigned holders = __ballot_sync(0xffffffffu, mask != 0);
          if (holders) {
            const int mine = __popc(mask);
            int       incl, total;
            if (__ballot_sync(0xffffffffu, mine > 1) == 0) {  // the usual case, one survivor per row: no scan needed
              incl  = __popc(holders & (0xffffffffu >> (31 - lane)));
              total = __popc(holders);
            } else {
              incl = mine;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
              }
              total = __shfl_sync(0xffffffffu, incl, 31);
            }
            if constexpr (FP4C) {
              if (total <= kCandStage) {
                if (nStaged + total > kCandStage) flushCandidates();
                int at = nStaged + incl - mine;
                while (mask) {
                  const int j = __ffs(mask) - 1;
                  mask &= mask - 1;
                  candStage[ew][at++] = make_int2(static_cast<int>(gr), static_cast<int>(tn * TN + cb * 32 + j));
                }
                nStaged += total;
                continue;
              }
            }
            // (a block with more survivors than the staging holds: straight to the list)
            unsigned long long base = 0;
            if (lane == 31) base = atomicAdd(p.candCursor, static_cast<unsigned long long>(total));
            base                  = __shfl_sync(0xffffffffu, base, 31);
            unsigned long long at = base + incl - mine;
            while (mask) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1;
              if (at < p.candCap) p.cand[at] = make_int2(static_cast<int>(gr), static_cast<int>(tn * TN + cb * 32 + j));
              ++at;
            }
          }
          continue;
        }
        while (maybe) {
          const int j = __ffs(maybe) - 1;
          maybe &= maybe - 1;
          const uint32_t gc = tn * TN + cb * 32 + j;
          bool           ok = gr < p.n && gc < p.nY;
          if (p.symmetric) ok = ok && gr < gc;
          if (ok) {
            // (dynamic register indexing is avoided: the accumulator is re-read through a shuffle-free select chain)
            int cij = 0;
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (q == j) cij = FP4 ? __float2int_rn(__uint_as_float(r[q])) : static_cast<int>(r[q]);
            if (cij >= threshT[pa + popB[as][cb * 32 + j]]) mask |= 1u << j;
          }
        }
        const unsigned any = __ballot_sync(0xffffffffu, mask != 0);
        if (any) {
          rowHits += __popc(mask);
          if (p.countsY) {
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
              const unsigned col = __ballot_sync(0xffffffffu, (mask >> j) & 1u);
              if (lane == 0 && col) atomicAdd(&colAcc[as][cb * 32 + j], __popc(col));
            }
          }
          if (p.edges) {
            const int mine = __popc(mask);
            int       incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int v = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += v;
            }
            const int          total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned long long base  = 0;
            if (lane == 31) base = atomicAdd(p.edgeCursor, static_cast<unsigned long long>(total));
            base                  = __shfl_sync(0xffffffffu, base, 31);
            unsigned long long at = base + incl - mine;
            uint32_t           m  = mask;
            while (m) {
              const int j = __ffs(m) - 1;
              m &= m - 1;
              if (at < p.edgeCap) p.edges[at] = make_int2(static_cast<int>(gr), static_cast<int>(tn * TN + cb * 32 + j));
              ++at;
            }
          }
        }
      }
      tcFenceBefore();
      __syncwarp();
      if (lane == 0) {  // accumulator may be overwritten
        if (P2 && rank != 0) mbarArriveRemote(&tmemEmpty[as], 0);
        else mbarArrive(&tmemEmpty[as]);
      }
      if (MODE == kTcCount && rowHits) atomicAdd(p.counts + gr, p.sign * rowHits);
      ++local;
      }
    }
#ifdef B200_TC_TIMING
    if (ew == 0 && lane == 0) {
      atomicAdd(&g_tcClk[3], static_cast<unsigned long long>(clock64() - tcStart));
      atomicAdd(&g_tcClk[4], static_cast<unsigned long long>(tcAcc[4]));
      atomicAdd(&g_tcClk[5], static_cast<unsigned long long>(tcAcc[5]));
    }
#endif
    flushCandidates();
    if (MODE != kTcCount && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this warp's TMA stores
    if (MODE == kTcCount && p.countsY && static_cast<int>(local) > part) {  // the column counts of this buffer's last tile
      asm volatile("bar.sync %0, %1;" ::"r"(barId), "n"(32 * kGrp) : "memory");
      for (int c = et; c < TN; c += 32 * kGrp) {
        const int v = colAcc[part][c];
        if (v) atomicAdd(p.countsY + tnOf[part] * TN + c, p.sign * v);
      }
    }
  }
  tcFenceBefore();
  __syncthreads();
  if constexpr (CL) clusterSync();  // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) {
    tcFenceAfter();
    if constexpr (P2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace

void launchRowPopcount(const uint32_t* fp, size_t n, int words, int32_t* pop, cudaStream_t s);
void launchThreshTable(int maxS, double cutoff, uint16_t* thresh, cudaStream_t s);

// Count / materialise modes on tensor cores. Returns false when the problem shape is not eligible (caller uses the SIMT tile).
#ifdef B200_TC_TIMING
extern "C" void b200mol_debug_clocks_tc(unsigned long long* out8) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_tcClk, sizeof(g_tcClk));
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_tcClk, z, sizeof(z));
}
#endif

int g_superpose = 4;      // fingerprints summed into one row of the count pass (option "similarity_superpose": 1, 2 or 4)
int g_superposeCols = 4;  // ... and into one column ("similarity_superpose_cols": 1, 2 or 4); 4 x 4 sums stay <= 16, exact
int g_superposeLast = 0;  // pairs per accumulator the last graph pass really ran with (1 after an overflow fallback)

constexpr int kMaxPipeline = 8;
int g_pipelineChunks = 4;  // chunks of a superposed pass whose verification overlaps the next chunk's tensor pass
                           // (option "similarity_pipeline_chunks"; 1 = off)
int g_superposeAuto = 1;  // 1: a large graph pass picks rows x cols from a pilot over one row group ("similarity_superpose_auto")
unsigned long long g_candidatesLast = 0;  // candidates the last superposed pass listed ("similarity_candidates_last")

static bool launchTensorImpl(SimMode mode, const SimLaunch& q, cudaStream_t s, int superS, int superC, bool* overflow,
                             unsigned long long* pilotCand = nullptr);

// Count / materialise modes on tensor cores. Returns false when the problem shape is not eligible (caller uses the SIMT tile).
bool launchSimilarityTensor(SimMode mode, const SimLaunch& q, cudaStream_t s) {
  // Superposition serves the Butina passes (symmetric and / or with an edge list): they synchronise for their edge
  // total anyway, and the superposed pass needs one host read of its candidate count. The plain thresholded count
  // (b200mol_tanimoto_count_ge) stays asynchronous on the unsuperposed tile.
  const bool graphPass = mode == kCountTanimoto && (q.symmetric || q.edges != nullptr);
  if (!graphPass || g_superpose * g_superposeCols == 1) {
    if (graphPass) g_superposeLast = 1;
    return launchTensorImpl(mode, q, s, 1, 1, nullptr);
  }
  int S = g_superpose, C = g_superposeCols;
  // How far superposition pays depends on the data: the sum of S C random intersections must stay below the threshold
  // ONE true neighbour pair reaches, or every accumulator is a candidate. A pilot over a prefix sample of the
  // fingerprints (at most 1/64 of the pairs; candidates listed but nothing counted) measures the candidate rate of each
  // column factor, widest first; the model
  //   time(S, C) = pairs / (S C) * tPair + candidates * S C * tVerify
  // (tile and verify rates measured on B200, profiles/r02_path_a_summary.md) picks the cheapest. A narrower factor can
  // only win while the wider one spends more time verifying than multiplying, so the search stops as soon as it does not.
  if (g_superposeAuto && q.nX >= 8 * static_cast<size_t>(kGroupRows) && C > 1) {
    const double nX = static_cast<double>(q.nX), nY = static_cast<double>(q.nY);
    SimLaunch pilot   = q;
    pilot.groupOffset = 0;
    pilot.groupStride = 1;
    pilot.nX          = std::min<size_t>(32768, std::max<size_t>(2 * kGroupRows, q.nX / 8));
    pilot.nY          = q.symmetric ? pilot.nX : std::min<size_t>(q.nY, std::max<size_t>(pilot.nX, q.nY / 8));
    const double sX = static_cast<double>(pilot.nX), sY = static_cast<double>(pilot.nY);
    const double pilotPairs = q.symmetric ? sX * (sX - 1) / 2.0 : sX * sY;
    const double totalPairs = (q.symmetric ? nX * (nX - 1) / 2.0 : nX * nY) / (q.groupStride < 1 ? 1 : q.groupStride);
    constexpr double tPair = 0.85e-12, tVerify = 0.1e-9;  // seconds per unsuperposed pair / per verified pair
    double bestT = totalPairs * tPair;  // unsuperposed
    int    bestS = 1, bestC = 1;
    for (int c = C; c >= 1; c >>= 1) {
      unsigned long long got = 0;
      if (!launchTensorImpl(mode, pilot, s, S, c, nullptr, &got)) return false;
      const double tPass = totalPairs / (S * c) * tPair;
      const double tVer  = static_cast<double>(got) * totalPairs / pilotPairs * S * c * tVerify;
      if (tPass + tVer < bestT) bestT = tPass + tVer, bestS = S, bestC = c;
      if (tVer <= tPass) break;
    }
    S = bestS, C = bestC;
  }
  while (S * C > 1) {
    bool overflow = false;
    if (!launchTensorImpl(mode, q, s, S, C, &overflow)) return false;
    g_superposeLast = S * C;
    if (!overflow) return true;  // else: the candidate list overflowed (dense graph / loose cutoff), nothing was counted yet
    if (C > 1) C = 1;            // second chance: rows only
    else S = 1;
  }
  g_superposeLast = 1;
  return launchTensorImpl(mode, q, s, 1, 1, nullptr);
}

static bool launchTensorImpl(SimMode mode, const SimLaunch& q, cudaStream_t s, int superS, int superC, bool* overflow,
                             unsigned long long* pilotCand) {
  if (mode == kCountCosine) return false;
  const int bits = q.words * 32;
  if (bits % kTK != 0 || bits > 4096) return false;
  const bool same = (q.x == q.y && q.nX == q.nY);
  if (q.symmetric && !same) return false;

  // block-scaled fp4 operands (twice the int8 MMA rate, half the operand bytes through shared memory) when the
  // fingerprint is a whole number of 256-bit chunks, in every mode; else the int8 tile
  const bool count = mode == kCountTanimoto;
  const bool fp4   = g_tensorFp4 && bits % (2 * kTK) == 0;
  const int  tn    = fp4 ? kTNFp4 : kTN;
  const int  rowBytes = fp4 ? bits / 2 : bits;  // bytes of one expanded fingerprint
  if (!fp4 || !count) superS = superC = 1;  // the superposed sums need the fp4 value set {0..4}; only the count mode verifies
  const bool   super  = superS * superC > 1;
  const size_t nSuper = (q.nX + superS - 1) / superS;   // rows of the X operand
  const size_t nSuperY = (q.nY + superC - 1) / superC;  // rows of the Y operand (tile columns)

  TcParams p{};
  p.n         = static_cast<uint32_t>(nSuper);
  p.nY        = static_cast<uint32_t>(nSuperY);
  p.kChunks   = rowBytes / kTK;
  p.tilesM    = static_cast<uint32_t>((nSuper + kTM - 1) / kTM);
  p.superS    = superS;
  p.superC    = superC;
  p.rowSpan   = static_cast<uint32_t>(kTM * superS);
  p.colSpan   = static_cast<uint32_t>(tn * superC);
  p.groupTiles = static_cast<uint32_t>(kGroupRows / (kTM * superS));
  p.tilesN    = static_cast<uint32_t>((nSuperY + tn - 1) / tn);
  p.symmetric = q.symmetric ? 1 : 0;
  p.groupOffset = q.groupOffset;
  p.groupStride = q.groupStride < 1 ? 1 : q.groupStride;
  p.sign      = q.sign;
  p.counts    = q.rowCounts;
  p.countsY   = q.symmetric ? q.rowCounts : nullptr;
  p.edges     = q.edges;
  p.edgeCursor = q.edgeCursor;
  p.edgeCap   = q.edgeCap;
  p.out       = q.out;
  p.recipLen  = 2 * bits;
  Scratch<double> recip(mode == kMaterialiseTanimoto ? static_cast<size_t>(p.recipLen) + 1 : 0, s);
  if (mode == kMaterialiseTanimoto) {
    recipTableKernel<<<(p.recipLen + 256) / 256, 256, 0, s>>>(recip.get(), p.recipLen);
    B200_LAUNCHED();
    p.recipG = recip.get();
  }
  {
    // a pair passes iff 1 - c / (|A| + |B| - c) <= cutoff, i.e. c >= alpha (|A| + |B|); the pre-filter's alpha is rounded
    // DOWN (and its products too), so it never rejects what the exact fp64 table accepts
    const double a  = q.cutoff < 2.0 ? (1.0 - q.cutoff) / (2.0 - q.cutoff) : 0.0;
    float        af = static_cast<float>(a);
    if (static_cast<double>(af) > a) af = nextafterf(af, -1.0f);
    p.alpha = nextafterf(af, -1.0f);  // (one more ulp: fp64 rounding inside the table's predicate)
  }

  // 0/1 expansion of the fingerprints: bytes (2 KB per 2048-bit row) or packed fp4 (1 KB); a superposed operand is the
  // sum of superS (superC) consecutive expansions. X and Y share one buffer when they are the same set, summed alike.
  const bool       ownY = !same || superS != superC;
  Scratch<uint8_t> expX(nSuper * static_cast<size_t>(rowBytes), s);
  Scratch<uint8_t> expYown(ownY ? nSuperY * static_cast<size_t>(rowBytes) : 0, s);
  auto expand = [&](const uint32_t* src, size_t rows, int S, size_t superRows, uint8_t* dst) {
    const size_t nw = superRows * static_cast<size_t>(q.words);
    const auto   grid = static_cast<unsigned>((nw + 255) / 256);
    if (S > 1) expandBitsFp4SuperKernel<<<grid, 256, 0, s>>>(src, rows, q.words, S, superRows, reinterpret_cast<uint4*>(dst));
    else if (fp4) expandBitsFp4Kernel<<<grid, 256, 0, s>>>(src, nw, reinterpret_cast<uint4*>(dst));
    else expandBitsKernel<<<grid, 256, 0, s>>>(src, nw, reinterpret_cast<uint4*>(dst));
    B200_LAUNCHED();
  };
  expand(q.x, q.nX, superS, nSuper, expX.get());
  if (ownY) expand(q.y, q.nY, superC, nSuperY, expYown.get());
  const uint8_t* expY = ownY ? expYown.get() : expX.get();

  Scratch<int32_t> popX(q.nX, s), popYown(same ? 0 : q.nY, s);
  Scratch<int32_t> popSuper(superS > 1 ? nSuper : 0, s), popSuperY(superC > 1 && ownY ? nSuperY : 0, s);
  launchRowPopcount(q.x, q.nX, q.words, popX.get(), s);
  if (!same) launchRowPopcount(q.y, q.nY, q.words, popYown.get(), s);
  const int32_t* popYExact = same ? popX.get() : popYown.get();
  p.popX = popX.get();
  p.popY = popYExact;
  if (superS > 1) {
    superMinPopKernel<<<static_cast<unsigned>((nSuper + 255) / 256), 256, 0, s>>>(popX.get(), q.nX, superS, nSuper, popSuper.get());
    B200_LAUNCHED();
    p.popX = popSuper.get();
  }
  if (superC > 1) {
    if (ownY) {
      superMinPopKernel<<<static_cast<unsigned>((nSuperY + 255) / 256), 256, 0, s>>>(popYExact, q.nY, superC, nSuperY, popSuperY.get());
      B200_LAUNCHED();
      p.popY = popSuperY.get();
    } else p.popY = popSuper.get();
  }
  // candidates of a superposed pass: (super row, super column) pairs the exact kernel re-examines. Sized for the
  // neighbour graphs this pass is used on (tens of edges per point); a denser graph overflows it and the caller falls back.
  unsigned long long          candCap = 0;
  Scratch<int2>               cand;
  Scratch<unsigned long long> candCursor;
  if (super) {
    const unsigned long long all = static_cast<unsigned long long>(nSuper) * nSuperY;
    candCap                      = std::min<unsigned long long>(all, std::max<unsigned long long>(1ull << 22, 64ull * q.nX));
    cand                         = Scratch<int2>(candCap, s);
    candCursor                   = Scratch<unsigned long long>(kMaxPipeline, s);
    B200_CUDA(cudaMemsetAsync(candCursor.get(), 0, kMaxPipeline * sizeof(unsigned long long), s));
    p.cand = cand.get(), p.candCursor = candCursor.get(), p.candCap = candCap;
  }
  const int         maxS = 2 * bits;
  Scratch<uint16_t> thresh(2 * static_cast<size_t>(maxS + 1), s);
  if (mode == kCountTanimoto) {
    launchThreshTable(maxS, q.cutoff, thresh.get(), s);
    threshSuffixMinKernel<<<1, 1024, 0, s>>>(thresh.get(), maxS + 1, thresh.get() + maxS + 1);
    B200_LAUNCHED();
    p.threshLen = maxS + 1;
  }
  p.thresh = thresh.get();

  CUtensorMap tmA, tmB;
  makeTensorMap2D(&tmA, expX.get(), nSuper, rowBytes, kTM, kTK, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1);
  const bool cluster = fp4 && count && g_tensorCluster != 0;  // CTA pairs: 1 = multicast column operand, 2 = cta_group::2 MMAs,
  const bool pairMma = cluster && g_tensorCluster == 2;  // 3 = multicast column operand + stationary row operand
  const bool stationary = cluster && g_tensorCluster == 3 && p.kChunks <= kMaxChunksStat;
  makeTensorMap2D(&tmB, expY, nSuperY, rowBytes, cluster ? tn / 2 : tn, kTK, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1);

  CUtensorMap tmOut = tmA;  // (unused in the count mode)
  if (!count) {
    p.outTma = (q.nY % 2 == 0 && (reinterpret_cast<uintptr_t>(q.out) & 15) == 0) ? 1 : 0;
    if (p.outTma) makeTensorMap2D(&tmOut, q.out, q.nX, q.nY, 32, 16, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8);
  }
  const size_t smemBytes =
    stationary ? static_cast<size_t>(kMaxChunksStat) * kABytes + static_cast<size_t>(kStagesStat) * tn * kTK + 1024 + 64
               : (pairMma ? static_cast<size_t>(kStagesPair) * (kABytes + tn / 2 * kTK)
                          : static_cast<size_t>(count ? (fp4 ? kStagesCountFp4 : kStagesCount) : kStagesMat) * (kABytes + tn * kTK)) +
                   (count ? static_cast<size_t>(maxS + 1) * 4 : static_cast<size_t>(1024 + kEpiWarpsMat * 4096)) + 1024 + 64;
  // each variant may use what its static shared memory leaves of the 227 KB a CTA can have
  static bool configured[kMaxDevices] = {};
  auto optIn = [](auto kernel) {
    cudaFuncAttributes a{};
    B200_CUDA(cudaFuncGetAttributes(&a, kernel));
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - static_cast<int>(a.sharedSizeBytes)));
  };
  if (!configured[currentDeviceSlot()]) {
    optIn(simTensorKernel<kTcCount, false, 0>);
    optIn(simTensorKernel<kTcCount, true, 0>);
    optIn(simTensorKernel<kTcCount, true, 1>);
    optIn(simTensorKernel<kTcCount, true, 2>);
    optIn(simTensorKernel<kTcCount, true, 3>);
    optIn(simTensorKernel<kTcTanimoto, false, 0>);
    optIn(simTensorKernel<kTcCosine, false, 0>);
    optIn(simTensorKernel<kTcTanimoto, true, 0>);
    optIn(simTensorKernel<kTcCosine, true, 0>);
    configured[currentDeviceSlot()] = true;
  }
  constexpr size_t kStaticMax = kStagesCountFp4 > 4 ? 9472 : 12288;  // static shared memory of the largest variant
  B200_REQUIRE(smemBytes + kStaticMax <= 227 * 1024, "tensor similarity tile does not fit shared memory");
  // units a call owns: tiles, or vertical tile pairs (same enumeration as the kernel's UnitWalk)
  auto unitsOf = [&](const TcParams& pk) -> uint64_t {
    return stationary ? countUnits<kTNFp4, true, kRunStat>(pk)
           : cluster  ? countUnits<kTNFp4, true>(pk)
                      : (fp4 ? countUnits<kTNFp4, false>(pk) : countUnits<kTN, false>(pk));
  };
  // the count kernel over the row groups `pk` selects (false: none of them is owned by this call)
  auto launchCount = [&](const TcParams& pk) -> bool {
    const uint64_t units = unitsOf(pk);
    if (units == 0) return false;
    if (cluster) {
      cudaLaunchConfig_t cfg{};
      cudaLaunchAttribute attr[1];
      attr[0].id               = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.blockDim         = dim3(threadsTC(kTcCount));
      cfg.dynamicSmemBytes = smemBytes;
      cfg.stream           = s;
      cfg.attrs            = attr;
      cfg.numAttrs         = 1;
      cfg.gridDim          = dim3(2);
      int maxClusters = 0;
      if (stationary) B200_CUDA(cudaOccupancyMaxActiveClusters(&maxClusters, simTensorKernel<kTcCount, true, 3>, &cfg));
      else if (pairMma) B200_CUDA(cudaOccupancyMaxActiveClusters(&maxClusters, simTensorKernel<kTcCount, true, 2>, &cfg));
      else B200_CUDA(cudaOccupancyMaxActiveClusters(&maxClusters, simTensorKernel<kTcCount, true, 1>, &cfg));
      B200_REQUIRE(maxClusters >= 1, "no CTA pair fits the device");
      uint64_t pairs = maxClusters;  // persistent: one resident cluster per schedulable SM pair
      if (pairs > units) pairs = units;
      cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
      if (stationary) B200_CUDA(cudaLaunchKernelEx(&cfg, simTensorKernel<kTcCount, true, 3>, tmA, tmB, tmOut, pk));
      else if (pairMma) B200_CUDA(cudaLaunchKernelEx(&cfg, simTensorKernel<kTcCount, true, 2>, tmA, tmB, tmOut, pk));
      else B200_CUDA(cudaLaunchKernelEx(&cfg, simTensorKernel<kTcCount, true, 1>, tmA, tmB, tmOut, pk));
    } else {
      const int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(smCount()), units));
      if (fp4) simTensorKernel<kTcCount, true, 0><<<grid, threadsTC(kTcCount), smemBytes, s>>>(tmA, tmB, tmOut, pk);
      else simTensorKernel<kTcCount, false, 0><<<grid, threadsTC(kTcCount), smemBytes, s>>>(tmA, tmB, tmOut, pk);
    }
    B200_LAUNCHED();
    return true;
  };
  auto launchVerify = [&](const int2* list, unsigned long long nCand, cudaStream_t on) {
    PhaseTimer         t("verify_candidates", on);
    const unsigned int blocks2 = static_cast<unsigned int>(std::min<unsigned long long>((nCand + 63) / 64, static_cast<unsigned long long>(smCount()) * 16));
    verifyCandidatesKernel<<<blocks2, 256, 0, on>>>(q.x, q.y, q.words, list, nCand, superS, superC, static_cast<uint32_t>(q.nX),
                                                    static_cast<uint32_t>(q.nY), q.symmetric ? 1 : 0, popX.get(), popYExact,
                                                    thresh.get(), q.sign, q.rowCounts, q.symmetric ? q.rowCounts : nullptr, q.edges,
                                                    q.edgeCursor, q.edgeCap);
    B200_LAUNCHED();
  };

  // Superposed pass in a PIPELINE of K chunks of the row groups (chunk k = groups k, k + K, ... of this call's): the exact
  // verification of chunk k runs on a second stream while the tensor pass of chunk k + 1 has the SMs - a pass CTA leaves
  // room for one verify block per SM - so only the last chunk's verification is exposed (8 ms of a 59 ms step were).
  // A chunk whose candidate list overflowed is redone on its own with fewer pairs per accumulator after the others.
  const uint64_t ownedGroups = ((q.nX + kGroupRows - 1) / kGroupRows + p.groupStride - 1) / p.groupStride;
  const int      K = (super && !pilotCand && g_pipelineChunks > 1 && ownedGroups >= 4ull * g_pipelineChunks) ? g_pipelineChunks : 1;
  if (K > 1) {
    static cudaStream_t side[kMaxDevices] = {};
    cudaStream_t&       s2 = side[currentDeviceSlot()];
    if (!s2) B200_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    const unsigned long long capK = candCap / K;
    cudaEvent_t              ev[kMaxPipeline];
    {
      PhaseTimer t("neighbor_pass_tc", s);  // the K tensor passes, back to back on the caller's stream
      for (int k = 0; k < K; ++k) {
        TcParams pk    = p;
        pk.groupOffset = p.groupOffset + static_cast<uint32_t>(k) * p.groupStride;
        pk.groupStride = static_cast<uint32_t>(K) * p.groupStride;
        pk.cand        = cand.get() + static_cast<size_t>(k) * capK;
        pk.candCursor  = candCursor.get() + k;
        pk.candCap     = capK;
        launchCount(pk);
        B200_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
        B200_CUDA(cudaEventRecord(ev[k], s));
      }
    }
    unsigned long long listed = 0;
    int                redo[kMaxPipeline], nRedo = 0;
    for (int k = 0; k < K; ++k) {
      unsigned long long nk = 0;
      B200_CUDA(cudaStreamWaitEvent(s2, ev[k], 0));
      B200_CUDA(cudaMemcpyAsync(&nk, candCursor.get() + k, sizeof(nk), cudaMemcpyDeviceToHost, s2));
      B200_CUDA(cudaStreamSynchronize(s2));  // (waits for chunk k's pass and for the verifications queued before it)
      listed += nk;
      if (nk > capK) redo[nRedo++] = k;
      else if (nk) launchVerify(cand.get() + static_cast<size_t>(k) * capK, nk, s2);
      cudaEventDestroy(ev[k]);
    }
    cudaEvent_t done;
    B200_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    B200_CUDA(cudaEventRecord(done, s2));
    B200_CUDA(cudaStreamWaitEvent(s, done, 0));  // the caller's stream continues after the last verification
    cudaEventDestroy(done);
    g_candidatesLast = listed;
    for (int r = 0; r < nRedo; ++r) {
      SimLaunch qk   = q;
      qk.groupOffset = p.groupOffset + static_cast<uint32_t>(redo[r]) * p.groupStride;
      qk.groupStride = static_cast<uint32_t>(K) * p.groupStride;
      bool again     = false;
      if (!launchTensorImpl(mode, qk, s, superS, 1, &again)) return false;
      if (again && !launchTensorImpl(mode, qk, s, 1, 1, nullptr)) return false;
    }
    return true;
  }

  if (unitsOf(p) == 0) return true;  // nothing owned by this rank (more ranks than row groups)
  int blocks = smCount();
  if (static_cast<uint64_t>(blocks) > unitsOf(p)) blocks = static_cast<int>(unitsOf(p));
  if (mode == kCountTanimoto) {
    PhaseTimer t("neighbor_pass_tc", s);
    launchCount(p);
  } else if (mode == kMaterialiseTanimoto) {
    PhaseTimer t("cross_tc", s);
    if (fp4) simTensorKernel<kTcTanimoto, true, 0><<<blocks, threadsTC(kTcTanimoto), smemBytes, s>>>(tmA, tmB, tmOut, p);
    else simTensorKernel<kTcTanimoto, false, 0><<<blocks, threadsTC(kTcTanimoto), smemBytes, s>>>(tmA, tmB, tmOut, p);
  } else {
    PhaseTimer t("cross_tc", s);
    if (fp4) simTensorKernel<kTcCosine, true, 0><<<blocks, threadsTC(kTcCosine), smemBytes, s>>>(tmA, tmB, tmOut, p);
    else simTensorKernel<kTcCosine, false, 0><<<blocks, threadsTC(kTcCosine), smemBytes, s>>>(tmA, tmB, tmOut, p);
  }
  if (mode != kCountTanimoto) B200_LAUNCHED();
  if (super) {
    // the one host read of a superposed pass: how many candidates (the callers synchronise for their edge total anyway)
    unsigned long long nCand = 0;
    B200_CUDA(cudaMemcpyAsync(&nCand, candCursor.get(), sizeof(nCand), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    if (pilotCand) {  // dry run over one row group: the candidate count is the result
      *pilotCand = nCand;
      return true;
    }
    g_candidatesLast = nCand;
    if (nCand > candCap) {
      if (overflow) *overflow = true;  // nothing has been counted yet: the caller reruns without superposition
      return true;
    }
    if (nCand) launchVerify(cand.get(), nCand, s);
  }
  return true;
}

}  // namespace b200
