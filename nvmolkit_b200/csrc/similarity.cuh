// Internal interface of the N x M popcount-similarity tile (used by tanimoto.cu and butina.cu).
#pragma once
#include <cstdint>

#include "common.cuh"

namespace b200 {

// What the tile kernel does with the 128x128 integer intersection counts.
enum SimMode : int {
  kMaterialiseTanimoto = 0,  // fp64 similarity matrix
  kMaterialiseCosine   = 1,
  kCountTanimoto       = 2,  // thresholded neighbour counts (+ optional edge list)
  kCountCosine         = 3,
};

struct SimLaunch {
  const uint32_t* x      = nullptr;  // [nX][words]
  const uint32_t* y      = nullptr;  // [nY][words]
  size_t          nX     = 0;
  size_t          nY     = 0;
  int             words  = 0;
  // materialise
  double* out = nullptr;  // [nX][nY]
  // count
  double   cutoff      = 0.0;      // neighbour iff 1 - sim <= cutoff (fp64)
  int      sign        = 1;        // counts += sign * hits
  int32_t* rowCounts   = nullptr;  // [nX]
  bool     symmetric   = false;    // x == y: visit tiles tn >= tm only, skip the diagonal pairs, update both endpoints
  uint32_t groupOffset = 0;        // multi-GPU row sharding: this call covers tile-row groups (32 x 128 rows each)
  uint32_t groupStride = 1;        //   groupOffset, groupOffset + groupStride, ...
  int2*    edges       = nullptr;  // optional (symmetric only): (i<j) neighbour pairs appended here
  unsigned long long* edgeCursor = nullptr;  // device counter; entries beyond edgeCap are dropped (cursor still advances)
  unsigned long long  edgeCap    = 0;
};

// Launch the tile kernel(s). Row popcounts and the threshold table are computed internally (stream-ordered scratch).
void launchSimilarity(SimMode mode, const SimLaunch& p, cudaStream_t stream);

}  // namespace b200
