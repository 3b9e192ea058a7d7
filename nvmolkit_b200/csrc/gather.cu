// Result exchange at the end of a molecule-sharded conformer job: the ONE collective of the path, through
// NCCL over NVLink / NVSwitch, in the C-ABI so that a C++ caller needs no torch. Replaces the reference's peer copies
// into one target GPU (DeviceCoordCollector::finalizeOnTarget, src/conformer/device_coord_collector.cpp:30-145,
// copyDeviceToDeviceAsync src/utils/p2p.cpp:56-86).
//
// NCCL is not linked: the entry points are looked up at run time (dlsym) in whatever NCCL the process has already loaded -
// torch's bundled copy inside Python, the system libnccl.so.2 otherwise - so the communicator the caller passes and the
// functions called on it always belong to the same library.
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "common.cuh"

namespace b200 {
namespace {

struct Nccl {
  ncclResult_t (*allGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)      = nullptr;
  ncclResult_t (*broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*groupStart)()                                                                         = nullptr;
  ncclResult_t (*groupEnd)()                                                                           = nullptr;
  const char* (*errorString)(ncclResult_t)                                                             = nullptr;
  ncclResult_t (*commCount)(const ncclComm_t, int*)                                                    = nullptr;
  ncclResult_t (*commUserRank)(const ncclComm_t, int*)                                                 = nullptr;
};

const Nccl& nccl() {
  static const Nccl api = [] {
    Nccl  a;
    void* h = RTLD_DEFAULT;
    if (!dlsym(h, "ncclAllGather")) {
      h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h) fail(B200MOL_ERR_CUDA, "NCCL is not loaded in this process and libnccl.so.2 cannot be opened: %s", dlerror());
    }
    auto get = [&](const char* name) {
      void* p = dlsym(h, name);
      if (!p) fail(B200MOL_ERR_CUDA, "NCCL symbol %s not found", name);
      return p;
    };
    a.allGather    = reinterpret_cast<decltype(a.allGather)>(get("ncclAllGather"));
    a.broadcast    = reinterpret_cast<decltype(a.broadcast)>(get("ncclBroadcast"));
    a.groupStart   = reinterpret_cast<decltype(a.groupStart)>(get("ncclGroupStart"));
    a.groupEnd     = reinterpret_cast<decltype(a.groupEnd)>(get("ncclGroupEnd"));
    a.errorString  = reinterpret_cast<decltype(a.errorString)>(get("ncclGetErrorString"));
    a.commCount    = reinterpret_cast<decltype(a.commCount)>(get("ncclCommCount"));
    a.commUserRank = reinterpret_cast<decltype(a.commUserRank)>(get("ncclCommUserRank"));
    return a;
  }();
  return api;
}

#define B200_NCCL(expr)                                                                                          \
  do {                                                                                                           \
    ncclResult_t r_ = (expr);                                                                                    \
    if (r_ != ncclSuccess) ::b200::fail(B200MOL_ERR_CUDA, "%s failed: %s", #expr, nccl().errorString(r_));       \
  } while (0)

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200mol_allgather_counts(void* nccl_comm, int64_t n_conf_local, int64_t n_atoms_local, int64_t* h_conf_counts,
                                        int64_t* h_atom_counts, void* stream) {
  return guarded([&] {
    B200_REQUIRE(nccl_comm && h_conf_counts && h_atom_counts, "null pointer");
    B200_REQUIRE(n_conf_local >= 0 && n_atoms_local >= 0, "negative count");
    ncclComm_t   comm = static_cast<ncclComm_t>(nccl_comm);
    cudaStream_t s    = asStream(stream);
    int          world = 0;
    B200_NCCL(nccl().commCount(comm, &world));
    Scratch<int64_t> d(static_cast<size_t>(2) * (world + 1), s);
    const int64_t    mine[2] = {n_conf_local, n_atoms_local};
    B200_CUDA(cudaMemcpyAsync(d.get(), mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    B200_NCCL(nccl().allGather(d.get(), d.get() + 2, 2, ncclInt64, comm, s));
    std::vector<int64_t> all(static_cast<size_t>(2) * world);
    B200_CUDA(cudaMemcpyAsync(all.data(), d.get() + 2, all.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    for (int r = 0; r < world; ++r) {
      h_conf_counts[r] = all[2 * r];
      h_atom_counts[r] = all[2 * r + 1];
    }
  });
}

extern "C" int b200mol_allgather_results(void* nccl_comm, const int64_t* h_conf_counts, const int64_t* h_atom_counts,
                                         const double* d_positions_local, const int32_t* d_conf_atoms_local,
                                         const double* d_energy_local, const int8_t* d_converged_local, double* d_positions,
                                         int32_t* d_conf_atoms, double* d_energy, int8_t* d_converged, void* stream) {
  return guarded([&] {
    B200_REQUIRE(nccl_comm && h_conf_counts && h_atom_counts, "null pointer");
    ncclComm_t   comm = static_cast<ncclComm_t>(nccl_comm);
    cudaStream_t s    = asStream(stream);
    int          world = 0, rank = 0;
    B200_NCCL(nccl().commCount(comm, &world));
    B200_NCCL(nccl().commUserRank(comm, &rank));
    // exact sizes, no padding: one broadcast per rank and array, all inside one NCCL group (a single fused launch)
    B200_NCCL(nccl().groupStart());
    int64_t confOff = 0, atomOff = 0;
    for (int r = 0; r < world; ++r) {
      const int64_t nc = h_conf_counts[r], na = h_atom_counts[r];
      B200_REQUIRE(nc >= 0 && na >= 0, "negative count for rank %d", r);
      const bool me = r == rank;
      if (na > 0 && d_positions)
        B200_NCCL(nccl().broadcast(me ? d_positions_local : nullptr, d_positions + 3 * atomOff, static_cast<size_t>(3 * na), ncclFloat64, r, comm, s));
      if (nc > 0) {
        if (d_conf_atoms)
          B200_NCCL(nccl().broadcast(me ? d_conf_atoms_local : nullptr, d_conf_atoms + confOff, static_cast<size_t>(nc), ncclInt32, r, comm, s));
        if (d_energy)
          B200_NCCL(nccl().broadcast(me ? d_energy_local : nullptr, d_energy + confOff, static_cast<size_t>(nc), ncclFloat64, r, comm, s));
        if (d_converged)
          B200_NCCL(nccl().broadcast(me ? d_converged_local : nullptr, d_converged + confOff, static_cast<size_t>(nc), ncclInt8, r, comm, s));
      }
      confOff += nc;
      atomOff += na;
    }
    B200_NCCL(nccl().groupEnd());
  });
}
