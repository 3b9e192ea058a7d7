// Optional per-phase CUDA-event timing inside the library (bench.py's roofline needs the duration of the dominant
// kernel alone, measured on the stream it is launched on). Off by default; b200mol_profile_enable(1) turns it on.
#pragma once
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace b200 {

struct PhaseEvents {
  cudaEvent_t start = nullptr, stop = nullptr;
  bool        recorded = false;
};
extern bool                               g_profileOn;
extern std::mutex                         g_profileMutex;
extern std::map<std::string, PhaseEvents> g_phases;

// RAII: records start now and stop at scope exit on `s` (only when profiling is enabled).
struct PhaseTimer {
  PhaseEvents* ev = nullptr;
  cudaStream_t s;
  PhaseTimer(const char* name, cudaStream_t stream) : s(stream) {
    if (!g_profileOn) return;
    std::lock_guard<std::mutex> lock(g_profileMutex);
    PhaseEvents&                e = g_phases[name];
    if (!e.start) {
      cudaEventCreate(&e.start);
      cudaEventCreate(&e.stop);
    }
    ev = &e;
    cudaEventRecord(e.start, s);
  }
  ~PhaseTimer() {
    if (ev) {
      cudaEventRecord(ev->stop, s);
      ev->recorded = true;
    }
  }
};

}  // namespace b200
