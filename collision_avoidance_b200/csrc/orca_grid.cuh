// orca_grid.cuh -- uniform-grid neighbor pipeline for large worlds (agents_per_env > 256).
// (first slice: interface only; the pipeline lands in a follow-up commit)
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "orca_step_small.cuh"

namespace orca {

struct GridScratch {
  int dummy = 0;
};

inline void grid_free(GridScratch&) {}

inline int launch_grid_step(GridScratch&, const StepArgs&, int, cudaStream_t, int64_t*, std::string* err) {
  *err = "agents_per_env > 256 needs the uniform-grid pipeline (not built yet)";
  return -3;
}

}  // namespace orca
