// d2d_plan.cuh -- Primitive planner kernels (placeholder until the A* kernel lands)
#pragma once
#include "d2d_state.cuh"
#define D2D_PLAN_SLOTS 1
__host__ __device__ inline size_t d2d_plan_workspace_bytes(int n_u) { return 16; }
