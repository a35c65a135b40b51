// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
//
// Host-side shim that lets the reference's own CUDA-C kernel strings
// (/root/reference/models/voting.py:4-172, extracted verbatim at build time by
// oracle/build_ref.py, never copied into this repo) compile with plain g++ and run
// on the CPU.  It supplies the handful of CUDA built-ins the strings use
// (blockIdx/blockDim/threadIdx, __global__, atomicAdd, float max) and nothing else;
// float3/make_float3 come from the CUDA toolkit's host headers, vector operators
// from the reference's own helper_math.cuh (include line 27 re-pointed to
// <cuda_runtime.h> in a temp copy).
#pragma once
#include <cmath>
#include <math.h>
#include <cuda_runtime.h>

#ifndef __global__
#define __global__
#endif

struct OracleDim3 { unsigned x, y, z; };
static thread_local OracleDim3 blockIdx = {0, 0, 0};
static thread_local OracleDim3 threadIdx = {0, 0, 0};
static thread_local OracleDim3 blockDim = {1, 1, 1};

// The reference header first (temp copy, on the include path), so that its own
// uint/int min/max calls bind before the float overloads below become visible.
#include "helper_math.cuh"

// CUDA resolves max(float,float) to the float overload; helper_math.cuh's host
// branch only declares the int one, which would truncate probabilities.
static inline float max(float a, float b) { return a > b ? a : b; }
static inline float min(float a, float b) { return a < b ? a : b; }

static inline float atomicAdd(float* addr, float v) {
    float old;
#pragma omp atomic capture
    { old = *addr; *addr += v; }
    return old;
}
