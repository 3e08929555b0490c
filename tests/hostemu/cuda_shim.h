// tests/hostemu/cuda_shim.h -- TEST INFRASTRUCTURE.  Minimal stand-ins that let g++ compile the
// __global__ bodies of fbpic_b200/csrc/b2_ext_kernels.cuh for the CPU: a launch becomes four nested loops
// over (blockIdx, threadIdx).  Only kernels without shared memory / warp intrinsics can be emulated
// this way; the product never links this.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#ifndef __restrict__
#define __restrict__
#endif

#define __grid_constant__
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct emu_dim3 { unsigned x, y, z; emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;

// explicit round-to-nearest mul/add: compile this unit with -ffp-contract=off
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
// launches are sequential loops here: a plain read-modify-write
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    const unsigned long long old = *p;
    *p = old + v;
    return old;
}

#define EMU_LAUNCH(grid, block, kernel, ...)                                             \
    do {                                                                                 \
        gridDim = (grid); blockDim = (block);                                            \
        for (unsigned by_ = 0; by_ < gridDim.y; ++by_)                                   \
        for (unsigned bx_ = 0; bx_ < gridDim.x; ++bx_)                                   \
        for (unsigned ty_ = 0; ty_ < blockDim.y; ++ty_)                                  \
        for (unsigned tx_ = 0; tx_ < blockDim.x; ++tx_) {                                \
            blockIdx = emu_dim3(bx_, by_); threadIdx = emu_dim3(tx_, ty_);               \
            kernel(__VA_ARGS__);                                                         \
        }                                                                                \
    } while (0)
