// Host stand-in for <cuda_runtime.h>: just enough of the CUDA device vocabulary for the kernels of
// porthamiltonian_fem_b200/csrc to compile with g++ and run inside the SPMD emulator (emu_rt.hpp).  Test infrastructure only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "emu_rt.hpp"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static thread_local          /* one OS thread per CTA: thread_local == per-CTA */
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { unsigned int x, y; };
struct uint4 { unsigned int x, y, z, w; };
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }

using phw_emu::threadIdx;
using phw_emu::blockIdx;
using phw_emu::blockDim;
using phw_emu::gridDim;

static inline void __syncthreads() { phw_emu::syncthreads(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) { return phw_emu::shfl_xor(v, lane_mask); }
static inline double __hiloint2double(int hi, int lo) {
    unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double d; std::memcpy(&d, &u, 8); return d;
}
static inline long long clock64() { return phw_emu::clock(); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
