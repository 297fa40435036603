// hostsim.h — TEST INFRASTRUCTURE ONLY.  Lets the device headers (mp.cuh, curve.cuh, kernels.cuh, keygroup.cuh)
// compile with plain g++ so that the limb arithmetic, the group law and the thread-per-item kernels can be run
// one "thread" at a time on the CPU and compared with Python big integers (tests/test_hostsim.py, -m "not gpu").
// It is never part of libsbv.so: the product has no CPU path.  Only defined when SBV_HOSTSIM is set and the
// compiler is not nvcc.
#pragma once
#if defined(SBV_HOSTSIM) && !defined(__CUDACC__)
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__
#define __constant__ static const
#define __launch_bounds__(...)

struct hostsim_dim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local hostsim_dim3 threadIdx, blockIdx, blockDim, gridDim;
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << (sh & 31)) >> 32);
}
static inline int __ffs(int x) { return x ? __builtin_ctz((unsigned)x) + 1 : 0; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}
// one simulated thread at a time: a "ballot" sees only the caller (kernels that need a real warp are not simulated)
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __activemask() { return 1u; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
typedef int cudaError_t;
typedef void *cudaStream_t;
#endif
