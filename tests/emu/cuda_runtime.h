// Stand-in for <cuda_runtime.h> used ONLY by the CPU test suite (tests/emu/): it lets g++
// compile the device headers of libsphmw (csrc/pair_list.cuh, wcsph_ops.cuh, kernels_sph.cuh,
// sphmw_internal.h) so that the kernels' logic — traversal order, pair list, closures — can be
// run thread by thread on the host and compared with the oracle.  Test infrastructure: the
// product never includes it.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define SPHMW_EMU 1
#define __host__
#define __device__
#define __global__
#define __shared__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 {
    unsigned x, y, z;
};
struct float4 {
    float x, y, z, w;
};
struct double2 {
    double x, y;
};
extern uint3 threadIdx, blockIdx, blockDim, gridDim;  // set by the harness before every "thread"

typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef int cudaError_t;
#define cudaSuccess 0
inline const char *cudaGetErrorString(cudaError_t) { return "emulation"; }

// dynamic shared memory of k_binary_build: one queue for the "block" being emulated
extern uint32_t nl_queue[];
inline uint32_t *nl_queue_emu() { return nl_queue; }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)((const char *)p - (const char *)nl_queue); }

inline uint32_t __ldcs(const uint32_t *p) { return *p; }
inline void __stcs(uint32_t *p, uint32_t v) { *p = v; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    unsigned long long o = *p;
    *p = o + v;
    return o;
}
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) {
    uint32_t o = *p;
    *p = o + v;
    return o;
}
// every emulated thread is its own "warp" for the pair counter
inline unsigned __activemask() { return 1u << (threadIdx.x & 31); }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
inline int __ffs(unsigned m) { return __builtin_ffs((int)m); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
