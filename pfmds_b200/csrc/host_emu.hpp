// TEST SUPPORT ONLY — included by common.cuh when the device sources are compiled by a plain C++ compiler (never by nvcc, never
// by the product build).  tests/*_host.cpp compile a .cu file of this directory for the host and run its kernels as ordinary
// functions, one call per (block, thread), to check the kernels' arithmetic and indexing against the CPU oracle without a GPU:
//   - threadIdx / blockIdx / blockDim / gridDim are globals set by emu_launch();
//   - the threads of a block run one after the other, from the last to thread 0, so the running total kept by block_sum()
//     reaches thread 0 — the only thread that uses it in these kernels — last;
//   - kernels that exchange data between the lanes of a warp (shuffles with SPLIT > 1, ballots, shared-memory scans) cannot be
//     run this way; the thread-per-atom variants can.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct double4 { double x, y, z, w; };
struct float4 { float x, y, z, w; };
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline int atomicCAS(int* a, int cmp, int val) { int old = *a; if (old == cmp) *a = val; return old; }
static inline int atomicAdd(int* a, int v) { int old = *a; *a += v; return old; }
// separately rounded product / sum: the host harnesses are built with -ffp-contract=off, so plain operators do that
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
static inline void __syncthreads() {}
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }  // only reached with SPLIT == 1 (no iterations) in the emulated kernels

static inline double4 ld256(const double4* p) { return *p; }
static inline double4 ld256_nc(const double4* p) { return *p; }

// running per-block totals, one per block_sum() call site order within a thread
static double emu_block_acc[64];
static int emu_block_call[2048];
static inline void emu_block_begin() {
    for (double& a : emu_block_acc) a = 0.;
    for (int& c : emu_block_call) c = 0;
}
static inline double block_sum(double v) {
    int c = emu_block_call[threadIdx.x]++;
    emu_block_acc[c] += v;
    return emu_block_acc[c];
}

template <class K, class... A>
static void emu_launch(K kernel, unsigned gx, unsigned gy, unsigned threads, A... args) {
    gridDim.x = gx; gridDim.y = gy; blockDim.x = threads;
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) {
            blockIdx.x = bx; blockIdx.y = by;
            emu_block_begin();
            for (int t = (int)threads - 1; t >= 0; --t) {
                threadIdx.x = (unsigned)t;
                kernel(args...);
            }
        }
}
