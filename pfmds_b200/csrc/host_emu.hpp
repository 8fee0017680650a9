// TEST SUPPORT ONLY — included by common.cuh when the device sources are compiled by a plain C++ compiler (never by nvcc, never
// by the product build: pfmds_b200/build.py compiles everything with nvcc and libpfmds_b200.so has no CPU path).
// The test suite compiles the .cu files of this directory for the host and runs their kernels as ordinary functions, one call
// per (block, thread), to check arithmetic, indexing and the host-side orchestration against the CPU oracle without a GPU:
//   - threadIdx / blockIdx / blockDim / gridDim are (thread-local) globals set by the launch loop behind LAUNCH();
//   - the threads of a block run one after the other, from the last to thread 0, so the running totals kept by block_sum()
//     reach thread 0 — the only thread that uses them in these kernels — last;
//   - the CUDA runtime calls the library makes are mapped onto malloc / memcpy / no-ops (one "device", synchronous "streams");
//   - kernels that exchange data between the lanes of a warp (shuffles with SPLIT > 1, ballots, scans) cannot be run this way:
//     the library's emulated build takes the thread-per-atom variants (SMALL_N = 0, no warp-per-atom list build) and the few
//     shuffle-based reductions have a serial twin under #ifndef __CUDACC__.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PFMDS_HOST_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __noinline__
#define __shared__ static thread_local

struct double4 { double x, y, z, w; };
struct float4 { float x, y, z, w; };
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline int atomicCAS(int* a, int cmp, int val) { int old = *a; if (old == cmp) *a = val; return old; }
static inline int atomicAdd(int* a, int v) { int old = *a; *a += v; return old; }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { unsigned old = *a; *a += v; return old; }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { unsigned long long old = *a; *a += v; return old; }
static inline int atomicMax(int* a, int v) { int old = *a; if (v > old) *a = v; return old; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
static inline int __double2loint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v & 0xffffffffll); }
// separately rounded product / sum: the emulated builds use -ffp-contract=off, so plain operators do that
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline void __syncthreads() {}
// only reached with SPLIT == 1 (no iterations) or in code paths the emulated build never launches
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }

static inline double4 ld256(const double4* p) { return *p; }
static inline double4 ld256_nc(const double4* p) { return *p; }

// running per-block totals, one per block_sum() call in the order a thread makes them
static thread_local double emu_block_mx;
static thread_local double emu_block_acc[64];
static thread_local int emu_block_call[2048];
static inline void emu_block_begin() {
    for (double& a : emu_block_acc) a = 0.;
    for (int& c : emu_block_call) c = 0;
    emu_block_mx = -1.0e300;
}
static inline double emu_block_max(double v) {  // running maximum of the block (one call site per kernel), complete at thread 0
    if (v > emu_block_mx) emu_block_mx = v;
    return emu_block_mx;
}
static inline double block_sum(double v) {
    int c = emu_block_call[threadIdx.x]++;
    emu_block_acc[c] += v;
    return emu_block_acc[c];
}

template <class K, class... A>
static void emu_launch_cfg(K kernel, dim3 grid, dim3 block, A... args) {
    gridDim = grid; blockDim = block;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx.x = bx; blockIdx.y = by;
            emu_block_begin();
            for (int t = (int)block.x - 1; t >= 0; --t) {
                threadIdx.x = (unsigned)t;
                kernel(args...);
            }
        }
}
#define LAUNCH(kernel, grid, block, stream, ...) emu_launch_cfg(kernel, dim3(grid), dim3(block), __VA_ARGS__)
// the standalone harnesses (tests/forces_host.cpp, tests/nl_host.cpp) launch with explicit grid sizes
template <class K, class... A>
static void emu_launch(K kernel, unsigned gx, unsigned gy, unsigned threads, A... args) { emu_launch_cfg(kernel, dim3(gx, gy), dim3(threads), args...); }

// ---- CUDA runtime calls made by the library, on host memory ----------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef struct emu_stream_s* cudaStream_t;
typedef struct emu_event_s* cudaEvent_t;
typedef struct emu_graph_s* cudaGraph_t;
typedef struct emu_graphexec_s* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "not available in the host replay"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { *p = (T*)calloc(bytes ? bytes : 1, 1); return *p ? cudaSuccess : cudaErrorEmu; }
template <class T> static inline cudaError_t cudaMallocAsync(T** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)calloc(1, 8); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)calloc(1, 8); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
// CUDA graphs are never used by the emulated build (use_graphs is forced off); the calls only have to compile
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorEmu; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorEmu; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorEmu; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorEmu; }
