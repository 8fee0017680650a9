// TEST SUPPORT ONLY — included by common.cuh when the device sources are compiled by a plain C++ compiler (never by nvcc, never
// by the product build: pfmds_b200/build.py compiles everything with nvcc and libpfmds_b200.so has no CPU path).
// The test suite compiles the .cu files of this directory for the host and runs their kernels as ordinary functions, one call
// per (block, thread), to check arithmetic, indexing and the host-side orchestration against the CPU oracle without a GPU:
//   - threadIdx / blockIdx / blockDim / gridDim are (thread-local) globals set by the launch loop behind LAUNCH();
//   - the threads of a block run one after the other, from the last to thread 0, so the running totals kept by block_sum()
//     reach thread 0 — the only thread that uses them in these kernels — last;
//   - the CUDA runtime calls the library makes are mapped onto malloc / memcpy / no-ops (one "device", synchronous "streams");
//   - kernels that exchange data between the lanes of a warp (shuffles with SPLIT > 1, ballots, scans) cannot be run this way:
//     the SERIAL flavour of the emulated library takes the thread-per-atom variants (SMALL_N = 0, no warp-per-atom list build)
//     and the few shuffle-based reductions have a serial twin under #ifndef PFMDS_COOP;
//   - the LOCK-STEP flavour (-DPFMDS_EMU_WARP, below) runs every thread of a block as a fiber and implements the warp
//     collectives and block barriers between them, so those kernels run as they are.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <time.h>

#include <functional>
#include <vector>

#define PFMDS_HOST_EMU 1
// dynamic shared memory of a launch (kernels that declare `extern __shared__` take this buffer instead)
static inline std::vector<unsigned char>& emu_dyn_buf() { static thread_local std::vector<unsigned char> b; return b; }
static inline void emu_dynamic_smem(size_t n) { if (emu_dyn_buf().size() < n) emu_dyn_buf().resize(n); }
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __noinline__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local

struct double4 { double x, y, z, w; };
struct double2 { double x, y; };
struct float4 { float x, y, z, w; };
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
// one instance per process (C++17 inline variables): non-static inline device functions of the shared headers are merged across
// translation units by the linker and must all see the same indices
inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline int atomicCAS(int* a, int cmp, int val) { int old = *a; if (old == cmp) *a = val; return old; }
static inline int atomicAdd(int* a, int v) { int old = *a; *a += v; return old; }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { unsigned old = *a; *a += v; return old; }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { unsigned long long old = *a; *a += v; return old; }
static inline int atomicMax(int* a, int v) { int old = *a; if (v > old) *a = v; return old; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
static inline int __double2loint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) { unsigned long long v = ((unsigned long long)(unsigned int)hi << 32) | (unsigned int)lo; double d; memcpy(&d, &v, 8); return d; }
// ranks of the slab decomposition are OS threads in the replay: fences are real, a sleeping spin loop gives up its time slice
#include <sched.h>
#include <atomic>
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { sched_yield(); }
// separately rounded product / sum: the emulated builds use -ffp-contract=off, so plain operators do that
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
// ---- stream capture: while a capture is open, launches and asynchronous copies / fills are recorded instead of executed, and a
// graph launch replays them (arguments by value, as CUDA does) -- so the library's CUDA-graph replay of steady-state steps and
// its host-side bookkeeping run in the CPU suite too
struct emu_graph_s { std::vector<std::function<void()>> ops; };
inline thread_local emu_graph_s* emu_capture = nullptr;

#ifndef PFMDS_EMU_WARP
static inline void __syncthreads() {}
// only reached with SPLIT == 1 (no iterations) or in code paths the emulated build never launches
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
#endif

static inline double4 ld256(const double4* p) { return *p; }
static inline double4 ld256_nc(const double4* p) { return *p; }

#ifndef PFMDS_EMU_WARP
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
static void emu_launch_now(K kernel, dim3 grid, dim3 block, A... args) {
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
#else
// ---- PFMDS_EMU_WARP: lock-step replay ---------------------------------------------------------------------------------------
// Every thread of a block is a fiber (ucontext) on the calling OS thread.  A fiber runs until it reaches a warp collective
// (__shfl_*_sync, __ballot_sync, __syncwarp) or a block barrier (__syncthreads) and yields; when every live lane of its warp
// (every live thread of the block) waits at the same kind of point the scheduler exchanges the values and resumes them.  This
// is what lets the CPU suite run the kernels the GPU actually launches for small systems: 8 lanes per atom with shuffle-tree
// sums, the warp-per-atom list build with ballot / popc compaction, the block scans and the shared-memory reductions.
// Reading a shuffle source lane that has exited returns a poison pattern (on hardware the value is undefined); lanes of one
// warp waiting at different kinds of points, or a barrier that part of the block can never reach, abort with a message.
#include <functional>
#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <vector>

enum { EMU_READY = 0, EMU_WARP = 1, EMU_BLOCK = 2, EMU_DONE = 3 };
enum { EMU_OP_IDX = 0, EMU_OP_UP, EMU_OP_DOWN, EMU_OP_XOR, EMU_OP_BALLOT, EMU_OP_SYNCWARP };
#if defined(__x86_64__) && !defined(PFMDS_EMU_UCONTEXT)
// Fiber switch without the two rt_sigprocmask system calls of swapcontext (a reduction kernel yields ~10^7 times): saves the
// callee-saved registers and the FP control words on the current stack, swaps stack pointers, restores.  System V x86-64 only.
#define EMU_FAST_SWITCH 1
extern "C" void emu_ctx_switch(void** save_sp, void* load_sp);
asm(".text\n.weak emu_ctx_switch\n.type emu_ctx_switch,@function\nemu_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n  subq $8, %rsp\n  stmxcsr (%rsp)\n  fnstcw 4(%rsp)\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  ldmxcsr (%rsp)\n  fldcw 4(%rsp)\n  addq $8, %rsp\n  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_ctx_switch,.-emu_ctx_switch\n");
struct EmuFiber { void* sp; int state, op, arg; unsigned long long val, out; };
inline thread_local void* emu_sched_sp = nullptr;
#define EMU_TO_SCHED(f) emu_ctx_switch(&(f).sp, emu_sched_sp)
#define EMU_TO_FIBER(f) emu_ctx_switch(&emu_sched_sp, (f).sp)
#else
struct EmuFiber { ucontext_t ctx; int state, op, arg; unsigned long long val, out; };
#define EMU_TO_SCHED(f) swapcontext(&(f).ctx, &emu_sched)
#define EMU_TO_FIBER(f) swapcontext(&emu_sched, &(f).ctx)
#endif
inline thread_local std::vector<EmuFiber>* emu_fib = nullptr;
inline thread_local std::vector<char*>* emu_stacks = nullptr;
inline thread_local ucontext_t emu_sched;
inline thread_local int emu_cur = 0;
inline thread_local std::function<void()>* emu_body = nullptr;
static const size_t EMU_STACK = 256 * 1024;
static const unsigned long long EMU_POISON = 0x7ff8deadbeefdeadull;  // a NaN as double, a large negative number as int

static void emu_tramp() {
    (*emu_body)();
    (*emu_fib)[(size_t)emu_cur].state = EMU_DONE;
    EMU_TO_SCHED((*emu_fib)[(size_t)emu_cur]);
    abort();  // a finished fiber is never resumed
}
static inline unsigned long long emu_yield(int state, int op, unsigned long long v, int arg) {
    EmuFiber& f = (*emu_fib)[(size_t)emu_cur];
    f.state = state; f.op = op; f.val = v; f.arg = arg;
    EMU_TO_SCHED(f);
    return f.out;
}
template <class T> static inline unsigned long long emu_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, "shuffle of a wide type"); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T emu_unbits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_unbits<T>(emu_yield(EMU_WARP, EMU_OP_IDX, emu_bits(v), lane)); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_unbits<T>(emu_yield(EMU_WARP, EMU_OP_UP, emu_bits(v), d)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_unbits<T>(emu_yield(EMU_WARP, EMU_OP_DOWN, emu_bits(v), d)); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_unbits<T>(emu_yield(EMU_WARP, EMU_OP_XOR, emu_bits(v), m)); }
static inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)emu_yield(EMU_WARP, EMU_OP_BALLOT, pred ? 1ull : 0ull, 0); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_yield(EMU_WARP, EMU_OP_SYNCWARP, 0, 0); }
static inline void __syncthreads() { emu_yield(EMU_BLOCK, 0, 0, 0); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }

static void emu_fail(const char* what) {
    fprintf(stderr, "host replay (lock-step): %s in block (%u,%u)\n", what, blockIdx.x, blockIdx.y);
    abort();
}
// all live lanes of warp [w0, w1) wait at a warp collective: exchange and resume
static void emu_resolve_warp(std::vector<EmuFiber>& F, int w0, int w1) {
    int op = -1;
    for (int t = w0; t < w1; ++t)
        if (F[(size_t)t].state == EMU_WARP) {
            if (op < 0) op = F[(size_t)t].op;
            else if (op != F[(size_t)t].op) emu_fail("lanes of one warp wait at different collectives");
        }
    unsigned ballot = 0;
    for (int t = w0; t < w1; ++t)
        if (F[(size_t)t].state == EMU_WARP && F[(size_t)t].val) ballot |= 1u << (t - w0);
    for (int t = w0; t < w1; ++t) {
        EmuFiber& f = F[(size_t)t];
        if (f.state != EMU_WARP) continue;
        const int l = t - w0;
        int src = l;
        switch (op) {
        case EMU_OP_IDX: src = f.arg & 31; break;
        case EMU_OP_UP: src = l - f.arg; break;
        case EMU_OP_DOWN: src = l + f.arg; break;
        case EMU_OP_XOR: src = l ^ f.arg; break;
        default: break;
        }
        if (op == EMU_OP_BALLOT) f.out = ballot;
        else if (op == EMU_OP_SYNCWARP) f.out = 0;
        else if (src < 0 || src > 31) f.out = f.val;                                  // out of range: own value (defined)
        else if (w0 + src >= w1 || F[(size_t)(w0 + src)].state != EMU_WARP) f.out = EMU_POISON;  // exited lane: undefined on hardware
        else f.out = F[(size_t)(w0 + src)].val;
    }
    for (int t = w0; t < w1; ++t)
        if (F[(size_t)t].state == EMU_WARP) F[(size_t)t].state = EMU_READY;
}
static void emu_run_block(int T) {
    if (!emu_fib) { emu_fib = new std::vector<EmuFiber>(); emu_stacks = new std::vector<char*>(); }
    std::vector<EmuFiber>& F = *emu_fib;
    if ((int)F.size() < T) F.resize((size_t)T);
    while ((int)emu_stacks->size() < T) {
        void* m = mmap(nullptr, EMU_STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) emu_fail("no memory for fiber stacks");
        emu_stacks->push_back((char*)m);
    }
    for (int t = 0; t < T; ++t) {
        EmuFiber& f = F[(size_t)t];
#ifdef EMU_FAST_SWITCH
        // initial frame as emu_ctx_switch leaves it: [mxcsr | x87 cw][r15 r14 r13 r12 rbx rbp][return address = emu_tramp]
        char* top = (*emu_stacks)[(size_t)t] + EMU_STACK;          // 16-byte aligned (mmap)
        void** slot = (void**)(top - 16);
        *slot = (void*)emu_tramp;                                   // `ret` lands there with rsp = top - 8, as after a call
        for (int k = 1; k <= 6; ++k) slot[-k] = nullptr;
        unsigned int* cw = (unsigned int*)(top - 16 - 56);
        cw[0] = 0x1f80u; cw[1] = 0x037fu;                           // default MXCSR and x87 control word
        f.sp = (void*)cw;
#else
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = (*emu_stacks)[(size_t)t];
        f.ctx.uc_stack.ss_size = EMU_STACK;
        f.ctx.uc_link = &emu_sched;
        makecontext(&f.ctx, emu_tramp, 0);
#endif
        f.state = EMU_READY;
    }
    for (;;) {
        bool ran = false;
        for (int t = 0; t < T; ++t)
            if (F[(size_t)t].state == EMU_READY) {
                emu_cur = t;
                threadIdx.x = (unsigned)t;
                EMU_TO_FIBER(F[(size_t)t]);
                ran = true;
            }
        int live = 0, at_block = 0;
        for (int t = 0; t < T; ++t) { live += F[(size_t)t].state != EMU_DONE; at_block += F[(size_t)t].state == EMU_BLOCK; }
        if (live == 0) return;
        bool released = false;
        for (int w0 = 0; w0 < T; w0 += 32) {
            const int w1 = w0 + 32 < T ? w0 + 32 : T;
            int nw = 0, nb = 0;
            for (int t = w0; t < w1; ++t) { nw += F[(size_t)t].state == EMU_WARP; nb += F[(size_t)t].state == EMU_BLOCK; }
            if (nw && nb) emu_fail("lanes of one warp wait at a warp collective and at __syncthreads");
            if (nw) { emu_resolve_warp(F, w0, w1); released = true; }
        }
        if (!released && at_block == live) {
            for (int t = 0; t < T; ++t)
                if (F[(size_t)t].state == EMU_BLOCK) F[(size_t)t].state = EMU_READY;
            released = true;
        }
        if (!ran && !released) emu_fail("deadlock");
    }
}
template <class K, class... A>
static void emu_launch_now(K kernel, dim3 grid, dim3 block, A... args) {
    gridDim = grid; blockDim = block;
#ifdef PFMDS_EMU_TRACE
    fprintf(stderr, "launch %s grid (%u,%u) block %u\n", __PRETTY_FUNCTION__, grid.x, grid.y, block.x);
#endif
    std::function<void()> body = [&] { kernel(args...); };
    std::function<void()>* saved = emu_body;
    emu_body = &body;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx.x = bx; blockIdx.y = by;
#ifdef PFMDS_EMU_TRACE
            if (bx < 3 || bx + 2 > grid.x) fprintf(stderr, "  block %u\n", bx);
#endif
            emu_run_block((int)block.x);
        }
#ifdef PFMDS_EMU_TRACE
    fprintf(stderr, "  done\n");
#endif
    emu_body = saved;
}
#endif  // PFMDS_EMU_WARP
template <class K, class... A>
static void emu_launch_cfg(K kernel, dim3 grid, dim3 block, A... args) {
    if (emu_capture) emu_capture->ops.push_back([=] { emu_launch_now(kernel, grid, block, args...); });
    else emu_launch_now(kernel, grid, block, args...);
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
typedef emu_graph_s* cudaGraph_t;
typedef emu_graph_s* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "not available in the host replay"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { if (emu_capture) return cudaErrorEmu; *p = (T*)calloc(bytes ? bytes : 1, 1); return *p ? cudaSuccess : cudaErrorEmu; }
template <class T> static inline cudaError_t cudaMallocAsync(T** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (emu_capture) return cudaErrorEmu; memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    if (emu_capture) emu_capture->ops.push_back([=] { memmove(d, s, n); });
    else memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    if (emu_capture) emu_capture->ops.push_back([=] { memset(d, v, n); });
    else memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)calloc(1, 8); return cudaSuccess; }
// as on the device: synchronising, allocating or copying synchronously while a capture is open is an error
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return emu_capture ? cudaErrorEmu : cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)calloc(1, 8); return cudaSuccess; }
enum { cudaEventBlockingSync = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
// "inter-process" handles of the slab decomposition: the ranks of the replay share one address space, a handle is the pointer
struct cudaIpcMemHandle_t { char reserved[64]; };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
// events carry the host clock: elapsed times of the replay are wall times of the serial loops (never zero)
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    if (emu_capture) return cudaSuccess;
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    *(double*)e = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return emu_capture ? cudaErrorEmu : cudaSuccess; }
// (the replay runs every launch at once, in call order: a cross-stream dependency is already satisfied when it is declared)
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { double d = *(double*)b - *(double*)a; *ms = (float)(d > 1e-6 ? d : 1e-6); return cudaSuccess; }
// CUDA graphs: a capture records closures (above), an executable graph is a copy of the list, a launch runs it
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { if (emu_capture) return cudaErrorEmu; emu_capture = new emu_graph_s; return cudaSuccess; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = emu_capture; emu_capture = nullptr; return *g ? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { *e = new emu_graph_s(*g); return cudaSuccess; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { for (auto& op : e->ops) op(); return cudaSuccess; }
