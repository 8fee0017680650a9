// pfmds_b200 — device-side data model shared by all kernels (sm_100a, FP64).
//
// Layout in HBM (one context = one simulation):
//   pos[N]  double4 {x,y,z,scratch}   32-byte records: one sector per neighbour gather
//   vel[N]  double4 {vx,vy,vz,mass}   the kick/KE kernels get the mass with the velocity
//   frc[N]  double4 {fx,fy,fz,-}
//   gmask[N] uint32  bit g-1 set when the atom is in settings-file group g
//   orig[N]  int     file-order index of the atom stored at this slot
// Atoms are stored in cell order (re-sorted whenever every neighbour list is rebuilt) so that the
// j-atoms of a warp's i-atoms sit in a few contiguous ranges; `orig` maps back for I/O.
// A neighbour list is an ELL block: nlist[p*stride + i] (stride = N rounded up to 32) so that the
// 32 lanes of a warp read slot p of 32 consecutive atoms with one coalesced request; entries are
// slot indices, only int32 indices are stored and distances are recomputed in registers every
// step (the reference caches dr(3,max,N) and |dr| instead: md_general.f90:30-35).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
// every kernel launch of the library goes through this macro (the host replay of the test suite substitutes a serial loop)
#define LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#else
#include "host_emu.hpp"  // tests/emu/host_emu.hpp (the test builds add -I tests/emu): these sources compiled for the host by the test suite; the product build is nvcc only
#endif

#if defined(__CUDACC__) || defined(PFMDS_EMU_WARP)
#define PFMDS_COOP 1  // warp-cooperative code (shuffles, ballots, shared-memory reductions) is compiled: nvcc, or the lock-step host replay
#endif
#if defined(__CUDACC__) || defined(PFMDS_EMU_LIB)
#define PFMDS_HAVE_CTX 1  // the context and the launch wrappers exist: the product (nvcc) and the emulated library of the test suite
#endif

#include "mathx.cuh"

#define PFMDS_MAX_GROUPS 32
#define PFMDS_GHOST 0x80000000u  // gmask bit of a ghost copy (slab decomposition); group 32 is unavailable in that mode
#define PFMDS_ERRW 4  // error word: code, detail a, detail b, spare

// the reference's FP64 literals (md_general.f90:165,304; md_integrators.f90:62,211; cut_off_function.f90:8)
#define PFMDS_MASS_COEF (1.6605389217 / 1.6021765654 * 100.0)
#define PFMDS_KB (1.3806488 / 1.6021765654 * 1.0e-4)
#define PFMDS_PI 3.14159265358979

enum { K_LJ = 0, K_LJ1G = 1, K_LJC = 2, K_MORSEC = 3, K_TB = 4, K_RJL = 5, K_REBOSC = 6 };

struct BoxD { double L[3]; double h[3]; };

// device error codes mirror include/pfmds_b200.h
enum { E_OUT_OF_CELL = 10, E_TOO_MANY = 11, E_GR_NEIB = 12 };

__device__ __forceinline__ void raise_error(int* err, int code, int a, int b) {
    if (atomicCAS(&err[0], 0, code) == 0) { err[1] = a; err[2] = b; }
}

// minimum image, bit-identical to dr - half*(sign(1,dr-half)+sign(1,dr+half))  (md_general.f90:423-441)
__device__ __forceinline__ double min_image(double d, double half, double L) {
    if (d >= half) d -= L;
    else if (d < -half) d += L;
    return d;
}

// cosine switch and its derivative divided by r (cut_off_function.f90:6-28); called only for r < R2
__device__ __forceinline__ void fcut_dfcut(double r, double R1, double R2, double& f, double& dfr) {
    if (r < R1) { f = 1.0; dfr = 0.0; return; }
    double s, c;
    mx::sincos_0pi(PFMDS_PI * (r - R1) / (R2 - R1), s, c);
    f = (1.0 + c) / 2;
    dfr = -s * PFMDS_PI / (R2 - R1) / r / 2;
}
__device__ __forceinline__ double fcut_only(double r, double R1, double R2) {
    if (r < R1) return 1.0;
    double s, c;
    mx::sincos_0pi(PFMDS_PI * (r - R1) / (R2 - R1), s, c);
    return (1.0 + c) / 2;
}
#ifdef __CUDACC__
// One 256-bit load per 32-byte record (sm_100a LDG.E.256): a gather of 32 records costs the L1 one
// request instead of the two LDG.128 the compiler emits for a double4.
__device__ __forceinline__ double4 ld256_nc(const double4* p) {  // read-only data
    double4 v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double4 ld256(const double4* p) {
    double4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
#endif  // __CUDACC__

#ifdef PFMDS_COOP
// block-wide sum of `v` (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

#endif  // PFMDS_COOP

// parameters of one interaction, passed by value to the kernels
struct LJp { double eps, sig, R1, R2; };
struct LJ1Gp { double R1, R2, c6, c12, c6t6, c12t12; };
struct LJCp { double eps, sig, delt, R1, R2; int simplified; };          // ljc
struct MORp { double d, r, a, delt, R1, R2; int simplified; };           // morsec
struct TBp { double d, s, b, r0, delt, a0, c02, d02, R1, R2; };
struct RJLp { double A0, xi, p, q, r0, R1, R2; };
struct REBp { double A, Q, alpha, B[3], beta[3], T, g[6], R1, R2; };  // rebosc, REBOsolidcarbon.f90:6-8

#define NHC_MAXF 4
struct NhcPack { int n; uint32_t bit[NHC_MAXF]; double* state[NHC_MAXF]; int M[NHC_MAXF]; int L[NHC_MAXF]; double T[NHC_MAXF]; };

// ---- Nose-Hoover chain half step: md_integrators.f90:200-245 ---------------------------------------
// state = x[M], v[M], q[M], s, ke_cached, s_pending.  Returns the velocity scale s = exp(-v1 dt/2).
// NOT inlined: the kernels that run it (k_nhc, k_nhc_open, k_nhc_close, k_sum_kick_ke, k_sl_ke_close) must produce the same bits, and
// inlined into different surroundings the compiler is free to contract a*b + c*d into an fma either way round and to hoist
// loop-invariant products (seen in round 2: chain x, v one ulp apart between two kernels that inlined it).
static __device__ __noinline__ double nhc_chain(double* state, int M, int L, double temperature, double ke, double ts2, double ts3, double ts4) {
    double* x = state;
    double* v = state + M;
    const double* q = state + 2 * M;
    double kt = PFMDS_KB * temperature;
    double kedif = 2. * ke - 3. * L * kt, b = 0.;
    if (M == 1) {
        v[0] = v[0] + kedif / q[0] * ts3;
    } else {
        v[M - 1] = v[M - 1] + (q[M - 2] * v[M - 2] * v[M - 2] - kt) / q[M - 1] * ts3;
        for (int i = M - 2; i >= 1; --i) {
            b = exp(-v[i + 1] * ts4);
            v[i] = v[i] * (b * b) + (q[i - 1] * v[i - 1] * v[i - 1] - kt) / q[i] * ts3 * b;
        }
        b = exp(-v[1] * ts4);
        v[0] = v[0] * (b * b) + kedif / q[0] * ts3 * b;
    }
    double s = exp(-v[0] * ts2);
    state[3 * M] = s;
    kedif = 2. * ke * (s * s) - 3. * L * kt;
    for (int i = 0; i < M; ++i) x[i] = x[i] + v[i] * ts2;
    if (M == 1) {
        v[0] = v[0] + kedif / q[0] * ts3;
    } else {
        v[0] = v[0] * (b * b) + kedif / q[0] * ts3 * b;  // the reference reuses the last b here (:236)
        for (int i = 1; i <= M - 2; ++i) {
            b = exp(-v[i + 1] * ts4);
            v[i] = v[i] * (b * b) + (q[i - 1] * v[i - 1] * v[i - 1] - kt) / q[i] * ts3 * b;
        }
        v[M - 1] = v[M - 1] + (q[M - 2] * v[M - 2] * v[M - 2] - kt) / q[M - 1] * ts3;
    }
    state[3 * M + 1] = ke * (s * s);  // kinetic energy of the group after the scaling
    return s;
}

// One thermostat half step on a LOCAL copy of the chain (state block of 3M+4 doubles: x, v, q, s, cached KE, pending scale, spare).
// nhc_chain reads and writes the chain link by link; on the state block in global memory every access is a dependent round trip
// to L2 (k_nhc_close 8.4 us, k_nhc_open 5.7 us of a 10^6-atom step; 16 of the 23 us of a small system's closing kernel, ncu).
// Same statements on the same numbers: bit-identical results.
//   mode 0: plain half step with `ke` (non-fused NVT path)        1: opening half step on the cached KE, pending *= s
//   mode 2: closing half step with `ke`, pending = s               3: closing, then the next step's opening (pending = s_c s_o)
#define NHC_MLOC 8
__device__ __forceinline__ void nhc_step(double* state, int M, int L, double temperature, double ke, double ts2, double ts3, double ts4, int mode) {
    if (M > NHC_MLOC) {  // long chains: in place
        if (mode == 0) { nhc_chain(state, M, L, temperature, ke, ts2, ts3, ts4); return; }
        if (mode == 1) { state[3 * M + 2] *= nhc_chain(state, M, L, temperature, state[3 * M + 1], ts2, ts3, ts4); return; }
        state[3 * M + 2] = nhc_chain(state, M, L, temperature, ke, ts2, ts3, ts4);
        if (mode == 3) state[3 * M + 2] *= nhc_chain(state, M, L, temperature, state[3 * M + 1], ts2, ts3, ts4);
        return;
    }
    double loc[3 * NHC_MLOC + 4];
    const int n = 3 * M + 4;
    for (int i = 0; i < n; ++i) loc[i] = state[i];
    if (mode == 0) nhc_chain(loc, M, L, temperature, ke, ts2, ts3, ts4);
    else if (mode == 1) loc[3 * M + 2] *= nhc_chain(loc, M, L, temperature, loc[3 * M + 1], ts2, ts3, ts4);
    else {
        loc[3 * M + 2] = nhc_chain(loc, M, L, temperature, ke, ts2, ts3, ts4);
        if (mode == 3) loc[3 * M + 2] *= nhc_chain(loc, M, L, temperature, loc[3 * M + 1], ts2, ts3, ts4);
    }
    for (int i = 0; i < n; ++i) state[i] = loc[i];
}

// Slab decomposition, fused halo: what a compute kernel needs to store its border atoms' results straight into
// the neighbours' ghost slots (IPC-mapped peer memory over NVLink), publish a sequence number when the whole grid
// is done, and/or wait for the neighbours' sequence number before it starts.  All zeros = single-GPU behaviour.
struct SlabDev {
    int push;                      // store to the peers and signal
    const int* rs_l; const int* rs_r;   // per slot: ghost slot of this atom in the left / right neighbour, -1 if it is not a border atom
    double4* peer_l; double4* peer_r;   // the neighbours' current position arrays
    int* sig_l; int* sig_r; int sig_seq; unsigned int* counter;
    const int* wait_a; const int* wait_b; int wait_seq;   // 0: nothing to wait for
    int* err;
    unsigned long long timeout_ns;  // give up (error 31) after this long: host-side skew between the ranks (module load, graph
                                    // instantiation, a rank writing a snapshot) must fit inside it; PFMDS_SLAB_TIMEOUT_S, default 120 s
};
#if defined(__CUDACC__) || defined(PFMDS_EMU_WARP)  // device, or the lock-step host replay (ranks as threads: the flags are real)
__device__ __forceinline__ unsigned long long pf_now_ns() {
#ifdef __CUDA_ARCH__
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#else
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
#endif
}
__device__ __forceinline__ void slab_wait(const SlabDev& S) {
    if (S.wait_seq > 0) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = pf_now_ns();
            while (*reinterpret_cast<const volatile int*>(S.wait_a) < S.wait_seq || *reinterpret_cast<const volatile int*>(S.wait_b) < S.wait_seq) {
                if (*reinterpret_cast<volatile int*>(S.err) != 0) break;  // the run is already failing: do not wait once per block
                __nanosleep(100);
                if (pf_now_ns() - t0 > S.timeout_ns) { raise_error(S.err, 31, S.wait_seq, 1); break; }
            }
            // acquire: one load per block (a system-scope FENCE here, once per block of a 7 800-block grid, cost the rjl kernels 30 us
            // per launch on two B200s); the block barrier below carries the ordering to the other threads
#ifdef __CUDA_ARCH__
            int a_, b_;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(a_) : "l"(S.wait_a) : "memory");
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(b_) : "l"(S.wait_b) : "memory");
#else
            __threadfence_system();
#endif
        }
        __syncthreads();
    }
}
// last block of the grid publishes the sequence number; only the threads that stored to a peer pay the system-scope
// fence (a few per cent of them), the block barrier and the device-scope counter order the rest
__device__ __forceinline__ void slab_signal(const SlabDev& S, bool pushed) {
    if (S.push) {
        if (pushed) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();  // release: the barrier ordered this block's (fenced) peer stores before this point; cumulativity carries them past the counter
            unsigned int t = atomicAdd(S.counter, 1u);
            if (t == gridDim.x * gridDim.y - 1) {
                *S.counter = 0;
                __threadfence_system();
                *reinterpret_cast<volatile int*>(S.sig_l) = S.sig_seq;
                *reinterpret_cast<volatile int*>(S.sig_r) = S.sig_seq;
                __threadfence_system();
            }
        }
    }
}
#else
static inline void slab_wait(const SlabDev&) {}
static inline void slab_signal(const SlabDev&, bool) {}
#endif

struct ListView { const int* nlist; const int* nnum; size_t stride; };

// ---- debug build only (-DPFMDS_STAMPS, tools/stamps_probe.py): per-step time stamps of the phases of a small system's step ----------
// Not part of the product: libpfmds_b200.so is built without the macro and every STAMP() below is empty.
#if defined(__CUDACC__) && defined(PFMDS_STAMPS)
#define STAMP_SLOTS 16
#define STAMP_STEPS 4096
static __device__ unsigned long long* g_stamp_buf;  // [0] = step counter, then per step STAMP_SLOTS minima and STAMP_SLOTS maxima
__device__ __forceinline__ void stamp_at(int slot, bool is_min) {
    unsigned long long* b = g_stamp_buf;
    if (!b) return;
    unsigned long long s = *reinterpret_cast<volatile unsigned long long*>(b);
    if (s >= STAMP_STEPS) return;
    unsigned long long t = pf_now_ns();
    unsigned long long* q = b + 1 + s * 2 * STAMP_SLOTS + slot + (is_min ? 0 : STAMP_SLOTS);
    if (is_min) atomicMin(q, t); else atomicMax(q, t);
}
#define STAMP_MIN(slot) do { if (threadIdx.x == 0) stamp_at(slot, true); } while (0)
#define STAMP_MAX(slot) do { if (threadIdx.x == 0) stamp_at(slot, false); } while (0)
#define STAMP_NEXT_STEP() do { if (g_stamp_buf) atomicAdd(g_stamp_buf, 1ull); } while (0)
#define STAMP_BIND_FN(name) void name(unsigned long long* p) { cudaMemcpyToSymbol(g_stamp_buf, &p, sizeof p); }
#else
#define STAMP_MIN(slot) do { } while (0)
#define STAMP_MAX(slot) do { } while (0)
#define STAMP_NEXT_STEP() do { } while (0)
#endif
