// pfmds_b200 — slab spatial decomposition of one large cell over the GPUs of a node (BASELINE.json
// configs[3]; absent from the reference, whose only multi-process mode is the ensemble).
//
// Rank r owns the atoms with x in [r W, (r+1) W), W = Lx / P.  Around them it keeps ghost copies of the
// neighbour ranks' atoms that lie within H = (largest r_cut, padded like a cell edge) of the shared face.
// All forces are gathers, so there is no reverse (force) communication:
//   every step      ghost positions  <- owners                (after the drift)
//   rjl, every step ghost 1/Eb       <- owners                (between the density and the force pass)
//   nvt / momentum  sum over ranks of the KE / momentum partial sums (ncclAllReduce, double)
//   rebuild steps   atoms that left the slab migrate to the neighbour, ghosts are re-selected
// The exchange is ncclSend/ncclRecv with the left and right neighbour inside one group, on the context's
// stream, between kernels that never leave the device.  NCCL is resolved with dlopen at run time, so the
// library has no link-time dependency on it (inside a torch process the bundled libnccl.so.2 is reused).
// Supported interactions in this mode: lj, lj1g, rjl (tb / ljc / morsec would need ghost bond orders and
// normals: refused).
#include "ctx.hpp"
#ifdef __CUDACC__
#include <dlfcn.h>
#include <nccl.h>
#else
#include "nccl_emu.hpp"  // test support: the lock-step host replay of the test suite runs the ranks as threads of one process
#endif

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.hpp"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x; } while (0)

namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    void load() {
        if (h) return;
#ifndef __CUDACC__   // host replay (tests/emu): in-process stand-in, see nccl_emu.hpp
        h = this;
        GetUniqueId = emu_ncclGetUniqueId; CommInitRank = emu_ncclCommInitRank; CommDestroy = emu_ncclCommDestroy; Send = emu_ncclSend; Recv = emu_ncclRecv;
        AllReduce = emu_ncclAllReduce; GroupStart = emu_ncclGroupStart; GroupEnd = emu_ncclGroupEnd; GetErrorString = emu_ncclGetErrorString;
        return;
#else
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw std::string("cannot load libnccl.so.2: ") + dlerror();
#define SYM(f) *(void**)(&f) = dlsym(h, "nccl" #f); if (!f) throw std::string("libnccl lacks nccl" #f)
        SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(Send); SYM(Recv); SYM(AllReduce); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
#undef SYM
#endif
    }
};
NcclApi g_nccl;
#define NK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) throw std::string("NCCL: ") + g_nccl.GetErrorString(r_) + " at " #x; } while (0)
}  // namespace

struct Slab {
    int rank = 0, nranks = 1, left = 0, right = 0;
    ncclComm_t comm = nullptr;
    int n_local = 0, n_ghost = 0, cap = 0;
    long long n_global = 0;
    double W = 0, H = 0;
    int *send_idx[2]{nullptr, nullptr}, *ghost_slot[2]{nullptr, nullptr};
    int n_send[2]{0, 0}, n_recv[2]{0, 0};
    double *sbuf[2]{nullptr, nullptr}, *rbuf[2]{nullptr, nullptr};
    int *cat = nullptr, *scan[3]{nullptr, nullptr, nullptr}, *flag = nullptr, *cnt_d = nullptr, *scan_tmp = nullptr;
    // direct peer-memory halo (NVLink stores into the neighbours' ghost slots, device-side flags instead of NCCL)
    bool p2p = false;
    int* flags = nullptr;                       // mine, written by the neighbours: [0,1] pos ready from left/right, [2,3] 1/Eb ready, [4,5] step done
    double4* pos_buf[2]{nullptr, nullptr};      // my two position buffers (they swap at every cell re-sort)
    double4* peer_pos[2][2]{{nullptr, nullptr}, {nullptr, nullptr}};  // [direction][parity] the neighbours' position buffers, IPC-mapped
    int* peer_flags[2]{nullptr, nullptr};
    int peer_parity[2]{0, 0};
    int* pslot[2]{nullptr, nullptr};            // ghost slot, in neighbour d, of my k-th border atom for that neighbour
    int seq_pos = 0, seq_w = 0, seq_done = 0;
    // fused halo: the compute kernels store into the neighbours' ghost slots themselves
    bool fused = false, pos_pushed = false;
    int wait_pos_seq = 0;
    int* rs[2]{nullptr, nullptr};   // per slot: ghost slot in the left / right neighbour, -1 otherwise
    unsigned int* counter = nullptr;
    // thermostat kinetic energies without NCCL: every rank stores its partial sums into every rank's mailbox (peer memory) and
    // sums all of them in rank order -- the same numbers in the same order on every rank (integ_nvt_kick_close)
    double* ke_box = nullptr;                   // mine: [2 parities][nranks][KE_W] doubles, last of each row = sequence number
    std::vector<double*> peer_box;              // every rank's mailbox (IPC mapped; [rank] = my own)
    double** peer_box_d = nullptr;              // the same pointers on the device
    int ke_seq = 0;
    unsigned long long timeout_ns = 120000000000ull;  // spin-wait limit of the flag waits (PFMDS_SLAB_TIMEOUT_S)
    std::vector<void*> ipc_opened;
};

#define GHOST_BIT 0x80000000u
#define MIG_W 9   // doubles per migrating atom: pos4, vel4, (mask, orig)
#define GH_W 5    // doubles per new ghost: pos4, (mask, orig)
#define KE_W (NHC_MAXF + 1)

__global__ void k_sl_scatter_rs(int n, const int* __restrict__ idx, const int* __restrict__ ps, int* __restrict__ rs);

// Host waits of the redistribution block on an event instead of spinning: with one process per GPU and the NCCL /
// watchdog threads of the host framework, eight spinning ranks oversubscribe a 16-thread host and one descheduled
// rank stalls all the others at the next exchange.
static void host_wait(pfmds_ctx* c) {
    static thread_local cudaEvent_t ev = nullptr;
    if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming));
    CK(cudaEventRecord(ev, c->st));
    CK(cudaEventSynchronize(ev));
}

// ---- small generic pieces ---------------------------------------------------------------------------
__global__ void k_sl_scan_block(int n, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ sums) {
    __shared__ int sh[32];
    int base = blockIdx.x * 2048 + threadIdx.x * 2;
    int a = base < n ? in[base] : 0, b = base + 1 < n ? in[base + 1] : 0;
    int v = a + b, incl = v;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sh[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = sh[lane], si = s;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
        sh[lane] = si - s;
        if (lane == 31) sums[blockIdx.x] = si;
    }
    __syncthreads();
    int excl = incl - v + sh[w];
    if (base < n) out[base] = excl;
    if (base + 1 < n) out[base + 1] = excl + a;
}
__global__ void k_sl_scan_sums(int nb, int* sums, int* total) {  // serial: nb is a few hundred at most
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < nb; ++i) { int v = sums[i]; sums[i] = run; run += v; }
        *total = run;
    }
}
__global__ void k_sl_scan_add(int n, int* out, const int* __restrict__ sums) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += sums[i / 2048];
}
static void exclusive_scan(pfmds_ctx* c, Slab* s, const int* in, int* out, int n, int* total_d) {
    if (n == 0) { CK(cudaMemsetAsync(total_d, 0, sizeof(int), c->st)); return; }
    int sb = (n + 2047) / 2048;
    LAUNCH((k_sl_scan_block), sb, 1024, c->st, n, in, out, s->scan_tmp);
    LAUNCH((k_sl_scan_sums), 1, 32, c->st, sb, s->scan_tmp, total_d);
    LAUNCH((k_sl_scan_add), (n + 255) / 256, 256, c->st, n, out, s->scan_tmp);
    c->launches += 3;
}

// ---- migration ---------------------------------------------------------------------------------------
// category of every slot after the drift: 0 stays, 1 goes to the left rank, 2 to the right rank, 3 ghost (dropped)
__global__ void k_sl_classify(int N, const double4* __restrict__ pos, const uint32_t* __restrict__ gmask, const int* __restrict__ orig, double W, int nranks,
                              int rank, int* __restrict__ cat, int* __restrict__ f0, int* __restrict__ f1, int* __restrict__ f2, int* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int k = 3;
    if (!(gmask[i] & GHOST_BIT)) {
        int owner = (int)floor(pos[i].x / W);
        owner = owner < 0 ? 0 : (owner >= nranks ? nranks - 1 : owner);
        int left = (rank + nranks - 1) % nranks, right = (rank + 1) % nranks;
        if (owner == rank) k = 0;
        else if (owner == left) k = 1;
        else if (owner == right) k = 2;
        else { raise_error(err, 30, orig[i], owner); k = 0; }  // moved farther than one slab between rebuilds
    }
    cat[i] = k;
    f0[i] = k == 0; f1[i] = k == 1; f2[i] = k == 2;
}
__global__ void k_sl_pack_migrants(int N, const double4* __restrict__ pos, const double4* __restrict__ vel, const uint32_t* __restrict__ gmask,
                                   const int* __restrict__ orig, const int* __restrict__ cat, const int* __restrict__ s0, const int* __restrict__ s1,
                                   const int* __restrict__ s2, double4* __restrict__ pos2, double4* __restrict__ vel2, uint32_t* __restrict__ gm2,
                                   int* __restrict__ orig2, double* __restrict__ bl, double* __restrict__ br) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int k = cat[i];
    if (k == 0) {
        int d = s0[i];
        pos2[d] = pos[i]; vel2[d] = vel[i]; gm2[d] = gmask[i]; orig2[d] = orig[i];
    } else if (k == 1 || k == 2) {
        double* b = (k == 1 ? bl : br) + (size_t)(k == 1 ? s1[i] : s2[i]) * MIG_W;
        double4 p = pos[i], v = vel[i];
        b[0] = p.x; b[1] = p.y; b[2] = p.z; b[3] = p.w; b[4] = v.x; b[5] = v.y; b[6] = v.z; b[7] = v.w;
        b[8] = __hiloint2double((int)gmask[i], orig[i]);
    }
}
__global__ void k_sl_unpack_migrants(int n, const double* __restrict__ buf, int at, double4* __restrict__ pos, double4* __restrict__ vel,
                                     uint32_t* __restrict__ gm, int* __restrict__ orig) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* b = buf + (size_t)k * MIG_W;
    pos[at + k] = make_double4(b[0], b[1], b[2], b[3]);
    vel[at + k] = make_double4(b[4], b[5], b[6], b[7]);
    gm[at + k] = (uint32_t)__double2hiint(b[8]);
    orig[at + k] = __double2loint(b[8]);
}
// ---- ghost selection ---------------------------------------------------------------------------------
__global__ void k_sl_border_flags(int N, const double4* __restrict__ pos, double x_lo, double x_hi, double H, int* __restrict__ fl, int* __restrict__ fr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double x = pos[i].x;
    fl[i] = (x - x_lo) < H;
    fr[i] = (x_hi - x) <= H;
}
__global__ void k_sl_pack_ghosts(int N, const double4* __restrict__ pos, const uint32_t* __restrict__ gmask, const int* __restrict__ orig,
                                 const int* __restrict__ fl, const int* __restrict__ fr, const int* __restrict__ sl, const int* __restrict__ sr,
                                 int* __restrict__ idx_l, int* __restrict__ idx_r, double* __restrict__ bl, double* __restrict__ br) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double4 p = pos[i];
    double tag = __hiloint2double((int)(gmask[i] | GHOST_BIT), orig[i]);
    if (fl[i]) { int d = sl[i]; idx_l[d] = i; double* b = bl + (size_t)d * GH_W; b[0] = p.x; b[1] = p.y; b[2] = p.z; b[3] = p.w; b[4] = tag; }
    if (fr[i]) { int d = sr[i]; idx_r[d] = i; double* b = br + (size_t)d * GH_W; b[0] = p.x; b[1] = p.y; b[2] = p.z; b[3] = p.w; b[4] = tag; }
}
__global__ void k_sl_unpack_ghosts(int n, const double* __restrict__ buf, int at, double4* __restrict__ pos, double4* __restrict__ vel,
                                   uint32_t* __restrict__ gm, int* __restrict__ orig, int* __restrict__ slot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* b = buf + (size_t)k * GH_W;
    pos[at + k] = make_double4(b[0], b[1], b[2], b[3]);
    vel[at + k] = make_double4(0., 0., 0., 1.);
    gm[at + k] = (uint32_t)__double2hiint(b[4]);
    orig[at + k] = __double2loint(b[4]);
    slot[k] = at + k;
}
__global__ void k_sl_remap(int n, int* __restrict__ idx, const int* __restrict__ newslot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) idx[k] = newslot[idx[k]];
}
// ---- per-step halo ------------------------------------------------------------------------------------
template <int FIELD>  // 0: x,y,z   1: w
__global__ void k_sl_pack_halo(int n, const int* __restrict__ idx, const double4* __restrict__ pos, double* __restrict__ buf) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double4 p = pos[idx[k]];
    if (FIELD == 0) { buf[3 * (size_t)k] = p.x; buf[3 * (size_t)k + 1] = p.y; buf[3 * (size_t)k + 2] = p.z; }
    else buf[k] = p.w;
}
template <int FIELD>
__global__ void k_sl_unpack_halo(int n, const int* __restrict__ slot, const double* __restrict__ buf, double4* __restrict__ pos) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double* p = reinterpret_cast<double*>(&pos[slot[k]]);
    if (FIELD == 0) { p[0] = buf[3 * (size_t)k]; p[1] = buf[3 * (size_t)k + 1]; p[2] = buf[3 * (size_t)k + 2]; }
    else p[3] = buf[k];
}

// ---- direct peer-memory halo ---------------------------------------------------------------------------
// One kernel stores the border atoms' x,y,z (or 1/Eb) straight into the ghost slots of both neighbours through
// their IPC-mapped position arrays: no pack buffer, no NCCL call, no unpack kernel on the other side.
template <int FIELD>
__global__ void k_sl_push(int nl, const int* __restrict__ idx_l, const int* __restrict__ ps_l, double4* peer_l, int nr, const int* __restrict__ idx_r,
                          const int* __restrict__ ps_r, double4* peer_r, const double4* __restrict__ pos) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int* idx; const int* ps; double4* peer;
    if (k < nl) { idx = idx_l; ps = ps_l; peer = peer_l; }
    else if (k < nl + nr) { k -= nl; idx = idx_r; ps = ps_r; peer = peer_r; }
    else return;
    const double4 p = pos[idx[k]];
    double* q = reinterpret_cast<double*>(&peer[ps[k]]);
    if (FIELD == 0) { q[0] = p.x; q[1] = p.y; q[2] = p.z; }
    else q[3] = p.w;
}
// publish `seq` in both neighbours' flag words (runs after the push kernel on the same stream: its stores are complete)
__global__ void k_sl_signal(int* to_left, int* to_right, int seq) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(to_left) = seq;
    *reinterpret_cast<volatile int*>(to_right) = seq;
    __threadfence_system();
}
__device__ __forceinline__ unsigned long long sl_now_ns() {
    unsigned long long t;
#ifdef __CUDACC__
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#else  // host replay
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    t = (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
#endif
    return t;
}
// wait until both neighbours have published at least `seq`; gives up after `timeout_ns` and records an error instead of hanging
__global__ void k_sl_wait(const int* from_left, const int* from_right, int seq, int* err, unsigned long long timeout_ns) {
    unsigned long long t0 = sl_now_ns();
    while (*reinterpret_cast<const volatile int*>(from_left) < seq || *reinterpret_cast<const volatile int*>(from_right) < seq) {
        if (*reinterpret_cast<volatile int*>(err) != 0) break;
        __nanosleep(200);
        unsigned long long t = sl_now_ns();
        if (t - t0 > timeout_ns) { raise_error(err, 31, seq, 0); break; }
    }
    __threadfence_system();
}

static bool slab_setup_p2p(pfmds_ctx* c, Slab* s) {
    // exchange IPC handles of my two position buffers and my flag words with both neighbours
    struct Pack { cudaIpcMemHandle_t pos0, pos1, fl; };
    Pack mine{}, from[2]{};
    bool ok = true;
    CK(cudaMalloc(&s->flags, 8 * sizeof(int)));
    CK(cudaMemset(s->flags, 0, 8 * sizeof(int)));
    s->pos_buf[0] = c->pos; s->pos_buf[1] = c->pos2;
    if (cudaIpcGetMemHandle(&mine.pos0, c->pos) != cudaSuccess || cudaIpcGetMemHandle(&mine.pos1, c->pos2) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.fl, s->flags) != cudaSuccess) { ok = false; cudaGetLastError(); }
    char *d_mine = nullptr, *d_from = nullptr;
    CK(cudaMalloc(&d_mine, sizeof(Pack)));
    CK(cudaMalloc(&d_from, 2 * sizeof(Pack)));
    CK(cudaMemcpy(d_mine, &mine, sizeof(Pack), cudaMemcpyHostToDevice));
    NK(g_nccl.GroupStart());
    NK(g_nccl.Send(d_mine, sizeof(Pack), ncclChar, s->left, s->comm, c->st));
    NK(g_nccl.Send(d_mine, sizeof(Pack), ncclChar, s->right, s->comm, c->st));
    NK(g_nccl.Recv(d_from + sizeof(Pack), sizeof(Pack), ncclChar, s->right, s->comm, c->st));
    NK(g_nccl.Recv(d_from, sizeof(Pack), ncclChar, s->left, s->comm, c->st));
    NK(g_nccl.GroupEnd());
    CK(cudaMemcpyAsync(from, d_from, 2 * sizeof(Pack), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    cudaFree(d_mine); cudaFree(d_from);
    const int ndist = (s->left == s->right) ? 1 : 2;
    for (int d = 0; d < ndist && ok; ++d) {
        void *p0 = nullptr, *p1 = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&p0, from[d].pos0, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&p1, from[d].pos1, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&pf, from[d].fl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
        s->ipc_opened.push_back(p0); s->ipc_opened.push_back(p1); s->ipc_opened.push_back(pf);
        s->peer_pos[d][0] = (double4*)p0; s->peer_pos[d][1] = (double4*)p1; s->peer_flags[d] = (int*)pf;
    }
    if (ndist == 1) { s->peer_pos[1][0] = s->peer_pos[0][0]; s->peer_pos[1][1] = s->peer_pos[0][1]; s->peer_flags[1] = s->peer_flags[0]; }
    // every rank must take the same path
    int h = ok ? 1 : 0, *dflag = s->cnt_d + 12;
    CK(cudaMemcpy(dflag, &h, sizeof(int), cudaMemcpyHostToDevice));
    NK(g_nccl.AllReduce(dflag, dflag, 1, ncclInt, ncclMin, s->comm, c->st));
    CK(cudaMemcpyAsync(&h, dflag, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int d = 0; d < 2; ++d) { CK(cudaMalloc(&s->pslot[d], sizeof(int) * c->stride)); CK(cudaMalloc(&s->rs[d], sizeof(int) * c->stride)); }
    CK(cudaMalloc(&s->counter, sizeof(unsigned int)));
    CK(cudaMemset(s->counter, 0, sizeof(unsigned int)));
    // Fusing the stores and flags into the compute kernels is implemented (SlabDev, common.cuh) but measured slower than
    // the separate push kernel on 2 x B200 (1.31 vs 1.18 ms/step: system-scope fences and NVLink store latency land on
    // the FP64-bound warps), so it is opt-in.
    const char* fz = std::getenv("PFMDS_SLAB_FUSED");
    s->fused = fz && fz[0] == '1';
    if (h == 1 && !(std::getenv("PFMDS_SLAB_KE_NCCL") && std::getenv("PFMDS_SLAB_KE_NCCL")[0] == '1')) {
        // mailboxes of the kinetic-energy all-gather: handles go to every rank
        const int P = s->nranks;
        const size_t bytes = sizeof(double) * 2 * (size_t)P * KE_W;
        CK(cudaMalloc(&s->ke_box, bytes));
        CK(cudaMemset(s->ke_box, 0, bytes));
        cudaIpcMemHandle_t mh{};
        bool okb = cudaIpcGetMemHandle(&mh, s->ke_box) == cudaSuccess;
        if (!okb) cudaGetLastError();
        char *d_m = nullptr, *d_all = nullptr;
        CK(cudaMalloc(&d_m, sizeof mh));
        CK(cudaMalloc(&d_all, sizeof mh * (size_t)P));
        CK(cudaMemcpy(d_m, &mh, sizeof mh, cudaMemcpyHostToDevice));
        NK(g_nccl.GroupStart());
        for (int r = 0; r < P; ++r) {
            if (r == s->rank) continue;
            NK(g_nccl.Send(d_m, sizeof mh, ncclChar, r, s->comm, c->st));
            NK(g_nccl.Recv(d_all + sizeof mh * (size_t)r, sizeof mh, ncclChar, r, s->comm, c->st));
        }
        NK(g_nccl.GroupEnd());
        std::vector<cudaIpcMemHandle_t> all((size_t)P);
        CK(cudaMemcpyAsync(all.data(), d_all, sizeof mh * (size_t)P, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        cudaFree(d_m); cudaFree(d_all);
        s->peer_box.assign((size_t)P, nullptr);
        for (int r = 0; r < P && okb; ++r) {
            if (r == s->rank) { s->peer_box[(size_t)r] = s->ke_box; continue; }
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { okb = false; cudaGetLastError(); break; }
            s->ipc_opened.push_back(q);
            s->peer_box[(size_t)r] = (double*)q;
        }
        int hb = okb ? 1 : 0;   // every rank must take the same path
        CK(cudaMemcpy(dflag, &hb, sizeof(int), cudaMemcpyHostToDevice));
        NK(g_nccl.AllReduce(dflag, dflag, 1, ncclInt, ncclMin, s->comm, c->st));
        CK(cudaMemcpyAsync(&hb, dflag, sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        if (hb == 1) {
            CK(cudaMalloc(&s->peer_box_d, sizeof(double*) * (size_t)P));
            CK(cudaMemcpy(s->peer_box_d, s->peer_box.data(), sizeof(double*) * (size_t)P, cudaMemcpyHostToDevice));
        } else s->peer_box.clear();
    }
    return h == 1;
}

// after the re-sort of a rebuild step: tell each neighbour where its border atoms now live here, and which of my
// two position buffers is current
static void slab_exchange_peer_slots(pfmds_ctx* c, Slab* s) {
    int par = (c->pos == s->pos_buf[0]) ? 0 : 1;
    CK(cudaMemcpyAsync(s->cnt_d + 13, &par, sizeof(int), cudaMemcpyHostToDevice, c->st));
    NK(g_nccl.GroupStart());
    if (s->n_recv[0] > 0) NK(g_nccl.Send(s->ghost_slot[0], (size_t)s->n_recv[0], ncclInt, s->left, s->comm, c->st));
    if (s->n_recv[1] > 0) NK(g_nccl.Send(s->ghost_slot[1], (size_t)s->n_recv[1], ncclInt, s->right, s->comm, c->st));
    NK(g_nccl.Send(s->cnt_d + 13, 1, ncclInt, s->left, s->comm, c->st));
    NK(g_nccl.Send(s->cnt_d + 13, 1, ncclInt, s->right, s->comm, c->st));
    if (s->n_send[1] > 0) NK(g_nccl.Recv(s->pslot[1], (size_t)s->n_send[1], ncclInt, s->right, s->comm, c->st));
    if (s->n_send[0] > 0) NK(g_nccl.Recv(s->pslot[0], (size_t)s->n_send[0], ncclInt, s->left, s->comm, c->st));
    NK(g_nccl.Recv(s->cnt_d + 15, 1, ncclInt, s->right, s->comm, c->st));
    NK(g_nccl.Recv(s->cnt_d + 14, 1, ncclInt, s->left, s->comm, c->st));
    NK(g_nccl.GroupEnd());
    CK(cudaMemcpyAsync(s->peer_parity, s->cnt_d + 14, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    host_wait(c);
    // per-slot view of the same lists, for the kernels that push while they compute
    for (int d = 0; d < 2; ++d) {
        CK(cudaMemsetAsync(s->rs[d], 0xff, sizeof(int) * c->stride, c->st));
        if (s->n_send[d] > 0) LAUNCH((k_sl_scatter_rs), (s->n_send[d] + 255) / 256, 256, c->st, s->n_send[d], s->send_idx[d], s->pslot[d], s->rs[d]);
    }
    s->pos_pushed = false; s->wait_pos_seq = 0;
}

// the neighbours may overwrite my ghost positions for the next step only after my force kernels of this step are done
void slab_step_done(pfmds_ctx* c) {
    Slab* s = c->slab;
    if (!s->p2p) return;
    s->seq_done += 1;
    LAUNCH((k_sl_signal), 1, 1, c->st, s->peer_flags[0] + 5, s->peer_flags[1] + 4, s->seq_done);
    c->launches += 1;
}

// ------------------------------------------------------------------------------------------------------
int slab_unique_id(char* id128) {
    g_nccl.load();
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

void slab_init(pfmds_ctx* c, int rank, int nranks, const char* id128, long long n_global, int n_local, int capacity) {
    g_nccl.load();
    Slab* s = new Slab;
    c->slab = s;
    s->rank = rank; s->nranks = nranks; s->left = (rank + nranks - 1) % nranks; s->right = (rank + 1) % nranks;
    s->n_local = n_local; s->n_ghost = 0; s->cap = capacity; s->n_global = n_global;
    s->W = c->box.L[0] / nranks;
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    NK(g_nccl.CommInitRank(&s->comm, nranks, id, rank));
    const size_t S = c->stride;
    for (int d = 0; d < 2; ++d) {
        CK(cudaMalloc(&s->send_idx[d], sizeof(int) * S));
        CK(cudaMalloc(&s->ghost_slot[d], sizeof(int) * S));
        CK(cudaMalloc(&s->sbuf[d], sizeof(double) * MIG_W * (S / 2 + 1024)));
        CK(cudaMalloc(&s->rbuf[d], sizeof(double) * MIG_W * (S / 2 + 1024)));
    }
    CK(cudaMalloc(&s->cat, sizeof(int) * S));
    CK(cudaMalloc(&s->flag, sizeof(int) * 3 * S));
    for (int k = 0; k < 3; ++k) CK(cudaMalloc(&s->scan[k], sizeof(int) * S));
    CK(cudaMalloc(&s->cnt_d, sizeof(int) * 16));
    CK(cudaMalloc(&s->scan_tmp, sizeof(int) * (S / 2048 + 2)));
    CK(cudaMalloc(&c->newslot, sizeof(int) * S));
    if (const char* to = std::getenv("PFMDS_SLAB_TIMEOUT_S")) { double v = std::atof(to); if (v >= 1.) s->timeout_ns = (unsigned long long)(v * 1e9); }
    const char* env = std::getenv("PFMDS_SLAB_P2P");
    s->p2p = !(env && env[0] == '0') && slab_setup_p2p(c, s);
}

void slab_destroy(pfmds_ctx* c) {
    Slab* s = c->slab;
    if (!s) return;
    for (int d = 0; d < 2; ++d) { cudaFree(s->send_idx[d]); cudaFree(s->ghost_slot[d]); cudaFree(s->sbuf[d]); cudaFree(s->rbuf[d]); }
    cudaFree(s->cat); cudaFree(s->flag); for (int k = 0; k < 3; ++k) cudaFree(s->scan[k]);
    cudaFree(s->cnt_d); cudaFree(s->scan_tmp); cudaFree(c->newslot);
    for (void* p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(s->ke_box); cudaFree(s->peer_box_d);
    cudaFree(s->flags); cudaFree(s->pslot[0]); cudaFree(s->pslot[1]); cudaFree(s->rs[0]); cudaFree(s->rs[1]); cudaFree(s->counter);
    if (s->comm) g_nccl.CommDestroy(s->comm);
    delete s;
    c->slab = nullptr;
}

bool slab_uses_p2p(pfmds_ctx* c) { return c->slab && c->slab->p2p; }
bool slab_fused(pfmds_ctx* c) {
    Slab* s = c->slab;
    if (!s || !s->p2p || !s->fused) return false;
    if (s->n_global / s->nranks < 400000) return false;  // small slabs run the lanes-per-atom kernels, which use the unfused halo
    for (auto& it : c->inter) if (it.kind != K_RJL) return false;
    return true;
}
__global__ void k_sl_scatter_rs(int n, const int* __restrict__ idx, const int* __restrict__ ps, int* __restrict__ rs) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) rs[idx[k]] = ps[k];
}
SlabDev slab_dev(pfmds_ctx* c, int stage) {
    Slab* s = c->slab;
    SlabDev S{};
    S.err = c->err;
    S.timeout_ns = s->timeout_ns;
    S.rs_l = s->rs[0]; S.rs_r = s->rs[1];
    S.peer_l = s->peer_pos[0][s->peer_parity[0]]; S.peer_r = s->peer_pos[1][s->peer_parity[1]];
    S.counter = s->counter;
    if (stage == 0) {         // kick+drift: wait until the neighbours are done with the old ghost positions, push the new ones
        S.push = 1; S.wait_a = s->flags + 4; S.wait_b = s->flags + 5; S.wait_seq = s->seq_done;
        s->seq_pos += 1;
        S.sig_l = s->peer_flags[0] + 1; S.sig_r = s->peer_flags[1] + 0; S.sig_seq = s->seq_pos;
        s->pos_pushed = true;
    } else if (stage == 1) {  // rjl density: wait for the ghost positions of this step, push 1/Eb
        S.push = 1; S.wait_a = s->flags + 0; S.wait_b = s->flags + 1; S.wait_seq = s->wait_pos_seq;
        s->wait_pos_seq = 0;
        s->seq_w += 1;
        S.sig_l = s->peer_flags[0] + 3; S.sig_r = s->peer_flags[1] + 2; S.sig_seq = s->seq_w;
    } else {                  // rjl force: wait for the ghost 1/Eb
        S.push = 0; S.wait_a = s->flags + 2; S.wait_b = s->flags + 3; S.wait_seq = s->seq_w;
    }
    return S;
}
bool slab_pos_pushed_by_kick(pfmds_ctx* c, bool rebuild_step) { return slab_fused(c) && !rebuild_step; }
int slab_rank(pfmds_ctx* c) { return c->slab->rank; }
int slab_nranks(pfmds_ctx* c) { return c->slab->nranks; }
int slab_n_local(pfmds_ctx* c) { return c->slab->n_local; }
long long slab_n_global(pfmds_ctx* c) { return c->slab->n_global; }

// One block: (1) this rank's KE partial sums in block order, (2) stored with their sequence number into every rank's mailbox,
// (3) wait for every rank's row of this sequence number in my mailbox, (4) sum the rows in rank order, (5) the chain update
// (k_nhc_close's work).  Every rank adds the same numbers in the same order: identical thermostat state on all ranks, bit for bit.
__global__ void k_sl_ke_close(int nparts, const double* __restrict__ part, NhcPack P, double ts2, double ts3, double ts4, int rank, int nranks,
                              double* const* __restrict__ box, int seq, int* err, unsigned long long timeout_ns) {
    __shared__ double mine[NHC_MAXF];
    for (int k = 0; k < P.n; ++k) {
        double ke = 0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) ke += part[i * NHC_MAXF + k];
        ke = block_sum(ke);
        if (threadIdx.x == 0) mine[k] = ke;
        __syncthreads();
    }
    const int par = seq & 1;
    if ((int)threadIdx.x < nranks) {   // thread r writes my row into rank r's mailbox
        double* row = box[threadIdx.x] + ((size_t)par * nranks + rank) * KE_W;
        for (int k = 0; k < P.n; ++k) reinterpret_cast<volatile double*>(row)[k] = mine[k];
        __threadfence_system();
        reinterpret_cast<volatile double*>(row)[NHC_MAXF] = (double)seq;
    }
    if ((int)threadIdx.x < nranks) {   // thread r waits for rank r's row in my mailbox
        const volatile double* row = box[rank] + ((size_t)par * nranks + threadIdx.x) * KE_W;
        const unsigned long long t0 = pf_now_ns();
        while (row[NHC_MAXF] != (double)seq) {
            if (*reinterpret_cast<volatile int*>(err) != 0) break;
            __nanosleep(100);
            if (pf_now_ns() - t0 > timeout_ns) { raise_error(err, 31, seq, 2); break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < P.n; ++k) {
            double ke = 0;
            for (int r = 0; r < nranks; ++r) ke += reinterpret_cast<const volatile double*>(box[rank])[((size_t)par * nranks + r) * KE_W + k];
            double* st = P.state[k];
            const int M = P.M[k];
            nhc_step(st, M, P.L[k], P.T[k], ke, ts2, ts3, ts4, 2);
        }
    }
}
// returns false when the mailboxes are not available (no peer access, PFMDS_SLAB_KE_NCCL=1): the caller takes the NCCL all-reduce
bool slab_ke_close(pfmds_ctx* c, const NhcPack& P, int nparts, const double* part, double ts2, double ts3, double ts4) {
    Slab* s = c->slab;
    if (!s || !s->peer_box_d || s->nranks > 64) return false;
    s->ke_seq += 1;
    LAUNCH((k_sl_ke_close), 1, 256, c->st, nparts, part, P, ts2, ts3, ts4, s->rank, s->nranks, (double* const*)s->peer_box_d, s->ke_seq, c->err, s->timeout_ns);
    c->launches += 1;
    return true;
}

// ---- this rank's atoms to / from the host (pfmds_slab_download / pfmds_slab_upload) ---------------------------------------------
// Owned atoms (not ghost copies) in slot order, compacted on the device: an exclusive scan of the owner flags gives every owned slot
// its row in dense [n_local][3] arrays (the migration / ghost buffers, free outside a rebuild), which then cross PCIe with one
// copy per array.  (Round 1 copied the padded double4 arrays and compacted on the host: 0.6-0.7 s per call at 1.3e7 atoms per rank.)
__global__ void k_sl_owner_flags(int N, const uint32_t* __restrict__ gmask, int* __restrict__ f) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) f[i] = (gmask[i] & GHOST_BIT) ? 0 : 1;
}
__global__ void k_sl_gather_owned(int N, const uint32_t* __restrict__ gmask, const int* __restrict__ row, const double4* __restrict__ a,
                                  double* __restrict__ out, const int* __restrict__ orig, int* __restrict__ gid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || (gmask[i] & GHOST_BIT)) return;
    const int r = row[i];
    if (a) { const double4 v = a[i]; out[3 * (size_t)r] = v.x; out[3 * (size_t)r + 1] = v.y; out[3 * (size_t)r + 2] = v.z; }
    if (gid) gid[r] = orig[i] + 1;
}
__global__ void k_sl_scatter_owned(int N, const uint32_t* __restrict__ gmask, const int* __restrict__ row, const double* __restrict__ in,
                                   double4* __restrict__ a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || (gmask[i] & GHOST_BIT)) return;
    const int r = row[i];
    double4 v = a[i];
    v.x = in[3 * (size_t)r]; v.y = in[3 * (size_t)r + 1]; v.z = in[3 * (size_t)r + 2];
    a[i] = v;
}
static int slab_owner_rows(pfmds_ctx* c, Slab* s) {   // rows of the owned slots in s->scan[0]; returns their number (known on the host)
    const int N = c->N, T = 256;
    LAUNCH((k_sl_owner_flags), (N + T - 1) / T, T, c->st, N, c->gmask, s->flag);
    exclusive_scan(c, s, s->flag, s->scan[0], N, s->cnt_d + 7);
    c->launches += 1;
    return s->n_local;
}
void slab_download(pfmds_ctx* c, int* n_local, int* gid, double* pos, double* vel, double* frc) {
    Slab* s = c->slab;
    const int N = c->N, T = 256, nl = slab_owner_rows(c, s);
    const size_t n3 = 3 * (size_t)nl;
    double* stage[3] = {s->sbuf[0], s->sbuf[1], s->rbuf[0]};
    const double4* src[3] = {c->pos, c->vel, c->frc};
    double* out[3] = {pos, vel, frc};
    int* gstage = reinterpret_cast<int*>(s->rbuf[1]);
    bool first = true;
    for (int a = 0; a < 3; ++a) {
        if (!out[a] && !(first && gid && a == 2)) continue;
        LAUNCH((k_sl_gather_owned), (N + T - 1) / T, T, c->st, N, c->gmask, s->scan[0], out[a] ? src[a] : (const double4*)nullptr, stage[a], c->orig,
               (first && gid) ? gstage : (int*)nullptr);
        first = false;
        c->launches += 1;
        if (out[a]) CK(cudaMemcpyAsync(out[a], stage[a], sizeof(double) * n3, cudaMemcpyDeviceToHost, c->st));
    }
    if (gid) CK(cudaMemcpyAsync(gid, gstage, sizeof(int) * (size_t)nl, cudaMemcpyDeviceToHost, c->st));
    if (n_local) *n_local = nl;
    CK(cudaStreamSynchronize(c->st));
}
void slab_upload(pfmds_ctx* c, int n_local, const double* pos, const double* vel) {
    Slab* s = c->slab;
    const int N = c->N, T = 256, nl = slab_owner_rows(c, s);
    if (nl != n_local) throw std::string("pfmds_slab_upload expects the atoms of the last pfmds_slab_download");
    const size_t n3 = 3 * (size_t)nl;
    if (pos) {
        CK(cudaMemcpyAsync(s->sbuf[0], pos, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st));
        LAUNCH((k_sl_scatter_owned), (N + T - 1) / T, T, c->st, N, c->gmask, s->scan[0], s->sbuf[0], c->pos);
    }
    if (vel) {
        CK(cudaMemcpyAsync(s->sbuf[1], vel, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st));
        LAUNCH((k_sl_scatter_owned), (N + T - 1) / T, T, c->st, N, c->gmask, s->scan[0], s->sbuf[1], c->vel);
    }
    c->launches += 2;
    CK(cudaStreamSynchronize(c->st));
}

void slab_allreduce_sum(pfmds_ctx* c, double* d, int n) {
    NK(g_nccl.AllReduce(d, d, (size_t)n, ncclDouble, ncclSum, c->slab->comm, c->st));
}
void slab_allreduce_max(pfmds_ctx* c, double* d, int n) {
    NK(g_nccl.AllReduce(d, d, (size_t)n, ncclDouble, ncclMax, c->slab->comm, c->st));
}
void slab_allreduce_max_int(pfmds_ctx* c, int* d, int n) {
    NK(g_nccl.AllReduce(d, d, (size_t)n, ncclInt, ncclMax, c->slab->comm, c->st));
}
void slab_allreduce_sum_ll(pfmds_ctx* c, unsigned long long* d, int n) {
    NK(g_nccl.AllReduce(d, d, (size_t)n, ncclUint64, ncclSum, c->slab->comm, c->st));
}

// send `ns[d]` items of width w from sbuf[d] to the neighbour in direction d (0 left, 1 right); what arrives from
// the right neighbour was sent to ITS left, and lands in rbuf[1]
static void neighbour_exchange(pfmds_ctx* c, Slab* s, const int* ns, const int* nr, int w) {
    NK(g_nccl.GroupStart());
    if (ns[0] > 0) NK(g_nccl.Send(s->sbuf[0], (size_t)ns[0] * w, ncclDouble, s->left, s->comm, c->st));
    if (ns[1] > 0) NK(g_nccl.Send(s->sbuf[1], (size_t)ns[1] * w, ncclDouble, s->right, s->comm, c->st));
    if (nr[1] > 0) NK(g_nccl.Recv(s->rbuf[1], (size_t)nr[1] * w, ncclDouble, s->right, s->comm, c->st));
    if (nr[0] > 0) NK(g_nccl.Recv(s->rbuf[0], (size_t)nr[0] * w, ncclDouble, s->left, s->comm, c->st));
    NK(g_nccl.GroupEnd());
}
// exchange two counts with the neighbours (host round trip: only at rebuild steps)
static void exchange_counts(pfmds_ctx* c, Slab* s, const int* ns, int* nr) {
    int h[2] = {ns[0], ns[1]};
    CK(cudaMemcpyAsync(s->cnt_d + 8, h, sizeof h, cudaMemcpyHostToDevice, c->st));
    NK(g_nccl.GroupStart());
    NK(g_nccl.Send(s->cnt_d + 8, 1, ncclInt, s->left, s->comm, c->st));
    NK(g_nccl.Send(s->cnt_d + 9, 1, ncclInt, s->right, s->comm, c->st));
    NK(g_nccl.Recv(s->cnt_d + 11, 1, ncclInt, s->right, s->comm, c->st));
    NK(g_nccl.Recv(s->cnt_d + 10, 1, ncclInt, s->left, s->comm, c->st));
    NK(g_nccl.GroupEnd());
    CK(cudaMemcpyAsync(nr, s->cnt_d + 10, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    host_wait(c);
}

// Rebuild step: migrate, then re-select and import ghosts.  Leaves c->N = n_local + n_ghost with the ghosts
// appended; the cell re-sort that follows permutes everything and slab_after_reorder() fixes the index lists.
void slab_redistribute(pfmds_ctx* c) {
    Slab* s = c->slab;
    const int T = 256;
    const int N0 = c->N;
    s->H = c->cell_rc;
    if ((s->nranks == 2 && s->W <= 2.0 * s->H) || s->W <= s->H)
        throw std::string("slab decomposition: slabs of width ") + std::to_string(s->W) + " A are too thin for a halo of " + std::to_string(s->H) + " A";
    int *f0 = s->flag, *f1 = s->flag + c->stride, *f2 = s->flag + 2 * c->stride;
    LAUNCH((k_sl_classify), (N0 + T - 1) / T, T, c->st, N0, c->pos, c->gmask, c->orig, s->W, s->nranks, s->rank, s->cat, f0, f1, f2, c->err);
    exclusive_scan(c, s, f0, s->scan[0], N0, s->cnt_d + 0);
    exclusive_scan(c, s, f1, s->scan[1], N0, s->cnt_d + 1);
    exclusive_scan(c, s, f2, s->scan[2], N0, s->cnt_d + 2);
    int h[3];
    CK(cudaMemcpyAsync(h, s->cnt_d, sizeof h, cudaMemcpyDeviceToHost, c->st));
    host_wait(c);
    int ns[2] = {h[1], h[2]}, nr[2] = {0, 0};
    exchange_counts(c, s, ns, nr);
    const int n_stay = h[0];
    if ((size_t)n_stay + nr[0] + nr[1] > c->stride) throw std::string("slab decomposition: capacity exceeded by migration");
    LAUNCH((k_sl_pack_migrants), (N0 + T - 1) / T, T, c->st, N0, c->pos, c->vel, c->gmask, c->orig, s->cat, s->scan[0], s->scan[1], s->scan[2], c->pos2,
                                                         c->vel2, c->gmask2, c->orig2, s->sbuf[0], s->sbuf[1]);
    std::swap(c->pos, c->pos2); std::swap(c->vel, c->vel2); std::swap(c->gmask, c->gmask2); std::swap(c->orig, c->orig2);
    neighbour_exchange(c, s, ns, nr, MIG_W);
    if (nr[0] > 0) LAUNCH((k_sl_unpack_migrants), (nr[0] + T - 1) / T, T, c->st, nr[0], s->rbuf[0], n_stay, c->pos, c->vel, c->gmask, c->orig);
    if (nr[1] > 0) LAUNCH((k_sl_unpack_migrants), (nr[1] + T - 1) / T, T, c->st, nr[1], s->rbuf[1], n_stay + nr[0], c->pos, c->vel, c->gmask, c->orig);
    s->n_local = n_stay + nr[0] + nr[1];
    c->launches += 4;
    // ghosts
    const int NL = s->n_local;
    const double x_lo = s->rank * s->W, x_hi = (s->rank + 1) * s->W;
    LAUNCH((k_sl_border_flags), (NL + T - 1) / T, T, c->st, NL, c->pos, x_lo, x_hi, s->H, f0, f1);
    exclusive_scan(c, s, f0, s->scan[0], NL, s->cnt_d + 3);
    exclusive_scan(c, s, f1, s->scan[1], NL, s->cnt_d + 4);
    CK(cudaMemcpyAsync(h, s->cnt_d + 3, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    host_wait(c);
    s->n_send[0] = h[0]; s->n_send[1] = h[1];
    exchange_counts(c, s, s->n_send, s->n_recv);
    if ((size_t)NL + s->n_recv[0] + s->n_recv[1] > c->stride) throw std::string("slab decomposition: capacity exceeded by ghosts");
    LAUNCH((k_sl_pack_ghosts), (NL + T - 1) / T, T, c->st, NL, c->pos, c->gmask, c->orig, f0, f1, s->scan[0], s->scan[1], s->send_idx[0], s->send_idx[1],
                                                       s->sbuf[0], s->sbuf[1]);
    neighbour_exchange(c, s, s->n_send, s->n_recv, GH_W);
    if (s->n_recv[0] > 0)
        LAUNCH((k_sl_unpack_ghosts), (s->n_recv[0] + T - 1) / T, T, c->st, s->n_recv[0], s->rbuf[0], NL, c->pos, c->vel, c->gmask, c->orig, s->ghost_slot[0]);
    if (s->n_recv[1] > 0)
        LAUNCH((k_sl_unpack_ghosts), (s->n_recv[1] + T - 1) / T, T, c->st, s->n_recv[1], s->rbuf[1], NL + s->n_recv[0], c->pos, c->vel, c->gmask, c->orig,
                                                                       s->ghost_slot[1]);
    s->n_ghost = s->n_recv[0] + s->n_recv[1];
    c->N = NL + s->n_ghost;
    c->launches += 4;
    CK(cudaGetLastError());
}

void slab_after_reorder(pfmds_ctx* c) {
    Slab* s = c->slab;
    const int T = 256;
    for (int d = 0; d < 2; ++d) {
        if (s->n_send[d] > 0) LAUNCH((k_sl_remap), (s->n_send[d] + T - 1) / T, T, c->st, s->n_send[d], s->send_idx[d], c->newslot);
        if (s->n_recv[d] > 0) LAUNCH((k_sl_remap), (s->n_recv[d] + T - 1) / T, T, c->st, s->n_recv[d], s->ghost_slot[d], c->newslot);
    }
    c->launches += 4;
    if (s->p2p) slab_exchange_peer_slots(c, s);
}

// field 0: ghost positions, field 1: ghost pos.w (1/Eb)
void slab_exchange(pfmds_ctx* c, int field) {
    Slab* s = c->slab;
    const int T = 256;
    KTimer kt(c, KS_OTHER);
    if (s->p2p) {
        const int nl = s->n_send[0], nr = s->n_send[1], n = nl + nr;
        double4 *pl = s->peer_pos[0][s->peer_parity[0]], *pr = s->peer_pos[1][s->peer_parity[1]];
        if (field == 0 && slab_fused(c)) {
            // positions: pushed by the kick+drift kernel itself when it was the fused variant, else pushed here; either way
            // the first density kernel waits for the neighbours' flag in its prologue
            if (!s->pos_pushed) {
                LAUNCH((k_sl_wait), 1, 1, c->st, s->flags + 4, s->flags + 5, s->seq_done, c->err, s->timeout_ns);
                if (n > 0) LAUNCH((k_sl_push<0>), (n + T - 1) / T, T, c->st, nl, s->send_idx[0], s->pslot[0], pl, nr, s->send_idx[1], s->pslot[1], pr, c->pos);
                s->seq_pos += 1;
                LAUNCH((k_sl_signal), 1, 1, c->st, s->peer_flags[0] + 1, s->peer_flags[1] + 0, s->seq_pos);
                c->launches += 3;
            }
            s->pos_pushed = false;
            s->wait_pos_seq = s->seq_pos;
            CK(cudaGetLastError());
            return;
        }
        if (field == 0) {
            // my neighbours' force kernels of the previous step must be done with the old ghost positions
            LAUNCH((k_sl_wait), 1, 1, c->st, s->flags + 4, s->flags + 5, s->seq_done, c->err, s->timeout_ns);
            if (n > 0) LAUNCH((k_sl_push<0>), (n + T - 1) / T, T, c->st, nl, s->send_idx[0], s->pslot[0], pl, nr, s->send_idx[1], s->pslot[1], pr, c->pos);
            s->seq_pos += 1;
            LAUNCH((k_sl_signal), 1, 1, c->st, s->peer_flags[0] + 1, s->peer_flags[1] + 0, s->seq_pos);   // I am my left neighbour's right neighbour
            LAUNCH((k_sl_wait), 1, 1, c->st, s->flags + 0, s->flags + 1, s->seq_pos, c->err, s->timeout_ns);
            c->launches += 4;
        } else {
            if (n > 0) LAUNCH((k_sl_push<1>), (n + T - 1) / T, T, c->st, nl, s->send_idx[0], s->pslot[0], pl, nr, s->send_idx[1], s->pslot[1], pr, c->pos);
            s->seq_w += 1;
            LAUNCH((k_sl_signal), 1, 1, c->st, s->peer_flags[0] + 3, s->peer_flags[1] + 2, s->seq_w);
            LAUNCH((k_sl_wait), 1, 1, c->st, s->flags + 2, s->flags + 3, s->seq_w, c->err, s->timeout_ns);
            c->launches += 3;
        }
        CK(cudaGetLastError());
        return;
    }
    for (int d = 0; d < 2; ++d) {
        if (s->n_send[d] == 0) continue;
        if (field == 0) LAUNCH((k_sl_pack_halo<0>), (s->n_send[d] + T - 1) / T, T, c->st, s->n_send[d], s->send_idx[d], c->pos, s->sbuf[d]);
        else LAUNCH((k_sl_pack_halo<1>), (s->n_send[d] + T - 1) / T, T, c->st, s->n_send[d], s->send_idx[d], c->pos, s->sbuf[d]);
    }
    neighbour_exchange(c, s, s->n_send, s->n_recv, field == 0 ? 3 : 1);
    for (int d = 0; d < 2; ++d) {
        if (s->n_recv[d] == 0) continue;
        if (field == 0) LAUNCH((k_sl_unpack_halo<0>), (s->n_recv[d] + T - 1) / T, T, c->st, s->n_recv[d], s->ghost_slot[d], s->rbuf[d], c->pos);
        else LAUNCH((k_sl_unpack_halo<1>), (s->n_recv[d] + T - 1) / T, T, c->st, s->n_recv[d], s->ghost_slot[d], s->rbuf[d], c->pos);
    }
    c->launches += 4;
    CK(cudaGetLastError());
}
