// pfmds_b200 -- persistent step kernel of small systems (included by forces.cu, after the per-thread bodies of its kernels).
//
// A small system's step is a chain of 5-us kernels (tools/stamps_probe.py on a B200, A/B gas of 10 648 atoms, CUDA-graph replay: of
// the 35 us of a step 10 us are gaps between dependent graph nodes, 8 us the tail of the closing kernel).  Between two events that need
// the host or another code path -- list rebuild, momentum removal, a step that reports energies, the end of the call -- the steps
// are therefore run by ONE cooperative launch: the phases of a step are separated by grid barriers (one atomic + a spin on L2)
// instead of kernel boundaries, the thermostat chain lives in shared memory (every block updates its own copy with the same
// numbers), and the grid stays resident.
//
// Every phase calls the per-thread body of the kernel it replaces (d_lj, d_lj1g, d_rjl_*_split, d_tb_*, d_normals, d_cos_*:
// forces.cu; d_kick_drift*, d_sum_kick_atom: integ_bodies.cuh) with the same virtual thread numbering, the per-interaction force
// buffers are summed in file order and the kinetic-energy partial sums keep their blocks of IT atoms and the order of
// k_sum_kick_ke's closing block: the results are bit-identical to the step-by-step path (tests/test_persist_gpu.py asserts it).
#pragma once
#if defined(__CUDACC__)
#include "integ_bodies.cuh"

#define PB IT            // threads per block: the blocks of the kinetic-energy partial sums
#define P_MAXOPS 16
#define P_STAGES 3
#define P_MAX_STEPS 4096  // per launch (the barrier counter is 32 bits: blocks x barriers per step x steps)

enum { PO_LJ = 0, PO_LJ1G, PO_RJL_D, PO_RJL_F, PO_TB_BOND, PO_TB_FORCE, PO_TB_REDUCE, PO_NORMALS, PO_COS_G, PO_COS_M, PO_COS_IND };
enum { PG_SPLIT = 0, PG_ATOM, PG_BOND };   // virtual threads of an op: SMALL_SPLIT lanes per owner / one per owner / one per (owner, slot)

struct POp {
    int kind, stage, geom, maxn;
    int morse, simplified, spare0, spare1;
    ListView lv;
    const int* owners; const int* n_owners;   // the slots with a non-empty row of `lv`, ascending (k_owner_compact): only they get threads
    double4* out;                // the interaction's force buffer
    double4* gnorm; double4* tvec; double4* fpart; double* aux; double* aux2;
    double pref;
    union U { LJp lj; LJ1Gp lj1g; RjlD rd; RjlF rf; TBp tb; CosP cos; } u;
};
struct PArgs {
    int N, nsteps, nvt, last_mode;   // last_mode: thermostat mode of the last step's closing half step (3: also opens the next step)
    double4 *pos, *vel, *frc;
    const uint32_t* gmask; const int* orig;
    uint32_t ball, bxyz, bz; int zero_all;
    double dt;
    BoxD box; WrapC W;
    int* err;
    FBufs F;
    NhcPack P;
    double* part;
    unsigned int* bar;               // grid barrier counter, zero at launch
    int nops, stage_used[P_STAGES];
    POp ops[P_MAXOPS];
};

// The slots that own a non-empty row, in ascending order, and their number: one block, ballot / popc compaction chunk by chunk.
__global__ void __launch_bounds__(1024) k_owner_compact(int N, const int* __restrict__ nnum, int* __restrict__ owners, int* __restrict__ n_owners) {
    __shared__ int wsum[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int start = 0; start < N; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const bool own = i < N && nnum[i] > 0;
        const unsigned m = __ballot_sync(0xffffffffu, own);
        if (lane == 0) wsum[w] = __popc(m);
        __syncthreads();
        int off = base;
        for (int q = 0; q < w; ++q) off += wsum[q];
        if (own) owners[off + __popc(m & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += wsum[q]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_owners = base;
}

// Grid barrier (all blocks are resident: cooperative launch).  The counter only grows during a launch; `target` is this block's
// copy of the value it reaches when every block has arrived.  Thread 0 arrives with a release (this block's writes, ordered before it by
// the block barrier) and spins with acquire loads (SASS: CCTL.IVALL after each, so the L1 holds no stale line afterwards); the block
// barriers carry both to the other threads.
__device__ __forceinline__ void p_grid_bar(unsigned int* bar, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        unsigned int v;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
    }
    __syncthreads();
}

// virtual blocks (PB threads) an op needs for `no` owners
__device__ __forceinline__ int p_op_blocks(const POp& o, int no) {
    const int lanes = o.geom == PG_SPLIT ? no * SMALL_SPLIT : (o.geom == PG_BOND ? ((no + PB - 1) / PB * PB) * o.maxn : no);
    return (lanes + PB - 1) / PB;
}
// FULL: every interaction kind; otherwise the pair potentials only (lj, lj1g, rjl: a third block per SM fits the register file)
template <bool FULL>
__device__ __forceinline__ void p_run_op(const POp& o, const int t, const int no, const PArgs& A) {
    const int N = A.N;
    if (o.geom == PG_SPLIT) {
        const int q = t / SMALL_SPLIT;
        const int tt = (q < no ? o.owners[q] : N) * SMALL_SPLIT + t % SMALL_SPLIT;   // thread number of the kernel this op replaces
        switch (o.kind) {
        case PO_LJ: d_lj<true, false, SMALL_SPLIT>(tt, N, A.pos, o.out, o.lv, o.u.lj, A.box, nullptr); break;
        case PO_LJ1G: d_lj1g<true, false, SMALL_SPLIT>(tt, N, A.pos, o.out, o.lv, o.u.lj1g, A.box, nullptr); break;
        case PO_RJL_D: d_rjl_density_split<false, SMALL_SPLIT, RjlD>(tt, N, A.pos, o.lv, o.u.rd, A.box, A.W, nullptr); break;
        case PO_RJL_F: d_rjl_force_split<SMALL_SPLIT, RjlF, false>(tt, N, A.pos, o.out, o.lv, o.u.rf, A.box, A.W); break;
        case PO_COS_G:
            if (FULL) {
                if (o.morse) d_cos_direct<true, true, true, false, SMALL_SPLIT>(tt, N, A.pos, o.out, o.lv, o.u.cos, A.box, o.gnorm, o.tvec, nullptr);
                else d_cos_direct<false, true, true, false, SMALL_SPLIT>(tt, N, A.pos, o.out, o.lv, o.u.cos, A.box, o.gnorm, o.tvec, nullptr);
            }
            break;
        case PO_COS_M:
            if (FULL) {
                if (o.morse) d_cos_direct<true, false, true, false, SMALL_SPLIT>(tt, N, A.pos, o.out, o.lv, o.u.cos, A.box, o.gnorm, o.tvec, nullptr);
                else d_cos_direct<false, false, true, false, SMALL_SPLIT>(tt, N, A.pos, o.out, o.lv, o.u.cos, A.box, o.gnorm, o.tvec, nullptr);
            }
            break;
        }
    } else if (FULL && o.geom == PG_ATOM) {
        const int i = t < no ? o.owners[t] : N;
        switch (o.kind) {
        case PO_TB_REDUCE: d_tb_reduce(i, N, o.fpart, o.out, o.lv); break;
        case PO_NORMALS: d_normals(i, N, A.pos, o.lv, A.box, o.simplified, o.gnorm); break;
        case PO_COS_IND: d_cos_indirect(i, N, A.pos, o.out, o.lv, o.pref, A.box, o.gnorm, o.tvec); break;
        }
    } else if (FULL) {   // PG_BOND
        const int npad = (no + PB - 1) / PB * PB, q = t % npad, p = t / npad;
        const int i = q < no ? o.owners[q] : N;
        if (o.kind == PO_TB_BOND) d_tb_bond(i, p, N, A.pos, o.lv, o.u.tb, A.box, o.aux, o.aux2);
        else d_tb_force<true, false>(i, p, N, A.pos, o.fpart, o.lv, o.u.tb, A.box, o.aux, o.aux2, nullptr);
    }
}

// cuobjdump -res-usage: FULL 128 registers (2 blocks per SM), pair potentials only: 3 blocks per SM
#define P_NHC_W (3 * NHC_MLOC + 4)
template <bool FULL>
__global__ void __launch_bounds__(PB, FULL ? 2 : 3) k_persist(const __grid_constant__ PArgs A) {
    __shared__ double chain[NHC_MAXF][P_NHC_W];   // this block's copy of the thermostat chains (x, v, q, s, cached KE, pending scale)
    __shared__ int op_no[P_MAXOPS], op_vb0[P_MAXOPS], stage_nvb[P_STAGES];
    NhcPack P = A.P;
    if (A.nvt) {
        for (int k = 0; k < A.P.n; ++k) {
            for (int q = threadIdx.x; q < 3 * A.P.M[k] + 4; q += blockDim.x) chain[k][q] = A.P.state[k][q];
            P.state[k] = chain[k];
        }
    }
    if (threadIdx.x == 0) {   // the ops' ranges of virtual blocks inside their stages, from the owner counts (constant during the launch)
        for (int st = 0; st < P_STAGES; ++st) stage_nvb[st] = 0;
        for (int k = 0; k < A.nops; ++k) {
            const int no = *A.ops[k].n_owners;
            op_no[k] = no;
            op_vb0[k] = stage_nvb[A.ops[k].stage];
            stage_nvb[A.ops[k].stage] += p_op_blocks(A.ops[k], no);
        }
    }
    __syncthreads();
    const int nvb_atoms = (A.N + IT - 1) / IT;
    const double ts2 = A.dt / 2;
    unsigned int target = 0;
    for (int s = 0; s < A.nsteps; ++s) {
        // ---- opening: pending thermostat scale, half kick, drift (k_kick_drift_nvt / k_kick_drift) ----
        if (blockIdx.x == 0) STAMP_MIN(0);
        for (int vb = blockIdx.x; vb < nvb_atoms; vb += gridDim.x) {
            const int i = vb * IT + threadIdx.x;
            if (A.nvt) { bool pushed = false; d_kick_drift_nvt(i, A.N, A.pos, A.vel, A.frc, A.gmask, A.orig, A.bxyz, A.bz, A.dt, ts2, A.box, P, A.err, SlabDev{}, pushed); }
            else d_kick_drift(i, A.N, A.pos, A.vel, A.frc, A.gmask, A.orig, A.bxyz, A.bz, A.dt, ts2, A.box, A.err);
        }
        if (blockIdx.x == 0) STAMP_MAX(1);
        p_grid_bar(A.bar, target);
        if (blockIdx.x == 0) STAMP_MAX(2);
        // ---- the interactions, by dependency stage; each accumulates into its own force buffer ----
        for (int st = 0; st < P_STAGES; ++st) {
            if (!A.stage_used[st]) continue;
            const int nvb = stage_nvb[st];
            for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
                int k = 0;
                for (; k < A.nops; ++k)
                    if (A.ops[k].stage == st && vb >= op_vb0[k] && vb < op_vb0[k] + p_op_blocks(A.ops[k], op_no[k])) break;
                if (k < A.nops) p_run_op<FULL>(A.ops[k], (vb - op_vb0[k]) * PB + (int)threadIdx.x, op_no[k], A);
            }
            STAMP_MAX(3 + st);
            p_grid_bar(A.bar, target);
        }
        if (blockIdx.x == 0) STAMP_MAX(6);
        // ---- closing: buffers summed in file order, half kick, kinetic-energy partial sums (k_sum_kick_ke) ----
        for (int vb = blockIdx.x; vb < nvb_atoms; vb += gridDim.x) {
            const int i = vb * IT + threadIdx.x;
            double ke[NHC_MAXF];
            for (int k = 0; k < NHC_MAXF; ++k) ke[k] = 0.;
            if (i < A.N) {
                if (A.nvt) d_sum_kick_atom<2>(i, A.vel, A.frc, A.gmask, A.F, A.zero_all, A.ball, A.bxyz, A.bz, ts2, P, ke);
                else d_sum_kick_atom<1>(i, A.vel, A.frc, A.gmask, A.F, A.zero_all, A.ball, A.bxyz, A.bz, ts2, P, ke);
            }
            if (A.nvt)
                for (int k = 0; k < P.n; ++k) {
                    double sk = block_sum(ke[k]);
                    if (threadIdx.x == 0) A.part[vb * NHC_MAXF + k] = sk;
                }
        }
        if (blockIdx.x == 0) STAMP_MAX(7);
        if (A.nvt) {
            p_grid_bar(A.bar, target);
            if (blockIdx.x == 0) STAMP_MAX(8);
            // k_sum_kick_ke's closing block, run by every block on its own copy of the chains: partial sums in block order
            // (thread t adds blocks t, t + IT, ...: the blocks beyond the last atom contribute 0.0 there and nothing here)
            const int mode = s + 1 < A.nsteps ? 3 : A.last_mode;
            for (int k = 0; k < P.n; ++k) {
                double sk = 0.;
                for (int b = threadIdx.x; b < nvb_atoms; b += blockDim.x) sk += __ldcg(&A.part[(size_t)b * NHC_MAXF + k]);
                sk = block_sum(sk);
                if (threadIdx.x == 0) nhc_step(chain[k], P.M[k], P.L[k], P.T[k], sk, ts2, ts2 / 2, ts2 / 4, mode);
                __syncthreads();
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { STAMP_MAX(9); STAMP_NEXT_STEP(); }
    }
    if (A.nvt && blockIdx.x == 0)
        for (int k = 0; k < A.P.n; ++k)
            for (int q = threadIdx.x; q < 3 * A.P.M[k] + 4; q += blockDim.x) A.P.state[k][q] = chain[k][q];
}
#endif  // __CUDACC__
