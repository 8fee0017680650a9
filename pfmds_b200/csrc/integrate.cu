// pfmds_b200 — velocity-Verlet kicks/drifts, Nose-Hoover chain, statics quench, group reductions.
// Replaces code_source/MOLECULAR_DYNAMICS/md_integrators.f90 (all) and the reductions / checks of
// md_general.f90:96-112,161-311,342-398.  Group membership is a per-atom bit mask, so one pass over
// the slot-ordered arrays serves any settings-file group.  All reductions are two-stage with a
// fixed grid and a fixed summation order: results are run-to-run reproducible (the reference's
// OpenMP partial sums are not).
#include <cstdio>
#include <string>

#include "ctx.hpp"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x; } while (0)
#include "integ_bodies.cuh"  // IT, outside(), the per-atom bodies of the kick / drift / sum kernels

__global__ void k_check_positions(int N, const double4* __restrict__ pos, const int* __restrict__ orig, BoxD box, int* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double4 p = pos[i];
    if (outside(p.x, box.L[0]) || outside(p.y, box.L[1]) || outside(p.z, box.L[2])) raise_error(err, E_OUT_OF_CELL, orig[i], 0);
}
void integ_check_positions(pfmds_ctx* c) {
    LAUNCH((k_check_positions), (c->N + IT - 1) / IT, IT, c->st, c->N, c->pos, c->orig, c->box, c->err);
    c->launches += 1;
}

// md_general.f90:382-398
__global__ void k_invert_z(int N, const double4* __restrict__ pos, double4* __restrict__ vel, double zl, double zh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;  // ghost copies carry zero velocities: the test below never fires for them
    double z = pos[i].z, vz = vel[i].z;
    if ((z > zl && z < (zl + zh) / 2 && vz > 0.) || (z < zh && z > (zl + zh) / 2 && vz < 0.)) vel[i].z = -vz;
}
void integ_invert_z(pfmds_ctx* c) {
    LAUNCH((k_invert_z), (c->N + IT - 1) / IT, IT, c->st, c->N, c->pos, c->vel, 0.8 * c->box.L[2], 0.9 * c->box.L[2]);
    c->launches += 1;
}

// ---- kinetic energy of a group: md_general.f90:161-182 ---------------------------------------------
__global__ void __launch_bounds__(IT) k_ke_partial(int N, const double4* __restrict__ vel, const uint32_t* __restrict__ gmask, uint32_t bit,
                                                   double* __restrict__ part) {
    double s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if ((gmask[i] & (bit | PFMDS_GHOST)) == bit) {
            double4 v = vel[i];
            s += v.w * (v.x * v.x + v.y * v.y + v.z * v.z) / 2 * PFMDS_MASS_COEF;
        }
    }
    s = block_sum(s);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void k_sum_to(int n, const double* __restrict__ part, double* out) {
    double s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    s = block_sum(s);
    if (threadIdx.x == 0) *out = s;
}
void integ_kinetic_energy(pfmds_ctx* c, int group, double* d_out) {
    LAUNCH((k_ke_partial), RED_BLOCKS, IT, c->st, c->N, c->vel, c->gmask, 1u << (group - 1), c->part);
    LAUNCH((k_sum_to), 1, 1024, c->st, RED_BLOCKS, c->part, d_out);
    c->launches += 2;
    if (c->slab) slab_allreduce_sum(c, d_out, 1);
}

// (nhc_chain, the Nose-Hoover chain half step of md_integrators.f90:200-245, lives in common.cuh)
// One block sums the KE partials (fixed order), thread 0 runs the chain.
__global__ void __launch_bounds__(1024) k_nhc(int nparts, const double* __restrict__ part, double* state, int M, int L, double temperature, double ts2, double ts3,
                      double ts4) {
    double ke = 0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) ke += part[i];
    ke = block_sum(ke);
    if (threadIdx.x != 0) return;
    nhc_step(state, M, L, temperature, ke, ts2, ts3, ts4, 0);
}
// scale_velocities, md_general.f90:96-112
__global__ void k_scale(int N, double4* __restrict__ vel, const uint32_t* __restrict__ gmask, uint32_t bit, const double* __restrict__ s_ptr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if ((gmask[i] & (bit | PFMDS_GHOST)) == bit) {
        double s = *s_ptr;
        double4 v = vel[i];
        v.x *= s; v.y *= s; v.z *= s;
        vel[i] = v;
    }
}
void integ_nhc_half(pfmds_ctx* c, Nhc& t, double dt) {
    uint32_t bit = 1u << (t.group - 1);
    KTimer kt(c, KS_NHC);
    LAUNCH((k_ke_partial), RED_BLOCKS, IT, c->st, c->N, c->vel, c->gmask, bit, c->part);
    if (c->slab) {  // sum over ranks first, then every rank runs the same chain update on the same number
        LAUNCH((k_sum_to), 1, 1024, c->st, RED_BLOCKS, c->part, c->red + 48);
        slab_allreduce_sum(c, c->red + 48, 1);
        LAUNCH((k_nhc), 1, 32, c->st, 1, c->red + 48, t.state, t.M, t.L, t.temperature, dt / 2, dt / 4, dt / 8);
        c->launches += 1;
    } else
    LAUNCH((k_nhc), 1, 1024, c->st, RED_BLOCKS, c->part, t.state, t.M, t.L, t.temperature, dt / 2, dt / 4, dt / 8);
    LAUNCH((k_scale), (c->N + IT - 1) / IT, IT, c->st, c->N, c->vel, c->gmask, bit, t.state + 3 * t.M);
    c->launches += 3;
}

// ---- velocity Verlet: md_integrators.f90:7-97 ------------------------------------------------------
// First half of a step: half-kick then drift + one wrap.  xyz group: all components; z group: z only.
// The position test of check_positions is applied to the positions the step starts from.
__global__ void __launch_bounds__(IT) k_kick_drift(int N, double4* __restrict__ pos, double4* __restrict__ vel, const double4* __restrict__ frc,
                                                   const uint32_t* __restrict__ gmask, const int* __restrict__ orig, uint32_t bxyz, uint32_t bz,
                                                   double ts1, double ts2, BoxD box, int* err) {
    d_kick_drift(blockIdx.x * blockDim.x + threadIdx.x, N, pos, vel, frc, gmask, orig, bxyz, bz, ts1, ts2, box, err);
}
void integ_kick_drift(pfmds_ctx* c, double dt) {
    KTimer kt(c, KS_KICK_DRIFT);
    LAUNCH((k_kick_drift), (c->N + IT - 1) / IT, IT, c->st, c->N, c->pos, c->vel, c->frc, c->gmask, c->orig, 1u << (c->xyz_moving - 1),
                                                         1u << (c->z_moving - 1), dt, dt / 2, c->box, c->err);
    c->launches += 1;
}
__global__ void __launch_bounds__(IT) k_kick(int N, double4* __restrict__ vel, const double4* __restrict__ frc, const uint32_t* __restrict__ gmask,
                                             uint32_t bxyz, uint32_t bz, double ts2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t g = gmask[i];
    double4 v = vel[i], f = frc[i];   // requested together with the mask (integ_bodies.cuh d_kick_drift)
    if (g & PFMDS_GHOST) return;
    bool mx = g & bxyz, mz = g & bz;
    if (!mx && !mz) return;
    if (mx) {
        v.x = v.x + f.x / v.w / PFMDS_MASS_COEF * ts2;
        v.y = v.y + f.y / v.w / PFMDS_MASS_COEF * ts2;
        v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
    }
    if (mz) v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
    vel[i] = v;
}
void integ_kick(pfmds_ctx* c, double dt) {
    KTimer kt(c, KS_KICK);
    LAUNCH((k_kick), (c->N + IT - 1) / IT, IT, c->st, c->N, c->vel, c->frc, c->gmask, 1u << (c->xyz_moving - 1), 1u << (c->z_moving - 1), dt / 2);
    c->launches += 1;
}

// ---- molecular statics quench: md_integrators.f90:99-145 -------------------------------------------
__global__ void k_quench(int N, double4* __restrict__ vel, const double4* __restrict__ frc, const uint32_t* __restrict__ gmask, uint32_t bxyz,
                         uint32_t bz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t g = gmask[i];
    if (g & PFMDS_GHOST) return;
    bool mx = g & bxyz, mz = g & bz;
    if (!mx && !mz) return;
    double4 v = vel[i], f = frc[i];
    if (mx) {
        double fv = f.x * v.x + f.y * v.y + f.z * v.z, ff = f.x * f.x + f.y * f.y + f.z * f.z;
        if (fv > 0. && ff > 1.0e-12) { v.x = fv / ff * f.x; v.y = fv / ff * f.y; v.z = fv / ff * f.z; }
        else { v.x = 0.; v.y = 0.; v.z = 0.; }
    }
    if (mz) {
        double fv = f.x * v.x + f.y * v.y + f.z * v.z;
        if (!(fv > 0.)) { v.x = 0.; v.y = 0.; v.z = 0.; }
    }
    vel[i] = v;
}
void integ_quench(pfmds_ctx* c) {
    LAUNCH((k_quench), (c->N + IT - 1) / IT, IT, c->st, c->N, c->vel, c->frc, c->gmask, 1u << (c->xyz_moving - 1), 1u << (c->z_moving - 1));
    c->launches += 1;
}

// ---- group sums: momentum removal and the stdout diagnostics ---------------------------------------
// part layout per block: [0..2] sum F, [3..5] sum m x, [6..8] sum m v, [9] sum m, [10] max |v|^2 (all atoms)
__global__ void __launch_bounds__(IT) k_sums_partial(int N, const double4* __restrict__ pos, const double4* __restrict__ vel,
                                                     const double4* __restrict__ frc, const uint32_t* __restrict__ gmask, uint32_t bit,
                                                     double* __restrict__ part) {
    double a[11];
    for (int k = 0; k < 11; ++k) a[k] = 0.;
    a[10] = -1.;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        uint32_t gm = gmask[i];
        if (gm & PFMDS_GHOST) continue;
        double4 v = vel[i];
        double v2 = v.x * v.x + v.y * v.y + v.z * v.z;
        if (a[10] < v2) a[10] = v2;
        if (gm & bit) {
            double4 p = pos[i], f = frc[i];
            a[0] += f.x; a[1] += f.y; a[2] += f.z;
            a[3] += v.w * p.x; a[4] += v.w * p.y; a[5] += v.w * p.z;
            a[6] += v.w * v.x; a[7] += v.w * v.y; a[8] += v.w * v.z;
            a[9] += v.w;
        }
    }
    for (int k = 0; k < 10; ++k) {
        double s = block_sum(a[k]);
        if (threadIdx.x == 0) part[blockIdx.x * 16 + k] = s;
    }
    // block max
#ifdef PFMDS_COOP
    __shared__ double mx[IT / 32];
    double m = a[10];
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) mx[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < IT / 32; ++w) m = fmax(m, mx[w]);
        part[blockIdx.x * 16 + 10] = m;
    }
#else  // host replay (host_emu.hpp): no lane exchange; the running maximum is complete when thread 0 runs
    double m = emu_block_max(a[10]);
    if (threadIdx.x == 0) part[blockIdx.x * 16 + 10] = m;
#endif
}
__global__ void k_sums_final(int nb, const double* __restrict__ part, double* out) {
    for (int k = 0; k < 10; ++k) {
        double s = 0;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) s += part[i * 16 + k];
        s = block_sum(s);
        if (threadIdx.x == 0) out[k] = s;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double m = -1.;
        for (int i = 0; i < nb; ++i) m = fmax(m, part[i * 16 + 10]);
        out[10] = m;
    }
}
void integ_diagnostics(pfmds_ctx* c, double* d_out) {
    LAUNCH((k_sums_partial), RED_BLOCKS, IT, c->st, c->N, c->pos, c->vel, c->frc, c->gmask, 1u << (c->all_atoms - 1), c->part);
    LAUNCH((k_sums_final), 1, 1024, c->st, RED_BLOCKS, c->part, d_out);
    c->launches += 2;
    if (c->slab) { slab_allreduce_sum(c, d_out, 10); slab_allreduce_max(c, d_out + 10, 1); }
}
// zero_momentum, md_general.f90:236-253
__global__ void k_sub_mcv(int N, double4* __restrict__ vel, const uint32_t* __restrict__ gmask, uint32_t bit, const double* __restrict__ sums) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if ((gmask[i] & (bit | PFMDS_GHOST)) == bit) {
        double totm = sums[9];
        double4 v = vel[i];
        v.x = v.x - sums[6] / totm; v.y = v.y - sums[7] / totm; v.z = v.z - sums[8] / totm;
        vel[i] = v;
    }
}
void integ_zero_momentum(pfmds_ctx* c) {
    integ_diagnostics(c, c->red + 16);
    LAUNCH((k_sub_mcv), (c->N + IT - 1) / IT, IT, c->st, c->N, c->vel, c->gmask, 1u << (c->all_atoms - 1), c->red + 16);
    c->launches += 1;
}

// ---- fused NVT path (thermostat groups pairwise disjoint, at most NHC_MAXF of them) ------------------
// The reference sums the kinetic energy twice per step (opening and closing half step).  Nothing but the
// closing scale changes the velocities between the closing chain update of step n and the opening one of
// step n+1 (invert_z only flips signs), so the opening KE is s_c^2 KE_c analytically: one reduction per
// step, fused into the closing kick; the two scalings are folded into the next kick+drift as one factor.
//   k_nhc_open  (1 thread / thermostat)  chain update with the cached KE  -> s_pending *= s_o
//   k_kick_drift                         v *= s_pending, kick, drift            (pending := 1 implicitly)
//   k_kick_ke                            closing kick + per-thermostat KE partials
//   k_nhc_close (1 block)                fixed-order sum, chain update         -> s_pending = s_c
__global__ void k_nhc_open(NhcPack P, double ts2, double ts3, double ts4) {
    int k = threadIdx.x;
    if (k >= P.n) return;
    double* st = P.state[k];
    int M = P.M[k];
    nhc_step(st, M, P.L[k], P.T[k], 0., ts2, ts3, ts4, 1);
}
__global__ void __launch_bounds__(IT) k_kick_drift_nvt(int N, double4* __restrict__ pos, double4* __restrict__ vel, const double4* __restrict__ frc,
                                                       const uint32_t* __restrict__ gmask, const int* __restrict__ orig, uint32_t bxyz, uint32_t bz,
                                                       double ts1, double ts2, BoxD box, NhcPack P, int* err, SlabDev S) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool pushed = false;
    STAMP_MIN(0);
    slab_wait(S);  // slab mode: the neighbours' force kernels are done with the ghost positions this kernel is about to overwrite
    d_kick_drift_nvt(i, N, pos, vel, frc, gmask, orig, bxyz, bz, ts1, ts2, box, P, err, S, pushed);
    STAMP_MAX(1);
    slab_signal(S, pushed);
}
__global__ void k_reset_pending(NhcPack P) {
    int k = threadIdx.x;
    if (k < P.n) P.state[k][3 * P.M[k] + 2] = 1.0;
}
__global__ void __launch_bounds__(IT) k_kick_ke(int N, double4* __restrict__ vel, const double4* __restrict__ frc, const uint32_t* __restrict__ gmask,
                                                uint32_t bxyz, uint32_t bz, double ts2, NhcPack P, double* __restrict__ part) {
    double ke[NHC_MAXF];
    for (int k = 0; k < NHC_MAXF; ++k) ke[k] = 0.;
    // fixed grid (RED_BLOCKS: the order of the partial sums), so a thread walks several atoms: the records of the next one are
    // requested before the current one is worked on (the loop was a chain of round trips to memory, one per atom)
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t g = 0;
    double4 v = make_double4(0., 0., 0., 0.), f = v;
    if (i < N) { g = gmask[i]; v = vel[i]; f = frc[i]; }
    while (i < N) {
        const int in = i + stride;
        uint32_t g2 = 0;
        double4 v2 = make_double4(0., 0., 0., 0.), f2 = v2;
        if (in < N) { g2 = gmask[in]; v2 = vel[in]; f2 = frc[in]; }
        do {
            if (g & PFMDS_GHOST) break;
            bool mx = g & bxyz, mz = g & bz;
            bool th = false;
            for (int k = 0; k < P.n; ++k) th |= (g & P.bit[k]) != 0;
            if (!mx && !mz && !th) break;
            if (mx || mz) {
                if (mx) {
                    v.x = v.x + f.x / v.w / PFMDS_MASS_COEF * ts2;
                    v.y = v.y + f.y / v.w / PFMDS_MASS_COEF * ts2;
                    v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
                }
                if (mz) v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
                vel[i] = v;
            }
            double e = v.w * (v.x * v.x + v.y * v.y + v.z * v.z) / 2 * PFMDS_MASS_COEF;
            for (int k = 0; k < P.n; ++k)
                if (g & P.bit[k]) ke[k] += e;
        } while (false);
        i = in; g = g2; v = v2; f = f2;
    }
    for (int k = 0; k < P.n; ++k) {
        double s = block_sum(ke[k]);
        if (threadIdx.x == 0) part[blockIdx.x * NHC_MAXF + k] = s;
    }
}
// also_open: the next step follows at once inside the same call (nothing reads the chain in between), so its opening half step
// -- k_nhc_open's statements, on the kinetic energy this kernel has just cached -- runs here too: one launch less on the critical
// path of every step of a small system.
__global__ void __launch_bounds__(1024) k_nhc_close(int nparts, const double* __restrict__ part, NhcPack P, double ts2, double ts3, double ts4, int also_open) {
    for (int k = 0; k < P.n; ++k) {
        double ke = 0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) ke += part[i * NHC_MAXF + k];
        ke = block_sum(ke);
        if (threadIdx.x == 0) {
            double* st = P.state[k];
            int M = P.M[k];
            nhc_step(st, M, P.L[k], P.T[k], ke, ts2, ts3, ts4, also_open ? 3 : 2);
        }
        __syncthreads();
    }
}
// ---- small systems: sum of the per-interaction force buffers (+ closing kick, + thermostat KE partials) ----------------------
// The interactions of a step ran concurrently, each accumulating into its own buffer (ctx.hpp fbuf).  Per atom: start from what
// zero_forces leaves (0 inside the all_atoms group, the old force outside it: md_integrators.f90:147-163), add the buffers in file
// order -- the sequence of additions of the one-after-the-other path, so the same bits -- store, clear the buffers for the next
// step, then k_kick / k_kick_ke's arithmetic unchanged (same grid, same partial sums: the same kinetic energy bits).
template <int MODE>  // 0: sum only (step 0, restore); 1: + closing half kick; 2: + KE partial sums of the thermostat groups
__global__ void __launch_bounds__(IT) k_sum_kick_ke(int N, double4* __restrict__ vel, double4* __restrict__ frc, const uint32_t* __restrict__ gmask, FBufs F,
                                                    int zero_all, uint32_t ball, uint32_t bxyz, uint32_t bz, double ts2, NhcPack P, double* __restrict__ part,
                                                    unsigned int* ticket, int also_open) {
    double ke[NHC_MAXF];
    STAMP_MIN(4);
    for (int k = 0; k < NHC_MAXF; ++k) ke[k] = 0.;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        d_sum_kick_atom<MODE>(i, vel, frc, gmask, F, zero_all, ball, bxyz, bz, ts2, P, ke);
    }
    STAMP_MAX(5);
    if (MODE == 2) {
        for (int k = 0; k < P.n; ++k) {
            double s = block_sum(ke[k]);
            if (threadIdx.x == 0) part[blockIdx.x * NHC_MAXF + k] = s;
        }
        STAMP_MAX(6);
#ifdef PFMDS_COOP
        // The block that finishes last closes the thermostat step (k_nhc_close's statements: partials summed in block order by one
        // block, chain update by its thread 0), so small systems, whose steps are chains of 5-us kernels, lose one link of the chain.
        if (ticket) {
            __shared__ bool last;
            if (threadIdx.x == 0) {
                __threadfence();                                   // this block's partials before its ticket
                last = atomicAdd(ticket, 1u) == gridDim.x - 1;
            }
            __syncthreads();
            if (last) {
                STAMP_MAX(7);
                __threadfence();                                   // every other block's partials after their tickets
                const volatile double* vp = part;
                for (int k = 0; k < P.n; ++k) {
                    double sk = 0.;
                    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) sk += vp[(size_t)b * NHC_MAXF + k];
                    sk = block_sum(sk);
                    STAMP_MAX(8);
                    if (threadIdx.x == 0) {
                        double* st = P.state[k];
                        const int M = P.M[k];
                        nhc_step(st, M, P.L[k], P.T[k], sk, ts2, ts2 / 2, ts2 / 4, also_open ? 3 : 2);
                    }
                    __syncthreads();
                }
                STAMP_MAX(9);
                if (threadIdx.x == 0) { *ticket = 0u; STAMP_NEXT_STEP(); }   // ready for the next launch
            }
        }
#endif
    }
}
#if defined(__CUDACC__) && defined(PFMDS_STAMPS)
STAMP_BIND_FN(integ_stamps_bind)
#endif
static NhcPack pack_of(pfmds_ctx* c) {
    NhcPack P{};
    P.n = (int)c->nhc.size();
    for (int k = 0; k < P.n; ++k) {
        P.bit[k] = 1u << (c->nhc[k].group - 1); P.state[k] = c->nhc[k].state; P.M[k] = c->nhc[k].M; P.L[k] = c->nhc[k].L; P.T[k] = c->nhc[k].temperature;
    }
    return P;
}
void integ_nvt_open_kick_drift(pfmds_ctx* c, double dt, bool rebuild_step) {
    NhcPack P = pack_of(c);
    KTimer kt(c, KS_KICK_DRIFT);
    if (!c->nhc_opened) { LAUNCH((k_nhc_open), 1, 32, c->st, P, dt / 2, dt / 4, dt / 8); c->launches += 1; }  // else: done by the previous step's k_nhc_close
    c->nhc_opened = false;
    SlabDev S{};
    if (c->slab && slab_pos_pushed_by_kick(c, rebuild_step)) S = slab_dev(c, 0);
    LAUNCH((k_kick_drift_nvt), (c->N + IT - 1) / IT, IT, c->st, c->N, c->pos, c->vel, c->frc, c->gmask, c->orig, 1u << (c->xyz_moving - 1),
                                                             1u << (c->z_moving - 1), dt, dt / 2, c->box, P, c->err, S);
    // s_pending has been consumed; the closing half step of this same step overwrites it (k_nhc_close), so no reset here
    c->launches += 1;
}
__global__ void k_reduce_ke_partials(int nparts, const double* __restrict__ part, int n, double* __restrict__ out) {
    for (int k = 0; k < n; ++k) {
        double ke = 0;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) ke += part[i * NHC_MAXF + k];
        ke = block_sum(ke);
        if (threadIdx.x == 0) out[k] = ke;
        __syncthreads();
    }
}
void integ_nvt_kick_close(pfmds_ctx* c, double dt) {
    NhcPack P = pack_of(c);
    KTimer kt(c, KS_KICK);
    LAUNCH((k_kick_ke), RED_BLOCKS, IT, c->st, c->N, c->vel, c->frc, c->gmask, 1u << (c->xyz_moving - 1), 1u << (c->z_moving - 1), dt / 2, P, c->part);
    if (c->slab && slab_ke_close(c, P, RED_BLOCKS, c->part, dt / 2, dt / 4, dt / 8)) {
        // one kernel: local sums, all-gather through the ranks' mailboxes (peer memory), rank-ordered sum, chain update
    } else if (c->slab) {  // rank-local sums, one all-reduce, then every rank runs the same chain update
        LAUNCH((k_reduce_ke_partials), 1, 1024, c->st, RED_BLOCKS, c->part, P.n, c->red + 48);
        slab_allreduce_sum(c, c->red + 48, P.n);
        LAUNCH((k_nhc_close), 1, 32, c->st, 1, c->red + 48, P, dt / 2, dt / 4, dt / 8, 0);
        c->launches += 1;
    } else {
        LAUNCH((k_nhc_close), 1, 1024, c->st, RED_BLOCKS, c->part, P, dt / 2, dt / 4, dt / 8, (int)c->pre_open);
        c->nhc_opened = c->pre_open;
    }
    c->launches += 2;
}
void integ_sum_forces(pfmds_ctx* c, int mode, double dt) {
    FBufs F{};
    F.n = (int)c->fbuf.size();
    for (int t = 0; t < F.n; ++t) F.b[t] = c->fbuf[(size_t)t];
    const uint32_t ball = 1u << (c->all_atoms - 1), bxyz = 1u << (c->xyz_moving - 1), bz = 1u << (c->z_moving - 1);
    NhcPack P{};
    if (mode == 2) P = pack_of(c);
    KTimer kt(c, KS_KICK);
    unsigned int* const none = nullptr;
    // only the blocks that have atoms (small systems: 42 of 592 at 10 648 atoms): fewer tickets, and the closing block adds the
    // partial sums of those blocks only -- the others contributed 0.0, so the kinetic energy keeps its bits
    const int nbk = (c->N + IT - 1) / IT < RED_BLOCKS ? ((c->N + IT - 1) / IT < 1 ? 1 : (c->N + IT - 1) / IT) : RED_BLOCKS;
    if (mode == 0) LAUNCH((k_sum_kick_ke<0>), nbk, IT, c->st, c->N, c->vel, c->frc, c->gmask, F, (int)c->zero_all, ball, bxyz, bz, 0., P, c->part, none, 0);
    else if (mode == 1) LAUNCH((k_sum_kick_ke<1>), nbk, IT, c->st, c->N, c->vel, c->frc, c->gmask, F, (int)c->zero_all, ball, bxyz, bz, dt / 2, P, c->part, none, 0);
    else {
#ifdef PFMDS_COOP
        unsigned int* tk = c->ticket;    // the kernel's last block runs the chain update itself
#else
        unsigned int* tk = none;         // serial host replay: no block can know it is the last
#endif
        LAUNCH((k_sum_kick_ke<2>), tk ? nbk : RED_BLOCKS, IT, c->st, c->N, c->vel, c->frc, c->gmask, F, (int)c->zero_all, ball, bxyz, bz, dt / 2, P, c->part, tk, (int)c->pre_open);
        if (!tk) { LAUNCH((k_nhc_close), 1, 1024, c->st, RED_BLOCKS, c->part, P, dt / 2, dt / 4, dt / 8, (int)c->pre_open); c->launches += 1; }
        c->nhc_opened = c->pre_open;
    }
    c->launches += 1;
    c->fbuf_active = false;
}
// apply scalings that are still pending (before anything else reads or changes velocities)
void integ_flush_pending(pfmds_ctx* c) {
    if (!c->nhc_pending) return;
    NhcPack P = pack_of(c);
    for (int k = 0; k < P.n; ++k)
        LAUNCH((k_scale), (c->N + IT - 1) / IT, IT, c->st, c->N, c->vel, c->gmask, P.bit[k], P.state[k] + 3 * P.M[k] + 2);
    LAUNCH((k_reset_pending), 1, 32, c->st, P);
    c->launches += P.n + 1;
    c->nhc_pending = false;
}
