// pfmds_b200 -- per-atom bodies of the integrator kernels (integrate.cu wraps each in its __global__ kernel: index from blockIdx /
// threadIdx).  Kept apart from the launch geometry so that another driver of the same statements -- the persistent step kernel tried
// in round 2, DESIGN section 12 -- executes them on the same numbers.
#pragma once
#include "common.cuh"

#define IT 256  // threads per block of the integrator kernels: the KE partial sums are per block of IT consecutive atoms

// md_general.f90:342-364 -- every atom, tolerance 1e-7, negated condition so NaN is caught too
__device__ __forceinline__ bool outside(double x, double L) { return !(x > (0. - 0.0000001) && x < (L + 0.0000001)); }

// velocity Verlet, first half of a step (md_integrators.f90:7-97): half kick, drift, one wrap
__device__ __forceinline__ void d_kick_drift(int i, int N, double4* pos, double4* vel, const double4* frc, const uint32_t* gmask, const int* orig,
                                             uint32_t bxyz, uint32_t bz, double ts1, double ts2, const BoxD& box, int* err) {
    if (i >= N) return;
    // the three records are requested together with the mask, not after it: one round trip to memory per atom instead of two
    // (these kernels are latency bound: ncu, 3.3 TB/s at 45 % of the warp slots with the loads behind the mask test)
    uint32_t g = gmask[i];
    double4 p = pos[i], v = vel[i], f = frc[i];
    if (g & PFMDS_GHOST) return;
    bool mx = g & bxyz, mz = g & bz;
    if (!mx && !mz) return;
    if (outside(p.x, box.L[0]) || outside(p.y, box.L[1]) || outside(p.z, box.L[2])) raise_error(err, E_OUT_OF_CELL, orig[i], 0);
    if (mx) {
        v.x = v.x + f.x / v.w / PFMDS_MASS_COEF * ts2;
        v.y = v.y + f.y / v.w / PFMDS_MASS_COEF * ts2;
        v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
    }
    if (mz) v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
    if (mx) {
        p.x = p.x + v.x * ts1; if (p.x > box.L[0]) p.x = p.x - box.L[0]; else if (p.x < 0.) p.x = p.x + box.L[0];
        p.y = p.y + v.y * ts1; if (p.y > box.L[1]) p.y = p.y - box.L[1]; else if (p.y < 0.) p.y = p.y + box.L[1];
        p.z = p.z + v.z * ts1; if (p.z > box.L[2]) p.z = p.z - box.L[2]; else if (p.z < 0.) p.z = p.z + box.L[2];
    }
    if (mz) { p.z = p.z + v.z * ts1; if (p.z > box.L[2]) p.z = p.z - box.L[2]; else if (p.z < 0.) p.z = p.z + box.L[2]; }
    pos[i] = p;
    vel[i] = v;
}

// fused NVT: the pending thermostat scale (state[3M+2] of the atom's chain), then the same kick + drift.  `pushed` (slab mode): the
// new position went straight into the neighbours' ghost slots.
__device__ __forceinline__ void d_kick_drift_nvt(int i, int N, double4* pos, double4* vel, const double4* frc, const uint32_t* gmask, const int* orig,
                                                 uint32_t bxyz, uint32_t bz, double ts1, double ts2, const BoxD& box, const NhcPack& P, int* err,
                                                 const SlabDev& S, bool& pushed) {
    uint32_t g = i < N ? gmask[i] : PFMDS_GHOST;
    // requested together with the mask (see d_kick_drift); slots past N are not read
    double4 v = make_double4(0., 0., 0., 0.), p = v, f = v;
    if (i < N) { v = vel[i]; p = pos[i]; f = frc[i]; }
    bool mx = g & bxyz, mz = g & bz;
    double sc = 1.0;
    bool th = false;
    for (int k = 0; k < P.n; ++k)
        if (g & P.bit[k]) { sc = P.state[k][3 * P.M[k] + 2]; th = true; }
    if (!(g & PFMDS_GHOST) && (mx || mz || th)) {  // one exit point: the kernel's slab_signal() holds a block barrier
        v.x *= sc; v.y *= sc; v.z *= sc;
        if (mx || mz) {
            if (outside(p.x, box.L[0]) || outside(p.y, box.L[1]) || outside(p.z, box.L[2])) raise_error(err, E_OUT_OF_CELL, orig[i], 0);
            if (mx) {
                v.x = v.x + f.x / v.w / PFMDS_MASS_COEF * ts2;
                v.y = v.y + f.y / v.w / PFMDS_MASS_COEF * ts2;
                v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
            }
            if (mz) v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
            if (mx) {
                p.x = p.x + v.x * ts1; if (p.x > box.L[0]) p.x = p.x - box.L[0]; else if (p.x < 0.) p.x = p.x + box.L[0];
                p.y = p.y + v.y * ts1; if (p.y > box.L[1]) p.y = p.y - box.L[1]; else if (p.y < 0.) p.y = p.y + box.L[1];
                p.z = p.z + v.z * ts1; if (p.z > box.L[2]) p.z = p.z - box.L[2]; else if (p.z < 0.) p.z = p.z + box.L[2];
            }
            if (mz) { p.z = p.z + v.z * ts1; if (p.z > box.L[2]) p.z = p.z - box.L[2]; else if (p.z < 0.) p.z = p.z + box.L[2]; }
            pos[i] = p;
            if (S.push) {  // the new position goes straight into the ghost copies of this atom on the neighbour GPUs
                int a = S.rs_l[i], b = S.rs_r[i];
                if (a >= 0) { double* q = reinterpret_cast<double*>(&S.peer_l[a]); q[0] = p.x; q[1] = p.y; q[2] = p.z; }
                if (b >= 0) { double* q = reinterpret_cast<double*>(&S.peer_r[b]); q[0] = p.x; q[1] = p.y; q[2] = p.z; }
                pushed = (a >= 0) || (b >= 0);
            }
        }
        vel[i] = v;
    }
}

// ---- small systems: sum of the per-interaction force buffers (+ closing kick, + thermostat KE contributions) --------------------
// Per atom: start from what zero_forces leaves (0 inside the all_atoms group, the old force outside it: md_integrators.f90:147-163),
// add the buffers in file order -- the sequence of additions of the one-after-the-other path, so the same bits -- store, clear the
// buffers for the next step, then k_kick / k_kick_ke's arithmetic unchanged.  ke[k] += this atom's kinetic energy for chain k (MODE 2).
#define FBUF_MAX 12
struct FBufs { int n; double4* b[FBUF_MAX]; };
template <int MODE>  // 0: sum only (step 0, restore); 1: + closing half kick; 2: + KE of the thermostat groups
__device__ __forceinline__ void d_sum_kick_atom(int i, double4* vel, double4* frc, const uint32_t* gmask, const FBufs& F, int zero_all, uint32_t ball,
                                                uint32_t bxyz, uint32_t bz, double ts2, const NhcPack& P, double* ke) {
    // every load of this atom is issued before the first store: with the loads and the clearing stores interleaved buffer by
    // buffer (the pointers may alias as far as the compiler knows) the loop was a chain of dependent L2 round trips, 8 us of a
    // 10 648-atom step (tools/stamps_probe.py)
    const uint32_t g = gmask[i];
    double4 a[FBUF_MAX];
#pragma unroll
    for (int t = 0; t < FBUF_MAX; ++t)
        if (t < F.n) a[t] = F.b[t][i];
    double4 v = make_double4(0., 0., 0., 0.);
    if (MODE != 0) v = vel[i];
    double4 f = make_double4(0., 0., 0., 0.);
    if (!zero_all && !(g & ball)) f = frc[i];
#pragma unroll
    for (int t = 0; t < FBUF_MAX; ++t)
        if (t < F.n) {
            f.x += a[t].x; f.y += a[t].y; f.z += a[t].z;
            F.b[t][i] = make_double4(0., 0., 0., 0.);
        }
    frc[i] = f;
    if (MODE == 0 || (g & PFMDS_GHOST)) return;
    const bool mx = g & bxyz, mz = g & bz;
    bool th = false;
    if (MODE == 2)
        for (int k = 0; k < P.n; ++k) th |= (g & P.bit[k]) != 0;
    if (!mx && !mz && !th) return;
    if (mx || mz) {
        if (mx) {
            v.x = v.x + f.x / v.w / PFMDS_MASS_COEF * ts2;
            v.y = v.y + f.y / v.w / PFMDS_MASS_COEF * ts2;
            v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
        }
        if (mz) v.z = v.z + f.z / v.w / PFMDS_MASS_COEF * ts2;
        vel[i] = v;
    }
    if (MODE == 2) {
        double e = v.w * (v.x * v.x + v.y * v.y + v.z * v.z) / 2 * PFMDS_MASS_COEF;
        for (int k = 0; k < P.n; ++k)
            if (g & P.bit[k]) ke[k] += e;
    }
}
