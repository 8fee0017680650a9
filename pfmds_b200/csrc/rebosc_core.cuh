// pfmds_b200 — `rebosc` (INTERACTION_POTENTIALS/REBOsolidcarbon.f90:27-91) on the device: the potential energy and the
// forces the reference obtains by central differences of that energy (calculate_forces_numerically,
// MOLECULAR_DYNAMICS/md_interactions.f90:273-311, dx = 1e-6).
//
// The reference copies a truncated neighbour list around every atom (three shells), shifts the cached dr by -+dx and calls
// the whole-list energy routine six times per atom: O(N) work per call, O(N^2) per step.  The difference E(-dx) - E(+dx)
// only feels the pair terms that depend on the position of the shifted atom m:
//     E_ab = f_c(r_ab) [ V_R(r_ab) - ((bsp_ab + bsp_ba)/2 + T bdh_ab) V_A(r_ab) ],
// which involves the bonds of a and of b, so E_ab depends on x_m iff a or b is m or a bonded (r < R2) neighbour of m.
// One thread per (atom, axis) evaluates exactly those pairs twice (x_m -+ dx) straight from the resident positions and the
// ELL list — no list copies, no O(N) sweeps — and takes the same central difference.  Everything else of the truncated-list
// energy cancels in the reference's subtraction (up to its rounding noise, eps * E_cluster / dx ~ 1e-8 eV/A; the parity tests
// state that tolerance).  FP64, latency bound (three bonds per carbon atom).
#pragma once
#include "common.cuh"

#define REB_SMAX 16  // bonded neighbours (r < R2) one atom may have

struct Shift { int atom; double d[3]; };  // atom == -1: nobody moves

__device__ __forceinline__ double4 reb_pos(const double4* __restrict__ pos, int a, const Shift& s) {
    double4 p = pos[a];
    if (a == s.atom) { p.x += s.d[0]; p.y += s.d[1]; p.z += s.d[2]; }
    return p;
}
__device__ __forceinline__ void reb_bond(const double4& pa, const double4& pb, const BoxD& box, double d[3], double& r) {
    d[0] = min_image(pb.x - pa.x, box.h[0], box.L[0]);
    d[1] = min_image(pb.y - pa.y, box.h[1], box.L[1]);
    d[2] = min_image(pb.z - pa.z, box.h[2], box.L[2]);
    r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}
__device__ __forceinline__ double reb_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double reb_g(const REBp& P, double c) {  // REBOsolidcarbon.f90:47-48
    double c2 = c * c, c3 = c2 * c, c4 = c2 * c2, c5 = c4 * c;
    return P.g[0] + P.g[1] * c + P.g[2] * c2 + P.g[3] * c3 + P.g[4] * c4 + P.g[5] * c5;
}

// One pair term E_ab (REBOsolidcarbon.f90:41-59 for both bond orders and the dihedral sum, :73-79 for the energy).
__device__ inline double reb_pair_energy(int a, int b, const double4* __restrict__ pos, const ListView& lv, const REBp& P, const BoxD& box, const Shift& sh) {
    const double4 pa = reb_pos(pos, a, sh), pb = reb_pos(pos, b, sh);
    double dab[3], r;
    reb_bond(pa, pb, box, dab, r);
    if (!(r < P.R2)) return 0.;
    const int na = lv.nnum[a], nb = lv.nnum[b];
    double sum_ab = 0., sum_ba = 0., bdh = 0.;
    const double aa = r * r;
    for (int q = 0; q < na; ++q) {
        const int qa = lv.nlist[(size_t)q * lv.stride + a];
        if (qa == b) continue;
        double daq[3], rq;
        reb_bond(pa, reb_pos(pos, qa, sh), box, daq, rq);
        if (!(rq < P.R2)) continue;
        const double fq = fcut_only(rq, P.R1, P.R2);
        const double ab = reb_dot(dab, daq);
        sum_ab += fq * reb_g(P, ab / (r * rq));
        for (int l = 0; l < nb; ++l) {
            const int la = lv.nlist[(size_t)l * lv.stride + b];
            if (la == a) continue;
            double dbl[3], rl;
            reb_bond(pb, reb_pos(pos, la, sh), box, dbl, rl);
            if (!(rl < P.R2)) continue;
            const double ac = reb_dot(dab, dbl), bc = reb_dot(daq, dbl);
            const double num = aa * bc - ab * ac;
            bdh += fq * fcut_only(rl, P.R1, P.R2) * (1. - num * num / (aa * (rq * rq) - ab * ab) / (aa * (rl * rl) - ac * ac));
        }
    }
    for (int l = 0; l < nb; ++l) {
        const int la = lv.nlist[(size_t)l * lv.stride + b];
        if (la == a) continue;
        double dbl[3], rl;
        reb_bond(pb, reb_pos(pos, la, sh), box, dbl, rl);
        if (!(rl < P.R2)) continue;
        sum_ba += fcut_only(rl, P.R1, P.R2) * reb_g(P, -reb_dot(dab, dbl) / (r * rl));  // the bond seen from b points the other way
    }
    const double bsp_ab = pow(1. + sum_ab, -0.5), bsp_ba = pow(1. + sum_ba, -0.5);
    return fcut_only(r, P.R1, P.R2) * ((1. + P.Q / r) * P.A * exp(-P.alpha * r) -
                                       ((bsp_ab + bsp_ba) / 2 + P.T * bdh) * (P.B[0] * exp(-P.beta[0] * r) + P.B[1] * exp(-P.beta[1] * r) + P.B[2] * exp(-P.beta[2] * r)));
}

// REBOsc_energy, one list-owner atom: every bonded pair once, from the end with the lower slot index.
__device__ __forceinline__ double reb_energy_thread(int i, const double4* __restrict__ pos, const ListView& lv, const REBp& P, const BoxD& box) {
    const int n = lv.nnum[i];
    Shift none;
    none.atom = -1; none.d[0] = none.d[1] = none.d[2] = 0.;
    double e = 0.;
    for (int p = 0; p < n; ++p) {
        const int j = lv.nlist[(size_t)p * lv.stride + i];
        if (j > i) e += reb_pair_energy(i, j, pos, lv, P, box, none);
    }
    return e;
}

// calculate_forces_numerically for rebosc, one (atom m, axis k):  F_k(m) += (E(x_m - dx e_k) - E(x_m + dx e_k)) / 2 / dx
__device__ __forceinline__ void reb_numforce_thread(int m, int k, const double4* __restrict__ pos, double4* __restrict__ frc, const ListView& lv,
                                                    const REBp& P, const BoxD& box, const int* __restrict__ orig, int* err) {
    const int nm = lv.nnum[m];
    if (nm == 0) return;
    const double dx = 1.0e-6;
    // bonded neighbours of m (with a margin far above dx, so that a shift cannot bring anybody else inside R2)
    int S[REB_SMAX];
    int ns = 0;
    {
        const double4 pm = pos[m];
        for (int p = 0; p < nm; ++p) {
            const int j = lv.nlist[(size_t)p * lv.stride + m];
            double d[3], r;
            reb_bond(pm, pos[j], box, d, r);
            if (r < P.R2 + 1.0e-4) {
                if (ns < REB_SMAX) S[ns] = j;
                ++ns;
            }
        }
        if (ns > REB_SMAX) { raise_error(err, E_TOO_MANY, orig[m], ns); return; }
    }
    double e12[2];
    for (int s = 0; s < 2; ++s) {
        Shift sh;
        sh.atom = m; sh.d[0] = sh.d[1] = sh.d[2] = 0.;
        sh.d[k] = s == 0 ? -dx : dx;
        double e = 0.;
        for (int p = 0; p < nm; ++p) e += reb_pair_energy(m, lv.nlist[(size_t)p * lv.stride + m], pos, lv, P, box, sh);
        for (int u = 0; u < ns; ++u) {
            const int a = S[u], na = lv.nnum[a];
            for (int p = 0; p < na; ++p) {
                const int b = lv.nlist[(size_t)p * lv.stride + a];
                if (b == m) continue;  // counted above as (m, a)
                bool in_s = false;
                for (int w = 0; w < ns; ++w) in_s |= S[w] == b;
                if (in_s && b < a) continue;  // both ends bonded to m: counted once, from the lower slot
                e += reb_pair_energy(a, b, pos, lv, P, box, sh);
            }
        }
        e12[s] = e;
    }
    reinterpret_cast<double*>(&frc[m])[k] += (e12[0] - e12[1]) / 2 / dx;
}
