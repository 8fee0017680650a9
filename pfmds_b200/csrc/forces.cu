// pfmds_b200 — force and energy kernels for lj, lj1g, ljc, morsec, rjl, tb (sm_100a, FP64).
//
// Every kernel is a gather: one thread owns one list-owner atom, walks its ELL row (coalesced int32
// loads), gathers the partner's 32-byte position record, recomputes the min-image dr and |dr| in
// registers and accumulates the force on its own atom only — no atomics, deterministic.  The
// reference does the same sums over cached dr/moddr arrays (INTERACTION_POTENTIALS/*.f90); parity
// is to 1e-9 relative, the summation order inside a row differs (cell order instead of ascending j).
//
// FP64-pipe bound: see DESIGN.md for the per-pair operation counts the roofline uses.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#if defined(__CUDACC__) || defined(PFMDS_EMU_LIB)
#include "ctx.hpp"
#else
#include "common.cuh"  // host emulation of the kernels (tests/forces_host.cpp): no context, no launchers
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x; } while (0)
#ifndef FT
#define FT 128  // threads per block for the force kernels (tools/rjl_variants.py builds other sizes for A/B runs)
#endif

struct Vec { double x, y, z; };
__device__ __forceinline__ double dot(const Vec& a, const Vec& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__device__ __forceinline__ Vec bond_vec(const double4& pi, const double4& pj, const BoxD& box, double& r2) {
    Vec d;
    d.x = min_image(pj.x - pi.x, box.h[0], box.L[0]);
    d.y = min_image(pj.y - pi.y, box.h[1], box.L[1]);
    d.z = min_image(pj.z - pi.z, box.h[2], box.L[2]);
    r2 = d.x * d.x + d.y * d.y + d.z * d.z;
    return d;
}

__device__ __forceinline__ void add_force(double4* frc, int i, double fx, double fy, double fz) {
    double4 f = frc[i];
    f.x += fx; f.y += fy; f.z += fz;
    frc[i] = f;
}

// Small systems: SPLIT lanes share one atom's row (slot p = sub, sub+SPLIT, ...) and combine with a fixed
// xor-shuffle tree, so a few thousand atoms still give the SMs enough warps.  SPLIT = 1 is thread-per-atom.
template <int SPLIT>
__device__ __forceinline__ double split_sum(double v) {
#pragma unroll
    for (int o = SPLIT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#define SMALL_SPLIT 8  // systems below pfmds_ctx::small_n atoms (runtime: PFMDS_SMALL_N) take the SPLIT kernels

// per-block partial of the per-thread energy; the final sum is done by k_sum_partials in block order
__device__ __forceinline__ void store_partial(double e, double* part) {
    double s = block_sum(e);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void k_sum_partials(int n, const double* __restrict__ part, double scale, double* out) {
    double s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    s = block_sum(s);
    if (threadIdx.x == 0) *out = s * scale;
}

// ---- lj : LennardJones.f90:23-69 ----------------------------------------------------------------
template <bool F, bool E, int SPLIT>
__global__ void __launch_bounds__(FT) k_lj(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, LJp P, BoxD box,
                                           double* part) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t / SPLIT, sub = t % SPLIT;
    double e = 0, fx = 0, fy = 0, fz = 0;
    STAMP_MIN(2);
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = pos[i];
        const double R22 = P.R2 * P.R2, s2 = P.sig * P.sig;
        for (int p = sub; p < n; p += SPLIT) {
            int j = lv.nlist[(size_t)p * lv.stride + i];
            double r2;
            Vec d = bond_vec(pi, pos[j], box, r2);
            double r = sqrt(r2);
            if (r < P.R2) {
                (void)R22;
                double ir2 = 1.0 / r2;
                double q = s2 * ir2, V = q * q * q;
                double f, dfr;
                fcut_dfcut(r, P.R1, P.R2, f, dfr);
                if (E) e += 4 * P.eps * V * (V - 1.) * f;
                if (F) {
                    double c = 4. * P.eps * (V * (12. * V - 6.) * ir2 * f - V * (V - 1.) * dfr);
                    fx -= c * d.x; fy -= c * d.y; fz -= c * d.z;
                }
            }
        }
    }
    if (F) {
        fx = split_sum<SPLIT>(fx); fy = split_sum<SPLIT>(fy); fz = split_sum<SPLIT>(fz);
        if (n > 0 && sub == 0) add_force(frc, i, fx, fy, fz);
    }
    STAMP_MAX(3);
    if (E) store_partial(e, part);
}

// ---- lj1g : LennardJones_1g.f90:28-117, cut_off_poly.f90 ------------------------------------------
// The reference visits each pair once (p <= lessnnum) and scatters +-F*dr to both atoms; with a
// symmetric full list the same sum is a gather over all neighbours.  The switch derivative keeps the
// reference's missing 1/(R2-R1) factor (cut_off_poly.f90:41).
template <bool F, bool E, int SPLIT>
__global__ void __launch_bounds__(FT) k_lj1g(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, LJ1Gp P, BoxD box,
                                             double* part) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t / SPLIT, sub = t % SPLIT;
    double e = 0, fx = 0, fy = 0, fz = 0;
    STAMP_MIN(2);
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = pos[i];
        const double R22 = P.R2 * P.R2, iw = 1.0 / (P.R2 - P.R1);
        for (int p = sub; p < n; p += SPLIT) {
            int j = lv.nlist[(size_t)p * lv.stride + i];
            double r2;
            Vec d = bond_vec(pi, pos[j], box, r2);
            if (r2 < R22) {  // beyond R2 the quintic switch and its derivative are exactly 0
                double invr2 = 1.0 / r2;
                double U = invr2 * invr2 * invr2;
                double f = 1.0, dfr = 0.0;
                if (r2 > P.R1 * P.R1) {
                    double r = sqrt(r2);
                    double x = (r - P.R1) * iw, x2 = x * x;
                    f = 1. + x2 * x * (-10. + 15. * x - 6. * x2);
                    dfr = r * x2 * (-30. + 60. * x - 30. * x2);
                }
                if (E) e += U * (P.c12 * U - P.c6) * f;
                if (F) {
                    double c = U * invr2 * ((P.c12t12 * U - P.c6t6) * f - (P.c12 * U - P.c6) * dfr);
                    fx -= c * d.x; fy -= c * d.y; fz -= c * d.z;
                }
            }
        }
    }
    if (F) {
        fx = split_sum<SPLIT>(fx); fy = split_sum<SPLIT>(fy); fz = split_sum<SPLIT>(fz);
        if (n > 0 && sub == 0) add_force(frc, i, fx, fy, fz);
    }
    STAMP_MAX(3);
    if (E) store_partial(e, part);
}

// ---- rjl : RosatoGuillopeLegrand.f90:23-94 --------------------------------------------------------
// Two passes over the same ELL row (the reference makes three and caches both exponentials per pair,
// :60-67; recomputing them costs less than the 16 B/pair round trip through HBM).
//   pass 1  Eb2_i = sum_j exp(-2q(r/r0-1)) f_c(r);  1/sqrt(Eb2_i) is written into pos[i].w so that pass 2
//           gets it with the same 32-byte gather that brings the partner's position.
//   pass 2  F_i -= [2 A0 (p/r0 f_c - f_c'/r r) e^{-p t} - xi (q/r0 f_c - f_c'/r r/2)(1/Eb_i + 1/Eb_j) e^{-2q t}]/r dr
// Rows are class-partitioned at build time (r < R1 | switch zone | beyond R2, nl.cu k_partition) so the
// lanes of a warp take the same branch; in a crystal the shells line up exactly.
struct RjlC {
    double R1, R22, R12, qa, qb, pa, pb, A0, xi, a1, a2, sw, pi_sw;
    static constexpr bool padded = false;
    // row epilogue of the density pass: 1/Eb from the row sum, and the owner's energy A0 sum_p - xi Eb
    __device__ __forceinline__ double inv_eb(double sq) const { return sq > 0. ? mx::rsqrt_fast(sq) : 0.; }
    __device__ __forceinline__ double energy(double sq, double sp, double ie) const { return A0 * sp - xi * (sq * ie); }
};
// exp arguments are affine in r: -2q(r/r0-1) = qa r + qb, -p(r/r0-1) = pa r + pb
static RjlC rjl_consts(const RJLp& P) {
    RjlC c;
    // exp_nc needs |argument| < 700 for every r in (0, R2): |2q|, |p| and their values at R2 bound it
    if (!(fabs(2. * P.q) < 600. && fabs(P.p) < 600. && fabs(2. * P.q * (P.R2 / P.r0 - 1.)) < 600. && fabs(P.p * (P.R2 / P.r0 - 1.)) < 600.))
        throw std::string("rjl parameters out of the supported range (|p|, |2q| and their products with R2/r0-1 must stay below 600)");
    c.R1 = P.R1; c.R12 = P.R1 * P.R1; c.R22 = P.R2 * P.R2;
    c.qa = -2. * P.q / P.r0; c.qb = 2. * P.q; c.pa = -P.p / P.r0; c.pb = P.p;
    c.A0 = P.A0; c.xi = P.xi; c.a1 = 2. * P.A0 * P.p / P.r0; c.a2 = P.xi * P.q / P.r0;
    c.sw = PFMDS_PI / (P.R2 - P.R1); c.pi_sw = c.sw / 2;
    return c;
}

// Both kernels are software-pipelined by hand, two slots per trip with ping-pong registers: the row
// indices are requested two to three slots ahead and each 32-byte record one slot ahead of its use,
// so a warp keeps both dependent memory levels (index -> gather) in flight under ~100 FP64
// instructions of work, without register-rotation moves.
// Minimum image: if the largest |d_k| high word is below the high word of the smallest half-box no
// component can need wrapping (exact, the comparison is conservative); otherwise the exact FP64 test of
// find_distance runs.  Only atoms within r_cut of a face ever take the second path.
#ifndef RJL_MINB
#define RJL_MINB 7  // measured on B200: 4: 0.997, 5: 1.004, 6: 0.968, 7: 0.948, 8: 0.957 ms/step
#endif
struct WrapC { int min_half_hi; };
static WrapC wrap_consts(const BoxD& b) {
    double hm = b.h[0] < b.h[1] ? (b.h[0] < b.h[2] ? b.h[0] : b.h[2]) : (b.h[1] < b.h[2] ? b.h[1] : b.h[2]);
    long long v;
    memcpy(&v, &hm, 8);
    WrapC w;
    w.min_half_hi = (int)(v >> 32);
    return w;
}
__device__ __forceinline__ void wrap3(double& dx, double& dy, double& dz, const BoxD& box, int min_half_hi) {
    int m = max(max(__double2hiint(dx) & 0x7fffffff, __double2hiint(dy) & 0x7fffffff), __double2hiint(dz) & 0x7fffffff);
    if (m >= min_half_hi) {
        dx = min_image(dx, box.h[0], box.L[0]);
        dy = min_image(dy, box.h[1], box.L[1]);
        dz = min_image(dz, box.h[2], box.L[2]);
    }
}

// ---- rjl, second generation pair routines (default; PFMDS_RJL_GEN=1 selects the first) -------------------------------
// ncu on the first generation (profiles/r1e_*): 111 warp instructions per listed pair, 57 of them FP64, the FP64 pipe 67 %
// busy and 63 % of the issue slots used: both limits bind, so instructions are removed, not overlapped.
//   * the minimum image is decided AFTER r^2: a pair inside R2 cannot have a wrapped component when R2 <= min half box
//     (|d_k| >= h_k gives r^2 >= h_k^2 >= R2^2 in rounded arithmetic too), so only pairs that fail r^2 < R2^2 pay the
//     high-word test, and only atoms near a face pay the exact one; dr and r^2 keep the first generation's bits;
//   * exp_m (13 FP64 instead of 16), rsqrt_q (3 instead of 5), minimax switch polynomials (mathx.cuh);
//   * the prefactors of both exponentials are folded into their arguments, a1 e^x = e^(x + ln a1), so that the common
//     class is  c = (e_p - (1/Eb_i + 1/Eb_j) e_q)/r  (3 FP64) and the switch zone only multiplies e_p and the 1/Eb sum by
//     f + sin(a) k_p and f + sin(a) k_q (k = -pi/(2 (R2-R1)) / slope of the exponent): see rjl_force_consts;
//   * the density pass needs the value of the switch only: one odd polynomial in a - pi/2 (half_switch).
// Per pair of the common class: 43 FP64 instructions in the force pass (57 before), 25 in the density pass (37 before).
struct RjlD {
    double R22, R12, qa, qb, pa, pb, sw, y0, A0, xi;
    static constexpr bool padded = true;
    __device__ __forceinline__ double inv_eb(double sq) const { return sq > 0. ? mx::rsqrt_fast(sq) : 0.; }
    __device__ __forceinline__ double energy(double sq, double sp, double ie) const { return A0 * sp - xi * (sq * ie); }
};
struct RjlF { double R22, R12, qa, qb, pa, pb, swh, u0, kp, kq, l2e, nln2; static constexpr bool padded = true; };
// the second generation needs positive prefactors (their logarithms) and R2 inside the half box
static bool rjl_gen2_ok(const RJLp& P, const BoxD& b) {
    double hm = b.h[0] < b.h[1] ? (b.h[0] < b.h[2] ? b.h[0] : b.h[2]) : (b.h[1] < b.h[2] ? b.h[1] : b.h[2]);
    double a1 = 2. * P.A0 * P.p / P.r0, a2 = P.xi * P.q / P.r0;
    return a1 > 0. && a2 > 0. && a1 < 1e300 && a2 < 1e300 && P.R2 <= hm && P.R2 > P.R1 && fabs(log(a1)) < 50. && fabs(log(a2)) < 50.;
}
static RjlD rjl_dens_consts(const RJLp& P) {
    const RjlC c = rjl_consts(P);
    RjlD d;
    d.R22 = c.R22; d.R12 = c.R12; d.qa = c.qa; d.qb = c.qb; d.pa = c.pa; d.pb = c.pb; d.A0 = c.A0; d.xi = c.xi;
    d.sw = c.sw;                                      // a = (r - R1) sw with the reference's pi literal; y = a - pi/2
    d.y0 = -P.R1 * c.sw - 1.57079632679489661923;
    return d;
}
static RjlF rjl_force_consts(const RJLp& P) {
    const RjlC c = rjl_consts(P);
    RjlF f;
    f.R22 = c.R22; f.R12 = c.R12; f.qa = c.qa; f.pa = c.pa;
    f.pb = c.pb + log(c.a1);                          // a1 e^{-p(r/r0-1)},  a1 = 2 A0 p / r0
    f.qb = c.qb + log(c.a2);                          // a2 e^{-2q(r/r0-1)}, a2 = xi q / r0
    f.swh = 0.5 * c.sw;                               // u = a/2 - pi/4
    f.u0 = -P.R1 * f.swh - 0.78539816339744830962;
    // switch zone: 2 A0 (-pa f + s pi_sw) e_p = a1 e_p (f + s kp), kp = -pi_sw / pa;  likewise kq = -pi_sw / qa
    f.kp = -c.pi_sw / c.pa;
    f.kq = -c.pi_sw / c.qa;
    f.l2e = 1.4426950408889634; f.nln2 = -6.93147180559945309417e-01;  // launch parameters stay in uniform registers (mx::exp_m2)
    return f;
}
// dr and r^2 of a pair; false when the pair is beyond R2 (or not a number).  Same bits as wrap3 + the first generation's sum.
__device__ __forceinline__ bool pair_inside(const double4& pi, const double4& pj, double R22, const BoxD& box, int mhh, double& dx, double& dy,
                                            double& dz, double& r2) {
    dx = pj.x - pi.x; dy = pj.y - pi.y; dz = pj.z - pi.z;
    r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    if (!(r2 < R22)) {
        int m = max(max(__double2hiint(dx) & 0x7fffffff, __double2hiint(dy) & 0x7fffffff), __double2hiint(dz) & 0x7fffffff);
        if (m < mhh) return false;
        dx = min_image(dx, box.h[0], box.L[0]);
        dy = min_image(dy, box.h[1], box.L[1]);
        dz = min_image(dz, box.h[2], box.L[2]);
        r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        if (!(r2 < R22)) return false;
    }
    return true;
}
template <bool E>
__device__ __forceinline__ void rjl_density_pair(const double4& pi, const double4& pj, const RjlD& C, const BoxD& box, int mhh, double& sq, double& sp) {
    double dx, dy, dz, r2;
    if (pair_inside(pi, pj, C.R22, box, mhh, dx, dy, dz, r2)) {
        double r = r2 * mx::rsqrt_q(r2);
        double eq = mx::exp_m(fma(C.qa, r, C.qb));
        double ep = E ? mx::exp_m(fma(C.pa, r, C.pb)) : 0.;
        if (r2 < C.R12) {
            sq += eq;
            if (E) sp += ep;
        } else {
            double f = mx::half_switch(fma(r, C.sw, C.y0));
            sq = fma(eq, f, sq);
            if (E) sp = fma(ep, f, sp);
        }
    }
}
// E: also accumulate a1 e_p f into `se` (the repulsive energy of the owner up to the factor A0 / a1 = r0 / 2p): the force pass has
// that exponential in hand, so a step that must report its energies pays one FP64 instruction per pair here instead of a second
// exponential per pair in the density pass.
template <bool E>
__device__ __forceinline__ void rjl_force_pair_g2(const double4& pi, const double4& pj, const RjlF& C, const BoxD& box, int mhh, double& fx, double& fy,
                                                  double& fz, double& se) {
    double dx, dy, dz, r2;
    if (pair_inside(pi, pj, C.R22, box, mhh, dx, dy, dz, r2)) {
        double ir = mx::rsqrt_q(r2);
        double r = r2 * ir;
        // switch zone first, then both exponentials side by side (mx::exp_m2), then the zone's two factors
        const bool zone = !(r2 < C.R12);
#ifdef __CUDA_ARCH__
        double gp, gq, fsw;  // read only when `zone`
#else
        double gp = 1., gq = 1., fsw = 1.;
#endif
        if (zone) {
            double f, s;
            mx::cos_switch_m(fma(r, C.swh, C.u0), f, s);
            gp = fma(s, C.kp, f);
            gq = fma(s, C.kq, f);
            fsw = f;
        }
        double A, eq;
        mx::exp_m2(fma(C.pa, r, C.pb), fma(C.qa, r, C.qb), A, eq, C.l2e, C.nln2);
        double w = pi.w + pj.w;
        if (zone) {
#ifdef __CUDA_ARCH__
            asm volatile("");  // keeps this a branch: if-converted it costs every pair two multiplications and four selects
#endif
            if (E) se = fma(A, fsw, se);
            A *= gp; w *= gq;
        } else if (E) se += A;
        double c = fma(-w, eq, A) * ir;
        fx = fma(-c, dx, fx); fy = fma(-c, dy, fy); fz = fma(-c, dz, fz);
    }
}
__device__ __forceinline__ void rjl_force_pair(const double4& pi, const double4& pj, const RjlF& C, const BoxD& box, int mhh, double& fx, double& fy,
                                               double& fz) {
    double unused = 0.;
    rjl_force_pair_g2<false>(pi, pj, C, box, mhh, fx, fy, fz, unused);
}

// ---- rjl, third generation pair routines (default; PFMDS_RJL_GEN=2 / 1 select the older ones) ---------------------------------
// ncu on the second generation (profiles/r2a_*): still bound by FP64 issue, 24 of the 43 FP64 instructions of a common-class pair
// are the two exponentials (two range reductions by ln 2 and two degree-8 polynomials).  Both exponents are affine in the SAME r,
// so one range reduction serves both: r = r_J + d with r_J = J g the nearest node of a grid of spacing g = 2^-m (J from the
// 1.5 x 2^52 rounding trick, d exact), and
//     a1 e^{pa r + pb} = [a1 e^{pa r_J + pb}] e^{pa d},   a2 e^{qa r + qb} = [a2 e^{qa r_J + qb}] e^{qa d}
// with the bracketed node values read from a table in global memory (16 bytes per node, a few tens of KB: L1 resident) and
// e^{x} for |x| <= max(|pa|,|qa|) g/2 <= 0.0045 from its degree-4 Taylor polynomial (truncation x^5/120 <= 1.6e-14 relative;
// the table entries are rounded from long double).  11 FP64 instructions replace 24: 30 per common-class pair in the force
// pass, 19 in the density pass.  Pairs closer than the table's first node (r < r_lo ~ 0.7 r0: none in a solid or liquid at any
// temperature MD is run at) take the second generation's analytic routine, so the result is defined for every input.
struct RjlG {
    double R22, R12;
    double inv_g, ng;             // 1/g and -g
    double cp1, cp2, cp3, cp4;    // pa^n / n!
    double cq1, cq2, cq3, cq4;    // qa^n / n!
    double swh, u0, kp, kq;       // force pass, switch zone (as RjlF)
    double sw, y0;                // density pass, value-only switch (as RjlD)
    double inv_a2, erep, xi;      // density epilogue: the table carries a2 e_q and a1 e_p; sum e_q = sq inv_a2, A0 sum e_p = erep sp
    const double2* tab;           // biased by -J_lo entries: tab[J] = {a1 e^{pa r_J + pb}, a2 e^{qa r_J + qb}}
    int J_lo;                     // first node of the table
    double pa, pb, qa, qb, l2e, nln2;  // exponents with the prefactors folded in (as RjlF): analytic exponentials for r below the table
    static constexpr bool padded = true;
    __device__ __forceinline__ double inv_eb(double sq) const { sq *= inv_a2; return sq > 0. ? mx::rsqrt_fast(sq) : 0.; }
    __device__ __forceinline__ double energy(double sq, double sp, double ie) const { return erep * sp - xi * ((sq * inv_a2) * ie); }
};
#define RJL_TAB_MAGIC 6755399441055744.0  // 1.5 x 2^52: fma(r, 1/g, MAGIC) rounds r/g to an integer held in the low word
__device__ __forceinline__ double2 rjl_tab_load(const double2* p) {
#ifdef __CUDA_ARCH__
    double2 v;
    asm("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return *p;
#endif
}
template <bool E>
__device__ __forceinline__ void rjl_force_pair_g3(const double4& pi, const double4& pj, const RjlG& C, const BoxD& box, int mhh, double& fx, double& fy,
                                                  double& fz, double& se) {
    double dx, dy, dz, r2;
    if (pair_inside(pi, pj, C.R22, box, mhh, dx, dy, dz, r2)) {
        double ir = mx::rsqrt_q(r2);
        double r = r2 * ir;
        const double u = fma(r, C.inv_g, RJL_TAB_MAGIC);
        const int J = mx::lo_word(u);
        double A, eq;
        if (J >= C.J_lo) {
            const double2 T = rjl_tab_load(C.tab + J);
            const double d = fma(u - RJL_TAB_MAGIC, C.ng, r);
            double Pp = fma(fma(fma(fma(C.cp4, d, C.cp3), d, C.cp2), d, C.cp1), d, 1.0);
            double Pq = fma(fma(fma(fma(C.cq4, d, C.cq3), d, C.cq2), d, C.cq1), d, 1.0);
            A = T.x * Pp; eq = T.y * Pq;
        } else {  // closer than the table's first node (never in a condensed phase): both exponentials analytically
            mx::exp_m2(fma(C.pa, r, C.pb), fma(C.qa, r, C.qb), A, eq, C.l2e, C.nln2);
        }
        double w = pi.w + pj.w;
        if (!(r2 < C.R12)) {
#ifdef __CUDA_ARCH__
            asm volatile("");  // keeps this a branch (rows are class-partitioned: warp-uniform in a crystal)
#endif
            double f, sn;
            mx::cos_switch_m(fma(r, C.swh, C.u0), f, sn);
            if (E) se = fma(A, f, se);
            A *= fma(sn, C.kp, f);
            w *= fma(sn, C.kq, f);
        } else if (E) se += A;
        double c = fma(-w, eq, A) * ir;
        fx = fma(-c, dx, fx); fy = fma(-c, dy, fy); fz = fma(-c, dz, fz);
    }
}
__device__ __forceinline__ void rjl_force_pair(const double4& pi, const double4& pj, const RjlG& C, const BoxD& box, int mhh, double& fx, double& fy,
                                               double& fz) {
    double unused = 0.;
    rjl_force_pair_g3<false>(pi, pj, C, box, mhh, fx, fy, fz, unused);
}
template <bool E>
__device__ __forceinline__ void rjl_force_pair_e(const double4& pi, const double4& pj, const RjlF& C, const BoxD& box, int mhh, double& fx, double& fy,
                                                 double& fz, double& se) {
    rjl_force_pair_g2<E>(pi, pj, C, box, mhh, fx, fy, fz, se);
}
template <bool E>
__device__ __forceinline__ void rjl_force_pair_e(const double4& pi, const double4& pj, const RjlG& C, const BoxD& box, int mhh, double& fx, double& fy,
                                                 double& fz, double& se) {
    rjl_force_pair_g3<E>(pi, pj, C, box, mhh, fx, fy, fz, se);
}
template <bool E>
__device__ __forceinline__ void rjl_density_pair(const double4& pi, const double4& pj, const RjlG& C, const BoxD& box, int mhh, double& sq, double& sp) {
    double dx, dy, dz, r2;
    if (pair_inside(pi, pj, C.R22, box, mhh, dx, dy, dz, r2)) {
        double r = r2 * mx::rsqrt_q(r2);
        const double u = fma(r, C.inv_g, RJL_TAB_MAGIC);
        const int J = mx::lo_word(u);
        double Pq, Pp = 0., tp = 0., tq;   // a2 e_q = tq Pq, a1 e_p = tp Pp
        if (J >= C.J_lo) {
            const double d = fma(u - RJL_TAB_MAGIC, C.ng, r);
            Pq = fma(fma(fma(fma(C.cq4, d, C.cq3), d, C.cq2), d, C.cq1), d, 1.0);
            if (E) {
                const double2 T = rjl_tab_load(C.tab + J);
                tp = T.x; tq = T.y;
                Pp = fma(fma(fma(fma(C.cp4, d, C.cp3), d, C.cp2), d, C.cp1), d, 1.0);
            } else {
#ifdef __CUDA_ARCH__
                asm("ld.global.nc.f64 %0, [%1];" : "=d"(tq) : "l"(reinterpret_cast<const double*>(C.tab + J) + 1));
#else
                tq = C.tab[J].y;
#endif
            }
        } else {  // closer than the table's first node: analytic exponentials, in the table's units
            mx::exp_m2(fma(C.pa, r, C.pb), fma(C.qa, r, C.qb), tp, tq, C.l2e, C.nln2);
            Pq = 1.; Pp = 1.;
        }
        if (r2 < C.R12) {
            sq = fma(tq, Pq, sq);
            if (E) sp = fma(tp, Pp, sp);
        } else {
            double f = mx::half_switch(fma(r, C.sw, C.y0));
            sq = fma(tq * Pq, f, sq);
            if (E) sp = fma(tp * Pp, f, sp);
        }
    }
}

// Node table of the third generation: spacing, range and entries (host side; pure arithmetic, shared with tests/forces_host.cpp).
struct RjlTabSpec { int m, J_lo, J_hi; double g; };
static bool rjl_gen3_ok(const RJLp& P, const BoxD& b) { return rjl_gen2_ok(P, b) && P.r0 > 0. && P.R2 > 0.; }
static RjlTabSpec rjl_tab_spec(const RJLp& P) {
    const RjlC c = rjl_consts(P);
    const double X = fmax(fabs(c.pa), fabs(c.qa));
    RjlTabSpec t;
    t.m = 0;
    while (X * ldexp(1., -t.m) * 0.5 > 0.0045 && t.m < 24) ++t.m;   // degree-4 Taylor of e^x on |x| <= X g / 2: x^5/120 <= 1.6e-14
    t.g = ldexp(1., -t.m);
    double r_lo = 0.7 * P.r0;
    if (r_lo > 0.5 * P.R2) r_lo = 0.5 * P.R2;
    if (P.R2 - r_lo > 8192. * t.g) r_lo = P.R2 - 8192. * t.g;      // at most 128 KB of table
    t.J_lo = (int)floor(r_lo / t.g);
    if (t.J_lo < 1) t.J_lo = 1;
    t.J_hi = (int)ceil(P.R2 / t.g) + 2;
    return t;
}
static void rjl_tab_fill(const RJLp& P, const RjlTabSpec& t, double2* out) {  // out[J - J_lo], J_lo <= J <= J_hi
    const long double a1 = 2.0L * P.A0 * P.p / P.r0, a2 = (long double)P.xi * P.q / P.r0;
    for (int J = t.J_lo; J <= t.J_hi; ++J) {
        const long double x = (long double)J * t.g / P.r0 - 1.0L;
        out[J - t.J_lo].x = (double)(a1 * expl(-(long double)P.p * x));
        out[J - t.J_lo].y = (double)(a2 * expl(-2.0L * P.q * x));
    }
}
static RjlG rjl_g3_consts(const RJLp& P, const RjlTabSpec& t, const double2* tab) {  // tab = the filled table (device or host memory)
    const RjlC c = rjl_consts(P);
    RjlG G;
    const RjlF f2 = rjl_force_consts(P);
    const RjlD d2 = rjl_dens_consts(P);
    G.pa = f2.pa; G.pb = f2.pb; G.qa = f2.qa; G.qb = f2.qb; G.l2e = f2.l2e; G.nln2 = f2.nln2;
    G.R22 = c.R22; G.R12 = c.R12;
    G.inv_g = ldexp(1., t.m); G.ng = -t.g;
    G.cp1 = c.pa; G.cp2 = c.pa * c.pa / 2.; G.cp3 = c.pa * c.pa * c.pa / 6.; G.cp4 = c.pa * c.pa * c.pa * c.pa / 24.;
    G.cq1 = c.qa; G.cq2 = c.qa * c.qa / 2.; G.cq3 = c.qa * c.qa * c.qa / 6.; G.cq4 = c.qa * c.qa * c.qa * c.qa / 24.;
    G.swh = f2.swh; G.u0 = f2.u0; G.kp = f2.kp; G.kq = f2.kq;
    G.sw = d2.sw; G.y0 = d2.y0;
    G.inv_a2 = 1. / c.a2; G.erep = c.A0 / c.a1; G.xi = c.xi;
    G.tab = tab - t.J_lo;
    G.J_lo = t.J_lo;
    return G;
}

template <bool E>
__device__ __forceinline__ void rjl_density_pair(const double4& pi, const double4& pj, const RjlC& C, const BoxD& box, int mhh, double& sq, double& sp) {
    double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
    wrap3(dx, dy, dz, box, mhh);
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    if (r2 < C.R22) {
        double r = r2 * mx::rsqrt_fast(r2);
        double eq = mx::exp_nc(fma(C.qa, r, C.qb));
        double f = 1.0;
        if (r2 >= C.R12) {
            double s;
            mx::cos_switch((r - C.R1) * C.sw, f, s);
        }
        sq = fma(eq, f, sq);
        if (E) sp = fma(mx::exp_nc(fma(C.pa, r, C.pb)), f, sp);
    }
}
// Row walk shared by both pipelined kernels.  PADDED rows (second generation; every list is allocated with two spare rows of
// valid slot numbers, capi.cu) are prefetched without bounds tests: the walk reads at most slot n+1 and never uses it.
template <bool PADDED>
struct RowWalk {
    const int* rp;  // PADDED: slot p+2;  otherwise slot p
    ptrdiff_t st;
    __device__ __forceinline__ RowWalk(const ListView& lv, int i) : rp(lv.nlist + i), st((ptrdiff_t)lv.stride) {}
    __device__ __forceinline__ int first(int n, int& j1) {
        int j0 = rp[0];
        j1 = (PADDED || n > 1) ? rp[st] : j0;
        if (PADDED) rp += 2 * st;
        return j0;
    }
    // slots p+2 and p+3 while slots p, p+1 are being worked on
    __device__ __forceinline__ void ahead(int p, int n, int j1, int& j2, int& j3) {
        if (PADDED) {
            j2 = rp[0]; j3 = rp[st];
        } else {
            j2 = p + 2 < n ? rp[2 * st] : j1;
            j3 = p + 3 < n ? rp[3 * st] : j1;
        }
        rp += 2 * st;
    }
};

#ifndef RJL_MINB_D
#define RJL_MINB_D 9  // density pass: 54 registers without a spill, 36 resident warps: 0.2769 -> 0.2722 ms per launch (tools/rjl_variants.py, B200; 6, 7, 8: 0.277)
#endif
template <bool E, class CT>  // CT = RjlC (first generation) or RjlD (second generation, below): picks the pair routine
__global__ void __launch_bounds__(FT, RJL_MINB_D) k_rjl_density(int N, double4* pos, ListView lv, CT C, BoxD box, WrapC W, double* part, SlabDev S) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0;
    bool pushed = false;
    slab_wait(S);  // slab mode: the neighbours' positions of this step have landed in my ghost slots
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        // The .w of every record is being written by this kernel (1/Eb of its owner) while x,y,z are
        // constant: the whole record is fetched with one request and .w is ignored here.
        const double4 pi = ld256(&pos[i]);
        double sq = 0, sp = 0;
        RowWalk<CT::padded> row(lv, i);
        int j1;
        double4 a = ld256(&pos[row.first(n, j1)]);
        int p = 0;
        for (; p + 1 < n; p += 2) {
            double4 b = ld256(&pos[j1]);
            int j2, j3;
            row.ahead(p, n, j1, j2, j3);
            rjl_density_pair<E>(pi, a, C, box, W.min_half_hi, sq, sp);
            a = ld256(&pos[j2]);
            rjl_density_pair<E>(pi, b, C, box, W.min_half_hi, sq, sp);
            j1 = j3;
        }
        if (p < n) rjl_density_pair<E>(pi, a, C, box, W.min_half_hi, sq, sp);
        double ie = C.inv_eb(sq);
        reinterpret_cast<double*>(&pos[i])[3] = ie;
        if (S.push) {  // the same 1/Eb goes straight into the ghost copies of this atom on the neighbour GPUs (NVLink stores)
            int a = S.rs_l[i], b = S.rs_r[i];
            if (a >= 0) reinterpret_cast<double*>(&S.peer_l[a])[3] = ie;
            if (b >= 0) reinterpret_cast<double*>(&S.peer_r[b])[3] = ie;
            pushed = (a >= 0) || (b >= 0);
        }
        if (E) e = C.energy(sq, sp, ie);
    }
    if (E) store_partial(e, part);
    slab_signal(S, pushed);
}

// small systems: SPLIT lanes per atom, plain loop (latency is hidden by the extra warps)
template <bool E, int SPLIT, class CT>
__global__ void __launch_bounds__(FT) k_rjl_density_split(int N, double4* pos, ListView lv, CT C, BoxD box, WrapC W, double* part) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t / SPLIT, sub = t % SPLIT;
    double e = 0, sq = 0, sp = 0;
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = ld256(&pos[i]);
        for (int p = sub; p < n; p += SPLIT) rjl_density_pair<E>(pi, ld256(&pos[lv.nlist[(size_t)p * lv.stride + i]]), C, box, W.min_half_hi, sq, sp);
    }
    sq = split_sum<SPLIT>(sq);
    if (E) sp = split_sum<SPLIT>(sp);
    if (n > 0 && sub == 0) {
        double ie = C.inv_eb(sq);
        reinterpret_cast<double*>(&pos[i])[3] = ie;
        if (E) e = C.energy(sq, sp, ie);
    }
    if (E) store_partial(e, part);
}

__device__ __forceinline__ void rjl_force_pair(const double4& pi, const double4& pj, const RjlC& C, const BoxD& box, int mhh, double& fx, double& fy,
                                               double& fz) {
    double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
    wrap3(dx, dy, dz, box, mhh);
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    if (r2 < C.R22) {
        double ir = mx::rsqrt_fast(r2);
        double r = r2 * ir;
        double ep = mx::exp_nc(fma(C.pa, r, C.pb)), eq = mx::exp_nc(fma(C.qa, r, C.qb));
        double ies = pi.w + pj.w;
        double c;
        if (r2 < C.R12) {
            c = (C.a1 * ep - C.a2 * ies * eq) * ir;
        } else {
            double f, s;
            mx::cos_switch((r - C.R1) * C.sw, f, s);
            double dfr_r = -s * C.pi_sw;  // f_c'(r) = df_cut * r
            c = (2. * C.A0 * (-C.pa * f - dfr_r) * ep - C.xi * (-0.5 * C.qa * f - 0.5 * dfr_r) * ies * eq) * ir;
        }
        fx = fma(-c, dx, fx); fy = fma(-c, dy, fy); fz = fma(-c, dz, fz);
    }
}
// RJL_MINB = blocks per SM the register allocation is held to.  7 (72 registers) is the measured optimum; 5 (94 registers, no
// constant reloads in the loop, 20 instead of 28 resident warps) measured slower on a B200 (0.425 against 0.358 ms, BENCH_r01).
template <class CT>  // CT = RjlC (first generation), RjlF (second) or RjlG (third)
__global__ void __launch_bounds__(FT, RJL_MINB) k_rjl_force(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, CT C, BoxD box,
                                                      WrapC W, SlabDev S, int overwrite) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    slab_wait(S);  // slab mode: the neighbours' 1/Eb have landed in my ghost slots
    if (i >= N) return;
    int n = lv.nnum[i];
    if (n == 0) { if (overwrite) frc[i] = make_double4(0., 0., 0., 0.); return; }
    const double4 pi = ld256_nc(&pos[i]);
    double fx = 0, fy = 0, fz = 0;
    RowWalk<CT::padded> row(lv, i);
    int j1;
    double4 a = ld256_nc(&pos[row.first(n, j1)]);
    int p = 0;
    for (; p + 1 < n; p += 2) {
        double4 b = ld256_nc(&pos[j1]);
        int j2, j3;
        row.ahead(p, n, j1, j2, j3);
        rjl_force_pair(pi, a, C, box, W.min_half_hi, fx, fy, fz);
        a = ld256_nc(&pos[j2]);
        rjl_force_pair(pi, b, C, box, W.min_half_hi, fx, fy, fz);
        j1 = j3;
    }
    if (p < n) rjl_force_pair(pi, a, C, box, W.min_half_hi, fx, fy, fz);
    // first interaction of the step and every atom is an owner: store instead of zero + accumulate (zero_forces fused away)
    if (overwrite) frc[i] = make_double4(fx, fy, fz, 0.);
    else add_force(frc, i, fx, fy, fz);
}

template <int SPLIT, class CT>
__global__ void __launch_bounds__(FT) k_rjl_force_split(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, CT C, BoxD box,
                                                        WrapC W) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t / SPLIT, sub = t % SPLIT;
    double fx = 0, fy = 0, fz = 0;
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = ld256_nc(&pos[i]);
        for (int p = sub; p < n; p += SPLIT) rjl_force_pair(pi, ld256_nc(&pos[lv.nlist[(size_t)p * lv.stride + i]]), C, box, W.min_half_hi, fx, fy, fz);
    }
    fx = split_sum<SPLIT>(fx); fy = split_sum<SPLIT>(fy); fz = split_sum<SPLIT>(fz);
    if (n > 0 && sub == 0) add_force(frc, i, fx, fy, fz);
}

// Force pass that also yields the interaction's energy (second generation only; steps that report their energies):
// e_i = (r0 / 2p) sum_j a1 e_p f  -  xi sqrt(sum_j e_q f),  the square root being 1 / (1/Eb_i) from the density pass.
// Same row walk and pair routine as k_rjl_force; no early exits, every thread reaches the block sum.
// Force pass that also yields the interaction's energy (steps that report their energies):
// e_i = (r0 / 2p) sum_j a1 e_p f  -  xi sqrt(sum_j e_q f),  the square root being 1 / (1/Eb_i) from the density pass.
// Same row walk and pair routine as k_rjl_force; no early exits, every thread reaches the block sum.
// (Tried and dropped, measured on a B200 at 10^6 atoms: the closing half kick + thermostat KE partials in this kernel's epilogue,
//  with and without the last block of the grid running the chain update: 0.375 / 0.398 ms against 0.359 ms + 0.034 ms for the
//  separate k_kick_ke -- the extra 64 B/atom of velocity traffic lands in a kernel whose L1 data pipe is already the limit.)
template <class CT>  // CT = RjlF (second generation) or RjlG (third)
__global__ void __launch_bounds__(FT, RJL_MINB) k_rjl_force_e(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, CT C, BoxD box,
                                                        WrapC W, SlabDev S, int overwrite, double erep, double xi, double* __restrict__ part) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    slab_wait(S);
    double e = 0;
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = ld256_nc(&pos[i]);
        double fx = 0, fy = 0, fz = 0, se = 0;
        RowWalk<true> row(lv, i);
        int j1;
        double4 a = ld256_nc(&pos[row.first(n, j1)]);
        int p = 0;
        for (; p + 1 < n; p += 2) {
            double4 b = ld256_nc(&pos[j1]);
            int j2, j3;
            row.ahead(p, n, j1, j2, j3);
            rjl_force_pair_e<true>(pi, a, C, box, W.min_half_hi, fx, fy, fz, se);
            a = ld256_nc(&pos[j2]);
            rjl_force_pair_e<true>(pi, b, C, box, W.min_half_hi, fx, fy, fz, se);
            j1 = j3;
        }
        if (p < n) rjl_force_pair_e<true>(pi, a, C, box, W.min_half_hi, fx, fy, fz, se);
        if (overwrite) frc[i] = make_double4(fx, fy, fz, 0.);
        else add_force(frc, i, fx, fy, fz);
        e = erep * se - (pi.w > 0. ? xi / pi.w : 0.);
    } else if (i < N && overwrite) frc[i] = make_double4(0., 0., 0., 0.);
    store_partial(e, part);
}
template <int SPLIT, class CT>
__global__ void __launch_bounds__(FT) k_rjl_force_split_e(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, CT C, BoxD box,
                                                          WrapC W, double erep, double xi, double* __restrict__ part) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t / SPLIT, sub = t % SPLIT;
    double fx = 0, fy = 0, fz = 0, se = 0, e = 0, ie = 0;
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = ld256_nc(&pos[i]);
        ie = pi.w;
        for (int p = sub; p < n; p += SPLIT)
            rjl_force_pair_e<true>(pi, ld256_nc(&pos[lv.nlist[(size_t)p * lv.stride + i]]), C, box, W.min_half_hi, fx, fy, fz, se);
    }
    fx = split_sum<SPLIT>(fx); fy = split_sum<SPLIT>(fy); fz = split_sum<SPLIT>(fz); se = split_sum<SPLIT>(se);
    if (n > 0 && sub == 0) {
        add_force(frc, i, fx, fy, fz);
        e = erep * se - (ie > 0. ? xi / ie : 0.);
    }
    store_partial(e, part);
}

// ---- lj1g, pipelined variant: the default from small_n atoms up (PFMDS_LJ1G_PIPE=0 selects k_lj1g); 0.174 -> 0.102 ms per launch on the 96^3 LJ fluid (BENCH_r01) ----
// Same sums as k_lj1g<F, E, 1>, restructured like the rjl kernels: ping-pong prefetch of the row indices and of the 32-byte
// records, the conservative exact wrap test instead of three unconditional selects, 1/r^2 and sqrt from the hardware seeds plus
// one third-order step (mathx.cuh) instead of the CUDA library's division and square root with their slow-path calls.
template <bool E>
__device__ __forceinline__ void lj1g_pair(const double4& pi, const double4& pj, const LJ1Gp& P, double R12, double R22, double iw, const BoxD& box,
                                          int mhh, double& fx, double& fy, double& fz, double& e) {
    double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
    wrap3(dx, dy, dz, box, mhh);
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    if (r2 < R22) {  // beyond R2 the quintic switch and its derivative are exactly 0
        double invr2 = mx::rcp_fast(r2);
        double U = invr2 * invr2 * invr2;
        double f = 1.0, dfr = 0.0;
        if (r2 > R12) {
            double r = r2 * mx::rsqrt_fast(r2);
            double x = (r - P.R1) * iw, x2 = x * x;
            f = 1. + x2 * x * (-10. + 15. * x - 6. * x2);
            dfr = r * x2 * (-30. + 60. * x - 30. * x2);
        }
        if (E) e += U * (P.c12 * U - P.c6) * f;
        double c = U * invr2 * ((P.c12t12 * U - P.c6t6) * f - (P.c12 * U - P.c6) * dfr);
        fx = fma(-c, dx, fx); fy = fma(-c, dy, fy); fz = fma(-c, dz, fz);
    }
}
template <bool E>
__global__ void __launch_bounds__(FT) k_lj1g_pipe(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, LJ1Gp P, BoxD box,
                                                  WrapC W, double* part, int overwrite) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0, fx = 0, fy = 0, fz = 0;
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = ld256_nc(&pos[i]);
        const double R12 = P.R1 * P.R1, R22 = P.R2 * P.R2, iw = 1.0 / (P.R2 - P.R1);
        const int* rp = lv.nlist + i;
        const size_t st = lv.stride;
        int j1 = n > 1 ? rp[st] : rp[0];
        double4 a = ld256_nc(&pos[rp[0]]);
        int p = 0;
        for (; p + 1 < n; p += 2) {
            double4 b = ld256_nc(&pos[j1]);
            int j2 = p + 2 < n ? rp[2 * st] : j1;
            int j3 = p + 3 < n ? rp[3 * st] : j1;
            rp += 2 * st;
            lj1g_pair<E>(pi, a, P, R12, R22, iw, box, W.min_half_hi, fx, fy, fz, e);
            a = ld256_nc(&pos[j2]);
            lj1g_pair<E>(pi, b, P, R12, R22, iw, box, W.min_half_hi, fx, fy, fz, e);
            j1 = j3;
        }
        if (p < n) lj1g_pair<E>(pi, a, P, R12, R22, iw, box, W.min_half_hi, fx, fy, fz, e);
        // first interaction of the step and every atom is an owner: store instead of zero + accumulate (zero_forces fused away, as in k_rjl_force)
        if (overwrite) frc[i] = make_double4(fx, fy, fz, 0.);
        else add_force(frc, i, fx, fy, fz);
    } else if (overwrite && i < N) frc[i] = make_double4(0., 0., 0., 0.);
    if (E) store_partial(e, part);
}

// ---- tb : TersoffBrenner.f90:26-150 -------------------------------------------------------------
__device__ __forceinline__ double tb_G(double c1, const TBp& T) { return 1. + T.c02 / T.d02 - T.c02 / (T.d02 + c1 * c1); }  // c1 = 1+cos

// One thread per (atom, slot) pair — blockIdx.y is the slot — so a graphene sheet of a few thousand atoms
// still fills the machine; the per-slot force contributions go to a scratch ELL block and are summed per atom
// in slot order by k_tb_reduce (deterministic, no atomics).
// pass A: bond orders B(p,i) = (1 + a0 sum_{q!=p} f_c(r_q) G(theta_pq))^-delt, 0 for r_p >= R2  (:83-96).
// Both B and B^(1/delt+1) = (1 + a0 zeta)^-(delt+1) are stored, from one logarithm: pass B then needs no pow().
__global__ void __launch_bounds__(FT) k_tb_bond(int N, const double4* __restrict__ pos, ListView lv, TBp T, BoxD box, double* __restrict__ B,
                                                double* __restrict__ Bx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int p = blockIdx.y;
    if (i >= N) return;
    int n = lv.nnum[i];
    if (p >= n) return;
    const double4 pi = pos[i];
    double rp2;
    Vec dp = bond_vec(pi, pos[lv.nlist[(size_t)p * lv.stride + i]], box, rp2);
    double rp = sqrt(rp2), b = 0., bx = 0.;
    if (rp < T.R2) {
        double z = 0.;
        for (int q = 0; q < n; ++q) {
            if (q == p) continue;
            double rq2;
            Vec dq = bond_vec(pi, pos[lv.nlist[(size_t)q * lv.stride + i]], box, rq2);
            double rq = sqrt(rq2);
            if (rq < T.R2) z += fcut_only(rq, T.R1, T.R2) * tb_G(1. + dot(dp, dq) / (rp * rq), T);
        }
        double L = mx::log_fast(1. + T.a0 * z);
        b = mx::exp_fast(-T.delt * L);
        bx = mx::exp_fast(-(T.delt + 1.) * L);
    }
    B[(size_t)p * lv.stride + i] = b;
    Bx[(size_t)p * lv.stride + i] = bx;
}
// pass B: forces (:99-146) and/or energy (:53-66) of one bond.  Bonds and partners beyond R2 contribute exactly
// zero in the reference (f_cut = df_cut = 0, B = 0) and are skipped.
// forces (:99-146) and / or energy (:53-66) of bond p of atom i (p < n = its row length): everything atom i feels through that bond --
// the pair term with both bond orders, the derivative of its own bond order and the cross terms through j's other bonds.
// Bonds and partners beyond R2 contribute exactly zero in the reference (f_cut = df_cut = 0, B = 0) and are skipped.
template <bool F, bool E>
__device__ __forceinline__ void tb_bond_terms(int i, int p, int n, const double4* __restrict__ pos, const ListView& lv, const TBp& T, const BoxD& box,
                                              const double* __restrict__ B, const double* __restrict__ Bx, double& e, double& fx, double& fy, double& fz) {
    const double4 pi = pos[i];
    const double s2s = sqrt(2. * T.s), s2is = sqrt(2. / T.s), dpre = T.d / (T.s - 1.);
    do {
        int j = lv.nlist[(size_t)p * lv.stride + i];
        const double4 pj = pos[j];
        double rp2;
        Vec dp = bond_vec(pi, pj, box, rp2);
        double rp = sqrt(rp2);
        if (!(rp < T.R2)) break;
        double Bip = B[(size_t)p * lv.stride + i];
        // reverse slot: i in j's row
        int nj = lv.nnum[j], l = 0;
        for (; l < nj; ++l)
            if (lv.nlist[(size_t)l * lv.stride + j] == i) break;
        if (l >= nj) break;  // cannot happen: same group, same r_cut, symmetric dr2
        double Bjl = B[(size_t)l * lv.stride + j];
        double f_c, dfr_p;
        fcut_dfcut(rp, T.R1, T.R2, f_c, dfr_p);
        double a = -s2s * T.b * (rp - T.r0);
        double ea = mx::exp_fast(a), eas = mx::exp_fast(a / T.s);
        if (E) {
            if (j > i) e += f_c * dpre * (ea - (Bip + Bjl) / 2 * T.s * eas);
        }
        if (F) {
            Vec dB = {0., 0., 0.};
            for (int q = 0; q < n; ++q) {  // :112-120 own row
                if (q == p) continue;
                double rq2;
                Vec dq = bond_vec(pi, pos[lv.nlist[(size_t)q * lv.stride + i]], box, rq2);
                double rq = sqrt(rq2);
                if (!(rq < T.R2)) continue;
                double fq, dfq;
                fcut_dfcut(rq, T.R1, T.R2, fq, dfq);
                double rr = 1. / rp / rq, cosi = dot(dp, dq) * rr, c1 = 1. + cosi, den = T.d02 + c1 * c1;
                double g1 = fq * 2. * T.a0 * T.c02 * c1 / (den * den);
                double g2 = dfq * T.a0 * tb_G(c1, T);
                dB.x += g1 * ((dp.x + dq.x) * rr - cosi * (dp.x / rp2 + dq.x / rq2)) + g2 * dq.x;
                dB.y += g1 * ((dp.y + dq.y) * rr - cosi * (dp.y / rp2 + dq.y / rq2)) + g2 * dq.y;
                dB.z += g1 * ((dp.z + dq.z) * rr - cosi * (dp.z / rp2 + dq.z / rq2)) + g2 * dq.z;
            }
            double bp = Bx[(size_t)p * lv.stride + i];
            dB.x *= bp; dB.y *= bp; dB.z *= bp;
            const Vec dl = {-dp.x, -dp.y, -dp.z};  // j -> i
            double bl = Bx[(size_t)l * lv.stride + j];
            double cx = 0, cy = 0, cz = 0;  // cross terms :142-152
            for (int q = 0; q < nj; ++q) {  // :126-133 and :142-152 share the geometry of j's row
                if (q == l) continue;
                double rq2;
                Vec dq = bond_vec(pj, pos[lv.nlist[(size_t)q * lv.stride + j]], box, rq2);
                double rq = sqrt(rq2);
                if (!(rq < T.R2)) continue;
                double fq = fcut_only(rq, T.R1, T.R2);
                double rr = 1. / rp / rq, cosi = dot(dl, dq) * rr, c1 = 1. + cosi, den = T.d02 + c1 * c1;
                double gg = 2. * T.a0 * T.c02 * c1 / (den * den);
                Vec w = {-dq.x * rr + cosi * dl.x / rp2, -dq.y * rr + cosi * dl.y / rp2, -dq.z * rr + cosi * dl.z / rp2};
                double g = bl * fq * gg;
                dB.x += g * w.x; dB.y += g * w.y; dB.z += g * w.z;
                double pre = T.delt / 2 * Bx[(size_t)q * lv.stride + j];
                double g1 = f_c * gg, g2 = dfr_p * T.a0 * tb_G(c1, T);
                double tail = fq * dpre * T.s * mx::exp_fast(-s2s * T.b * (rq - T.r0) / T.s);
                cx += pre * (g1 * w.x + dp.x * g2) * tail;
                cy += pre * (g1 * w.y + dp.y * g2) * tail;
                cz += pre * (g1 * w.z + dp.z * g2) * tail;
            }
            double h = -T.delt / 2.;
            dB.x *= h; dB.y *= h; dB.z *= h;
            double dff = dfr_p / f_c;
            double k1 = (dff - s2s * T.b / rp) * ea;       // multiplies dp
            double k2 = (Bip + Bjl) / 2 * (dff - s2is * T.b / rp);
            double A = f_c * dpre, se = T.s * eas;
            fx = A * (dp.x * k1 - (dB.x + k2 * dp.x) * se) + cx;
            fy = A * (dp.y * k1 - (dB.y + k2 * dp.y) * se) + cy;
            fz = A * (dp.z * k1 - (dB.z + k2 * dp.z) * se) + cz;
        }
    } while (false);
}
template <bool F, bool E>
__global__ void __launch_bounds__(FT) k_tb_force(int N, const double4* __restrict__ pos, double4* __restrict__ fpart, ListView lv, TBp T, BoxD box,
                                                 const double* __restrict__ B, const double* __restrict__ Bx, double* part) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int p = blockIdx.y;
    double e = 0, fx = 0, fy = 0, fz = 0;
    int n = i < N ? lv.nnum[i] : 0;
    if (p < n) {
        tb_bond_terms<F, E>(i, p, n, pos, lv, T, box, B, Bx, e, fx, fy, fz);
        if (F) fpart[(size_t)p * lv.stride + i] = make_double4(fx, fy, fz, 0.);
    }
    if (E) {
        double s = block_sum(e);
        if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}
// (Tried and dropped, measured on a B200: pass B and the per-atom sum in one kernel, four lanes per atom taking its bonds and lane 0
//  adding them in slot order -- bit-identical, one launch less, but the lanes of a warp then skip different q in the loops below:
//  0.057 against 0.035 ms per step for the tb force pass of graphene on Cu, 1.02e8 against 1.37e8 atom-steps/s.
//  Also tried and dropped: eight lanes per (atom, bond) sharing the two partner loops of a bond (9 row entries each in graphene: 3
//  bonds + 6 second neighbours the switch zeroes) with a shuffle-tree sum -- parity with the oracle 1e-9, but eight times the
//  blocks, most of them empty: tb force pass 0.034 -> 0.041 ms, graphene on Cu 1.37e8 -> 1.13e8, 64 replicas on one GPU 2.9e8 -> 1.8e8.)
__global__ void __launch_bounds__(FT) k_tb_reduce(int N, const double4* __restrict__ fpart, double4* __restrict__ frc, ListView lv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int n = lv.nnum[i];
    if (n == 0) return;
    double fx = 0, fy = 0, fz = 0;
    for (int p = 0; p < n; ++p) {
        double4 f = fpart[(size_t)p * lv.stride + i];
        fx += f.x; fy += f.y; fz += f.z;
    }
    add_force(frc, i, fx, fy, fz);
}

// ---- graphene normals : graphenenorm.f90:38-56 ----------------------------------------------------
__global__ void k_normals(int N, const double4* __restrict__ pos, ListView nn, BoxD box, int simplified, double4* __restrict__ gnorm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (nn.nnum[i] != 3) return;
    if (simplified) { gnorm[i] = make_double4(0., 0., 1., 0.); return; }  // LennardJonesCosine.f90:62
    const double4 pi = pos[i];
    double r2;
    Vec d1 = bond_vec(pi, pos[nn.nlist[i]], box, r2);
    Vec d2 = bond_vec(pi, pos[nn.nlist[nn.stride + i]], box, r2);
    Vec d3 = bond_vec(pi, pos[nn.nlist[2 * nn.stride + i]], box, r2);
    Vec a = {d2.x - d1.x, d2.y - d1.y, d2.z - d1.z}, b = {d1.x - d3.x, d1.y - d3.y, d1.z - d3.z};
    Vec n = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    if (n.z < 0.) { n.x = -n.x; n.y = -n.y; n.z = -n.z; }
    double len = sqrt(dot(n, n));
    gnorm[i] = make_double4(n.x / len, n.y / len, n.z / len, 0.);
}

// ---- ljc / morsec : LennardJonesCosine.f90, MorseCosine.f90 ---------------------------------------
struct CosP { double pe, sig, a, r0, delt, R1, R2; };  // pe = 4 eps (ljc) or d (morsec)
template <bool MORSE>
__device__ __forceinline__ double cos_V2(double r, double r2, const CosP& P) {
    if (MORSE) return mx::exp_fast(-P.a * (r - P.r0));
    double q = P.sig * P.sig / r2;
    return q * q * q;
}
// Direct term for the atoms of one side.  GRAPHENE: owner i is a carbon atom with its own normal and
// the kernel also accumulates T_i = sum_p V2 V3 f_c /(n_i.dr) dr for the normal-derivative term.
// Otherwise the owner is a metal atom and the normal is that of the carbon partner (:112-139).
template <bool MORSE, bool GRAPHENE, bool F, bool E, int SPLIT>
__global__ void __launch_bounds__(FT) k_cos_direct(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, CosP P, BoxD box,
                                                   const double4* __restrict__ gnorm, double4* __restrict__ tvec, double* part) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t / SPLIT, sub = t % SPLIT;
    double e = 0, fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
    int n = i < N ? lv.nnum[i] : 0;
    if (n > 0) {
        const double4 pi = pos[i];
        double4 nv = GRAPHENE ? gnorm[i] : make_double4(0, 0, 0, 0);
        for (int p = sub; p < n; p += SPLIT) {
            int j = lv.nlist[(size_t)p * lv.stride + i];
            double r2;
            Vec d = bond_vec(pi, pos[j], box, r2);
            double r = sqrt(r2);
            if (r < P.R2) {
                if (!GRAPHENE) nv = gnorm[j];
                double nd = nv.x * d.x + nv.y * d.y + nv.z * d.z;
                double V2 = cos_V2<MORSE>(r, r2, P), V1 = V2 * V2;
                double cosn = fabs(nd) / r;
                if (MORSE && !GRAPHENE) cosn = fabs(nd) / (sqrt(nv.x * nv.x + nv.y * nv.y + nv.z * nv.z) * r);  // MorseCosine.f90:127
                double V3 = mx::pow_pos(cosn, P.delt);
                double f, dfr;
                fcut_dfcut(r, P.R1, P.R2, f, dfr);
                if (E) e += MORSE ? P.pe * (V1 - 2. * V2 * V3) * f : P.pe * (V1 - V2 * V3) * f;
                if (F) {
                    double cr, cn;
                    if (MORSE) {
                        cr = 2. * (P.a * V1 - (P.a + P.delt / r) * V2 * V3) / r * f - (V1 - 2. * V2 * V3) * dfr;
                        cn = 2. * P.delt * V2 * V3 / nd * f;
                    } else {
                        cr = (12. * V1 - (6. + P.delt) * V2 * V3) / r2 * f - (V1 - V2 * V3) * dfr;
                        cn = P.delt * V2 * V3 / nd * f;
                    }
                    fx -= P.pe * (cr * d.x + cn * nv.x);
                    fy -= P.pe * (cr * d.y + cn * nv.y);
                    fz -= P.pe * (cr * d.z + cn * nv.z);
                    if (GRAPHENE) {
                        double w = V2 * V3 * f / nd;
                        tx += w * d.x; ty += w * d.y; tz += w * d.z;
                    }
                }
            }
        }
    }
    if (F) {
        fx = split_sum<SPLIT>(fx); fy = split_sum<SPLIT>(fy); fz = split_sum<SPLIT>(fz);
        if (GRAPHENE) { tx = split_sum<SPLIT>(tx); ty = split_sum<SPLIT>(ty); tz = split_sum<SPLIT>(tz); }
        if (sub == 0 && i < N) {
            if (n > 0) add_force(frc, i, fx, fy, fz);
            if (GRAPHENE) tvec[i] = make_double4(tx, ty, tz, 0.);
        }
    }
    if (E) store_partial(e, part);
}
// Normal-derivative term (LennardJonesCosine.f90:81-106): for each of the three nearest carbons j of
// i, with i in slot l1 of j's row and l2,l3 the cyclic successors,
//   F_i -= pref*delt * n_j * (T_j . (d12 (d23.d31) - d31 (d23.d12))) / (|d12|^2|d31|^2 - (d12.d31)^2)
// where the reference's inner sum over j's metal neighbours has been collected into T_j.
__global__ void __launch_bounds__(FT) k_cos_indirect(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView nn, double pref_delt,
                                                     BoxD box, const double4* __restrict__ gnorm, const double4* __restrict__ tvec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (nn.nnum[i] != 3) return;
    double fx = 0, fy = 0, fz = 0;
    for (int q = 0; q < 3; ++q) {
        int j = nn.nlist[(size_t)q * nn.stride + i];
        int a0 = nn.nlist[j], a1 = nn.nlist[nn.stride + j], a2 = nn.nlist[2 * nn.stride + j];
        int l1 = a0 == i ? 0 : (a1 == i ? 1 : (a2 == i ? 2 : 3));
        if (l1 > 2) continue;  // reference stops with 'l1>3'; the nearest-3 relation is symmetric
        int k1 = l1 == 0 ? a0 : (l1 == 1 ? a1 : a2), k2 = l1 == 0 ? a1 : (l1 == 1 ? a2 : a0), k3 = l1 == 0 ? a2 : (l1 == 1 ? a0 : a1);
        const double4 pj = pos[j];
        double r2;
        Vec e1 = bond_vec(pj, pos[k1], box, r2), e2 = bond_vec(pj, pos[k2], box, r2), e3 = bond_vec(pj, pos[k3], box, r2);
        Vec d12 = {e2.x - e1.x, e2.y - e1.y, e2.z - e1.z}, d31 = {e1.x - e3.x, e1.y - e3.y, e1.z - e3.z}, d23 = {e3.x - e2.x, e3.y - e2.y, e3.z - e2.z};
        double s2331 = dot(d23, d31), s2312 = dot(d23, d12);
        Vec w = {d12.x * s2331 - d31.x * s2312, d12.y * s2331 - d31.y * s2312, d12.z * s2331 - d31.z * s2312};
        double d1231 = dot(d12, d31);
        double den = dot(d12, d12) * dot(d31, d31) - d1231 * d1231;
        double4 t = tvec[j], nj = gnorm[j];
        double c = pref_delt * (t.x * w.x + t.y * w.y + t.z * w.z) / den;
        fx -= c * nj.x; fy -= c * nj.y; fz -= c * nj.z;
    }
    add_force(frc, i, fx, fy, fz);
}

// ------------------------------------------------------------------------------------------------
// zero_forces touches the all_atoms group only (md_integrators.f90:147-163): forces of atoms outside it keep accumulating
__global__ void k_zero_group(int N, double4* __restrict__ frc, const uint32_t* __restrict__ gmask, uint32_t bit) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && (gmask[i] & bit)) frc[i] = make_double4(0., 0., 0., 0.);
}
#if defined(__CUDACC__) && defined(PFMDS_STAMPS)
STAMP_BIND_FN(forces_stamps_bind)
#endif
#ifdef PFMDS_HAVE_CTX  // ---- launchers (host side of the library) ----------------------------------------
void forces_zero(pfmds_ctx* c) {  // zero_forces, md_integrators.f90:147-163
    if (c->first_overwrites && c->N >= c->small_n) return;  // the first force kernel stores instead of accumulating
    if (c->fbuf_active) return;                             // per-interaction buffers: k_sum_kick_ke starts every atom's sum from zero_forces' value
    KTimer kt(c, KS_ZERO_FORCES);
    if (c->zero_all) { CK(cudaMemsetAsync(c->frc, 0, sizeof(double4) * (size_t)c->N, c->st)); return; }
    LAUNCH((k_zero_group), (c->N + 255) / 256, 256, c->st, c->N, c->frc, c->gmask, 1u << (c->all_atoms - 1));
    c->launches += 1;
}

static CosP cosp_of(const Inter& it) {
    CosP P{};
    if (it.kind == K_LJC) { P.pe = 4. * it.ljc.eps; P.sig = it.ljc.sig; P.delt = it.ljc.delt; P.R1 = it.ljc.R1; P.R2 = it.ljc.R2; }
    else { P.pe = it.mor.d; P.a = it.mor.a; P.r0 = it.mor.r; P.delt = it.mor.delt; P.R1 = it.mor.R1; P.R2 = it.mor.R2; }
    return P;
}

// third-generation rjl: build the node table of an interaction once (called when the description is finalized)
void rjl_prepare(pfmds_ctx* c, Inter& it) {
    if (it.kind != K_RJL || c->rjl_gen != 3 || it.aux || !rjl_gen3_ok(it.rjl, c->box)) return;
    const RjlTabSpec t = rjl_tab_spec(it.rjl);
    std::vector<double2> h((size_t)(t.J_hi - t.J_lo + 1));
    rjl_tab_fill(it.rjl, t, h.data());
    CK(cudaMalloc(&it.aux, sizeof(double2) * h.size()));
    CK(cudaMemcpy(it.aux, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
}

void normals_interaction(pfmds_ctx* c, int k, cudaStream_t st) {  // update_norm_in_graphene, md_interactions.f90:195-208
    Inter& it = c->inter[k];
    if (it.kind != K_LJC && it.kind != K_MORSEC) return;
    const int N = c->N, nb = (N + FT - 1) / FT;
    int simp = it.kind == K_LJC ? it.ljc.simplified : it.mor.simplified;
    KTimer kt(c, KS_NORMALS);
    LAUNCH((k_normals), nb, FT, st ? st : c->st, N, c->pos, it.nl[2].view(c->stride), c->box, simp, it.gnorm);
    c->launches += 1;
}

// with_energy: the same pass also yields the interaction's potential energy (what pfmds_energies would compute with a
// second sweep over the lists): fused into the force kernels for rjl, lj1g and lj, a follow-up kernel for the others.
void forces_interaction(pfmds_ctx* c, int k, bool with_energy) {  // calculate_forces, md_interactions.f90:210-242
    Inter& it = c->inter[k];
    // small systems, steps without energies: every interaction accumulates into its own force buffer on its own stream (parallel
    // branches of the step's CUDA graph), summed per atom in file order by integrate.cu k_sum_kick_ke -- see compute_forces (capi.cu)
    cudaStream_t fs = c->fst ? c->fst : c->st;
    double4* fo = c->fout ? c->fout : c->frc;
    double* epart = with_energy ? c->part : nullptr;
    int e_parts = 0;
    double e_scale = 1.0;
    const int N = c->N, nb = (N + FT - 1) / FT;
    const bool small = N < c->small_n;
    const int nbs = (int)(((size_t)N * SMALL_SPLIT + FT - 1) / FT);
    const size_t st = c->stride;
    switch (it.kind) {
    case K_LJ:
        if (with_energy) {
            KTimer kt(c, KS_LJ);
            if (small) LAUNCH((k_lj<true, true, SMALL_SPLIT>), nbs, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj, c->box, epart);
            else LAUNCH((k_lj<true, true, 1>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj, c->box, epart);
            e_parts = small ? nbs : nb;
        } else
        { KTimer kt(c, KS_LJ); if (small) LAUNCH((k_lj<true, false, SMALL_SPLIT>), nbs, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj, c->box, nullptr); else LAUNCH((k_lj<true, false, 1>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj, c->box, nullptr); }
        if (e_parts) { LAUNCH((k_sum_partials), 1, 1024, fs, e_parts, c->part, 1.0, c->energy + k); c->launches += 1; if (c->slab) slab_allreduce_sum(c, c->energy + k, 1); e_parts = 0; with_energy = false; }
        {   // converse list: owners are the atoms of group 2 (small systems: a branch and a buffer of its own)
            cudaStream_t fs2 = c->fout2 ? (c->fst2 ? c->fst2 : c->st) : fs;
            double4* fo2 = c->fout2 ? c->fout2 : fo;
            KTimer kt(c, KS_LJ);
            if (small) LAUNCH((k_lj<true, false, SMALL_SPLIT>), nbs, FT, fs2, N, c->pos, fo2, it.nl[1].view(st), it.lj, c->box, nullptr);
            else LAUNCH((k_lj<true, false, 1>), nb, FT, fs2, N, c->pos, fo2, it.nl[1].view(st), it.lj, c->box, nullptr);
        }
        c->launches += 2;
        break;
    case K_LJ1G:
        if (c->lj1g_pipe && !small) {  // pipelined variant (thread per atom)
            KTimer kt(c, KS_LJ1G);
            const int ow = (k == 0 && c->first_overwrites) ? 1 : 0;
            if (with_energy) { LAUNCH((k_lj1g_pipe<true>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj1g, c->box, wrap_consts(c->box), epart, ow); e_parts = nb; e_scale = 0.5; }
            else LAUNCH((k_lj1g_pipe<false>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj1g, c->box, wrap_consts(c->box), (double*)nullptr, ow);
            c->launches += 1;
            break;
        }
        if (with_energy) {
            KTimer kt(c, KS_LJ1G);
            if (small) LAUNCH((k_lj1g<true, true, SMALL_SPLIT>), nbs, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj1g, c->box, epart);
            else LAUNCH((k_lj1g<true, true, 1>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj1g, c->box, epart);
            e_parts = small ? nbs : nb; e_scale = 0.5;
        } else
        { KTimer kt(c, KS_LJ1G); if (small) LAUNCH((k_lj1g<true, false, SMALL_SPLIT>), nbs, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj1g, c->box, nullptr); else LAUNCH((k_lj1g<true, false, 1>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), it.lj1g, c->box, nullptr); }
        c->launches += 1;
        break;
    case K_RJL:
    {
        const WrapC W = wrap_consts(c->box);
        const bool fused = c->slab && !small && slab_fused(c);  // density stores 1/Eb into the neighbours' ghosts itself
        const bool gen2 = c->rjl_gen != 1 && rjl_gen2_ok(it.rjl, c->box);
        const bool gen3 = gen2 && c->rjl_gen == 3 && it.aux != nullptr;  // node table built by rjl_prepare
        const ListView lv = it.nl[0].view(st);
        const int ow = (k == 0 && c->first_overwrites && !small) ? 1 : 0;
        // one launch sequence for both generations: CD / CF are the constant packs that select the pair routines.
        // Energies of the step: the first generation evaluates them in its density pass (a second exponential per pair), the
        // second takes them from the force pass, which has that exponential in hand (k_rjl_force_e).
        const bool e_in_force = with_energy && gen2;
        auto run = [&](auto CD, auto CF) {
            using TD = decltype(CD);
            using TF = decltype(CF);
            {
                KTimer kt(c, KS_RJL_DENSITY);
                if (with_energy && !e_in_force) {
                    if (small) LAUNCH((k_rjl_density_split<true, SMALL_SPLIT, TD>), nbs, FT, fs, N, c->pos, lv, CD, c->box, W, epart);
                    else LAUNCH((k_rjl_density<true, TD>), nb, FT, fs, N, c->pos, lv, CD, c->box, W, epart, fused ? slab_dev(c, 1) : SlabDev{});
                    e_parts = small ? nbs : nb;
                } else
                if (small) LAUNCH((k_rjl_density_split<false, SMALL_SPLIT, TD>), nbs, FT, fs, N, c->pos, lv, CD, c->box, W, (double*)nullptr);
                else LAUNCH((k_rjl_density<false, TD>), nb, FT, fs, N, c->pos, lv, CD, c->box, W, (double*)nullptr, fused ? slab_dev(c, 1) : SlabDev{});
            }
            if (c->slab && !fused) slab_exchange(c, 1);  // ghost 1/Eb from their owners
            if (!e_in_force) {
                KTimer kt(c, KS_RJL_FORCE);
                if (small) LAUNCH((k_rjl_force_split<SMALL_SPLIT, TF>), nbs, FT, fs, N, c->pos, fo, lv, CF, c->box, W);
                else LAUNCH((k_rjl_force<TF>), nb, FT, fs, N, c->pos, fo, lv, CF, c->box, W, fused ? slab_dev(c, 2) : SlabDev{}, ow);
            }
        };
        const double erep = it.rjl.r0 / (2. * it.rjl.p);   // A0 / a1, a1 = 2 A0 p / r0
        auto force_e = [&](auto CF) {   // force pass that also yields the step's energy
            using TF = decltype(CF);
            KTimer kt(c, KS_RJL_FORCE);
            if (small) { LAUNCH((k_rjl_force_split_e<SMALL_SPLIT, TF>), nbs, FT, fs, N, c->pos, fo, lv, CF, c->box, W, erep, it.rjl.xi, epart); e_parts = nbs; return; }
            const SlabDev SD = fused ? slab_dev(c, 2) : SlabDev{};
            LAUNCH((k_rjl_force_e<TF>), nb, FT, fs, N, c->pos, fo, lv, CF, c->box, W, SD, ow, erep, it.rjl.xi, epart);
            e_parts = nb;
        };
        if (gen3) {
            const RjlG G = rjl_g3_consts(it.rjl, rjl_tab_spec(it.rjl), reinterpret_cast<const double2*>(it.aux));
            run(G, G);
            if (e_in_force) force_e(G);
        } else if (e_in_force) {
            run(rjl_dens_consts(it.rjl), rjl_force_consts(it.rjl));
            force_e(rjl_force_consts(it.rjl));
        } else
        if (gen2) run(rjl_dens_consts(it.rjl), rjl_force_consts(it.rjl));
        else { const RjlC C = rjl_consts(it.rjl); run(C, C); }
    }
        c->launches += 2;
        break;
    case K_REBOSC:  // no analytic force in the reference: central differences of the energy (rebosc.cu)
        rebosc_forces(c, it);
        break;
    case K_TB:
    {
        dim3 grid(nb, it.nl[0].maxn);
        { KTimer kt(c, KS_TB_BOND); LAUNCH((k_tb_bond), grid, FT, fs, N, c->pos, it.nl[0].view(st), it.tb, c->box, it.aux, it.aux2); }
        {
            KTimer kt(c, KS_TB_FORCE);
            LAUNCH((k_tb_force<true, false>), grid, FT, fs, N, c->pos, it.fpart, it.nl[0].view(st), it.tb, c->box, it.aux, it.aux2, nullptr);
            LAUNCH((k_tb_reduce), nb, FT, fs, N, it.fpart, fo, it.nl[0].view(st));
            c->launches += 3;
        }
    }
        break;
    case K_LJC:
    case K_MORSEC: {
        if (c->fbuf_on) normals_interaction(c, k, fs);   // small systems: the normals open this interaction's branch (otherwise update_lists ran them)
        // the metal side needs the normals only: small systems give it a branch and a buffer of its own, next to graphene's direct + indirect chain
        cudaStream_t fs2 = c->fout2 ? (c->fst2 ? c->fst2 : c->st) : fs;
        double4* fo2 = c->fout2 ? c->fout2 : fo;
        if (c->fout2 && c->fst2) { CK(cudaEventRecord(c->aux_ev_mid, fs)); CK(cudaStreamWaitEvent(fs2, c->aux_ev_mid, 0)); }
        CosP P = cosp_of(it);
        bool simp = it.kind == K_LJC ? it.ljc.simplified : it.mor.simplified;
        if (it.kind == K_LJC) {
            { KTimer kt(c, KS_COS_GRAPHENE); if (small) LAUNCH((k_cos_direct<false, true, true, false, SMALL_SPLIT>), nbs, FT, fs, N, c->pos, fo, it.nl[0].view(st), P, c->box, it.gnorm, it.tvec, nullptr); else LAUNCH((k_cos_direct<false, true, true, false, 1>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), P, c->box, it.gnorm, it.tvec, nullptr); }
            if (!simp) { KTimer kt(c, KS_COS_INDIRECT); LAUNCH((k_cos_indirect), nb, FT, fs, N, c->pos, fo, it.nl[2].view(st), P.pe * P.delt, c->box, it.gnorm, it.tvec); }
            { KTimer kt(c, KS_COS_METAL); if (small) LAUNCH((k_cos_direct<false, false, true, false, SMALL_SPLIT>), nbs, FT, fs2, N, c->pos, fo2, it.nl[1].view(st), P, c->box, it.gnorm, it.tvec, nullptr); else LAUNCH((k_cos_direct<false, false, true, false, 1>), nb, FT, fs2, N, c->pos, fo2, it.nl[1].view(st), P, c->box, it.gnorm, it.tvec, nullptr); }
        } else {
            { KTimer kt(c, KS_COS_GRAPHENE); if (small) LAUNCH((k_cos_direct<true, true, true, false, SMALL_SPLIT>), nbs, FT, fs, N, c->pos, fo, it.nl[0].view(st), P, c->box, it.gnorm, it.tvec, nullptr); else LAUNCH((k_cos_direct<true, true, true, false, 1>), nb, FT, fs, N, c->pos, fo, it.nl[0].view(st), P, c->box, it.gnorm, it.tvec, nullptr); }
            if (!simp) { KTimer kt(c, KS_COS_INDIRECT); LAUNCH((k_cos_indirect), nb, FT, fs, N, c->pos, fo, it.nl[2].view(st), 2. * P.pe * P.delt, c->box, it.gnorm, it.tvec); }
            { KTimer kt(c, KS_COS_METAL); if (small) LAUNCH((k_cos_direct<true, false, true, false, SMALL_SPLIT>), nbs, FT, fs2, N, c->pos, fo2, it.nl[1].view(st), P, c->box, it.gnorm, it.tvec, nullptr); else LAUNCH((k_cos_direct<true, false, true, false, 1>), nb, FT, fs2, N, c->pos, fo2, it.nl[1].view(st), P, c->box, it.gnorm, it.tvec, nullptr); }
        }
        c->launches += simp ? 2 : 3;
        break;
    }
    }
    if (with_energy) {
        if (e_parts) {
            LAUNCH((k_sum_partials), 1, 1024, fs, e_parts, c->part, e_scale, c->energy + k);
            c->launches += 1;
            if (c->slab) slab_allreduce_sum(c, c->energy + k, 1);
        } else {
            energy_interaction(c, k);  // tb, ljc, morsec: separate sweep, still inside the step (no host round trip)
        }
    }
    CK(cudaGetLastError());
}

void energy_interaction(pfmds_ctx* c, int k) {  // energy(), md_interactions.f90:244-261: always on nl(1)
    Inter& it = c->inter[k];
    const int N = c->N, nb = (N + FT - 1) / FT;
    const size_t st = c->stride;
    double scale = 1.0;
    const bool small = N < c->small_n;
    const int nbs = (int)(((size_t)N * SMALL_SPLIT + FT - 1) / FT);
    int nparts = small ? nbs : nb;
    switch (it.kind) {
    case K_LJ: if (small) LAUNCH((k_lj<false, true, SMALL_SPLIT>), nbs, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), it.lj, c->box, c->part); else LAUNCH((k_lj<false, true, 1>), nb, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), it.lj, c->box, c->part); break;
    case K_LJ1G: if (small) LAUNCH((k_lj1g<false, true, SMALL_SPLIT>), nbs, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), it.lj1g, c->box, c->part); else LAUNCH((k_lj1g<false, true, 1>), nb, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), it.lj1g, c->box, c->part); scale = 0.5; break;
    case K_RJL: {
        const WrapC W = wrap_consts(c->box);
        const ListView lv = it.nl[0].view(st);
        auto run = [&](auto CD) {
            using TD = decltype(CD);
            if (small) LAUNCH((k_rjl_density_split<true, SMALL_SPLIT, TD>), nbs, FT, c->st, N, c->pos, lv, CD, c->box, W, c->part);
            else LAUNCH((k_rjl_density<true, TD>), nb, FT, c->st, N, c->pos, lv, CD, c->box, W, c->part, SlabDev{});
        };
        if (c->rjl_gen == 3 && it.aux != nullptr && rjl_gen2_ok(it.rjl, c->box)) run(rjl_g3_consts(it.rjl, rjl_tab_spec(it.rjl), reinterpret_cast<const double2*>(it.aux)));
        else if (c->rjl_gen != 1 && rjl_gen2_ok(it.rjl, c->box)) run(rjl_dens_consts(it.rjl));
        else run(rjl_consts(it.rjl));
        break;
    }
    case K_REBOSC: nparts = rebosc_energy_partials(c, it); c->launches -= 1; break;
    case K_TB: {
        dim3 grid(nb, it.nl[0].maxn);
        LAUNCH((k_tb_bond), grid, FT, c->st, N, c->pos, it.nl[0].view(st), it.tb, c->box, it.aux, it.aux2);
        LAUNCH((k_tb_force<false, true>), grid, FT, c->st, N, c->pos, it.fpart, it.nl[0].view(st), it.tb, c->box, it.aux, it.aux2, c->part);
        nparts = nb * it.nl[0].maxn;
        c->launches += 1;
        break;
    }
    case K_LJC: if (small) LAUNCH((k_cos_direct<false, true, false, true, SMALL_SPLIT>), nbs, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), cosp_of(it), c->box, it.gnorm, it.tvec, c->part); else LAUNCH((k_cos_direct<false, true, false, true, 1>), nb, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), cosp_of(it), c->box, it.gnorm, it.tvec, c->part); break;
    case K_MORSEC: if (small) LAUNCH((k_cos_direct<true, true, false, true, SMALL_SPLIT>), nbs, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), cosp_of(it), c->box, it.gnorm, it.tvec, c->part); else LAUNCH((k_cos_direct<true, true, false, true, 1>), nb, FT, c->st, N, c->pos, c->frc, it.nl[0].view(st), cosp_of(it), c->box, it.gnorm, it.tvec, c->part); break;
    }
    LAUNCH((k_sum_partials), 1, 1024, c->st, nparts, c->part, scale, c->energy + k);
    c->launches += 2;
    if (c->slab) slab_allreduce_sum(c, c->energy + k, 1);
    CK(cudaGetLastError());
}
#endif  // PFMDS_HAVE_CTX
