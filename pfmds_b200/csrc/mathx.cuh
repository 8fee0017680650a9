// pfmds_b200 — FP64 elementary functions for the force kernels.
//
// The CUDA math library's exp/sincos/sqrt materialise every 64-bit polynomial coefficient with
// MOV/IMAD pairs inside the pair loop (28 % of the issued instructions of the first rjl kernels,
// profiles/r1_*): here the coefficients live in the constant bank and enter DFMA as c[bank][off]
// operands, range reduction is specialised to what the potentials need, and there are no slow paths.
// Accuracy (tests/test_mathx.py, compiled for the host from this same header): relative error
// < 3e-14 for exp on [-700, 700], absolute error < 4e-16 for sin/cos on [0, pi], relative error
// < 5e-16 for rsqrt — three to six orders of magnitude inside the 1e-9 parity bar.
#pragma once
#include <math.h>
#include <string.h>

#ifdef __CUDACC__
#define MX_HD __host__ __device__ __forceinline__
#define MX_CONST __constant__
#else
#define MX_HD inline
#define MX_CONST static const
#endif

namespace mx {

// exp(r) = sum r^n/n!  on |r| <= ln2/4 after halving: coefficients of (r/2)^n folded in
MX_CONST double EXP_C[10] = {1.0,
                             1.0 / 2.0,
                             1.0 / 8.0,
                             1.0 / 48.0,
                             1.0 / 384.0,
                             1.0 / 3840.0,
                             1.0 / 46080.0,
                             1.0 / 645120.0,
                             1.0 / 10321920.0,
                             1.0 / 185794560.0};
// sin(y) = y + y^3 S(y^2), cos(y) = 1 + y^2 C(y^2) on |y| <= pi/4
MX_CONST double SIN_C[7] = {-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0, 1.0 / 6227020800.0, -1.0 / 1307674368000.0};
MX_CONST double COS_C[8] = {-1.0 / 2.0, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0, -1.0 / 87178291200.0,
                            1.0 / 20922789888000.0};

MX_HD double as_double(long long v) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(v);
#else
    double d; memcpy(&d, &v, 8); return d;
#endif
}
MX_HD long long as_ll(double d) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(d);
#else
    long long v; memcpy(&v, &d, 8); return v;
#endif
}

// exp(x), |x| clamped to 700.  k = round(x/ln2) by the 1.5*2^52 trick, r = x - k ln2 (two-part ln2),
// e^r = (P9(r/2))^2, result scaled by adding k to the exponent field.
MX_HD double exp_fast(double x) {
    x = fmin(fmax(x, -700.0), 700.0);
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    double t = fma(x, 1.4426950408889634, MAGIC);
    double kd = t - MAGIC;
    long long k = (long long)(int)(unsigned int)(as_ll(t) & 0xffffffffll);  // low word of t holds k (two's complement)
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = EXP_C[9];
    p = fma(p, r, EXP_C[8]);
    p = fma(p, r, EXP_C[7]);
    p = fma(p, r, EXP_C[6]);
    p = fma(p, r, EXP_C[5]);
    p = fma(p, r, EXP_C[4]);
    p = fma(p, r, EXP_C[3]);
    p = fma(p, r, EXP_C[2]);
    p = fma(p, r, EXP_C[1]);
    p = fma(p, r, EXP_C[0]);
    p = p * p;
    return as_double(as_ll(p) + (long long)((unsigned long long)k << 52));
}

// exp(x) without the clamp, for arguments the caller has bounded to |x| < 700
MX_HD double exp_nc(double x) {
    const double MAGIC = 6755399441055744.0;
    double t = fma(x, 1.4426950408889634, MAGIC);
    double kd = t - MAGIC;
    long long k = (long long)(int)(unsigned int)(as_ll(t) & 0xffffffffll);
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = EXP_C[9];
    p = fma(p, r, EXP_C[8]);
    p = fma(p, r, EXP_C[7]);
    p = fma(p, r, EXP_C[6]);
    p = fma(p, r, EXP_C[5]);
    p = fma(p, r, EXP_C[4]);
    p = fma(p, r, EXP_C[3]);
    p = fma(p, r, EXP_C[2]);
    p = fma(p, r, EXP_C[1]);
    p = fma(p, r, EXP_C[0]);
    p = p * p;
    return as_double(as_ll(p) + (long long)((unsigned long long)k << 52));
}

// Cosine switch without selects: for a in [0, pi], u = a/2 - pi/4 in [-pi/4, pi/4],
//   (1 + cos a)/2 = cos^2(a/2) = (cos u - sin u)^2 / 2,   sin a = (cos u - sin u)(cos u + sin u).
#ifdef MX_NOINLINE_SWITCH
__device__ __noinline__ void cos_switch(double a, double& half_one_plus_cos, double& sin_a) {
#else
MX_HD void cos_switch(double a, double& half_one_plus_cos, double& sin_a) {
#endif
    const double PIO4_HI = 7.85398163397448279e-01, PIO4_LO = 3.06161699786838302e-17;
    double u = (fma(a, 0.5, -PIO4_HI)) - PIO4_LO;
    double u2 = u * u;
    double ps = SIN_C[6];
    ps = fma(ps, u2, SIN_C[5]);
    ps = fma(ps, u2, SIN_C[4]);
    ps = fma(ps, u2, SIN_C[3]);
    ps = fma(ps, u2, SIN_C[2]);
    ps = fma(ps, u2, SIN_C[1]);
    ps = fma(ps, u2, SIN_C[0]);
    double su = fma(u * u2, ps, u);
    double pc = COS_C[7];
    pc = fma(pc, u2, COS_C[6]);
    pc = fma(pc, u2, COS_C[5]);
    pc = fma(pc, u2, COS_C[4]);
    pc = fma(pc, u2, COS_C[3]);
    pc = fma(pc, u2, COS_C[2]);
    pc = fma(pc, u2, COS_C[1]);
    pc = fma(pc, u2, COS_C[0]);
    double cu = fma(u2, pc, 1.0);
    double d = cu - su;
    half_one_plus_cos = 0.5 * d * d;
    sin_a = d * (cu + su);
}

// sin and cos of a in [0, pi] (slightly outside is fine): y = a - pi/2 folded to |y| <= pi/4
MX_HD void sincos_0pi(double a, double& s, double& c) {
    const double PIO2_HI = 1.57079632679489655800e+00, PIO2_LO = 6.12323399573676603587e-17;
    double y = (a - PIO2_HI) - PIO2_LO;  // in [-pi/2, pi/2]: sin a = cos y, cos a = -sin y
    // fold: |y| > pi/4  ->  z = pi/2 - |y|, cos y = sin z, sin|y| = cos z
    double ay = fabs(y);
    bool big = ay > 0.78539816339744830962;
    double z = big ? (PIO2_HI - ay) + PIO2_LO : ay;
    double z2 = z * z;
    double ps = SIN_C[6];
    ps = fma(ps, z2, SIN_C[5]);
    ps = fma(ps, z2, SIN_C[4]);
    ps = fma(ps, z2, SIN_C[3]);
    ps = fma(ps, z2, SIN_C[2]);
    ps = fma(ps, z2, SIN_C[1]);
    ps = fma(ps, z2, SIN_C[0]);
    double sz = fma(z * z2, ps, z);
    double pc = COS_C[7];
    pc = fma(pc, z2, COS_C[6]);
    pc = fma(pc, z2, COS_C[5]);
    pc = fma(pc, z2, COS_C[4]);
    pc = fma(pc, z2, COS_C[3]);
    pc = fma(pc, z2, COS_C[2]);
    pc = fma(pc, z2, COS_C[1]);
    pc = fma(pc, z2, COS_C[0]);
    double cz = fma(z2, pc, 1.0);
    double sin_ay = big ? cz : sz, cos_y = big ? sz : cz;
    s = cos_y;                            // sin a
    c = (y < 0.0) ? sin_ay : -sin_ay;     // cos a = -sin y
}

// 1/sqrt(x) for normal positive x: hardware seed (MUFU.RSQ64H, relative error 9e-7 measured on B200)
// and one third-order step  y <- y + y e (1/2 + 3/8 e),  e = 1 - x y^2  (error ~ 5/16 e^3 < 1e-18).
MX_HD double rsqrt_fast(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    double y = (double)(1.0f / sqrtf((float)x)) * (1.0 + 9e-7);  // host stand-in with the device seed's error
#endif
    double e = fma(-(x * y), y, 1.0);
    double t = fma(e, 0.375, 0.5);
    return fma(y * e, t, y);
}

// ---- short forms for the rjl pair loops (second generation kernels, forces.cu) -------------------------------------
// Those loops are bound by the FP64 pipe (every warp instruction holds the 16-lane pipe of an SMSP for two cycles) and by
// issue slots at the same time, so these variants spend fewer instructions at a precision still 10^3..10^5 inside the
// 1e-9 parity bar.  Coefficients are minimax fits (tools/minimax.py, Remez exchange in 60-digit arithmetic) with the
// leading terms kept exact, so those enter DFMA as immediates.
//   e^(r/2) = 1 + r/2 + r^2 Q(r),  |r| <= ln2/2:  relative error 2.3e-15
MX_CONST double EXP_M[7] = {0x1.0000000000601p-3, 0x1.55555555fb059p-6, 0x1.5555555062fc2p-9, 0x1.11110a0e2e82cp-12,
                            0x1.6c16e2fd4e000p-16, 0x1.a0755f88e73a1p-20, 0x1.9fb49a4f91900p-24};
//   sin u = u + u^3 S(u^2), cos u = 1 + u^2 C(u^2),  |u| <= pi/4:  absolute errors 5.0e-15 and 9.8e-17
MX_CONST double SIN_M[5] = {-0x1.5555555552d7dp-3, 0x1.1111110ccb02ap-7, -0x1.a019f92d8d0adp-13, 0x1.71d7452e064bfp-19, -0x1.a94b85bcf213fp-26};
MX_CONST double COS_M[6] = {-0x1.fffffffffffbcp-2, 0x1.555555555023bp-5, -0x1.6c16c164a5ccbp-10, 0x1.a019f8578ec11p-16, -0x1.27df47e14bfc5p-22,
                            0x1.1b87a0dbfd873p-29};
//   (1 + cos a)/2 = 1/2 - y/2 + y^3 H(y^2),  y = a - pi/2 in [-pi/2, pi/2]:  absolute error 1.4e-16
MX_CONST double HSW_M[7] = {0x1.55555555554f5p-4, -0x1.111111110c0eep-8, 0x1.a01a019b327e4p-14, -0x1.71de38306b5fcp-20, 0x1.ae635d685da20p-27,
                            -0x1.60e748c48ee26p-34, 0x1.9f157524660aap-42};

MX_HD int hi_word(double d) {
#ifdef __CUDA_ARCH__
    return __double2hiint(d);
#else
    return (int)(as_ll(d) >> 32);
#endif
}
MX_HD int lo_word(double d) {
#ifdef __CUDA_ARCH__
    return __double2loint(d);
#else
    return (int)(as_ll(d) & 0xffffffffll);
#endif
}
MX_HD double from_words(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    return as_double((long long)(((unsigned long long)(unsigned int)hi << 32) | (unsigned int)lo));
#endif
}

// exp(x) for |x| < 700 (no clamp): k = round(x log2 e) by the 1.5*2^52 trick, r = x - k ln2 with ONE fused step (the product
// is exact inside the FMA; what is lost is |k| * (ln2 - fl(ln2)) <= 1024 * 2.3e-17 absolute in r), e^r = (1 + r/2 + r^2 Q(r))^2,
// and 2^k enters through one integer add on the high word.  Relative error < 3e-14 on the whole range, < 8e-15 for |x| < 40 (measured 6.2e-15).
MX_HD double exp_m(double x) {
    const double MAGIC = 6755399441055744.0;
    double t = fma(x, 1.4426950408889634, MAGIC);
    double kd = t - MAGIC;
    double r = fma(kd, -6.93147180559945309417e-01, x);
    double p = EXP_M[6];
    p = fma(p, r, EXP_M[5]);
    p = fma(p, r, EXP_M[4]);
    p = fma(p, r, EXP_M[3]);
    p = fma(p, r, EXP_M[2]);
    p = fma(p, r, EXP_M[1]);
    p = fma(p, r, EXP_M[0]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * p;
    return from_words(hi_word(p) + (int)((unsigned int)lo_word(t) << 20), lo_word(p));  // low word of t holds k (two's complement)
}

// two independent exp_m evaluated side by side: every coefficient is fetched once for both Horner chains, and the two
// dependent chains fill each other's FP64 latency slots
// (log2 e and -ln 2 are arguments so that a kernel can hand them over as launch parameters: those stay in uniform
// registers across the pair loop, literals are rebuilt from immediates for every pair)
MX_HD void exp_m2(double xa, double xb, double& ea, double& eb, double l2e = 1.4426950408889634, double nln2 = -6.93147180559945309417e-01) {
    const double MAGIC = 6755399441055744.0;
    double ta = fma(xa, l2e, MAGIC), tb = fma(xb, l2e, MAGIC);
    double ka = ta - MAGIC, kb = tb - MAGIC;
    double ra = fma(ka, nln2, xa), rb = fma(kb, nln2, xb);
    double pa = EXP_M[6], pb = EXP_M[6];
    pa = fma(pa, ra, EXP_M[5]); pb = fma(pb, rb, EXP_M[5]);
    pa = fma(pa, ra, EXP_M[4]); pb = fma(pb, rb, EXP_M[4]);
    pa = fma(pa, ra, EXP_M[3]); pb = fma(pb, rb, EXP_M[3]);
    pa = fma(pa, ra, EXP_M[2]); pb = fma(pb, rb, EXP_M[2]);
    pa = fma(pa, ra, EXP_M[1]); pb = fma(pb, rb, EXP_M[1]);
    pa = fma(pa, ra, EXP_M[0]); pb = fma(pb, rb, EXP_M[0]);
    pa = fma(pa, ra, 0.5); pb = fma(pb, rb, 0.5);
    pa = fma(pa, ra, 1.0); pb = fma(pb, rb, 1.0);
    pa = pa * pa; pb = pb * pb;
    ea = from_words(hi_word(pa) + (int)((unsigned int)lo_word(ta) << 20), lo_word(pa));
    eb = from_words(hi_word(pb) + (int)((unsigned int)lo_word(tb) << 20), lo_word(pb));
}

// 1/sqrt(x) for normal positive x: hardware seed y (relative error d <= 9e-7) and one second-order step written so that it
// costs three FP64 instructions, y (3/2 - x y^2 / 2) = y + y (1 - x y^2)/2: -y/2 is the seed with 0x7ff00000 added to its
// high word (exponent - 1, sign set; the seed's low word is zero).  Relative error 3/2 d^2 <= 1.3e-12.
MX_HD double rsqrt_q(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    double y = from_words(hi_word((double)(1.0f / sqrtf((float)x)) * (1.0 + 9e-7)), 0);
#endif
    double yh = from_words((int)((unsigned int)hi_word(y) + 0x7ff00000u), 0);  // -y/2
    double b = fma(x * yh, y, 1.5);
    return y * b;
}

// Cosine switch from the half angle u = a/2 - pi/4 in [-pi/4, pi/4] (the caller forms u with one FMA from r):
//   (1 + cos a)/2 = (cos u - sin u)^2 / 2,   sin a = (cos u - sin u)(cos u + sin u).
MX_HD void cos_switch_m(double u, double& half_one_plus_cos, double& sin_a) {
    double u2 = u * u;
    double ps = SIN_M[4];
    ps = fma(ps, u2, SIN_M[3]);
    ps = fma(ps, u2, SIN_M[2]);
    ps = fma(ps, u2, SIN_M[1]);
    ps = fma(ps, u2, SIN_M[0]);
    double su = fma(u * u2, ps, u);
    double pc = COS_M[5];
    pc = fma(pc, u2, COS_M[4]);
    pc = fma(pc, u2, COS_M[3]);
    pc = fma(pc, u2, COS_M[2]);
    pc = fma(pc, u2, COS_M[1]);
    pc = fma(pc, u2, COS_M[0]);
    double cu = fma(u2, pc, 1.0);
    double d = cu - su;
    half_one_plus_cos = (0.5 * d) * d;
    sin_a = d * (cu + su);
}

// Value of the switch alone, (1 + cos a)/2 = (1 - sin y)/2 with y = a - pi/2 in [-pi/2, pi/2] (one FMA from r in the caller).
MX_HD double half_switch(double y) {
    double y2 = y * y;
    double p = HSW_M[6];
    p = fma(p, y2, HSW_M[5]);
    p = fma(p, y2, HSW_M[4]);
    p = fma(p, y2, HSW_M[3]);
    p = fma(p, y2, HSW_M[2]);
    p = fma(p, y2, HSW_M[1]);
    p = fma(p, y2, HSW_M[0]);
    return fma(y * y2, p, fma(y, -0.5, 0.5));
}

// 1/x for normal x: hardware seed (MUFU.RCP64H) and one third-order step y <- y + y (e + e^2), e = 1 - x y.
MX_HD double rcp_fast(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    double y = (double)(1.0f / (float)x) * (1.0 + 9e-7);
#endif
    double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// log(x) for normal positive x: x = m 2^k with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(s), s = (m-1)/(m+1),
// |s| <= 0.1716, odd series to s^19 (truncation 4e-18).
MX_CONST double LOG_C[10] = {2.0, 2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0, 2.0 / 17.0, 2.0 / 19.0};
MX_HD double log_fast(double x) {
    long long b = as_ll(x);
    int k = (int)(b >> 52) - 1023;
    long long mb = (b & 0x000fffffffffffffll) | 0x3ff0000000000000ll;  // m in [1,2)
    double m = as_double(mb);
    if (m > 1.4142135623730951) { m *= 0.5; k += 1; }
    double s = (m - 1.0) * rcp_fast(m + 1.0);
    double s2 = s * s;
    double p = LOG_C[9];
    p = fma(p, s2, LOG_C[8]);
    p = fma(p, s2, LOG_C[7]);
    p = fma(p, s2, LOG_C[6]);
    p = fma(p, s2, LOG_C[5]);
    p = fma(p, s2, LOG_C[4]);
    p = fma(p, s2, LOG_C[3]);
    p = fma(p, s2, LOG_C[2]);
    p = fma(p, s2, LOG_C[1]);
    p = fma(p, s2, LOG_C[0]);
    double kd = (double)k;
    return fma(kd, 6.93147180369123816490e-01, fma(s, p, kd * 1.90821492927058770002e-10));
}
// x^y for x >= 0 (0^y = 0 for y > 0), as the potentials use it: exp(y log x)
MX_HD double pow_pos(double x, double y) {
    if (!(x > 0.0)) return 0.0;
    return exp_fast(y * log_fast(x));
}

}  // namespace mx
