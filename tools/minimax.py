#!/usr/bin/env python
"""Minimax (Remez exchange, mpmath) coefficients of the short polynomials in pfmds_b200/csrc/mathx.cuh (the *_M tables).

  python tools/minimax.py            prints the tables as C initialisers with the error of the double-rounded coefficients

Each fit is  target(x) ~ sum_i c_i basis_i(x)  minimising  max |weight(x) (fit - target)|  on [lo, hi]; the fixed leading
terms (1 + r/2 for exp, u for sin, 1 for cos) stay exact so that they enter DFMA as immediates.
"""
from mpmath import mp, mpf, matrix, lu_solve, findroot, sin, cos, exp, pi, log, sqrt, taylor

mp.dps = 60


def remez(target, weight, nbasis, lo, hi, iters=40):
    """c_0..c_{n-1} for sum c_i x^i; returns (coefficients, levelled error)."""
    n = nbasis
    xs = [(lo + hi) / 2 - (hi - lo) / 2 * cos(pi * k / n) for k in range(n + 1)]
    c, E = None, None
    for _ in range(iters):
        A = matrix(n + 1, n + 1)
        b = matrix(n + 1, 1)
        for j, x in enumerate(xs):
            w = weight(x)
            for i in range(n):
                A[j, i] = w * x ** i
            A[j, n] = (-1) ** j
            b[j] = w * target(x)
        sol = lu_solve(A, b)
        c, E = [sol[i] for i in range(n)], sol[n]

        def err(x):
            return weight(x) * (sum(ci * x ** i for i, ci in enumerate(c)) - target(x))

        # new reference: extremum of the error in each interval between sign changes (dense scan + golden refinement)
        M = 4000
        grid = [lo + (hi - lo) * k / M for k in range(M + 1)]
        ev = [err(x) for x in grid]
        ext = []
        for k in range(M + 1):
            l = ev[k - 1] if k > 0 else None
            r = ev[k + 1] if k < M else None
            v = ev[k]
            if (l is None or abs(v) >= abs(l)) and (r is None or abs(v) >= abs(r)):
                if ext and (ext[-1][1] > 0) == (v > 0):
                    if abs(v) > abs(ext[-1][1]):
                        ext[-1] = (grid[k], v)
                else:
                    ext.append((grid[k], v))
        if len(ext) < n + 1:
            break
        # keep the n+1 consecutive extrema with the largest minimum magnitude
        best = max(range(len(ext) - n), key=lambda s: min(abs(e[1]) for e in ext[s:s + n + 1]))
        new = [e[0] for e in ext[best:best + n + 1]]
        if max(abs(a - b_) for a, b_ in zip(new, xs)) < (hi - lo) * mpf(10) ** -12:
            xs = new
            break
        xs = new
    return c, abs(E)


def max_err(fn, lo, hi, M=20000):
    return max(abs(fn(lo + (hi - lo) * k / M)) for k in range(M + 1))


def as_c(name, coeffs, note):
    body = ",\n".join("    %s" % float(c).hex() for c in coeffs)
    print("// %s\nMX_CONST double %s[%d] = {\n%s};" % (note, name, len(coeffs), body))
    print("//   decimal: " + ", ".join("%.17g" % float(c) for c in coeffs))


def main():
    # ---- exp: e^(r/2) = 1 + r/2 + r^2 Q(r) on |r| <= ln2/2 (+ margin), relative error ----
    L = log(2) / 2 * (1 + mpf(2) ** -9)
    tgt = lambda r: (exp(r / 2) - 1 - r / 2) / r ** 2 if abs(r) > mpf(10) ** -8 else mpf(1) / 8 + r / 48
    wgt = lambda r: r ** 2 / exp(r / 2)
    for nq in (6, 7):
        c, E = remez(tgt, wgt, nq, -L, L)
        cd = [mpf(float(x)) for x in c]
        e = max_err(lambda r: (1 + r / 2 + r ** 2 * sum(ci * r ** i for i, ci in enumerate(cd))) / exp(r / 2) - 1, -L, L)
        as_c("EXP_M%d" % nq, cd, "e^(r/2) = 1 + r/2 + r^2 Q(r), Q of degree %d: relative error %.2e (levelled %.2e); e^r = (.)^2 doubles it" % (nq - 1, float(e), float(E)))
    # ---- sin u = u + u^3 S(u^2), cos u = 1 + u^2 C(u^2) on |u| <= pi/4 (+ margin), absolute error ----
    Z = (pi / 4 * (1 + mpf(2) ** -9)) ** 2
    tgs = lambda z: (sin(sqrt(z)) - sqrt(z)) / sqrt(z) ** 3 if z > mpf(10) ** -12 else -mpf(1) / 6 + z / 120
    wgs = lambda z: sqrt(z) ** 3
    for ns in (5, 6):
        c, E = remez(tgs, wgs, ns, mpf(0), Z)
        cd = [mpf(float(x)) for x in c]
        e = max_err(lambda u: u + u ** 3 * sum(ci * u ** (2 * i) for i, ci in enumerate(cd)) - sin(u), -sqrt(Z), sqrt(Z))
        as_c("SIN_M%d" % ns, cd, "sin u = u + u^3 S(u^2) on |u| <= pi/4, %d coefficients: absolute error %.2e" % (ns, float(e)))
    tgc = lambda z: (cos(sqrt(z)) - 1) / z if z > mpf(10) ** -12 else -mpf(1) / 2 + z / 24
    wgc = lambda z: z
    for nc in (6, 7):
        c, E = remez(tgc, wgc, nc, mpf(0), Z)
        cd = [mpf(float(x)) for x in c]
        e = max_err(lambda u: 1 + u ** 2 * sum(ci * u ** (2 * i) for i, ci in enumerate(cd)) - cos(u), -sqrt(Z), sqrt(Z))
        as_c("COS_M%d" % nc, cd, "cos u = 1 + u^2 C(u^2) on |u| <= pi/4, %d coefficients: absolute error %.2e" % (nc, float(e)))
    # ---- value-only switch: (1 + cos a)/2 = 1/2 - y/2 + y^3 H(y^2), y = a - pi/2, |y| <= pi/2 (+ margin), absolute error ----
    Z2 = (pi / 2 * (1 + mpf(2) ** -9)) ** 2
    tgh = lambda z: -(sin(sqrt(z)) - sqrt(z)) / (2 * sqrt(z) ** 3) if z > mpf(10) ** -12 else mpf(1) / 12 - z / 240
    for nh in (6, 7):
        c, E = remez(tgh, wgs, nh, mpf(0), Z2)
        cd = [mpf(float(x)) for x in c]
        e = max_err(lambda y: mpf(1) / 2 - y / 2 + y ** 3 * sum(ci * y ** (2 * i) for i, ci in enumerate(cd)) - (1 - sin(y)) / 2, -sqrt(Z2), sqrt(Z2))
        as_c("HSW_M%d" % nh, cd, "(1 + cos a)/2 = 1/2 - y/2 + y^3 H(y^2), y = a - pi/2 in [-pi/2, pi/2], %d coefficients: absolute error %.2e" % (nh, float(e)))


if __name__ == "__main__":
    main()
