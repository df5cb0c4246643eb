#!/usr/bin/env python
"""Generate chomp_b200/csrc/special_coeffs.cuh: polynomial coefficient tables
for the device sine/cosine integrals Si(x), Ci(x) (used by the NFW Fourier
profile, reference halo.py:561-585 via scipy.special.sici) and for the Bessel
functions J0, J2 (reference kernel.py:707-712, 834-839 via scipy.special.j0/jn).

The approximations are derived here from scratch with mpmath (40 digits):
Chebyshev interpolants converted to the monomial basis of the scaled variable
s in [-1, 1], evaluated by Horner's rule on the device.

  x <= 4      Si(x) = x * P_si(x^2)                      P on t = x^2 in [0, 16]
              Ci(x) = gamma + ln x + x^2 * P_ci(x^2)
  x  > 4      Si(x) = pi/2 - f(x) cos x - g(x) sin x
              Ci(x) = f(x) sin x - g(x) cos x
              f(x) = F_r(u)/x, g(x) = G_r(u)/x^2, u = 1/x^2, on three ranges
              r = [4, 7], [7, 14], [14, inf)

Run:  python chomp_b200/csrc/tools/gen_special.py   (rewrites the header)
"""
import os
import sys

import mpmath as mp
import numpy as np
from numpy.polynomial import chebyshev as C

mp.mp.dps = 40
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "special_coeffs.cuh")


def cheb_fit(func, lo, hi, deg):
    """Monomial coefficients (in s, x = mid + half*s) of the degree-`deg`
    Chebyshev interpolant of func on [lo, hi]."""
    k = np.arange(deg + 1)
    nodes = np.cos(np.pi*(k + 0.5)/(deg + 1))
    mid, half = (mp.mpf(hi) + lo)/2, (mp.mpf(hi) - lo)/2
    vals = [func(mid + half*mp.mpf(float(s))) for s in nodes]
    # Chebyshev coefficients in mp precision
    n = deg + 1
    coef = []
    for j in range(n):
        s = mp.mpf(0)
        for i in range(n):
            s += vals[i]*mp.cos(mp.pi*j*(i + mp.mpf(0.5))/n)
        coef.append(s*(2 if j else 1)/n)
    # convert to monomials in mp precision via recurrence T_{j+1} = 2 s T_j - T_{j-1}
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        a = [mp.mpf(0)] + [2*c for c in T[j - 1]]
        b = T[j - 2] + [mp.mpf(0)]*(len(a) - len(T[j - 2]))
        T.append([x - y for x, y in zip(a, b)])
    mono = [mp.mpf(0)]*n
    for j in range(n):
        for i, c in enumerate(T[j]):
            mono[i] += coef[j]*c
    return [float(c) for c in mono], float(mid), float(half)


def horner(c, s):
    r = np.zeros_like(s) + c[-1]
    for a in c[-2::-1]:
        r = r*s + a
    return r


def f_aux(x):
    return mp.ci(x)*mp.sin(x) + (mp.pi/2 - mp.si(x))*mp.cos(x)


def g_aux(x):
    return -mp.ci(x)*mp.cos(x) + (mp.pi/2 - mp.si(x))*mp.sin(x)


def si_small(t):
    if t == 0:
        return mp.mpf(1)
    x = mp.sqrt(t)
    return mp.si(x)/x


def ci_small(t):
    if t == 0:
        return mp.mpf(-1)/4
    x = mp.sqrt(t)
    return (mp.ci(x) - mp.euler - mp.log(x))/t


def F_of_u(u):
    if u == 0:
        return mp.mpf(1)
    x = 1/mp.sqrt(u)
    return x*f_aux(x)


def G_of_u(u):
    if u == 0:
        return mp.mpf(1)
    x = 1/mp.sqrt(u)
    return x*x*g_aux(x)


def check(name, c, mid, half, func, lo, hi, n=400):
    xs = np.linspace(lo, hi, n)
    s = (xs - mid)/half
    approx = horner(c, s)
    exact = np.array([float(func(mp.mpf(float(x)))) for x in xs])
    err = np.max(np.abs(approx - exact)/np.maximum(np.abs(exact), 1e-300))
    print("%-10s deg %2d  max rel err %.2e" % (name, len(c) - 1, err))
    return err


def emit(f, name, c):
    """Host copy with the values + an *uninitialised* __constant__ array that
    chomp_b200_create uploads: with an initialiser in sight nvcc folds the
    coefficients into immediates (two UMOVs per DFMA); left opaque they become
    constant-bank operands of the DFMA itself."""
    f.write("static const double h_%s[%d] = {\n" % (name, len(c)))
    for a in c:
        f.write("    %s,\n" % repr(a))
    f.write("};\n")
    f.write("__constant__ double %s[%d];\n" % (name, len(c)))


def main():
    tabs = {}
    SMALL = 4.0
    DEG_S = 10
    tabs["SI_SMALL"] = cheb_fit(si_small, 0.0, SMALL**2, DEG_S)
    tabs["CI_SMALL"] = cheb_fit(ci_small, 0.0, SMALL**2, DEG_S)
    check("si_small", *tabs["SI_SMALL"], si_small, 0.0, SMALL**2)
    check("ci_small", *tabs["CI_SMALL"], ci_small, 0.0, SMALL**2)
    # x <= 1: most (k, M) pairs of the halo tables live here; degree 6 is already exact to rounding
    TINY = 1.0
    DEG_T = 6
    tabs["SI_TINY"] = cheb_fit(si_small, 0.0, TINY**2, DEG_T)
    tabs["CI_TINY"] = cheb_fit(ci_small, 0.0, TINY**2, DEG_T)
    check("si_tiny", *tabs["SI_TINY"], si_small, 0.0, TINY**2)
    check("ci_tiny", *tabs["CI_TINY"], ci_small, 0.0, TINY**2)
    ranges = [(4.0, 7.0), (7.0, 14.0), (14.0, None)]
    DEG_L = 12
    fg = []
    for i, (a, b) in enumerate(ranges):
        ulo = 0.0 if b is None else 1.0/b**2
        uhi = 1.0/a**2
        F = cheb_fit(F_of_u, ulo, uhi, DEG_L)
        G = cheb_fit(G_of_u, ulo, uhi, DEG_L)
        check("F[%d]" % i, *F, F_of_u, ulo, uhi)
        check("G[%d]" % i, *G, G_of_u, ulo, uhi)
        fg.append((F, G))
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_special.py -- do not edit.\n")
        f.write("// Chebyshev-derived monomial coefficients (variable s in "
                "[-1,1]) for Si/Ci.\n#pragma once\n\n")
        f.write("#define CHOMP_SICI_SMALL_X %r\n" % SMALL)
        f.write("#define CHOMP_SICI_TINY_X %r\n#define CHOMP_SICI_DEG_T %d\n" % (TINY, DEG_T))
        f.write("#define CHOMP_SICI_DEG_S %d\n#define CHOMP_SICI_DEG_L %d\n"
                % (DEG_S, DEG_L))
        f.write("#define CHOMP_SICI_X1 %r\n#define CHOMP_SICI_X2 %r\n"
                % (ranges[1][0], ranges[2][0]))
        for nm in ("SI_SMALL", "CI_SMALL", "SI_TINY", "CI_TINY"):
            c, mid, half = tabs[nm]
            f.write("#define CHOMP_%s_MID %r\n#define CHOMP_%s_IHALF %r\n"
                    % (nm, mid, nm, 1.0/half))
            emit(f, "k_" + nm.lower(), c)
        f.write("// [range][0]=mid, [1]=1/half of u = 1/x^2\n")
        f.write("static const double h_k_sici_urange[3][2] = {\n")
        for F, G in fg:
            f.write("    {%r, %r},\n" % (F[1], 1.0/F[2]))
        f.write("};\n__constant__ double k_sici_urange[3][2];\n")
        for nm, j in (("F", 0), ("G", 1)):
            f.write("static const double h_k_sici_%s[3][%d] = {\n" % (nm, DEG_L + 1))
            for pair in fg:
                f.write("    {" + ", ".join(repr(a) for a in pair[j][0]) + "},\n")
            f.write("};\n__constant__ double k_sici_%s[3][%d];\n" % (nm, DEG_L + 1))
        f.write("""
// upload the tables to the current device (called from chomp_b200_create)
static inline cudaError_t chomp_upload_special_tables() {
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(k_si_small, h_k_si_small, sizeof h_k_si_small)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(k_ci_small, h_k_ci_small, sizeof h_k_ci_small)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(k_si_tiny, h_k_si_tiny, sizeof h_k_si_tiny)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(k_ci_tiny, h_k_ci_tiny, sizeof h_k_ci_tiny)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(k_sici_urange, h_k_sici_urange, sizeof h_k_sici_urange)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(k_sici_F, h_k_sici_F, sizeof h_k_sici_F)) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(k_sici_G, h_k_sici_G, sizeof h_k_sici_G);
}
""")
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
