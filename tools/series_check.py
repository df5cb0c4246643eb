"""Check the small-argument series of the NFW profile numerator and its coefficient recurrence."""
import numpy as np
from scipy import special, integrate
from math import factorial
D = 7
def coeffs(c, D=D):
    ic = 1.0/c
    l = np.log1p(c); j = c/(1+c)
    a = []
    for m in range(1, 2*D+2):
        jn = (l - j)*ic
        l = 1.0/m - l*ic
        j = jn
        if m & 1:
            n = (m-1)//2
            a.append((-1)**n*j*c/factorial(2*n+1))
    return np.array(a)
def exact(z, c):
    cp = 1+c
    si1, ci1 = special.sici(z); si2, ci2 = special.sici(cp*z)
    return np.cos(z)*(ci2-ci1) + np.sin(z)*(si2-si1) - np.sin(c*z)/(cp*z)
worst = 0
for c in (1.0, 1.3, 2.0, 3.2, 5.0, 9.0, 17.0, 30.0, 80.0):
    a = coeffs(c)
    # direct quadrature of J_m for a check of the coefficients
    for n in (0, 3, 7):
        J = integrate.quad(lambda x: x**(2*n+1)/(1+x)**2, 0, c, epsabs=0, epsrel=1e-13)[0]
        ref = (-1)**n*J/(factorial(2*n+1)*c**(2*n))
        worst = max(worst, abs(a[n]/ref-1))
    for zc in (1e-3, 0.1, 0.5, 0.9, 1.0):
        z = zc/c
        t = zc*zc
        s = np.polyval(a[::-1], t)
        e = exact(z, c)
        print("c=%5.1f zc=%.3f series=%.16e exact=%.16e rel=%.2e" % (c, zc, s, e, s/e-1))
print("worst coefficient error", worst)
print("---- larger X")
for X, DD in ((1.5, 8), (2.0, 9), (2.0, 10), (3.0, 12), (4.0, 14)):
    w = 0
    for c in (1.0, 1.7, 3.2, 9.0, 30.0):
        a = coeffs(c, DD)
        for zc in np.linspace(0.05, X, 40):
            s = np.polyval(a[::-1], zc*zc); e = exact(zc/c, c)
            w = max(w, abs(s/e-1))
    print(X, DD, "max rel err %.2e" % w)
import mpmath as mp
mp.mp.dps = 30
def exact_mp(z, c):
    z = mp.mpf(z); c = mp.mpf(c); cp = 1+c
    return float(mp.cos(z)*(mp.ci(cp*z)-mp.ci(z)) + mp.sin(z)*(mp.si(cp*z)-mp.si(z)) - mp.sin(c*z)/(cp*z))
print("---- small concentrations, X = 3, D = 11 (forward recurrence below c = 1)")
for c in (0.2, 0.3, 0.4, 0.5, 0.7, 0.9):
    a = coeffs(c, 11)
    w = 0
    for zc in np.linspace(0.05, 3.0, 60):
        s = np.polyval(a[::-1], zc*zc); e = exact_mp(zc/c, c) if 'exact_mp' in globals() else exact(zc/c, c)
        w = max(w, abs(s/e-1))
    print("c=%.1f max rel err %.2e" % (c, w))
