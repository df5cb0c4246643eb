"""Explore degrees needed for range-wise polynomial fits of the Si/Ci auxiliary functions."""
import mpmath as mp, numpy as np, sys
sys.path.insert(0, 'chomp_b200/csrc/tools')
mp.mp.dps = 40
def f_aux(x): return mp.ci(x)*mp.sin(x) + (mp.pi/2 - mp.si(x))*mp.cos(x)
def g_aux(x): return -mp.ci(x)*mp.cos(x) + (mp.pi/2 - mp.si(x))*mp.sin(x)
def F_of_u(u):
    if u == 0: return mp.mpf(1)
    x = 1/mp.sqrt(u); return x*f_aux(x)
def G_of_u(u):
    if u == 0: return mp.mpf(1)
    x = 1/mp.sqrt(u); return x*x*g_aux(x)
def H_small(z):
    if z == 0: return -mp.euler
    return g_aux(z) + mp.log(z)*mp.cos(z)
def cheb_err(func, lo, hi, deg, n=200, rel=True):
    k = np.arange(deg+1); nodes = np.cos(np.pi*(k+0.5)/(deg+1))
    mid, half = (mp.mpf(hi)+lo)/2, (mp.mpf(hi)-lo)/2
    vals = np.array([float(func(mid+half*mp.mpf(float(s)))) for s in nodes])
    c = np.polynomial.chebyshev.chebfit(nodes, vals, deg)
    xs = np.linspace(-1, 1, n)
    ex = np.array([float(func(mid+half*mp.mpf(float(s)))) for s in xs])
    ap = np.polynomial.chebyshev.chebval(xs, c)
    return np.max(np.abs(ap-ex)/(np.abs(ex) if rel else 1.0))
print("large-x ranges, u = 1/x^2")
for (a, b) in ((2, 2.8284), (2.8284, 4), (4, 5.657), (5.657, 8), (8, 16), (16, None), (4, 8), (8, None), (2,4)):
    ulo = 0.0 if b is None else 1.0/b**2; uhi = 1.0/a**2
    for deg in (8, 10, 12, 14):
        print("x in [%s,%s] deg %d  F %.1e  G %.1e" % (a, b, deg, cheb_err(F_of_u, ulo, uhi, deg), cheb_err(G_of_u, ulo, uhi, deg)))
print("small z: H(z) = g + ln z cos z, and cos z, variable z")
for (a, b) in ((0, 0.5), (0.5, 1), (1, 2), (0, 1), (0, 2)):
    for deg in (8, 10, 12, 14):
        print("z in [%s,%s] deg %d  H %.1e  cos %.1e" % (a, b, deg, cheb_err(H_small, a, b, deg, rel=False), cheb_err(mp.cos, a, b, deg, rel=False)))
