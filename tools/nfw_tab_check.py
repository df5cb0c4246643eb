"""Emulate nfw_rho_tab with the generated tables and compare with scipy's sici formula."""
import re, numpy as np
from scipy import special
src = open('chomp_b200/csrc/nfw_coeffs.cuh').read()
def table(name):
    body = src[src.index('h_k_nfw_%s' % name):]
    body = body[body.index('= {')+3:body.index('};')]
    v = np.array([float(x) for x in re.findall(r'[-+]?(?:\d+\.\d*|\d+)(?:e[-+]?\d+)?', body)])
    return v.reshape(14, 8, 2)
A, B = table('A'), table('B')
def rl(x2):
    e = int(np.floor(np.log2(x2))) - 2
    return e if e < 4 else (4 if e < 6 else 5)
def horner(T, r, s):
    p = T[12, r, 0]; q = T[12, r, 1]
    for j in range(11, -1, -1):
        p = p*s + T[j, r, 0]; q = q*s + T[j, r, 1]
    return p, q
def tab(z, c):
    cp = 1+c; z2 = cp*z
    assert z2 >= 2
    iz2 = 1/z2; iz = iz2*cp; u2 = iz2*iz2
    sc, cc = np.sin(z2-z), np.cos(z2-z)
    ra = rl(z2*z2); mid, ih = A[13, ra]
    ft, g2 = horner(A, ra, (u2-mid)*ih)
    small = z < 2; u1 = iz*iz
    rb = (0 if z < 1 else 1) if small else 2 + rl(z*z)
    mid, ih = B[13, rb]
    p1, p2 = horner(B, rb, ((z if small else u1)-mid)*ih)
    g1 = p1 - np.log(z)*p2 if small else u1*p1
    return g1 - u2*(g2*cc - ft*iz2*sc)
def exact(z, c):
    cp = 1+c
    si1, ci1 = special.sici(z); si2, ci2 = special.sici(cp*z)
    return np.cos(z)*(ci2-ci1) + np.sin(z)*(si2-si1) - np.sin(c*z)/(cp*z)
rng = np.random.default_rng(1)
worst = 0
for _ in range(20000):
    c = 10**rng.uniform(-0.3, 1.6); z2 = 10**rng.uniform(np.log10(2.0), 3.2); z = z2/(1+c)
    a, e = tab(z, c), exact(z, c)
    # error relative to the scale of the profile at this z (1/z^2 decay for large z)
    err = abs(a-e)/max(abs(e), 1e-3*min(1.0, 1/z**2))
    if err > worst: worst = err; wz = (z, c, a, e)
print("worst rel err", worst, wz)
import mpmath as mp
mp.mp.dps = 30
def exact_mp(z, c):
    z = mp.mpf(z); c = mp.mpf(c); cp = 1+c
    return float(mp.cos(z)*(mp.ci(cp*z)-mp.ci(z)) + mp.sin(z)*(mp.si(cp*z)-mp.si(z)) - mp.sin(c*z)/(cp*z))
worst = 0
for _ in range(3000):
    c = 10**rng.uniform(-0.3, 1.6); z2 = 10**rng.uniform(np.log10(2.0), 3.2); z = z2/(1+c)
    a, e = tab(z, c), exact_mp(z, c)
    err = abs(a-e)/abs(e)
    if err > worst: worst = err; wz = (z, c, a, e)
print("vs mpmath: worst rel err", worst, wz)
