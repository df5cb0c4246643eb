"""Accuracy of GL-16 on fixed 2-e-fold panels in ln x for the x < 1 part of sigma^2(R)."""
import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import chomp_oracle as O
from oracle.quadrature import Tight
from common import C_DICT
from scipy import integrate
se = O.SingleEpoch(0.5, C_DICT, O.precision(), Tight(40))
gx, gw = np.polynomial.legendre.leggauss(16)
def W2(x):
    W = np.where(x < 0.1, 1 - x*x/10 + x**4/280, 3*(np.sin(x) - x*np.cos(x))/x**3)
    return W*W
def f(lx, R):
    x = np.exp(lx); return se.delta_k(x/R)*W2(x)
for R in (0.05, 0.1, 0.5, 2.0, 8.0, 30.0, 90.0):
    x_lo = 1e-3*R
    l_lo = np.log(x_lo)
    ref = integrate.quad(f, l_lo, 0.0, args=(R,), epsabs=0, epsrel=1e-13, limit=400)[0]
    # old: 4 equal panels
    old = 0.0
    e = np.linspace(l_lo, 0, 5)
    for a, b in zip(e[:-1], e[1:]):
        h = 0.5*(b-a); old += np.sum(h*gw*f(0.5*(a+b)+h*gx, R))
    # new: fixed panels [-2(j+1), -2j] + partial
    j_lo = int(np.floor(-l_lo/2.0)); new = 0.0
    for j in range(j_lo):
        new += np.sum(gw*f(-2*j-1+gx, R))
    a, b = l_lo, -2.0*j_lo; h = 0.5*(b-a); new += np.sum(h*gw*f(0.5*(a+b)+h*gx, R))
    tot = integrate.quad(f, l_lo, np.log(100*R), args=(R,), epsabs=0, epsrel=1e-12, limit=2000)[0]
    print("R=%5.2f low/total=%.3f  old err %.1e  new err %.1e (relative to total sigma^2)" % (R, ref/tot, abs(old-ref)/tot, abs(new-ref)/tot))
print("--- lower orders on the fixed lattice (full panels) / partial panel")
for nq in (8, 10, 12):
    x8, w8 = np.polynomial.legendre.leggauss(nq)
    for R in (0.05, 0.5, 2.0, 30.0):
        l_lo = np.log(1e-3*R)
        ref = integrate.quad(f, l_lo, 0.0, args=(R,), epsabs=0, epsrel=1e-13, limit=400)[0]
        tot = integrate.quad(f, l_lo, np.log(100*R), args=(R,), epsabs=0, epsrel=1e-12, limit=2000)[0]
        j_lo = int(np.floor(-l_lo/2.0)); new = 0.0
        for j in range(j_lo): new += np.sum(w8*f(-2*j-1+x8, R))
        a, b = l_lo, -2.0*j_lo; h = 0.5*(b-a); new += np.sum(h*w8*f(0.5*(a+b)+h*x8, R))
        print("nq=%d R=%5.2f err %.1e" % (nq, R, abs(new-ref)/tot))
