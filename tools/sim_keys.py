"""Distribution of Si/Ci regimes over (k, node) for a typical point (CPU simulation)."""
import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import chomp_oracle as O
from oracle.quadrature import Tight
from common import C_DICT, H_DICT, HOD_DICT
prec = O.precision(); integ = Tight(40)
z = 0.5
se = O.SingleEpoch(z, C_DICT, prec, integ)
mf = O.MassFunction(se, H_DICT)
nu = mf.nu_nodes; lnm = mf.ln_mass_nodes
print("logM range", lnm[0]/np.log(10), lnm[-1]/np.log(10), "nu", nu[0], nu[-1])
h = O.Halo(se, mf, O.HODZheng(HOD_DICT, prec["halo_precision"]), H_DICT)
gx = {n: np.polynomial.legendre.leggauss(n) for n in (4, 8, 16)}
edges = np.log(np.concatenate([[1.001*nu[0]], nu[1:-1], [0.999*nu[-1]]]))
def nodes(order):
    x, w = gx[order]
    a, b = edges[:-1], edges[1:]
    return (0.5*(a+b)[:, None] + 0.5*(b-a)[:, None]*x[None, :]).ravel()
nk = 200
lnk = np.linspace(np.log(1e-3), np.log(1e2), nk); k = np.exp(lnk)
M_max = np.exp(lnm[-1]); rv_max = h.virial_radius(M_max)
phi = k*rv_max
cls = np.where(phi < 45, 0, np.where(phi < 180, 1, 2))
print("rv_max", rv_max, "k classes", [(cls == c).sum() for c in range(3)])
tot = 0; hist = {}; nonuni = 0; warp_it = 0
def key(z1, z2):
    r = lambda x: np.where(x >= 14, 2, np.where(x >= 7, 1, 0))
    return np.where(z2 <= 1, 13, np.where(z2 <= 4, 0, np.where(z1 <= 4, 1 + r(z2), 4 + 3*r(z1) + r(z2))))
zc_hist = np.zeros(6)
for c, order in enumerate((4, 8, 16)):
    x = nodes(order); M = np.exp(mf.ln_mass(np.exp(x)))
    con = h.concentration(M); rs = h.virial_radius(M)/con; cp = 1 + con
    pad = (-len(x)) % 32
    rs = np.concatenate([rs, np.full(pad, rs[-1])]); cp = np.concatenate([cp, np.full(pad, cp[-1])])
    for kk in k[cls == c]:
        z1 = kk*rs; z2 = z1*cp
        ky = key(z1, z2).reshape(-1, 32)
        uni = (ky == ky[:, :1]).all(axis=1)
        warp_it += len(uni); nonuni += (~uni).sum()
        for v, u in zip(ky[:, 0], uni):
            kname = int(v) if u else -1
            hist[kname] = hist.get(kname, 0) + 1
        zc = (z2 - z1)
        zc_hist += np.array([(zc <= 1).sum(), ((zc > 1) & (zc <= 2)).sum(), ((zc > 2) & (zc <= 3)).sum(), ((zc > 3) & (zc <= 4)).sum(), ((zc > 4) & (zc <= 30)).sum(), (zc > 30).sum()])
    print("class", c, "nodes", len(x), "con range", con.min(), con.max(), "rs", rs.min(), rs.max())
print("warp iterations", warp_it, "non-uniform", nonuni)
for kname in sorted(hist): print(kname, hist[kname], "%.1f%%" % (100*hist[kname]/warp_it))
print("zc hist (<=1,2,3,4,30,inf) %", 100*zc_hist/zc_hist.sum())
