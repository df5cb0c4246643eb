import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/scratch')
import numpy as np
from scipy import special
from oracle.quadrature import Tight, gl_nodes
from try_cov import build
cov = build(Tight(16)); kc = cov.kernel
ref = np.load('/root/repo/tools/cov_tight16.npz')['K']
x = kc.ln_ktheta_nodes
edges = np.unique(np.concatenate([kc.windows[0].chi_nodes, kc.cosmo.chi_nodes, [kc.chi_min, kc.chi_max]]))
edges = edges[(edges >= kc.chi_min) & (edges <= kc.chi_max)]
def rule(i, j, nq, phase):
    kti, ktj = np.exp(x[i]), np.exp(x[j])
    top = min(kc.chi_max, kc.j0_limit/min(kti, ktj))
    gx, gw = gl_nodes(nq)
    tot = 0.0
    for a, b in zip(edges[:-1], edges[1:]):
        ns = max(1, int(np.ceil(max(kti, ktj)*(b-a)/phase)))
        d = (b-a)/ns
        lo = a + d*np.arange(ns); hi = lo + d
        full = hi <= top
        part = (lo < top) & ~full
        for l, h in list(zip(lo[full], hi[full])) + [(l, top) for l in lo[part]]:
            c = 0.5*(l+h) + 0.5*(h-l)*gx
            tot += 0.5*(h-l)*np.sum(gw*kc.weight_ng(c)*special.j0(kti*c)*special.j0(ktj*c))
    return tot
scale = np.max(np.abs(ref))
for nq, phase in ((4, 2.0), (4, 1.0), (6, 2.0), (6, 3.0), (8, 4.0), (8, 3.0)):
    errs = []
    for i, j in ((49, 49), (49, 40), (49, 10), (45, 45), (42, 30), (38, 38)):
        errs.append(abs(rule(i, j, nq, phase) - ref[i, j])/scale)
    print(nq, phase, ' '.join('%.1e' % e for e in errs))
print('ref entries rel scale', [ref[i,j]/scale for i,j in ((49,49),(49,40),(49,10),(45,45),(42,30),(38,38))])
print("---- cheaper rules, mid entries")
imin = np.unravel_index(np.argmin(ref), ref.shape); print('argmin', imin, x[imin[0]])
pairs = ((49, 49), (49, 10), (42, 30), (38, 38), (35, 35), (35, 20), (32, 32), (int(imin[0]), int(imin[1])), (30, 5))
for nq, phase in ((4, 4.0), (4, 6.0), (3, 2.0), (3, 3.0), (2, 1.0), (5, 6.0), (6, 8.0)):
    errs = [abs(rule(i, j, nq, phase) - ref[i, j])/scale for i, j in pairs]
    print(nq, phase, ' '.join('%.1e' % e for e in errs))
print([ '%.2e' % (ref[i,j]/scale) for i,j in pairs])
