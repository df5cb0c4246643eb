import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/scratch')
import numpy as np
from oracle.quadrature import Tight, gl_nodes, gl_panels
from oracle import chomp_oracle as O
from common import C_DICT, H_DICT, HOD_DICT, oracle_covariance
for n_halo in (50, 200):
    cov = oracle_covariance(C_DICT, H_DICT, HOD_DICT, theta_deg=(0.01, 1.0), tri_z=0.5, integ=Tight(16),
                            prec=O.precision(halo_npoints=n_halo))
    K = np.load('/root/repo/tools/cov_tight16.npz')['K']; cov.kernel.set_table(K)
    if n_halo == 50: cov.tri._i04 = np.load('/root/repo/tools/cov_tight16.npz')['tri']
    else:
        # smooth stand-in: interpolate the 50-node table (log-space bicubic) to 200 nodes
        from scipy.interpolate import RectBivariateSpline
        t50 = np.load('/root/repo/tools/cov_tight16.npz')['tri']; x50 = np.linspace(np.log(1e-3), np.log(100.), 50)
        sp = RectBivariateSpline(x50, x50, np.log(t50)); x200 = cov.tri.ln_k_nodes
        cov.tri._i04 = np.exp(sp(x200, x200))
    bins = cov.bins[:, 2]
    def ng_with_rule(ta, tb, width, order):
        npz = int(np.ceil((cov.ln_k_max - cov.ln_k_min)/width - 1e-9))
        x, w = gl_panels(np.linspace(cov.ln_k_min, cov.ln_k_max, npz + 1), order)
        I = np.array([np.sum(w*cov._kb_integrand(x, lka, ta, tb)) for lka in cov.ln_k_array])/cov.D_z_NG**4
        sp = O._spline(cov.ln_k_array, I)
        xx, ww = gl_panels(cov.ln_k_array, 8)
        return np.sum(ww*np.exp(2*xx)*sp(xx))/(4*np.pi**2*cov.area)
    pairs = [(0, 0), (0, 5), (3, 9), (9, 9), (5, 6)]
    ref = {p: cov.covariance_NG(bins[p[0]], bins[p[1]]) for p in pairs}
    diag = {i: cov.covariance_NG(bins[i], bins[i]) for i in range(10)}
    for width, order in ((0.0625, 4), (0.125, 4), (0.25, 4), (0.25, 6), (0.5, 6), (0.5, 8), (1.0, 8)):
        errs = [abs(ng_with_rule(bins[a], bins[b], width, order) - ref[(a, b)])/np.sqrt(abs(diag[a]*diag[b])) for a, b in pairs]
        print(n_halo, width, order, ' '.join('%.1e' % e for e in errs))
