"""Per-panel Gauss-Legendre orders from the local phase k * r_vir(panel top): error of the five mass
integrals at high k against a converged rule, for a few parameter points."""
import numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import chomp_oracle as O
from oracle.quadrature import Tight
from common import C_DICT, H_DICT, HOD_DICT, HOD_DICT_2
GL = {n: np.polynomial.legendre.leggauss(n) for n in (4, 6, 8, 10, 12, 16, 24, 32)}
def run(cd, hd, gd, z, thr1, thr2, k_classes=((45., 4), (180., 8), (1e9, 16)), verbose=False):
    prec = O.precision(halo_npoints=200); integ = Tight(40)
    se = O.SingleEpoch(z, cd, prec, integ); mf = O.MassFunction(se, hd)
    h = O.Halo(se, mf, O.HODZheng(gd, prec["halo_precision"]), hd)
    nu = mf.nu_nodes
    lo_all = np.log(mf.nu_min); hi = np.log(mf.nu_max)
    worst = {}
    for name in ("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg"):
        moment = None
        if name in ("h_m", "pp_mm"): lo = lo_all
        elif name in ("h_g", "pp_gm"):
            lo = np.log(h._lower_limit(h.hod.first_moment_zero)); moment = h.hod.first_moment if name == "pp_gm" else None
        else:
            lo = np.log(h._lower_limit(h.hod.second_moment_zero)); moment = h.hod.second_moment
        br, sg = h._panel_hints(lo, hi, moment)
        edges = np.unique(np.concatenate([[lo, hi], np.log(nu[1:-1]), np.atleast_1d(br)]))
        edges = edges[(edges >= lo) & (edges <= hi)]
        rv_top = h.virial_radius(np.exp(mf.ln_mass(np.exp(edges[1:]))))
        rv_max = h.virial_radius(np.exp(mf.ln_mass_nodes[-1]))
        lnk = np.linspace(np.log(1e-3), np.log(1e2), 200)
        for lk in lnk[150:]:
            k = np.exp(lk)
            # class maximum k (the list is shared by all k of the class)
            phi = k*rv_max
            kmax_cls = [c for c in k_classes if phi < c[0]][0]
            k_cls_top = min(kmax_cls[0]/rv_max, 100.0)
            def quad(order_of_panel):
                tot = 0.0
                for a, b, o in zip(edges[:-1], edges[1:], order_of_panel):
                    x, w = GL[o]; hlf = 0.5*(b-a)
                    tot += np.sum(hlf*w*h._integrand(name, 0.5*(a+b)+hlf*x, lk))
                return tot
            ref = quad([32]*len(rv_top))
            cur = quad([kmax_cls[1]]*len(rv_top))
            phi_p = k_cls_top*rv_top
            new = quad([min(kmax_cls[1], 4 if p < thr1 else (8 if p < thr2 else 16)) for p in phi_p])
            nodes_new = sum(min(kmax_cls[1], 4 if p < thr1 else (8 if p < thr2 else 16)) for p in phi_p)
            e1, e2 = abs(cur/ref-1), abs(new/ref-1)
            w = worst.setdefault(name, [0, 0, 0, 0])
            w[0] = max(w[0], e1); w[1] = max(w[1], e2); w[2] = kmax_cls[1]*len(rv_top); w[3] = max(w[3], nodes_new)
    return worst
from common import C_DICT_2, H_DICT_2
for label, args in (("base z0.5", (C_DICT, H_DICT, HOD_DICT, 0.5)), ("cosmo2/halo2 z0.0", (C_DICT_2, H_DICT_2, HOD_DICT, 0.0)), ("hod2 z1.0", (C_DICT, H_DICT, HOD_DICT_2, 1.0))):
    for thr1, thr2 in ((20., 90.),):
        w = run(*args, thr1, thr2)
        print(label, "thr", thr1, thr2, {k: ("cur %.1e new %.1e nodes<= %d" % (v[0], v[1], v[3])) for k, v in w.items() if k in ("h_m", "pp_mm", "pp_gg")})
