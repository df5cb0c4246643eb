import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from chomp_b200 import _lib, defaults, design, engine
from common import cov_err, w_err
B = 8
cosmo, halo, hod = design.synthetic_batch(B, seed=11)
theta = (0.001, 1.0)
res = {}
for n_halo in (50, 200):
    for nq in (4, 3, 2):
        prec = dict(defaults.default_precision, halo_npoints=n_halo)
        quad = dict(defaults.default_quadrature, hankel=nq)
        survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), theta_deg=theta, bins_per_decade=10.0,
                               power_spec="power_gg", precision=prec, quadrature=quad)
        cfg = survey.config(); cfg.tri_moment = 4
        eng = engine.Engine(cfg)
        setup = engine.CovarianceSetup(survey, theta, 10.0, 25.0, [1e10, 1e10], [1e10, 1e10], 1.0, True, "power_gg")
        cov = eng.covariance(cosmo, halo, hod, setup).cpu().numpy()
        w = eng.wtheta(cosmo, halo, hod, survey.theta, _lib.P_GG).cpu().numpy()
        res[(n_halo, nq)] = (cov, w)
    for nq in (3, 2):
        ce = max(cov_err(res[(n_halo, nq)][0][i], res[(n_halo, 4)][0][i]) for i in range(B))
        we = max(w_err(res[(n_halo, nq)][1][i], res[(n_halo, 4)][1][i]) for i in range(B))
        print("n_halo=%d nq_hankel=%d vs 4: cov err %.2e  w err %.2e" % (n_halo, nq, ce, we))
