import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from chomp_b200 import _lib, defaults, design, engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_halo = int(sys.argv[2]) if len(sys.argv) > 2 else 200
bpd = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
theta = (0.001, 1.0)
prec = dict(defaults.default_precision, halo_npoints=n_halo)
survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), theta_deg=theta, bins_per_decade=bpd,
                       power_spec="power_gg", precision=prec)
cfg = survey.config(); cfg.tri_moment = 4
eng = engine.Engine(cfg)
setup = engine.CovarianceSetup(survey, theta, bpd, 25.0, [1e10, 1e10], [1e10, 1e10], 1.0, True, "power_gg")
cosmo, halo, hod = design.synthetic_batch(B)
c, h, g = eng._dev(cosmo), eng._dev(halo), eng._dev(hod)
status = torch.zeros(B, dtype=torch.int32, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.covariance(c, h, g, setup, status=status); e1.record(); torch.cuda.synchronize()
    print("B=%d n_halo=%d bins=%d: %.1f ms  (%.0f cov-points/s) status!=0: %d finite: %s" % (
        B, n_halo, setup.bins.shape[0], e0.elapsed_time(e1), B/e0.elapsed_time(e1)*1e3, int((status != 0).sum()),
        bool(torch.isfinite(out).all())))
