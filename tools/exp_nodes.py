import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from chomp_b200 import _lib, design, engine
import bench
B=1024
cosmo, halo, hod = design.synthetic_batch(B)
s = bench.make_survey(); eng = engine.Engine(s); eng.reserve(B)
w = eng.wtheta(cosmo, halo, hod, s.theta, 3); torch.cuda.synchronize()
nn = eng.table(_lib.T_NU_QUAD_COUNT, B).cpu().numpy()
ep = eng.table(_lib.T_EPOCH, B).cpu().numpy()
F = _lib.EPOCH_FIELDS
rv = np.cbrt(3*np.exp(ep[:,F.index("ln_mass_max")])/(4*np.pi*ep[:,F.index("delta_v")]*ep[:,F.index("rho_bar")]))
print("nodes per class mean", nn.mean(0), "max", nn.max(0))
print("rv_max mean %.3f min %.3f max %.3f" % (rv.mean(), rv.min(), rv.max()))
nk=200; l0=np.log(1e-3); hk=(np.log(100)-l0)/(nk-1)
i1 = np.clip(np.ceil((np.log(45/rv)-l0)/hk),0,nk); i2=np.clip(np.ceil((np.log(180/rv)-l0)/hk),0,nk)
print("k per class mean", i1.mean(), (i2-i1).mean(), (nk-i2).mean())
work = i1*nn[:,0] + (i2-i1)*nn[:,1] + (nk-i2)*nn[:,2]
print("node-k pairs per point mean", work.mean(), " (uniform 424 nodes would be", 424*200, ")")
print("zbar mean", ep[:,0].mean())
