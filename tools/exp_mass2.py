import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from chomp_b200 import _lib, design, engine
import bench
B=256
cosmo, halo, hod = design.synthetic_batch(B)
s = bench.make_survey(); eng = engine.Engine(s); eng.reserve(B)
z = np.full(B, 0.5)
eng.mass_tables(cosmo, halo, z); torch.cuda.synchronize()
ep = eng.table(_lib.T_EPOCH, B).cpu().numpy()
w = ep[:, _lib.EPOCH_FIELDS.index("walk_steps")]
print("walk steps: min %d max %d mean %.1f" % (w.min(), w.max(), w.mean()))
print("ln_mass_min range", ep[:,6].min(), ep[:,6].max(), "ln_mass_max", ep[:,7].min(), ep[:,7].max())
import ctypes
lib = _lib.load()
