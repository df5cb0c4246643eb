import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from chomp_b200 import _lib, design, engine, defaults
import bench
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
B=4096
cosmo, halo, hod = design.synthetic_batch(B)
for name, lim in (("walk", None), ("fixed", dict(defaults.default_limits, mass_min=2e8, mass_max=1.5e16))):
    s = bench.make_survey()
    if lim: s.limits = lim
    eng = engine.Engine(s); eng.reserve(B)
    dc, dh = eng._dev(cosmo), eng._dev(halo)
    z = eng._dev(np.full(B, 0.5))
    print(name, "mass_tables ms", round(t(lambda: eng.mass_tables(dc, dh, z)),3))
