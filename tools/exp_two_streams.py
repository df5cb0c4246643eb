import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from chomp_b200 import _lib, design, engine
survey = bench.make_survey()
B = 4096
cosmo, halo, hod = design.synthetic_batch(B)
dev = torch.device("cuda", 0)
theta = torch.as_tensor(survey.theta, device=dev)
which = _lib.POWER_SPEC[survey.power_spec]
def run(nsplit):
    engs = [engine.Engine(survey, device=0) for _ in range(nsplit)]
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    n = B//nsplit
    parts = [tuple(torch.as_tensor(a[i*n:(i+1)*n], device=dev) for a in (cosmo, halo, hod)) for i in range(nsplit)]
    outs = [torch.empty((n, theta.numel()), dtype=torch.float64, device=dev) for _ in range(nsplit)]
    for e in engs: e.reserve(n)
    def step():
        cur = torch.cuda.current_stream()
        for s in streams: s.wait_stream(cur)
        for e, s, p, o in zip(engs, streams, parts, outs):
            with torch.cuda.stream(s):
                e.wtheta(p[0], p[1], p[2], theta, which, out=o)
        for s in streams: cur.wait_stream(s)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print("splits %d: %.3f ms/step  %.0f points/s" % (nsplit, ms, B/ms*1e3))
    return torch.cat(outs).cpu().numpy()
w1 = run(1); w2 = run(2); w4 = run(4)
print("identical:", np.array_equal(w1, w2), np.array_equal(w1, w4))
