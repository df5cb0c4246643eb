"""Where does a synchronised step spend its time?  Engine.wtheta vs ShardedEngine.wtheta, with and without
the NVML sampler thread of bench.py."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from chomp_b200 import _lib, design, distributed
survey = bench.make_survey()
B = 4096
cosmo, halo, hod = design.synthetic_batch(B)
dev = torch.device("cuda", 0)
sh = distributed.ShardedEngine(survey, device=0)
eng = sh.engine
theta = torch.as_tensor(survey.theta, device=dev)
which = _lib.POWER_SPEC[survey.power_spec]
d = [torch.as_tensor(a, device=dev) for a in (cosmo, halo, hod)]
out = torch.empty((B, theta.numel()), dtype=torch.float64, device=dev)
status = torch.zeros(B, dtype=torch.int32, device=dev)
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
def timed(fn, n=10, do_flush=True):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ms, host = [], []
    for s in range(n):
        if do_flush: flush.fill_(s & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter(); fn(); host.append(time.perf_counter() - t0); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return "%.3f ms (event)  %.3f ms (host call)" % (np.mean(ms), 1e3*np.mean(host))
direct = lambda: eng.wtheta(d[0], d[1], d[2], theta, which, out=out, status=status)
front = lambda: sh.wtheta(d[0], d[1], d[2])
for timing in (False, True):
    eng.set_timing(timing)
    print("timing marks", timing, "| direct:", timed(direct), "| front:", timed(front), "| direct no flush:", timed(direct, do_flush=False))
smp = bench.ClockSampler(0); smp.start(); time.sleep(0.2)
print("with sampler | direct:", timed(direct), "| front:", timed(front))
smp.stop_flag = True; smp.join()
