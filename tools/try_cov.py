import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import chomp_oracle as O
from oracle import covariance_oracle as CO
from oracle.quadrature import Tight, Romberg
from common import C_DICT, H_DICT, HOD_DICT, D2R

def build(integ, prec=None, tri_spec="power_gggg", cov_spec="power_gg", tri_z=0.5):
    prec = prec or O.precision()
    cm = O.MultiEpoch(0.0, 5.0, C_DICT, prec, integ)
    dist = O.dNdzGaussian(0.0, 2.0, 0.5, 0.1, prec=prec, integ=integ)
    wa = O.WindowFunctionGalaxy(dist, cm); wb = O.WindowFunctionGalaxy(dist, cm)
    kern = O.Kernel(1e-6*D2R, 100*D2R, wa, wb, cm)
    def factory(z, cls=O.Halo, **kw):
        se = O.SingleEpoch(z, C_DICT, prec, integ)
        mf = O.MassFunction(se, H_DICT)
        return cls(se, mf, O.HODZheng(HOD_DICT, prec["halo_precision"]), H_DICT, **kw)
    corr = O.Correlation(0.01, 1.0, kern, factory, "power_gg", bins_per_decade=5.0)
    tri = factory(tri_z, O.HaloTrispectrumOneHalo, power_spec=tri_spec)
    cov = CO.Covariance(corr, (0.01, 1.0), 5.0, 25.0, [1e10,1e10], [1e10,1e10], 1.0, True, tri, cov_spec)
    return cov

if __name__ == "__main__":
    t0=time.time()
    cov = build(Tight(16))
    print("bins", cov.bins[:, 2], "zbarNG", cov.kernel.z_bar_NG, cov.D_z_NG)
    T = cov.kernel.table(); print("K_NG table %.1fs"%(time.time()-t0), T.min(), T.max(), T[0,0], T[-1,-1], T[0,-1])
    cov.tri.table_i04(); print("tri %.1fs"%(time.time()-t0))
    total,P,G,NG = cov.get_covariance(parts=True); print("cov %.1fs"%(time.time()-t0))
    np.set_printoptions(linewidth=200, precision=4)
    print(np.diag(P)); print(np.diag(G)); print(np.diag(NG)); print(G[0], NG[0])
    np.savez('/root/repo/tools/cov_tight16.npz', K=T, P=P, G=G, NG=NG, proj=cov.proj_nodes, tri=cov.tri._i04)
