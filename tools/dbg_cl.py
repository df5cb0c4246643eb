import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, json
from test_oracle_extra import cl_oracle
from oracle.quadrature import Tight
G = json.load(open('/root/repo/tests/golden/reference_extra.json'))
ell = np.array(G['cl_tables']['ell'])
cf = cl_oracle('power_gg', False, Tight(24))
print(ell[13:], cf.kernel.chi_min, cf.kernel.chi_max, cf.kernel.z_bar)
for l in ell[13:]:
    chi = np.linspace(max(l/100.0, 1.0), cf.kernel.chi_max, 7)
    print(l, l/100.0, cf._integrand(chi, l), cf.correlation(l))
print(G['cl_tables']['power_gg'][12:])
print('---- mm')
cf = cl_oracle('power_mm', False, Tight(24))
for l in ell[13:15]:
    chi = np.linspace(max(l/100.0, 1.0), cf.kernel.chi_max, 7)
    print(l, cf._integrand(chi, l), cf.correlation(l), cf.halo.power('power_mm', l/chi))
