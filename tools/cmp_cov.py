import json, numpy as np
g = json.load(open('/root/repo/tests/golden/reference_covariance.json'))
t = np.load('/root/repo/tools/cov_tight16.npz')
c = g['cases']['power_gggg']
n = len(c['bins_center'])
K = np.array(c['kernel_NG_table']).reshape(50,50)
print('zbarNG', c['z_bar_NG'], c['D_z_NG'], 'Kmin', c['kernel_NG_min'])
print('K rel-to-scale err', np.max(np.abs(K - t['K']))/np.max(np.abs(K)))
print('proj err', np.max(np.abs(np.array(c['projected_a'])-t['proj'])/np.max(np.abs(t['proj']))))
for nm in ('cov_P','cov_G','cov_NG'):
    R = np.array(c[nm]).reshape(n,n); O = t[nm[4:]]
    d = np.sqrt(np.abs(np.outer(np.diag(R), np.diag(R))))
    print(nm, 'max err rel diag-geomean', np.max(np.abs(R-O)/np.where(d>0,d,1)), 'diag', np.diag(R)[:4])
tt = np.array(c['tri_table']).reshape(50,50)
print('tri', np.max(np.abs(tt - t['tri'])/np.abs(tt)))
