# critical-path length of the obs dependency DAG (ob j depends on k<j iff dist(k,j) < cutoff_k)
import numpy as np, time, sys
from scipy.spatial import cKDTree
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
cut = float(sys.argv[2]) if len(sys.argv) > 2 else 2000.0
rng = np.random.default_rng(0)
lat = np.arcsin(rng.uniform(-0.98, 0.98, N)); lon = rng.uniform(0, 2*np.pi, N)
u = np.stack([np.cos(lat)*np.cos(lon), np.cos(lat)*np.sin(lon), np.sin(lat)], 1)
theta = cut/6371.0; chord = 2*np.sin(theta/2)
t = cKDTree(u)
dp = np.zeros(N, dtype=np.int32); nnz = 0
t0 = time.time()
B = 2000
for b0 in range(0, N, B):
    nb = t.query_ball_point(u[b0:b0+B], chord, return_sorted=False)
    for i, lst in enumerate(nb):
        j = b0 + i
        a = np.asarray(lst); a = a[a < j]
        nnz += a.size
        dp[j] = 1 + (dp[a].max() if a.size else 0)
print('N', N, 'cutoff', cut, 'nnz', nnz, 'density', nnz/(N*N/2), 'critical path', dp.max(), 'time', time.time()-t0)
