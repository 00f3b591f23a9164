import sys, json, numpy as np, torch, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ssrs_b200 import movmodel as mm
class A: pass
a = A(); a.rows=5000; a.cols=6000; a.resolution=10.0; a.tracks_per_gpu=100000; a.seed=2021; a.no_solve=False
sr, sc = bench.start_cells(a, 100000)
up, pot, info = bench.build_fields_gpu(a, torch)
f = mm.interleave_fields(up, pot)
res = mm.simulate_tracks_batch(0.0, sr, sc, (5000,6000), fields=f, seed=2021)
L = res.traj_len.cpu().numpy().astype(np.int64)-1
print('steps total', L.sum(), 'mean', L.mean(), 'pcts', {p: int(np.percentile(L,p)) for p in (1,10,50,90,99,99.9,99.99,100)})
print('n > 50k', (L>50000).sum(), 'n>100k', (L>100000).sum(), 'steps in >50k tracks', L[L>50000].sum())
np.save('gpurun_out/lens.npy', L)
# potential plateau stats
p = pot
d = (p[1:,:]-p[:-1,:])
print('frac zero vertical diffs', float((d==0).float().mean()), 'frac positive (uphill north)', float((d>0).float().mean()))
