"""Workload for the ncu captures of round 2 (profiles/r02_*): configs[1] engine after a burn-in, then one call of
every kernel family -- rows, columns (aux / stats / scan / install), hypers, logpdf and simulate batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
cfg = bench.C1
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
X, cct, da = bench.make_table(cfg['rows'], cfg['cols'], cfg['views'], cfg['clusters'])
Zv, Zrv = bench.initial_partitions(cfg['rows'], cfg['cols'], 12345)
eng = Engine(X, num_states=S, rng=gen_rng(1), cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv},
             stream='philox', Kcap=cfg['kcap'], Vcap=cfg['vcap'])
eng.transition(N=burn, kernels=['rows', 'columns'], progress=False)
torch.cuda.synchronize()
print('burn-in done', flush=True)
os.environ['CGPM_B200_PYTHON_SCHEDULE'] = '1'
eng.transition(N=1, kernels=['rows', 'columns', 'alpha', 'view_alphas', 'column_hypers'], progress=False)
rng = np.random.RandomState(5)
Q = 200000
tv = rng.normal(2.0, 2.0, size=(Q, 2)); cv = rng.normal(2.0, 2.0, size=(Q, 2))
lp = eng.logpdf_bulk_arrays([0, 1], tv, [2, 3], cv)
sm = eng.simulate_bulk_arrays(50000, [0, 1], [2, 3], cv[:50000], N=1)
torch.cuda.synchronize()
print('done', float(np.mean(lp)), float(np.mean(sm)))
