"""Localise failures at very large cluster counts: a view with K0 clusters supplied, each kernel followed by a sync."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
R = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
K0 = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
C, S = 6, 2
X, cct, da = bench.make_table(R, C, 2, 4)
rng = np.random.RandomState(0)
Zv = {c: c % 2 for c in range(C)}
z = np.concatenate([np.arange(K0), rng.randint(K0, size=R - K0)])
Zrv = {0: [int(a) for a in z], 1: [int(a % 3) for a in np.arange(R)]}
def step(name, f):
    t0 = time.time(); f(); torch.cuda.synchronize(); print('%s ok %.2fs' % (name, time.time() - t0), flush=True)
eng = None
def mk():
    global eng
    eng = Engine(X, num_states=S, rng=gen_rng(3), cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={0: float(K0), 1: 1.0}, stream='philox')
step('init', mk)
dev = eng._ch.dev
print('Kcap', dev.Kcap, 'K', dev.t['nclu'].cpu().numpy()[:, :2].tolist(), flush=True)
step('score', lambda: eng.logpdf_score())
step('check', lambda: eng._ch.check_partitions())
step('rows', lambda: eng.transition(N=1, kernels=['rows'], progress=False))
print('K', dev.t['nclu'].cpu().numpy()[:, :2].tolist(), 'Kcap', dev.Kcap, flush=True)
step('check', lambda: eng._ch.check_partitions())
step('columns', lambda: eng.transition(N=1, kernels=['columns'], progress=False))
step('check', lambda: eng._ch.check_partitions())
step('hypers', lambda: eng.transition(N=1, kernels=['alpha', 'view_alphas', 'column_hypers'], progress=False))
step('rows2', lambda: eng.transition(N=1, kernels=['rows'], progress=False))
step('check', lambda: eng._ch.check_partitions())
print('done', eng.logpdf_score())
