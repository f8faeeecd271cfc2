"""Per-(chain, column) timing of the auxiliary-view kernel (not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['CGPM_B200_AUX_PROF'] = '1'
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
R, C, S = 100000, 50, 64
X, cct, da = bench.make_table(R, C, 4, 4)
eng = Engine(X, num_states=0, rng=gen_rng(1), cctypes=cct, stream='philox', Kcap=256, Vcap=24)
eng._make_chains(S, 1000)
for s in range(S):
    Zv, Zrv = bench.initial_partitions(R, C, s)
    eng._ch.init_chain(s, gen_rng(7 + s), Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv})
eng._ch.finish_init()
eng.transition(N=2, kernels=['rows', 'columns'], progress=False)
seq = eng._ch.dev.t['seq'].cpu().numpy().reshape(-1)[:S * C * 4].reshape(S * C, 4)
order = np.argsort(-(seq[:, 0] + seq[:, 1]))
print('slowest pairs (pass1 kcyc, pass2 kcyc, K, alpha idx):', seq[order[:10]].tolist())
for ai in range(0, 30, 3):
    m = seq[:, 3] == ai
    if m.any():
        print('alpha idx %2d  alpha %.3g  n %3d  K mean %7.0f  pass1 %8.0f  pass2 %8.0f kcycles'
              % (ai, eng._ch.dev.row_alpha_grid_h[ai], m.sum(), seq[m, 2].mean(), seq[m, 0].mean(), seq[m, 1].mean()))
