"""Host-side profile of the column sweep (not the bench)."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for it in range(8):
    t1 = time.time(); eng.transition(N=1, kernels=['rows'], progress=False); torch.cuda.synchronize(); t2 = time.time()
    eng._ch.transition_dims(list(range(S))); torch.cuda.synchronize()
    print('iter', it, 'rows %.0f ms' % ((t2 - t1) * 1e3), 'cols %.0f ms' % ((time.time() - t2) * 1e3), 'rounds', eng._ch.last_col_rounds, 'views', int((eng._ch.dev.t['nv'] > 0).sum()))
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
for _ in range(3):
    eng._ch.transition_dims(list(range(S)))
    print('rounds', eng._ch.last_col_rounds)
torch.cuda.synchronize()
pr.disable()
print('3 column sweeps: %.1f ms each' % ((time.time() - t0) / 3 * 1e3))
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
