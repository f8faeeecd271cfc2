"""Dict-style bulk queries (the reference's Engine.logpdf_bulk / simulate_bulk shapes) on a configs[1]-like table:
host encode + kernel + decode, per-query loop against the vectorised encoder."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200 import query
from cgpm_b200.engine import Engine, gen_rng
R, C, S = 20000, 50, 64
X, cct, da = bench.make_table(R, C, 4, 4)
Zv, Zrv = bench.initial_partitions(R, C, 12345)
eng = Engine(X, num_states=S, rng=gen_rng(1), cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv},
             stream='philox', Kcap=64, Vcap=24)
eng.transition(N=2, kernels=['rows', 'columns'], progress=False)
rng = np.random.RandomState(5)
Q = 20000
tv = rng.normal(2.0, 2.0, size=(Q, 2)); cv = rng.normal(2.0, 2.0, size=(Q, 2))
tl = [{0: float(a), 1: float(b)} for a, b in tv]
cl = [{2: float(a), 3: float(b)} for a, b in cv]
ref = eng.logpdf_bulk_arrays([0, 1], tv, [2, 3], cv)
for name, thr in (('per-query loop', 10 ** 9), ('vectorised', 32)):
    query.FAST_ENCODE_MIN = thr
    eng.logpdf_bulk([-1] * 64, tl[:64], cl[:64])
    torch.cuda.synchronize()
    t0 = time.time()
    lp = eng.logpdf_bulk([-1] * Q, tl, cl)
    dt = time.time() - t0
    t0 = time.time()
    sm = eng.simulate_bulk([-1] * 2000, [[0, 1]] * 2000, cl[:2000], Ns=[5] * 2000)
    ds = time.time() - t0
    print('%-15s logpdf_bulk %d queries x %d chains: %.3f s -> %.3g query-chains/s (max |diff| vs arrays %.1e); '
          'simulate_bulk 2000 x 5 x %d: %.3f s' % (name, Q, S, dt, Q * S / dt, float(np.max(np.abs(np.asarray(lp) - ref))), S, ds),
          flush=True)
