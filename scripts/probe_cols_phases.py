"""Per-phase GPU time of the column sweep via the launch-time log of the C-ABI calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['CGPM_B200_AUX_PROF'] = '1'
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
from cgpm_b200 import device as D
R, C, S = 100000, 50, 64
X, cct, da = bench.make_table(R, C, 4, 4)
eng = Engine(X, num_states=0, rng=gen_rng(1), cctypes=cct, stream='philox', Kcap=256, Vcap=24)
eng._make_chains(S, 1000)
for s in range(S):
    Zv, Zrv = bench.initial_partitions(R, C, s)
    eng._ch.init_chain(s, gen_rng(7 + s), Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv})
eng._ch.finish_init()
# wrap every C-ABI call with a sync + wall clock
lib = eng._ch.dev.lib
log = []
class Timed(object):
    def __getattr__(self, name):
        f = getattr(lib, name)
        if not name.startswith('cc_') or name in ('cc_last_error',):
            return f
        def call(*a):
            torch.cuda.synchronize(); t0 = time.time()
            r = f(*a)
            torch.cuda.synchronize(); log.append((name, (time.time() - t0) * 1e3))
            return r
        return call
eng._ch.dev.lib = Timed()
for it in range(9):
    eng.transition(N=1, kernels=['rows'], progress=False)
    log.clear()
    t0 = time.time()
    eng._ch.transition_dims(list(range(S)))
    torch.cuda.synchronize()
    tot = (time.time() - t0) * 1e3
    nclu = eng._ch.dev.t['nclu'].cpu().numpy(); nv = eng._ch.dev.t['nv'].cpu().numpy()
    seq = eng._ch.dev.t['seq'].cpu().numpy().reshape(-1)[:S * C * 4].reshape(S * C, 4)
    order = np.argsort(-(seq[:, 0] + seq[:, 1]))
    print('   slowest pairs (pass1 kcyc, pass2 kcyc, K, alpha idx):', seq[order[:4]].tolist())
    print('iter %d total %.0f ms  Kmax %d  ' % (it, tot, nclu[nv > 0].max()) + ' '.join('%s:%.0f' % (n[3:], ms) for n, ms in log))
