"""Repeat identical launches and print the duration distribution (environment jitter check)."""
import os, sys, time
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
eng.transition(N=2, kernels=['rows', 'columns'], progress=False)
ch = eng._ch
def timeit(f, n):
    out = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return np.array(out)
t = timeit(lambda: ch.dev.rebuild(), 40)
print('rebuild x40 ms: min %.1f med %.1f max %.1f' % (t.min(), np.median(t), t.max()), np.round(np.sort(t)[-5:], 1))
big = torch.empty(1 << 28, dtype=torch.float32, device='cuda')
t = timeit(lambda: big.mul_(1.0001), 200)
print('1GiB mul x200 ms: min %.2f med %.2f max %.2f' % (t.min(), np.median(t), t.max()), np.round(np.sort(t)[-5:], 2))
