"""Per-view cost of the rows sweep in the bench.py state (configs[1], supplied CRP(1) init, warm-up sweeps):
cycles per row, width, clusters, move rate of every (chain, view) CTA.  CGPM_B200_ROWS_PROF must be set."""
import os, sys
os.environ['CGPM_B200_ROWS_PROF'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng

R, C, S = 100000, 50, 64
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 3
X, cct, da = bench.make_table(R, C, 4, 4)
eng = Engine(X, num_states=0, rng=gen_rng(1), cctypes=cct, stream='philox', Kcap=256, Vcap=24)
eng._make_chains(S, 1000)
ch = eng._ch
for s in range(S):
    Zv, Zrv = bench.initial_partitions(R, C, s)
    ch.init_chain(s, gen_rng(7 + s), Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv})
ch.finish_init()
for it in range(warm + 2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.transition(N=1, kernels=['rows'], progress=False)
    e1.record(); torch.cuda.synchronize()
    seq = ch.dev.t['seq'].cpu().numpy(); nv = ch.dev.t['nv'].cpu().numpy()
    top = sorted((int(seq[s, v, 0]) * 1024 // R, int(seq[s, v, 1]), int(seq[s, v, 2]), round(int(seq[s, v, 3]) / R, 3), int(seq[s, v, 4]), int(seq[s, v, 5]))
                 for s in range(S) for v in range(ch.dev.Vcap) if nv[s, v] > 0)[-4:]
    print('rows sweep %d: %.1f ms; slowest (cycles/row, D, K after, move rate, mode, level): %s' % (it, e0.elapsed_time(e1), top))
    eng.transition(N=1, kernels=['columns'], progress=False)
eng.transition(N=1, kernels=['rows'], progress=False)
torch.cuda.synchronize()
seq = ch.dev.t['seq'].cpu().numpy(); nv = ch.dev.t['nv'].cpu().numpy()
rec = sorted((int(seq[s, v, 0]) * 1024 // R, int(seq[s, v, 1]), int(seq[s, v, 2]), round(int(seq[s, v, 3]) / R, 3), int(seq[s, v, 5]))
             for s in range(S) for v in range(ch.dev.Vcap) if nv[s, v] > 0)
print('views', len(rec))
print('slowest (cycles/row, D, K, move rate, level):', rec[-12:])
print('median', rec[len(rec) // 2])
import collections
by = collections.defaultdict(list)
for cyc, D, K, mv, lv in rec:
    by[D if D < 5 else (D // 8) * 8 + 4].append((cyc, mv))
for d in sorted(by):
    v = by[d]
    print('D~%d: n=%d cycles/row mean %d max %d, move rate mean %.3f' % (d, len(v), np.mean([a for a, _ in v]), max(a for a, _ in v), np.mean([b for _, b in v])))
