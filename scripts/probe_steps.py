"""configs[1] engine of bench.py: per-step rows / columns times and the slowest views of every step
(cycles/row, D, K, move rate, mode, level), to see what sets the step time after the burn-in."""
import os, sys, time
os.environ['CGPM_B200_ROWS_PROF'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
burn = int(sys.argv[1]) if len(sys.argv) > 1 else 15
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cfg = bench.C1
R, S = cfg['rows'], cfg['chains']
X, cct, da = bench.make_table(R, cfg['cols'], cfg['views'], cfg['clusters'])
Zv, Zrv = bench.initial_partitions(R, cfg['cols'], 12345)
eng = Engine(X, num_states=S, rng=gen_rng(1), cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv},
             stream='philox', Kcap=cfg['kcap'], Vcap=cfg['vcap'])
eng.transition(N=burn, kernels=['rows', 'columns'], progress=False)
torch.cuda.synchronize()
ch = eng._ch
for it in range(n):
    ch.dev.t['seq'][:, :, 0].zero_()
    eng.start_kernel_timing()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.transition(N=1, kernels=['rows', 'columns'], progress=False)
    e1.record(); torch.cuda.synchronize()
    km = eng.stop_kernel_timing()
    seq = ch.dev.t['seq'].cpu().numpy(); nv = ch.dev.t['nv'].cpu().numpy()
    top = sorted((int(seq[s, v, 0]) * 1024 // R, int(seq[s, v, 1]), int(seq[s, v, 2]), round(int(seq[s, v, 3]) / R, 3),
                  int(seq[s, v, 4]), int(seq[s, v, 5])) for s in range(S) for v in range(ch.dev.Vcap) if seq[s, v, 0] > 0)
    print('step %d: %.1f ms (rows %.1f columns %.1f); views %d; slowest (cycles/row, D, K, move rate, mode, level): %s; median %s' % (
        it, e0.elapsed_time(e1), sum(km.get('rows', [0])), sum(km.get('columns', [0])), len(top), top[-4:], top[len(top) // 2]),
        flush=True)
