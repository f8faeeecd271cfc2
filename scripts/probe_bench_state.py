"""bench.py's configs[1] engine, sweep by sweep: rows / columns time and the slowest (chain, view) CTAs of each rows
sweep (cycles per row, width, clusters, move rate, table mode, level).  Usage: probe_bench_state.py [sweeps]"""
import os, sys
os.environ['CGPM_B200_ROWS_PROF'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
cfg = bench.C1
R, C, S = cfg['rows'], cfg['cols'], cfg['chains']
n = int(sys.argv[1]) if len(sys.argv) > 1 else 45
X, cct, da = bench.make_table(R, C, cfg['views'], cfg['clusters'])
Zv, Zrv = bench.initial_partitions(R, C, 12345)
eng = Engine(X, num_states=S, rng=gen_rng(1), cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv},
             stream='philox', Kcap=cfg['kcap'], Vcap=cfg['vcap'])
ch = eng._ch
for it in range(n):
    ch.dev.t['seq'][:, :, 0].zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    eng.transition(N=1, kernels=['rows'], progress=False)
    ev[1].record()
    seq = ch.dev.t['seq'][:, :, :6].cpu().numpy()
    eng.transition(N=1, kernels=['columns'], progress=False)
    ev[2].record(); torch.cuda.synchronize()
    rec = sorted((int(seq[s, v, 0]) * 1024 // R, int(seq[s, v, 1]), int(seq[s, v, 2]), round(int(seq[s, v, 3]) / R, 3), int(seq[s, v, 4]), int(seq[s, v, 5]))
                 for s in range(S) for v in range(ch.dev.Vcap) if seq[s, v, 0] > 0)
    print('sweep %2d: rows %6.1f ms columns %5.1f ms; views %d; slowest %s; median %s' % (
        it, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), len(rec), rec[-4:], rec[len(rec) // 2]), flush=True)
