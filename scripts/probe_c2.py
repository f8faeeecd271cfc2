"""configs[2] shape on one GPU (1M x 200 mixed), S chains: per-sweep times and per-view costs."""
import os, sys, time
os.environ['CGPM_B200_ROWS_PROF'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = bench.C2
t0 = time.time()
X, cct, da = bench.make_table(R, cfg['cols'], cfg['views'], cfg['clusters'], True)
Zv, Zrv = bench.initial_partitions(R, cfg['cols'], 12345)
eng = Engine(X, num_states=S, rng=gen_rng(1), cctypes=cct, distargs=da, Zv=Zv, Zrv=Zrv, alpha=1.0,
             view_alphas={v: 1.0 for v in Zrv}, stream='philox', Kcap=cfg['kcap'], Vcap=cfg['vcap'])
torch.cuda.synchronize()
print('setup %.1fs' % (time.time() - t0), flush=True)
ch = eng._ch
for it in range(n):
    eng.start_kernel_timing()
    eng._ch.events = []
    eng.transition(N=1, kernels=['rows', 'columns'], progress=False)
    km = eng.stop_kernel_timing()
    seq = ch.dev.t['seq'].cpu().numpy(); nv = ch.dev.t['nv'].cpu().numpy()
    top = sorted((int(seq[s, v, 0]) * 1024 // R, int(seq[s, v, 1]), int(seq[s, v, 2]), round(int(seq[s, v, 3]) / R, 3), int(seq[s, v, 4]), int(seq[s, v, 5]))
                 for s in range(S) for v in range(ch.dev.Vcap) if nv[s, v] > 0)
    print('sweep %d: rows %.0f ms columns %.0f ms; views %d; slowest (cycles/row, D, K, move rate, mode, level): %s; median %s' % (
        it, sum(km.get('rows', [0])), sum(km.get('columns', [0])), len(top), top[-3:], top[len(top) // 2]), flush=True)
