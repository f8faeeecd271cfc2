"""configs[2] shape on one GPU (R x 200 mixed, supplied hypers as in bench.py), S chains: per-sweep times and per-view costs."""
import os, sys, time
os.environ['CGPM_B200_ROWS_PROF'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
full = len(sys.argv) > 4
cfg = bench.C2
t0 = time.time()
X, cct, da = bench.make_table(R, cfg['cols'], cfg['views'], cfg['clusters'], True)
Zv, Zrv = bench.initial_partitions(R, cfg['cols'], int(os.environ.get('PROBE_PSEED', '12345')))
eng = Engine(X, num_states=S, rng=gen_rng(int(os.environ.get('PROBE_SEED', '1'))), cctypes=cct, distargs=da, Zv=Zv, Zrv=Zrv, alpha=1.0,
             view_alphas={v: 1.0 for v in Zrv}, hypers=bench.supplied_hypers(X, cct), stream='philox', Kcap=cfg['kcap'],
             Vcap=cfg['vcap'])
torch.cuda.synchronize()
print('S %d R %d setup %.1fs' % (S, R, time.time() - t0), flush=True)
ch = eng._ch
for it in range(n):
    eng.start_kernel_timing()
    eng._ch.events = []
    eng.transition(N=1, kernels=['rows', 'columns'], progress=False)
    km = eng.stop_kernel_timing()
    seq = ch.dev.t['seq'].cpu().numpy(); nv = ch.dev.t['nv'].cpu().numpy()
    top = sorted((int(seq[s, v, 0]) * 1024 // R, int(seq[s, v, 1]), int(seq[s, v, 2]), round(int(seq[s, v, 3]) / R, 3), int(seq[s, v, 4]), int(seq[s, v, 5]))
                 for s in range(S) for v in range(ch.dev.Vcap) if nv[s, v] > 0)
    print('sweep %d: rows %.0f ms columns %.0f ms; views %d; (cycles/row, D, K, move rate, mode, level) slowest: %s; median %s' % (
        it, sum(km.get('rows', [0])), sum(km.get('columns', [0])), len(top), top[-3:], top[len(top) // 2]), flush=True)
    if full and it in (0, n - 1):
        print('   all:', top[::max(1, len(top) // 40)], flush=True)
    ph = ch.dev.t['seq'][:, :, 16:26].cpu().numpy().astype(np.float64)
    if ph[..., 8].sum() > 0:          # a -DCC_ROWS_PHASE=1 build (CGPM_B200_LIB): cycles per round of each phase, thread 0
        names = ['setup', 'wait+loads', 'score', 'draw+publish', 'barrier1', 'commit+replay', 'apply', 'barrier2+release']
        recs = sorted(((int(seq[s, v, 0]) * 1024 // R, s, v) for s in range(S) for v in range(ch.dev.Vcap) if nv[s, v] > 0))
        for cyc, s, v in recs[-3:] + recs[len(recs) // 2: len(recs) // 2 + 2]:
            p = ph[s, v]
            print('   view D %d K %d cyc/row %d: rounds/row %.3f moves/row %.3f; cycles per round: %s' % (
                seq[s, v, 1], seq[s, v, 2], cyc, p[8] / R, p[9] / R,
                ', '.join('%s %.0f' % (nm, 64.0 * p[i] / max(p[8], 1)) for i, nm in enumerate(names))), flush=True)
    ch.dev.t['seq'][:, :, 16:26].zero_()
