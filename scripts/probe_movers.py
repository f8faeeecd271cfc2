"""Rows sweep over views whose rows move at every other step: D-column views of i.i.d. noise, K initial clusters.
Usage: probe_movers.py [D] [K] [chains] [sweeps]"""
import os, sys
os.environ['CGPM_B200_ROWS_PROF'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cgpm_b200.engine import Engine, gen_rng
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = int(sys.argv[3]) if len(sys.argv) > 3 else 32
n = int(sys.argv[4]) if len(sys.argv) > 4 else 3
R, V = 100000, 8
C = V * D
rng = np.random.RandomState(0)
X = rng.normal(0.0, 1.0, size=(R, C))
Zv = {c: c // D for c in range(C)}
Zrv = {v: rng.randint(0, K, size=R).tolist() for v in range(V)}
hyp = [{'m': 0.0, 'r': 1.0, 's': 1.0, 'nu': 1.0} for _ in range(C)]
eng = Engine(X, num_states=S, rng=gen_rng(1), cctypes=['normal'] * C, Zv=Zv, Zrv=Zrv, alpha=1.0,
             view_alphas={v: float(K) for v in range(V)}, hypers=hyp, stream='philox', Kcap=64, Vcap=V + 1)
ch = eng._ch
ch.dev.t['seq'].zero_()
for it in range(n):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.transition(N=1, kernels=['rows'], progress=False)
    e1.record(); torch.cuda.synchronize()
    seq = ch.dev.t['seq'][:, :, :6].cpu().numpy()
    rec = sorted((int(seq[s, v, 0]) * 1024 // R, int(seq[s, v, 1]), int(seq[s, v, 2]), round(int(seq[s, v, 3]) / R, 3), int(seq[s, v, 5]))
                 for s in range(S) for v in range(V))
    print('sweep %d: %.1f ms; cycles/row min %s median %s max %s' % (it, e0.elapsed_time(e1), rec[0], rec[len(rec) // 2], rec[-1]), flush=True)
    ph = ch.dev.t['seq'][:, :V, 16:26].cpu().numpy().astype(np.float64)
    if ph[..., 8].sum() > 0:
        tot = ph.sum(axis=(0, 1))
        rounds, moves = tot[8], tot[9]
        names = ['setup', 'wait+loads', 'score', 'draw+publish', 'barrier1|solo:-', 'commit', 'apply', 'barrier2+release']
        print('   rounds/row %.3f moves/row %.3f; cycles per round: %s' % (
            rounds / (S * V * R), moves / (S * V * R), ', '.join('%s %.0f' % (nm, 64.0 * tot[i] / rounds) for i, nm in enumerate(names))))
    ch.dev.t['seq'][:, :, 16:26].zero_()
