"""Timing probe of the rows kernel at benchmark shapes (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import OrderedDict
from cgpm_b200.device import DeviceChains
from cgpm_b200 import codec, synth, _lib as L

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 50
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
V = int(sys.argv[4]) if len(sys.argv) > 4 else 2
sweeps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cct = ['normal'] * C; da = [None] * C
X, Zv, Zrv = synth.gen_table(R, cct, da, n_views=V, clusters_per_view=4, seed=0)
ctype = np.zeros(C, np.int32); ncat = np.zeros(C, np.int32)
t0 = time.time()
dev = DeviceChains(X, ctype, ncat, S, Vcap=16, Kcap=256)
hyp = np.zeros((C, 4))
for c in range(C):
    g = dev.grids_h[c]; hyp[c] = [g[0][10], g[1][15], g[2][15], g[3][15]]
for s in range(S):
    rng = np.random.RandomState(100 + s)
    zv = synth.crp_partition(C, 1.0, rng)
    views = OrderedDict()
    for v in np.unique(zv):
        views[int(v)] = {'alpha': 1.0, 'Zr': synth.crp_partition(R, 1.0, rng), 'order': None}
    spec = {'Zv': [int(z) for z in zv], 'views': views, 'alpha': 1.0, 'hyp': hyp}
    dev.upload_chain(s, codec.pack_chain(spec, R, C))
dev.rebuild(); dev.sync()
print('setup %.1fs' % (time.time() - t0))
nv = dev.t['nv'].cpu().numpy()
work = [dict(chain=s, view=int(v), n=R, n_cols=int(nv[s][v]), n_normal=int(nv[s][v])) for s in range(S) for v in np.nonzero(nv[s] > 0)[0]]
print('work items', len(work), 'views/chain', len(work) / S)
print('score0', dev.logpdf_score().cpu().numpy()[:3])
for it in range(sweeps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.transition_rows(work, seed=1234, counter=it)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    K = dev.t['nclu'].cpu().numpy()
    print('sweep %d: %.1f ms  %.3g cell-updates/s  K mean %.1f max %d' % (it, ms, R * C * S / ms * 1e3, K[nv > 0].mean(), K.max()))
import os
if os.environ.get('CGPM_B200_ROWS_PROF'):
    seq = dev.t['seq'].cpu().numpy()
    rec = sorted([(int(seq[w['chain'], w['view'], 0]), int(seq[w['chain'], w['view'], 1]), int(seq[w['chain'], w['view'], 2]), int(seq[w['chain'], w['view'], 3])) for w in work])
    print('final levels', collections.Counter(int(seq[w['chain'], w['view'], 5]) for w in work) if False else sorted(set(int(seq[w['chain'], w['view'], 5]) for w in work)))
    print('per-CTA (kcycles, D, K, moves): slowest', rec[-8:])
    print('median', rec[len(rec)//2], 'fastest', rec[:3])
    import collections
    byD = collections.defaultdict(list)
    for kc, D, K, mv in rec: byD[D if D < 4 else D // 8 * 8 + 4].append(kc)
    mvD = collections.defaultdict(list)
    for kc, D, K, mv in rec: mvD[D if D < 4 else D // 8 * 8 + 4].append(mv / R)
    print('move rate by width', {d: round(float(np.mean(v)), 4) for d, v in sorted(mvD.items())})
    print({d: (len(v), int(np.mean(v))) for d, v in sorted(byD.items())})
dev.check_errors()
print('score', dev.logpdf_score().cpu().numpy()[:3])
