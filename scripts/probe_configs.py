"""Sanity + timing of the other BASELINE.json configs at reduced chain counts (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cgpm_b200 import synth, hostmath as hm
from cgpm_b200.engine import Engine, gen_rng

def run(name, R, C, S, cct, da, views, sweeps, kernels, Vcap, Kcap):
    t0 = time.time()
    X, _, _ = synth.gen_table(R, cct, da, n_views=views, clusters_per_view=4, seed=0)
    eng = Engine(X, num_states=0, rng=gen_rng(1), cctypes=cct, distargs=da, stream='philox', Kcap=Kcap, Vcap=Vcap)
    eng._make_chains(S, 99)
    for s in range(S):
        rng = np.random.RandomState(s)
        zv = hm.crp_simulate_fast(C, 1.0, rng)
        Zv = {c: int(z) for c, z in enumerate(zv)}
        Zrv = {int(v): hm.crp_simulate_fast(R, 1.0, rng) for v in set(Zv.values())}
        eng._ch.init_chain(s, gen_rng(7 + s), Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv})
    eng._ch.finish_init(); torch.cuda.synchronize()
    print('%s: setup %.1f s, views/chain %.1f' % (name, time.time() - t0, float((eng._ch.dev.t['nv'] > 0).sum()) / S))
    for it in range(sweeps):
        eng._ch.events = []
        torch.cuda.synchronize(); t1 = time.time()
        eng.transition(N=1, kernels=kernels, progress=False)
        torch.cuda.synchronize(); dt = time.time() - t1
        ev = {n: a.elapsed_time(b) for n, a, b in eng._ch.events}
        nv = eng._ch.dev.t['nv'].cpu().numpy(); nclu = eng._ch.dev.t['nclu'].cpu().numpy()
        print('  sweep %d: %.0f ms %s  %.3g cell-updates/s  views %d  K mean %.1f max %d  score %.6g'
              % (it, dt * 1e3, {k: round(v) for k, v in ev.items()}, R * C * S / dt, (nv > 0).sum(), nclu[nv > 0].mean(), nclu.max(),
                 np.mean(eng.logpdf_score())))
    return eng, X

which = sys.argv[1] if len(sys.argv) > 1 else 'c4'
if which == 'c4':
    C = 5000
    run('c4 10k x 5000 normal, 4 chains', 10000, C, 4, ['normal'] * C, [None] * C, 50, 3, ['columns', 'rows'], Vcap=64, Kcap=128)
elif which == 'c3':
    C = 200
    cct = ['normal', 'categorical'] * 100
    da = [None if i % 2 == 0 else {'k': [2, 4, 8, 16][(i // 2) % 4]} for i in range(C)]
    run('c3 1M x 200 mixed, 8 chains', 1000000, C, 8, cct, da, 5, 2, ['rows', 'columns'], Vcap=16, Kcap=128)
elif which == 'c5':
    C = 50
    eng, X = run('c5 base 100k x 50, 64 chains', 100000, C, 64, ['normal'] * C, [None] * C, 4, 3, ['rows', 'columns'], Vcap=24, Kcap=256)
    from cgpm_b200 import query
    Q = 100000
    rng = np.random.RandomState(0)
    rows = rng.randint(0, 100000, size=Q)
    targets = [{0: float(X[r, 0]), 1: float(X[r, 1])} for r in rows]
    cons = [{2: float(X[r, 2]), 3: float(X[r, 3])} for r in rows]
    t0 = time.time(); enc = query._encode(eng._ch, [-1] * Q, targets, cons, simulate=False); t1 = time.time()
    torch.cuda.synchronize(); t2 = time.time()
    out = query.logpdf_bulk_encoded(eng._ch, list(range(64)), enc)
    torch.cuda.synchronize(); t3 = time.time()
    print('c5 logpdf_bulk: %d queries x 64 chains: encode %.1f s, device+copy %.3f s -> %.3g query-chains/s; ensemble mean %.4f'
          % (Q, t1 - t0, t3 - t2, Q * 64 / (t3 - t2), float(np.mean(out))))
