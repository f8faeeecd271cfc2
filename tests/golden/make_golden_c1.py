#!/usr/bin/env python
"""Generate tests/golden/trajectory_c1_1000x10.json.gz from the REFERENCE itself
(oracle/_ref, see oracle/build_ref.py): BASELINE.json configs[0] -- a 1 000 x 10
mixed table (5 normal + 5 categorical(k=4) columns, the reference's own
gen_data_table), Engine(num_states=4, rng=gen_rng(1)) with the reference-faithful
CRP-prior initialisation (views start with up to several hundred clusters), and
N iterations of kernels=['rows','columns','view_alphas','alpha','column_hypers'].

Run in the build container only:   python tests/golden/make_golden_c1.py [N]

Every iteration stores a sha256 digest of the canonical snapshot of each state
(partitions, member order, alphas, hypers, sufficient statistics as float.hex);
the full snapshots are kept for the initial state, the first and the last
iteration.  BASELINE.md section 3.5 gate (3).
"""
import gzip
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg          # noqa: E402  (builds / imports the reference)

KERNELS = ['rows', 'columns', 'view_alphas', 'alpha', 'column_hypers']


def digest(snap):
    body = {k: snap[k] for k in ('Zv', 'Zrv', 'row_order', 'alpha', 'view_alphas', 'hypers', 'suffstats')}
    return hashlib.sha256(json.dumps(body, sort_keys=True).encode()).hexdigest()


def main(N):
    cct = ['normal'] * 5 + ['categorical'] * 5
    da = [None] * 5 + [{'k': 4}] * 5
    T, _Zv, _Zc = mg.tu.gen_data_table(
        1000, [.5, .5], [[.25, .25, .5], [.3, .7]], cct, da, [.9] * 10, rng=mg.gu.gen_rng(0))
    X = T.T.copy()
    rng = np.random.RandomState(5)
    for _ in range(50):                                   # a few missing cells in both column types
        X[rng.randint(1000), rng.randint(10)] = np.nan
    seed, S = 1, 4
    t0 = time.time()
    eng = mg.Engine(X, num_states=S, rng=mg.gu.gen_rng(seed), cctypes=cct, distargs=da, multiprocess=0)
    g = {
        'X': [[None if np.isnan(v) else float(v) for v in row] for row in X],
        'cctypes': cct, 'distargs': [d or {} for d in da], 'engine_seed': seed, 'num_states': S,
        'kernels': KERNELS, 'iterations': N, 'digests': [], 'full': {},
        'K_init': [[len(set(v.Zr().values())) for v in s.views.values()] for s in eng.states],
    }
    snaps = [mg.snapshot(s) for s in eng.states]
    g['full']['init'] = snaps
    g['digest_init'] = [digest(x) for x in snaps]
    print('init done %.1fs, K per view: %s' % (time.time() - t0, g['K_init']), flush=True)
    for it in range(N):
        eng.transition(N=1, kernels=KERNELS, progress=False, multiprocess=0)
        snaps = [mg.snapshot(s) for s in eng.states]
        g['digests'].append([digest(x) for x in snaps])
        if it in (0, N - 1):
            g['full'][str(it)] = snaps
        print('iteration %d done %.1fs' % (it, time.time() - t0), flush=True)
    g['logpdf_score'] = [float(s.logpdf_score()) for s in eng.states]
    g['stream_check'] = [float(s.rng.random_sample()) for s in eng.states]
    with gzip.open(os.path.join(HERE, 'trajectory_c1_1000x10.json.gz'), 'wt') as f:
        json.dump(g, f)
    print('written')


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
