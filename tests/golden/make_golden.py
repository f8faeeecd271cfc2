#!/usr/bin/env python
"""Generate tests/golden/*.json from the REFERENCE itself (the py3-shimmed build
of /root/reference under oracle/_ref, see oracle/build_ref.py).  Run in the
build container only:   python tests/golden/make_golden.py

Floats that must match bit-for-bit are stored as float.hex() strings.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, ROOT)
import build_ref  # noqa: E402

ref = build_ref.build(quiet=True)
assert ref, 'reference tree not present'
sys.path.insert(0, ref)
from cgpm.crosscat.engine import Engine          # noqa: E402
from cgpm.crosscat.state import State            # noqa: E402
from cgpm.primitives.normal import Normal        # noqa: E402
from cgpm.primitives.categorical import Categorical  # noqa: E402
from cgpm.primitives.crp import Crp              # noqa: E402
from cgpm.utils import general as gu             # noqa: E402
from cgpm.utils import test as tu                # noqa: E402

KERNELS = ['alpha', 'view_alphas', 'column_hypers', 'rows', 'columns']


def hx(x):
    return float(x).hex()


def snapshot(state):
    md = state.to_metadata()
    out = {
        'Zv': sorted([int(c), int(v)] for c, v in md['Zv']),
        'Zrv': [[int(v), [int(z) for z in zr]] for v, zr in md['Zrv']],
        'row_order': [[int(v), [int(r) for r in view.Zr().keys()]] for v, view in state.views.items()],
        'alpha': hx(md['alpha']),
        'view_alphas': [[int(v), hx(a)] for v, a in md['view_alphas']],
        'hypers': [{k: hx(v) for k, v in h.items()} for h in md['hypers']],
        'suffstats': [],
        'logpdf_score': float(state.logpdf_score()),
    }
    for col in md['suffstats']:
        items = []
        for k, st in col:
            if 'counts' in st:
                items.append([int(k), {'N': hx(st['N']), 'counts': [hx(c) for c in st['counts']]}])
            elif 'sum_log_x' in st:      # lognormal: the same slots hold the sums of log x
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_log_x']), 'sum_x_sq': hx(st['sum_log_x_sq'])}])
            else:
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_x']), 'sum_x_sq': hx(st['sum_x_sq'])}])
        out['suffstats'].append(sorted(items))
    return out


def trajectory(lognormal=False):
    cct = ['normal', 'categorical', 'normal', 'categorical', 'normal', 'normal']
    da = [None, {'k': 4}, None, {'k': 3}, None, None]
    T, _Zv, _Zc = tu.gen_data_table(
        40, [.5, .5], [[.3, .7], [.2, .3, .5]], cct, da, [.9] * 6, rng=gu.gen_rng(0))
    X = T.T.copy()
    X[3, 0] = X[7, 1] = X[10, 2] = X[11, 3] = X[20, 5] = np.nan
    if lognormal:                  # columns 2 and 5 become positive and lognormal
        cct = list(cct)
        for c in (2, 5):
            X[:, c] = np.exp(0.5 * X[:, c])
            cct[c] = 'lognormal'
    seed, S, N = 17, 2, 4
    eng = Engine(X, num_states=S, rng=gu.gen_rng(seed), cctypes=cct, distargs=da, multiprocess=0)
    g = {
        'X': [[None if np.isnan(v) else float(v) for v in row] for row in X],
        'cctypes': cct, 'distargs': [d or {} for d in da], 'engine_seed': seed, 'num_states': S,
        'kernels': KERNELS, 'iterations': N, 'init': [snapshot(s) for s in eng.states], 'after': [],
        'queries': [],
    }
    for _ in range(N):
        eng.transition(N=1, kernels=KERNELS, progress=False, multiprocess=0)
        g['after'].append([snapshot(s) for s in eng.states])
    queries = [[-1, {'0': 1.2}, {}], [-1, {'0': 1.2, '1': 2}, {'2': 0.3}],
               [-1, {'3': 1, '5': -0.5 if not lognormal else 0.6}, {'0': 2., '4': 1., '1': 0}], [3, {'0': 0.5}, {}]]
    for rowid, t, c in queries:
        tt = {int(k): v for k, v in t.items()}
        cc = {int(k): v for k, v in c.items()}
        g['queries'].append({'rowid': rowid, 'targets': t, 'constraints': c,
                             'logpdf': [float(x) for x in eng.logpdf(rowid, tt, cc, multiprocess=0)]})
    g['stream_check'] = [float(s.rng.random_sample()) for s in eng.states]
    return g


def kats():
    """Known answers of the primitives (reference static methods)."""
    k = {}
    k['normal_predictive'] = [
        [[1.5, 0, 0, 0, 0., 1., 1., 1.], Normal.calc_predictive_logp(1.5, 0, 0, 0, 0., 1., 1., 1.)],
        [[1.5, 3, 3.3, 4.85, .5, 2., 3., 4.], Normal.calc_predictive_logp(1.5, 3, 3.3, 4.85, .5, 2., 3., 4.)],
        [[-7.25, 1000, 512.5, 1900.75, 1.5, .01, 20., 3.], Normal.calc_predictive_logp(-7.25, 1000, 512.5, 1900.75, 1.5, .01, 20., 3.)],
    ]
    k['normal_marginal'] = [
        [[3, 3.3, 4.85, .5, 2., 3., 4.], Normal.calc_logpdf_marginal(3, 3.3, 4.85, .5, 2., 3., 4.)],
        [[0, 0, 0, .5, 2., 3., 4.], Normal.calc_logpdf_marginal(0, 0, 0, .5, 2., 3., 4.)],
    ]
    k['normal_posterior'] = [[[3, 3.3, 4.85, .5, 2., 3., 4.], list(Normal.posterior_hypers(3, 3.3, 4.85, .5, 2., 3., 4.))]]
    k['categorical_predictive'] = [[[2, 6, [1, 2, 3, 0], .7], Categorical.calc_predictive_logp(2, 6, np.array([1., 2, 3, 0]), .7)]]
    k['categorical_marginal'] = [[[6, [1, 2, 3, 0], .7], float(Categorical.calc_logpdf_marginal(6, np.array([1., 2, 3, 0]), .7))]]
    from collections import OrderedDict
    counts = OrderedDict([(0, 2), (2, 3), (6, 1)])
    k['crp_predictive'] = [[[2, 6, 1.5], Crp.calc_predictive_logp(2, 6, counts, 1.5)],
                           [[7, 6, 1.5], Crp.calc_predictive_logp(7, 6, counts, 1.5)]]
    k['crp_marginal'] = [[[6, 1.5], float(Crp.calc_logpdf_marginal(6, counts, 1.5))]]
    k['log_pflip'] = [int(gu.log_pflip([-1, -2, -.5], array=[5, 7, 9], rng=r)) for r in [np.random.RandomState(42)] for _ in range(8)]
    rng = np.random.RandomState(42)
    k['log_pflip'] = [int(gu.log_pflip([-1, -2, -.5], array=[5, 7, 9], rng=rng)) for _ in range(8)]
    k['logsumexp'] = [[[999.0, 998.0, 997.5], gu.logsumexp([999.0, 998.0, 997.5])],
                      [[-1000.0, -1000.0], gu.logsumexp([-1000.0, -1000.0])]]
    k['normal_grids'] = {n: [float(x) for x in v] for n, v in Normal.construct_hyper_grids([1.0, 2.5, -0.5, 4.0]).items()}
    k['crp_grid'] = [float(x) for x in Crp.construct_hyper_grids([1] * 7)['alpha']]
    return k


if __name__ == '__main__':
    with open(os.path.join(HERE, 'trajectory_40x6.json'), 'w') as f:
        json.dump(trajectory(), f)
    with open(os.path.join(HERE, 'primitive_kats.json'), 'w') as f:
        json.dump(kats(), f, indent=1)
    with open(os.path.join(HERE, 'trajectory_lognormal_40x6.json'), 'w') as f:
        json.dump(trajectory(lognormal=True), f)
    print('golden vectors written')
