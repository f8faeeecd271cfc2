"""CPU: live replay of the oracle restatement against the reference itself
(oracle/_ref = the py3-shimmed /root/reference).  Skipped where the shimmed
reference has not been built (it needs /root/reference)."""
import os
import sys

import numpy as np
import pytest

import crosscat_oracle as co
from helpers import small_table, snapshot_oracle, count_table

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'cgpm', '.built')),
                                reason='oracle/_ref not built (reference tree absent)')


def ref_snapshot(state):
    sys.path.insert(0, os.path.join(os.path.dirname(REF), '..', 'tests', 'golden'))
    md = state.to_metadata()
    hx = lambda x: float(x).hex()
    out = {
        'Zv': sorted([int(c), int(v)] for c, v in md['Zv']),
        'Zrv': [[int(v), [int(z) for z in zr]] for v, zr in md['Zrv']],
        'row_order': [[int(v), [int(r) for r in view.Zr().keys()]] for v, view in state.views.items()],
        'alpha': hx(md['alpha']), 'view_alphas': [[int(v), hx(a)] for v, a in md['view_alphas']],
        'hypers': [{k: hx(v) for k, v in h.items()} for h in md['hypers']], 'suffstats': [],
    }
    for col in md['suffstats']:
        items = []
        for k, st in col:
            if 'counts' in st:
                items.append([int(k), {'N': hx(st['N']), 'counts': [hx(c) for c in st['counts']]}])
            elif 'sum_log_x' in st:        # lognormal: the same slots hold the sums of log x
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_log_x']), 'sum_x_sq': hx(st['sum_log_x_sq'])}])
            elif 'x_sum' in st:            # bernoulli
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['x_sum']), 'sum_x_sq': hx(0)}])
            elif 'sum_x_sq' not in st:     # poisson (third slot: sum_log_fact_x), exponential, geometric
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_x']), 'sum_x_sq': hx(st.get('sum_log_fact_x', 0))}])
            else:
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_x']), 'sum_x_sq': hx(st['sum_x_sq'])}])
        out['suffstats'].append(sorted(items))
    return out


@pytest.mark.parametrize('seed', [3, 4])
def test_oracle_equals_reference_over_transitions(seed):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from cgpm.crosscat.state import State
    X, cct, da = small_table(50, seed)
    ref = State(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seed))
    o = co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seed))
    kernels = ['alpha', 'view_alphas', 'column_hypers', 'rows', 'columns']
    for it in range(3):
        ref.transition(N=1, kernels=kernels, progress=False)
        o.transition(N=1, kernels=kernels)
        assert snapshot_oracle(o) == ref_snapshot(ref)
        assert o.logpdf_score() == ref.logpdf_score()
    assert o.logpdf(-1, {0: 1.0, 1: 2}, {2: 0.5}) == ref.logpdf(-1, {0: 1.0, 1: 2}, {2: 0.5})
    assert o.rng.random_sample() == ref.rng.random_sample()


def test_incremental_edits():
    """incorporate / unincorporate / force_cell / incorporate_dim / unincorporate_dim
    (state.py:187-320): oracle == reference, bit for bit, including the stream."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from cgpm.crosscat.state import State
    X, cct, da = small_table(40, 3)
    ref = State(X, cctypes=cct, distargs=da, rng=np.random.RandomState(1))
    o = co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(1))

    def both(f):
        f(ref)
        f(o)
        assert snapshot_oracle(o) == ref_snapshot(ref)
        assert o.logpdf_score() == ref.logpdf_score()

    def tr(kernels, N=1):
        return lambda s: s.transition(N=N, kernels=kernels, **({'progress': False} if isinstance(s, State) else {}))
    both(tr(['rows', 'columns']))
    both(lambda s: s.incorporate(40, {0: 1.5, 1: 2, 3: 0}))
    both(lambda s: s.incorporate(41, {2: -0.5, 5: 3.0, 4: 0.1}))
    both(lambda s: s.force_cell(41, {0: 0.25}))
    both(tr(['rows', 'column_hypers']))
    both(lambda s: s.unincorporate(41))
    T = list(np.random.RandomState(9).normal(size=41))
    both(lambda s: s.incorporate_dim(T, [77], cctype='normal', distargs=None))
    both(lambda s: s.incorporate_dim([float(i % 3) for i in range(41)], [78], cctype='categorical', distargs={'k': 3}, v=9))
    both(tr(['rows', 'columns', 'alpha', 'view_alphas', 'column_hypers'], N=2))
    both(lambda s: s.unincorporate_dim(2))
    both(tr(['rows', 'columns']))
    assert o.rng.random_sample() == ref.rng.random_sample()


@pytest.mark.parametrize('seed', [5, 6])
def test_oracle_equals_reference_with_lognormal_columns(seed):
    """SURVEY 8(f) rank 3, first step: the oracle's lognormal column (normal on log x with its own hyper
    grids and the -log x Jacobian, lognormal.py:56-102,151-157) replayed against the reference."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from cgpm.crosscat.state import State
    X, cct, da = small_table(50, seed)
    X = np.array(X)
    for c in (2, 5):                                   # two of the normal columns become positive, skewed
        X[:, c] = np.exp(0.4 * X[:, c])
        cct[c] = 'lognormal'
    ref = State(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seed))
    o = co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seed))
    assert snapshot_oracle(o) == ref_snapshot(ref)
    kernels = ['alpha', 'view_alphas', 'column_hypers', 'rows', 'columns']
    for it in range(3):
        ref.transition(N=1, kernels=kernels, progress=False)
        o.transition(N=1, kernels=kernels)
        assert snapshot_oracle(o) == ref_snapshot(ref)
        assert o.logpdf_score() == ref.logpdf_score()
    q = ({0: 1.0, 2: 1.7}, {5: 0.8, 1: 2})
    assert o.logpdf(-1, *q) == ref.logpdf(-1, *q)
    assert o.logpdf(-1, {2: -1.0}) == ref.logpdf(-1, {2: -1.0}) == float('-inf')
    # incremental edits on a table with lognormal columns

    def both(f):
        f(ref); f(o)
        assert snapshot_oracle(o) == ref_snapshot(ref)
    both(lambda s: s.incorporate(50, {0: 0.5, 2: 2.5, 5: 0.3}))
    both(lambda s: s.incorporate(51, {2: 0.7, 4: 0.1}))
    both(lambda s: s.force_cell(51, {5: 1.9}))
    both(lambda s: s.unincorporate(51))
    T = list(np.exp(np.random.RandomState(3).normal(size=51)))
    both(lambda s: s.incorporate_dim(T, [90], cctype='lognormal', distargs=None, v=99))
    ref.transition(N=1, kernels=kernels, progress=False)
    o.transition(N=1, kernels=kernels)
    assert snapshot_oracle(o) == ref_snapshot(ref)
    assert o.logpdf_score() == ref.logpdf_score()
    assert o.rng.random_sample() == ref.rng.random_sample()


def test_relevance_probability_with_hypothetical_rows():
    """state.py:599-629 against the oracle: same answers, same statistics and stream position afterwards."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from cgpm.crosscat.state import State
    X, cct, da = small_table(60, 7)
    ref = State(X, cctypes=cct, distargs=da, rng=np.random.RandomState(7))
    o = co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(7))
    for s in (ref, o):
        s.transition(N=2, kernels=['rows', 'columns'], **({'progress': False} if s is ref else {}))
    hyp = [{0: 0.3, 2: -0.2}, {4: 1.0}, {1: 2, 3: 0}]
    for tgt, qry, col in [(3, [5, 8], 0), (3, [], 2), (10, [11], 4), (0, [1, 2, 3], 1)]:
        assert o.relevance_probability(tgt, qry, col, hyp) == ref.relevance_probability(tgt, list(qry), col, hyp)
        assert o.relevance_probability(tgt, qry, col) == ref.relevance_probability(tgt, list(qry), col)
    assert snapshot_oracle(o) == ref_snapshot(ref)
    assert o.rng.random_sample() == ref.rng.random_sample()


@pytest.mark.parametrize('seed', [8, 9])
def test_oracle_equals_reference_with_count_primitives(seed):
    """bernoulli, poisson, exponential and geometric columns (bernoulli.py:141-155,
    poisson.py:147-172, exponential.py:140-166, geometric.py:140-164 and their hyper grids) replayed against the
    reference: whole transitions, queries, incremental edits -- bit for bit, including the stream position."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from cgpm.crosscat.state import State
    X, cct, da = count_table(50, seed)
    ref = State(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seed))
    o = co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seed))
    assert snapshot_oracle(o) == ref_snapshot(ref)
    assert o.logpdf_score() == ref.logpdf_score()
    kernels = ['alpha', 'view_alphas', 'column_hypers', 'rows', 'columns']
    for it in range(3):
        ref.transition(N=1, kernels=kernels, progress=False)
        o.transition(N=1, kernels=kernels)
        assert snapshot_oracle(o) == ref_snapshot(ref)
        assert o.logpdf_score() == ref.logpdf_score()
    q = ({1: 1, 2: 4, 3: 0.7}, {4: 2, 0: 1.5, 7: 9})
    assert o.logpdf(-1, *q) == ref.logpdf(-1, *q)
    assert o.logpdf(-1, {2: 2.5}) == ref.logpdf(-1, {2: 2.5}) == float('-inf')      # not a count
    assert o.logpdf(-1, {3: -1.0}) == ref.logpdf(-1, {3: -1.0}) == float('-inf')
    assert o.logpdf(-1, {1: 2}) == ref.logpdf(-1, {1: 2}) == float('-inf')

    def both(f):
        f(ref); f(o)
        assert snapshot_oracle(o) == ref_snapshot(ref)
    both(lambda s: s.incorporate(50, {0: 0.5, 1: 1, 2: 7, 3: 0.25, 4: 3, 7: 12}))
    both(lambda s: s.incorporate(51, {2: 0, 4: 0, 5: 1}))
    both(lambda s: s.force_cell(51, {3: 1.9, 1: 0}))
    both(lambda s: s.unincorporate(51))
    T = [float(v) for v in np.random.RandomState(3).poisson(5.0, size=51)]
    both(lambda s: s.incorporate_dim(T, [90], cctype='poisson', distargs=None, v=99))
    ref.transition(N=1, kernels=kernels, progress=False)
    o.transition(N=1, kernels=kernels)
    assert snapshot_oracle(o) == ref_snapshot(ref)
    assert o.logpdf_score() == ref.logpdf_score()
    assert o.rng.random_sample() == ref.rng.random_sample()
