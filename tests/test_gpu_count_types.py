"""the remaining collapsed primitives on the device -- bernoulli, poisson, exponential, geometric
(src/primitives/bernoulli.py:141-155, poisson.py:147-172, exponential.py:140-166, geometric.py:140-164).  The oracle's
restatement of them is pinned against the reference on CPU (tests/test_oracle_vs_reference.py::
test_oracle_equals_reference_with_count_primitives); here the GPU engine follows the oracle."""
import numpy as np
import pytest

import crosscat_oracle as co
from helpers import count_table, rel_tol
from test_gpu_transition import compare

pytestmark = pytest.mark.gpu
KERNELS = ['alpha', 'view_alphas', 'column_hypers', 'rows', 'columns']


def make(n, seed, S=2):
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = count_table(n, seed)
    eng = Engine(X, num_states=S, rng=gen_rng(seed), cctypes=cct, distargs=da, stream='numpy')
    seeds = gen_rng(seed).randint(low=1, high=2 ** 32 - 1, size=S)
    oracles = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd)) for sd in seeds]
    return X, cct, da, eng, oracles


@pytest.mark.parametrize('n,seed', [(40, 21), (90, 22)])
def test_engine_with_count_columns_follows_oracle(n, seed):
    """Injected numpy stream: partitions, hypers, alphas and the sufficient statistics (N, sum_x and poisson's
    sum_log_fact_x) bit-identical to the oracle after every kernel; scores to 1e-10; stream position equal."""
    X, cct, da, eng, oracles = make(n, seed)
    compare(eng, oracles, 'init')
    for it in range(4):
        for k in KERNELS:
            eng.transition(N=1, kernels=[k], progress=False)
            for o in oracles:
                o.transition(N=1, kernels=[k])
            compare(eng, oracles, (it, k))
    for s, o in enumerate(oracles):
        assert eng._ch.rngs[s].random_sample() == o.rng.random_sample()
    # densities: every type as target and as constraint, valid and invalid values
    queries = [(-1, {1: 1, 2: 4, 3: 0.7}, {4: 2, 0: 1.5, 7: 9}), (-1, {4: 0, 7: 30}, {1: 0, 5: 2}), (-1, {3: 12.5}, {}),
               (-1, {2: 0}, {3: 0.0}), (-1, {2: 2.5}, {}), (-1, {3: -1.0}, {}), (-1, {1: 2}, {}), (-1, {4: -3}, {0: 0.0})]
    for rowid, t, c in queries:
        got = eng.logpdf(rowid, t, c)
        for s, o in enumerate(oracles):
            ref = o.logpdf(rowid, t, c)
            if np.isinf(ref):
                assert got[s] == ref
            else:
                assert abs(got[s] - ref) <= 1e-9 * max(1.0, abs(ref)), (t, c, got[s], ref)
    # metadata carries the primitives' own statistic names
    md = eng.to_metadata()['states'][0]
    names = [set(dict(col)[max(dict(col))].keys()) for col in md['suffstats']]
    assert names[1] == {'N', 'x_sum'} and names[2] == {'N', 'sum_x', 'sum_log_fact_x'} and names[3] == names[4] == {'N', 'sum_x'}


def test_rows_conditionals_with_count_columns():
    """Per-step conditionals of the rows kernel on a table holding every primitive: within 1e-5 relative of the oracle's."""
    X, cct, da, eng, oracles = make(70, 23)
    ch = eng._ch
    for sweep in range(3):
        otr = [dict() for _ in oracles]
        for o, tr in zip(oracles, otr):
            o.transition_view_rows(trace=tr)
        trace = {'stride': 3 + ch.dev.Kcap}
        ch.transition_view_rows(list(range(ch.S)), trace=trace)
        for wi, w in enumerate(trace['work']):
            o = oracles[w['chain']]
            vid = list(o.views.keys())[ch.views_order[w['chain']].index(w['view'])]
            for i, (rowid, K, logp, zb) in enumerate(otr[w['chain']][vid]):
                rec = trace['data'][wi, i]
                assert int(rec[0]) == rowid and int(rec[1]) == len(K) and int(rec[2]) == zb
                got, ref = rec[3:3 + len(K)], np.asarray(logp)
                assert np.all(np.abs(got - ref) <= rel_tol(ref)), (sweep, wi, i, got, ref)
        compare(eng, oracles, ('rows', sweep))


def test_incremental_edits_with_count_columns():
    X, cct, da, eng, oracles = make(50, 24)

    def both(tag, fe, fo):
        fe(eng)
        for o in oracles:
            fo(o)
        compare(eng, oracles, tag)
    both('t0', lambda e: e.transition(N=1, kernels=['rows', 'columns'], progress=False), lambda o: o.transition(N=1, kernels=['rows', 'columns']))
    both('inc1', lambda e: e.incorporate(50, {0: 0.5, 1: 1, 2: 7, 3: 0.25, 4: 3, 7: 12}), lambda o: o.incorporate(50, {0: 0.5, 1: 1, 2: 7, 3: 0.25, 4: 3, 7: 12}))
    both('inc2', lambda e: e.incorporate(51, {2: 0, 4: 0, 5: 1}), lambda o: o.incorporate(51, {2: 0, 4: 0, 5: 1}))
    both('force', lambda e: e.force_cell(51, {3: 1.9, 1: 0, 7: 4}), lambda o: o.force_cell(51, {3: 1.9, 1: 0, 7: 4}))
    both('t1', lambda e: e.transition(N=1, kernels=KERNELS, progress=False), lambda o: o.transition(N=1, kernels=KERNELS))
    both('uninc', lambda e: e.unincorporate(51), lambda o: o.unincorporate(51))
    T = [float(v) for v in np.random.RandomState(3).poisson(5.0, size=51)]
    both('dim+', lambda e: e.incorporate_dim(T, [90], cctype='poisson', v=99), lambda o: o.incorporate_dim(T, [90], cctype='poisson', distargs=None, v=99))
    both('t2', lambda e: e.transition(N=2, kernels=KERNELS, progress=False), lambda o: o.transition(N=2, kernels=KERNELS))
    for s, o in enumerate(oracles):
        assert eng._ch.rngs[s].random_sample() == o.rng.random_sample()
    with pytest.raises(ValueError):
        eng.incorporate(51, {1: 3})               # Invalid Bernoulli
    with pytest.raises(ValueError):
        eng.incorporate(51, {2: 1.5})             # Invalid Poisson
    with pytest.raises(ValueError):
        eng.incorporate(51, {3: -0.1})            # Invalid Exponential


def test_simulate_and_device_stream_with_count_columns():
    """Predictive samples of every new type have the right support and match the predictive mean implied by logpdf;
    the device-stream (Philox) schedule -- through cc_transition -- keeps the partitions consistent."""
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = count_table(200, 25, nan=False)
    eng = Engine(X, num_states=3, rng=gen_rng(5), cctypes=cct, distargs=da, stream='philox')
    eng.transition(N=6, progress=False)
    eng._ch.check_partitions()
    assert all(np.isfinite(eng.logpdf_score()))
    N = 4000
    sm = eng.simulate_bulk_arrays(1, [1, 2, 3, 4], N=N)[:, 0]            # [chains, N, 4]
    assert set(np.unique(sm[..., 0])) <= {0.0, 1.0}
    assert np.all(sm[..., 1] >= 0) and np.all(sm[..., 1] == np.floor(sm[..., 1]))
    assert np.all(sm[..., 2] >= 0)
    assert np.all(sm[..., 3] >= 0) and np.all(sm[..., 3] == np.floor(sm[..., 3]))
    for s in range(3):
        # bernoulli: P(x = 1) from logpdf; poisson: mean from the pmf over 0..200
        p1 = np.exp(eng.logpdf(-1, {1: 1})[s])
        assert abs(sm[s, :, 0].mean() - p1) < 5 * np.sqrt(p1 * (1 - p1) / N) + 1e-3
        pm = np.exp(eng.logpdf_bulk_arrays([2], np.arange(0, 400.0)[:, None], statenos=[s])[0])
        assert abs(pm.sum() - 1) < 1e-6
        mean = float((pm * np.arange(0, 400.0)).sum())
        var = float((pm * np.arange(0, 400.0) ** 2).sum()) - mean ** 2
        assert abs(sm[s, :, 1].mean() - mean) < 5 * np.sqrt(var / N)
