"""End-to-end parity: Engine(stream='numpy') follows the oracle's trajectory for
the same seeds -- initial state, then every kernel of State.transition."""
import numpy as np
import pytest

import crosscat_oracle as co
from helpers import small_table, oracle_suffstats, device_suffstats, assert_suffstats_equal

pytestmark = pytest.mark.gpu

KERNELS = ['alpha', 'view_alphas', 'column_hypers', 'rows', 'columns']


def compare(eng, oracles, tag):
    ch = eng._ch
    scores = ch.logpdf_score()
    for s, o in enumerate(oracles):
        md = eng._ch and __import__('cgpm_b200.engine', fromlist=['x'])._chain_metadata(ch, s)
        om = o.to_metadata()
        assert sorted(md['Zv']) == sorted(om['Zv']), (tag, s, md['Zv'], om['Zv'])
        assert md['Zrv'] == om['Zrv'], (tag, s)
        assert md['alpha'] == om['alpha'], (tag, s)
        assert md['view_alphas'] == om['view_alphas'], (tag, s)
        for a, b in zip(md['hypers'], om['hypers']):
            assert dict(a) == dict(b), (tag, s, a, b)
        got, dl = device_suffstats(ch.dev, s, ch.ctype)
        assert_suffstats_equal(oracle_suffstats(o), got)
        ref = o.logpdf_score()
        assert abs(scores[s] - ref) <= 1e-10 * abs(ref), (tag, s, scores[s], ref)


@pytest.mark.parametrize('n_rows,seed', [(30, 11), (80, 12)])
def test_engine_numpy_stream_follows_oracle(n_rows, seed):
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = small_table(n_rows, seed)
    S = 3
    eng = Engine(X, num_states=S, rng=gen_rng(seed), cctypes=cct, distargs=da, stream='numpy')
    seeds = gen_rng(seed).randint(low=1, high=2 ** 32 - 1, size=S)
    oracles = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd)) for sd in seeds]
    compare(eng, oracles, 'init')
    for it in range(5):
        for k in KERNELS:
            eng.transition(N=1, kernels=[k], progress=False)
            for o in oracles:
                o.transition(N=1, kernels=[k])
            compare(eng, oracles, (it, k))
    for s, o in enumerate(oracles):
        assert eng._ch.rngs[s].random_sample() == o.rng.random_sample()
        assert eng._ch.diagnostics[s]['iterations'] == o.iterations


def test_columns_trace_logps(n_rows=60, seed=21):
    """Per-step column conditionals within 1e-5 of the oracle's."""
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = small_table(n_rows, seed)
    eng = Engine(X, num_states=2, rng=gen_rng(seed), cctypes=cct, distargs=da, stream='numpy')
    seeds = gen_rng(seed).randint(low=1, high=2 ** 32 - 1, size=2)
    oracles = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd)) for sd in seeds]
    for it in range(4):
        tr = {'stride': 3 + eng._ch.dev.Vcap + 1}
        eng._ch.transition_dims([0, 1], trace=tr)
        for s, o in enumerate(oracles):
            otr = []
            o.transition_dims(trace=otr)
            for j, (col, tables, logp, vs) in enumerate(otr):
                rec = tr['data'][s, j]
                assert int(rec[0]) == col and int(rec[1]) == len(tables) and int(rec[2]) == vs
                got, ref = rec[3:3 + len(tables)], np.asarray(logp)
                assert np.all(np.abs(got - ref) <= 1e-5 * np.maximum(1.0, np.abs(ref))), (got, ref)
        eng.transition(N=1, kernels=['rows', 'column_hypers'], progress=False)
        for o in oracles:
            o.transition(N=1, kernels=['rows', 'column_hypers'])
        compare(eng, oracles, ('trace', it))
