"""GPU vs vectors generated from the reference itself (no oracle in the loop):
Engine(stream='numpy') with the reference's seeds reproduces the reference's
partitions, hyperparameters and sufficient statistics bit for bit."""
import numpy as np
import pytest

from helpers import load_golden, golden_table, snapshot_engine, assert_snapshot_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('fixture', ['trajectory_40x6.json', 'trajectory_lognormal_40x6.json'])
def test_engine_reproduces_reference_golden_trajectory(fixture):
    from cgpm_b200.engine import Engine, gen_rng
    g = load_golden(fixture)
    X, cct, da = golden_table(g)
    S = g['num_states']
    eng = Engine(X, num_states=S, rng=gen_rng(g['engine_seed']), cctypes=cct, distargs=da, stream='numpy')
    for s in range(S):
        assert_snapshot_equal(snapshot_engine(eng, s), g['init'][s], ('init', s))
    sc = eng.logpdf_score()
    for s in range(S):
        assert abs(sc[s] - g['init'][s]['logpdf_score']) <= 1e-10 * abs(sc[s])
    for it in range(g['iterations']):
        eng.transition(N=1, kernels=g['kernels'], progress=False)
        sc = eng.logpdf_score()
        for s in range(S):
            assert_snapshot_equal(snapshot_engine(eng, s), g['after'][it][s], (it, s))
            assert abs(sc[s] - g['after'][it][s]['logpdf_score']) <= 1e-10 * abs(sc[s])
    for q in g['queries']:
        t = {int(k): v for k, v in q['targets'].items()}
        c = {int(k): v for k, v in q['constraints'].items()}
        got = eng.logpdf(q['rowid'], t, c)
        for s in range(S):
            assert abs(got[s] - q['logpdf'][s]) <= 1e-9 * max(1.0, abs(q['logpdf'][s]))
    for s in range(S):
        assert eng._ch.rngs[s].random_sample() == g['stream_check'][s]
