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


def test_engine_reproduces_reference_c1_trajectory():
    """BASELINE.md section 3.5 gate (3) on BASELINE.json configs[0]: 1 000 x 10 mixed table (5 normal +
    5 categorical(k=4)), Engine(num_states=4, rng=gen_rng(1)) with the reference-faithful CRP-prior
    initialisation (views start with up to 504 clusters), ten iterations of rows / columns / view_alphas /
    alpha / column_hypers under the injected numpy stream: partitions, member order, alphas, hypers and
    sufficient statistics equal the reference's after every iteration (sha256 of the canonical snapshot;
    full snapshots at the start, after the first and after the last iteration), and the stream position too."""
    from cgpm_b200.engine import Engine, gen_rng
    from helpers import snapshot_digest
    g = load_golden('trajectory_c1_1000x10.json.gz')
    X, cct, da = golden_table(g)
    S = g['num_states']
    eng = Engine(X, num_states=S, rng=gen_rng(g['engine_seed']), cctypes=cct, distargs=da, stream='numpy')
    for s in range(S):
        assert_snapshot_equal(snapshot_engine(eng, s), g['full']['init'][s], ('init', s))
    for it in range(g['iterations']):
        eng.transition(N=1, kernels=g['kernels'], progress=False)
        for s in range(S):
            snap = snapshot_engine(eng, s)
            if str(it) in g['full']:
                assert_snapshot_equal(snap, g['full'][str(it)][s], (it, s))
            assert snapshot_digest(snap) == g['digests'][it][s], (it, s)
    sc = eng.logpdf_score()
    for s in range(S):
        assert abs(sc[s] - g['logpdf_score'][s]) <= 1e-10 * abs(sc[s])
        assert eng._ch.rngs[s].random_sample() == g['stream_check'][s]
