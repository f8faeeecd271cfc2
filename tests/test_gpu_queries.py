"""GPU parity of logpdf / simulate / scores / dependence queries vs the oracle."""
import numpy as np
import pytest

import crosscat_oracle as co
from helpers import small_table

pytestmark = pytest.mark.gpu


def make(n_rows=80, seed=31, S=3, iters=3):
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = small_table(n_rows, seed)
    eng = Engine(X, num_states=S, rng=gen_rng(seed), cctypes=cct, distargs=da, stream='numpy')
    seeds = gen_rng(seed).randint(low=1, high=2 ** 32 - 1, size=S)
    oracles = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd)) for sd in seeds]
    eng.transition(N=iters, progress=False)
    for o in oracles:
        o.transition(N=iters)
    return X, cct, da, eng, oracles


def test_logpdf_bulk_matches_oracle():
    X, cct, da, eng, oracles = make()
    nanrow = int(np.argwhere(np.isnan(X))[0][0]); nancol = int(np.argwhere(np.isnan(X))[0][1])
    val = 1.0 if cct[nancol] == 'categorical' else 0.3
    queries = [
        (-1, {0: 1.2}, {}),
        (-1, {0: 1.2, 1: 2}, {2: 0.3}),
        (-1, {3: 1, 5: -0.5}, {0: 2., 4: 1., 1: 0}),
        (None, {2: 4.0, 4: 0.0, 0: 1.0}, None),
        (nanrow, {nancol: val}, None),
        (-1, {1: 7}, {}),                # invalid category -> -inf
        (-1, {0: float('nan'), 2: 1.0}, {}),
    ]
    rowids = [q[0] for q in queries]; T = [q[1] for q in queries]; Cn = [q[2] for q in queries]
    got = eng.logpdf_bulk(rowids, T, Cn)
    for s, o in enumerate(oracles):
        for i, (r, t, c) in enumerate(queries):
            ref = o.logpdf(r, t, c)
            if np.isinf(ref):
                assert got[s][i] == ref
            else:
                assert abs(got[s][i] - ref) <= 1e-9 * max(1.0, abs(ref)), (s, i, got[s][i], ref)
    one = eng.logpdf(-1, {0: 1.2, 1: 2}, {2: 0.3})
    assert one == [g[1] for g in got]
    st = eng.get_state(0)
    assert abs(st.logpdf(-1, {0: 1.2}) - got[0][0]) < 1e-12


def test_logpdf_zero_density_constraints_raise():
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = small_table(50, 5)
    Zv = {c: 0 for c in range(len(cct))}              # one view: constraints share the targets' view
    eng = Engine(X, num_states=1, rng=gen_rng(2), cctypes=cct, distargs=da, Zv=Zv, stream='numpy')
    o = co.OState(X, cctypes=cct, distargs=da, Zv=Zv, rng=np.random.RandomState(3))
    with pytest.raises(ValueError):
        o.logpdf(-1, {0: 1.0}, {1: 9})
    with pytest.raises(ValueError):
        eng.logpdf(-1, {0: 1.0}, {1: 9})               # sampling.py:78-79
    with pytest.raises(ValueError):
        eng.logpdf(0, {0: 1.0})                        # observed, non-NaN cell (state.py:463-466)
    with pytest.raises(ValueError):
        eng.logpdf(-1, {0: 1.0}, {0: 1.0})
    with pytest.raises(ValueError):
        Engine(X, num_states=1, cctypes=['poisson'] + cct[1:], distargs=da)


def test_simulate_statistics():
    """Predictive samples agree in distribution with the oracle's predictive:
    compare the sample mean log-density and a KS test against a Monte-Carlo
    draw from the oracle's posterior-predictive mixture."""
    from scipy import stats
    X, cct, da, eng, oracles = make(n_rows=120, seed=41, S=2, iters=4)
    N = 4000
    samples = eng.simulate(-1, [0, 1], {2: float(np.nanmean(X[:, 2]))}, N=N)
    rng = np.random.RandomState(0)
    for s, o in enumerate(oracles):
        xs = np.array([d[0] for d in samples[s]]); cs = np.array([d[1] for d in samples[s]])
        assert set(np.unique(cs)) <= set(range(4))
        # density check: E_p[log p(x)] estimated from our samples vs importance-free estimate by the oracle
        grid = np.linspace(xs.min() - 1, xs.max() + 1, 400)
        pdf = np.exp([o.logpdf(-1, {0: g}, {2: float(np.nanmean(X[:, 2]))}) for g in grid])
        cdf = np.cumsum(pdf) * (grid[1] - grid[0]); cdf /= cdf[-1]
        ks = stats.kstest(xs, lambda v: np.interp(v, grid, cdf))
        assert ks.pvalue > 1e-3, ks
        pc = np.exp([o.logpdf(-1, {1: k}, {2: float(np.nanmean(X[:, 2]))}) for k in range(4)])
        freq = np.bincount(cs, minlength=4) / float(N)
        assert np.all(np.abs(freq - pc) < 5 * np.sqrt(pc * (1 - pc) / N) + 5e-3), (freq, pc)
    a = eng.simulate(-1, [0], N=3); b = eng.simulate(-1, [0], N=3)
    assert a != b                                        # engine.py:232: reseeded every call
    single = eng.simulate(-1, [0, 3])
    assert isinstance(single[0], dict) and set(single[0]) == {0, 3}
    obs = eng.simulate(5, [0], N=2)
    assert len(obs) == 2 and len(obs[0]) == 2


def test_scores_and_dependence():
    X, cct, da, eng, oracles = make()
    sc = eng.logpdf_score()
    for s, o in enumerate(oracles):
        assert abs(sc[s] - o.logpdf_score()) <= 1e-10 * abs(sc[s])
        D = eng.dependence_probability_pairwise()[s]
        for i in range(len(cct)):
            for j in range(len(cct)):
                assert D[i, j] == o.dependence_probability(i, j)
    assert eng.dependence_probability(0, 1) == [o.dependence_probability(0, 1) for o in oracles]
    lw = eng._likelihood_weighted_integrate(eng.logpdf(-1, {0: 1.0}), -1)
    assert np.isfinite(lw)


def test_metadata_round_trip_and_reference_schema():
    from cgpm_b200.engine import Engine, State, gen_rng
    X, cct, da, eng, oracles = make(S=2, iters=2)
    md = eng.to_metadata()
    assert set(md) == {'X', 'states', 'factory'} and 'X' not in md['states'][0]
    keys = {'outputs', 'alpha', 'Zv', 'cctypes', 'hypers', 'distargs', 'suffstats', 'Cd', 'Ci', 'Zrv',
            'view_alphas', 'diagnostics', 'hooked_cgpms', 'loom_path', 'factory'}
    assert set(md['states'][0]) == keys
    om = oracles[0].to_metadata()
    m0 = md['states'][0]
    assert sorted(m0['Zv']) == sorted(om['Zv']) and m0['Zrv'] == om['Zrv']
    for a, b in zip(m0['suffstats'], om['suffstats']):            # full sufficient statistics, bit for bit
        da_, db_ = dict(a), dict(b)
        assert set(da_) == set(db_)
        for k in da_:
            assert set(da_[k]) == set(db_[k]), (k, da_[k], db_[k])
            for f in da_[k]:
                assert np.array_equal(np.asarray(da_[k][f], dtype=float), np.asarray(db_[k][f], dtype=float)), (k, f)
    eng2 = Engine.from_metadata(md, rng=gen_rng(3), stream='numpy')
    assert np.allclose(eng2.logpdf_score(), eng.logpdf_score(), rtol=1e-12)
    import io, pickle, json
    json.dumps(md['states'][0]['Zrv'])
    buf = io.BytesIO(); eng.to_pickle(buf); buf.seek(0)
    eng3 = Engine.from_pickle(buf, rng=gen_rng(3))
    assert np.allclose(eng3.logpdf_score(), eng.logpdf_score(), rtol=1e-12)
    st = State.from_metadata(eng.get_state(1).to_metadata())
    assert abs(st.logpdf_score() - eng.logpdf_score()[1]) <= 1e-12 * abs(st.logpdf_score())
    eng.add_state(count=2)
    assert eng.num_states() == 4
    eng.drop_state(0)
    assert eng.num_states() == 3
    eng.transition(N=1, progress=False)


def test_row_similarity_relevance_and_mutual_information():
    """SURVEY 8(f) rank 4, built on the query kernels."""
    X, cct, da, eng, oracles = make(n_rows=100, seed=51, S=2, iters=6)
    for s, o in enumerate(oracles):
        views = set(o.Zv[c] for c in o.outputs)
        ref = np.mean([o.views[v].Zr[3] == o.views[v].Zr[7] for v in views])
        assert eng.row_similarity(3, 7)[s] == ref
        P = eng.row_similarity_pairwise()[s]
        assert P.shape == (100, 100) and np.allclose(np.diag(P), 1.0) and P[3, 7] == ref and np.allclose(P, P.T)
        v = o.Zv[0]
        same = [r for r in range(100) if o.views[v].Zr[r] == o.views[v].Zr[5]][:3]
        assert eng.relevance_probability(5, same, 0)[s] == 1
    # mutual information: zero across views, positive entropy, and close to a brute-force
    # integral of the oracle's densities for a dependent pair of normal columns
    st = eng.get_state(0)
    o = oracles[0]
    zv = o.Zv
    normal_cols = [c for c in range(len(cct)) if cct[c] == 'normal']
    dep = [(a, b) for a in normal_cols for b in normal_cols if a < b and zv[a] == zv[b]]
    ind = [(a, b) for a in normal_cols for b in normal_cols if a < b and zv[a] != zv[b]]
    if ind:
        assert st.mutual_information([ind[0][0]], [ind[0][1]], N=50) == 0
    assert st.mutual_information([normal_cols[0]], [normal_cols[0]], N=200) > 0          # entropy
    if dep:
        a, b = dep[0]
        est = np.mean([st.mutual_information([a], [b], N=2000) for _ in range(3)])
        xa = X[:, a][~np.isnan(X[:, a])]; xb = X[:, b][~np.isnan(X[:, b])]
        ga = np.linspace(xa.min() - 4, xa.max() + 4, 70); gb = np.linspace(xb.min() - 4, xb.max() + 4, 70)
        pa = np.array([o.logpdf(-1, {a: x}) for x in ga]); pb = np.array([o.logpdf(-1, {b: y}) for y in gb])
        pab = np.array([[o.logpdf(-1, {a: x, b: y}) for y in gb] for x in ga])
        w = (ga[1] - ga[0]) * (gb[1] - gb[0])
        ref = float(np.sum(np.exp(pab) * (pab - pa[:, None] - pb[None, :])) * w)
        assert abs(est - ref) < 0.12 + 0.15 * abs(ref), (est, ref)


def test_relevance_probability_with_hypothetical_rows():
    """state.py:599-629: hypothetical rows are incorporated into the column's view with one Gibbs step each; the
    engine (numpy stream) gives the oracle's answers and consumes the chains' RandomStates identically."""
    X, cct, da, eng, oracles = make(n_rows=60, seed=71, S=2, iters=3)
    hyp = [{0: 0.3, 2: -0.2}, {4: 1.0}, {1: 2, 3: 0}, {0: float('nan')}]
    cases = [(3, [5, 8], 0), (3, [], 2), (10, [11], 4), (0, [1, 2, 3], 1), (7, [7], 5)]
    for tgt, qry, col in cases:
        got = eng.relevance_probability(tgt, qry, col, hypotheticals=hyp)
        for s, o in enumerate(oracles):
            assert got[s] == o.relevance_probability(tgt, qry, col, hyp), (tgt, qry, col, s)
    for s, o in enumerate(oracles):
        assert eng._ch.rngs[s].random_sample() == o.rng.random_sample()
    st = eng.get_state(0)
    assert st.relevance_probability(3, [5], 0, hypotheticals=[{0: 0.1}]) in (0, 1)
    assert eng.num_states() == 2 and eng.X.shape == (60, len(cct))          # nothing was left behind
