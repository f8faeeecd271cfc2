"""End-to-end parity: Engine(stream='numpy') follows the oracle's trajectory for
the same seeds -- initial state, then every kernel of State.transition."""
import numpy as np
import pytest

import crosscat_oracle as co
from helpers import small_table, oracle_suffstats, device_suffstats, assert_suffstats_equal, rel_tol

pytestmark = pytest.mark.gpu

KERNELS = ['alpha', 'view_alphas', 'column_hypers', 'rows', 'columns']


def compare(eng, oracles, tag):
    ch = eng._ch
    # (with Ci the reference refuses to score the column CRP, state.py:370-372; the oracle returns the plain score)
    scores = ch.logpdf_score() if not ch.Ci else ch.dev.logpdf_score().cpu().numpy() + ch._ln_const
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
                assert np.all(np.abs(got - ref) <= rel_tol(ref)), (got, ref)
        eng.transition(N=1, kernels=['rows', 'column_hypers'], progress=False)
        for o in oracles:
            o.transition(N=1, kernels=['rows', 'column_hypers'])
        compare(eng, oracles, ('trace', it))


def test_aux_member_draw_partition_law_and_marginal():
    """Device stream: the auxiliary view's partition is drawn by 'join the table of a uniformly
    chosen earlier row' (columns.cu).  Check the structure (table ids in order of appearance,
    stored K), the law (E[K] = sum_i alpha / (alpha + i), Var[K] = sum_i p_i (1 - p_i)), and the
    column's marginal likelihood under the drawn partition against the oracle's Dim."""
    import ctypes as C
    from cgpm_b200.core import Chains
    from cgpm_b200 import _lib as L
    from cgpm_b200.synth import gen_table
    R, S = 3000, 24
    cct = ['normal', 'categorical', 'normal', 'normal']
    da = [None, {'k': 3}, None, None]
    X, _, _ = gen_table(R, cct, da, n_views=2, clusters_per_view=3, seed=3)
    X[5, 0] = np.nan
    ch = Chains(X, cctypes=cct, distargs=da, S=S, seed=11, stream='philox')
    for s in range(S):
        ch.init_chain(s, np.random.RandomState(50 + s), Zv={c: 0 for c in range(len(cct))}, Zrv={0: [0] * R},
                      alpha=1.0, view_alphas={0: 1.0})
    ch.finish_init()
    dev, st, lib = ch.dev, C.byref(ch.dev.st), ch.dev.lib
    ncol = len(cct)
    ch._ensure_arena(S * ncol)
    assert ch._ensure_aux_store()
    _p = lambda a: a.ctypes.data_as(C.c_void_p)
    hyp = dev.t['hyp'].cpu().numpy().reshape(S, ncol, 4)
    for ai in (8, 15, 22):
        a_ch = np.repeat(np.arange(S, dtype=np.int32), ncol)
        a_co = np.tile(np.arange(ncol, dtype=np.int32), S)
        a_ai = np.full(S * ncol, ai, np.int32)
        L.check(lib.cc_col_aux(st, len(a_ch), _p(a_ch), _p(a_co), _p(a_ai), None, 2, C.c_uint64(123 + ai),
                               C.c_uint64(7), dev.stream_ptr()), 'cc_col_aux')
        dev.sync()
        zr = dev.aux_zr_all.cpu().numpy().reshape(S, ncol, R)
        K = dev.t['aux_K_all'].cpu().numpy().reshape(S, ncol)
        alpha = dev.t['aux_alpha'].cpu().numpy().reshape(S, ncol)
        colL = dev.t['colL'].cpu().numpy().reshape(S, ncol, dev.Vcap + 1)[:, :, dev.Vcap]
        a = float(alpha[0, 0])
        assert np.all(alpha == a)
        i = np.arange(R)
        pnew = a / (a + i)
        mean, var = pnew.sum(), (pnew * (1 - pnew)).sum()
        for s in range(S):
            for c in range(ncol):
                z = zr[s, c]
                # ids appear in increasing order: a new maximum is always the next integer
                run_max = np.maximum.accumulate(z)
                assert z[0] == 0 and np.all(np.diff(run_max) <= 1) and run_max[-1] == K[s, c] - 1
        ks = K.reshape(-1).astype(float)
        assert abs(ks.mean() - mean) < 4.5 * np.sqrt(var / ks.size) + 1e-9, (ai, ks.mean(), mean)
        assert 0.5 * var < ks.var() + 1e-9 < 1.8 * var + 1.0, (ai, ks.var(), var)
        # marginal of a few (chain, column) pairs under their partition: the oracle's Dim
        for s, c in [(0, 0), (1, 1), (2, 2), (S - 1, 3)]:
            ref = oracle_marginal(X[:, c], zr[s, c], cct[c], da[c], hyp[s, c])
            assert abs(colL[s, c] - ref) <= 1e-9 * max(1.0, abs(ref)), (ai, s, c, colL[s, c], ref)


def oracle_marginal(x, z, cctype, distargs, hyp):
    """Sum over tables of the cluster marginal likelihood (normal.py:175-181, categorical.py:144-155)."""
    from math import lgamma, log, pi
    tot = 0.0
    for k in np.unique(z):
        xs = x[z == k]
        xs = xs[~np.isnan(xs)]
        if cctype == 'normal':
            m, r, s, nu = [float(v) for v in hyp]
            N = len(xs); sx = float(np.sum(xs)); sxx = float(np.sum(xs * xs))
            rn = r + N; nun = nu + N; mn = (r * m + sx) / rn
            sn = s + sxx + r * m * m - rn * mn * mn
            if sn == 0: sn = s
            Z = lambda r_, s_, nu_: ((nu_ + 1.0) / 2.0) * log(2) + 0.5 * log(pi) - 0.5 * log(r_) - (nu_ / 2.0) * log(s_) + lgamma(nu_ / 2.0)
            tot += -(N / 2.0) * log(2 * pi) + Z(rn, sn, nun) - Z(r, s, nu)
        else:
            a = float(hyp[0]); kc = int(distargs['k'])
            cnt = np.bincount(xs.astype(int), minlength=kc)
            N = cnt.sum(); A = kc * a
            tot += lgamma(A) - lgamma(A + N) + sum(lgamma(c_ + a) for c_ in cnt) - kc * lgamma(a)
    return tot


@pytest.mark.parametrize('n_rows,seed', [(40, 31), (70, 32)])
def test_engine_with_lognormal_columns_follows_oracle(n_rows, seed):
    """SURVEY 8(f) rank 3 (lognormal): the column is a normal column on log x on the device; hyper grids, the order in
    which a fresh Dim draws its hyperparameters, the -log x Jacobian of logpdf / logpdf_score, simulate on the
    original scale and the metadata schema follow src/primitives/lognormal.py.  Engine(stream='numpy') must follow
    the oracle (itself pinned against the reference on CPU) kernel by kernel."""
    from cgpm_b200.engine import Engine, gen_rng, _chain_metadata
    X, cct, da = small_table(n_rows, seed)
    X = np.array(X)
    for c in (2, 5):
        X[:, c] = np.exp(0.4 * X[:, c])
        cct[c] = 'lognormal'
    S = 2
    eng = Engine(X, num_states=S, rng=gen_rng(seed), cctypes=cct, distargs=da, stream='numpy')
    seeds = gen_rng(seed).randint(low=1, high=2 ** 32 - 1, size=S)
    oracles = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd)) for sd in seeds]
    compare(eng, oracles, 'init')
    for it in range(3):
        for k in KERNELS:
            eng.transition(N=1, kernels=[k], progress=False)
            for o in oracles:
                o.transition(N=1, kernels=[k])
            compare(eng, oracles, (it, k))
    # densities on the original scale: targets and constraints in lognormal columns, a non-positive target
    queries = [({2: 1.7, 0: 0.3}, {5: 0.8}), ({5: 2.5}, {2: 0.4, 1: 2}), ({0: 0.1}, {2: 3.0})]
    for t, cns in queries:
        got = eng.logpdf(-1, t, cns)
        for s, o in enumerate(oracles):
            ref = o.logpdf(-1, t, cns)
            assert abs(got[s] - ref) <= 1e-9 * max(1.0, abs(ref)), (t, cns, got[s], ref)
    assert all(v == float('-inf') for v in eng.logpdf(-1, {2: -1.0}))
    # metadata: lognormal suffstat names, original X, cctypes
    md = eng.to_metadata()
    assert md['states'][0]['cctypes'][2] == 'lognormal'
    assert set(md['states'][0]['suffstats'][2][0][1]) == {'N', 'sum_log_x', 'sum_log_x_sq'}
    assert np.allclose(np.array(md['X'], dtype=float)[:, 2][~np.isnan(X[:, 2])], X[:, 2][~np.isnan(X[:, 2])], rtol=0, atol=0)
    eng2 = Engine.from_metadata(md, rng=gen_rng(1))
    assert np.allclose(eng2.logpdf_score(), eng.logpdf_score(), rtol=1e-12)
    # incremental edits keep the original scale on the host and log x on the device
    def both(tag, fe, fo):
        fe(eng)
        for o in oracles:
            fo(o)
        compare(eng, oracles, tag)
    R = n_rows
    both('inc', lambda e: e.incorporate(R, {0: 0.5, 2: 2.5, 5: 0.3}), lambda o: o.incorporate(R, {0: 0.5, 2: 2.5, 5: 0.3}))
    both('inc2', lambda e: e.incorporate(R + 1, {2: 0.7, 4: 0.1}), lambda o: o.incorporate(R + 1, {2: 0.7, 4: 0.1}))
    both('force', lambda e: e.force_cell(R + 1, {5: 1.9}), lambda o: o.force_cell(R + 1, {5: 1.9}))
    both('t', lambda e: e.transition(N=1, kernels=KERNELS, progress=False), lambda o: o.transition(N=1, kernels=KERNELS))
    both('uninc', lambda e: e.unincorporate(R + 1), lambda o: o.unincorporate(R + 1))
    T = list(np.exp(np.random.RandomState(3).normal(size=R + 1)))
    both('dim+', lambda e: e.incorporate_dim(T, [90], cctype='lognormal', v=99), lambda o: o.incorporate_dim(T, [90], cctype='lognormal', v=99))
    both('t2', lambda e: e.transition(N=1, kernels=KERNELS, progress=False), lambda o: o.transition(N=1, kernels=KERNELS))
    both('dim-', lambda e: e.unincorporate_dim(2), lambda o: o.unincorporate_dim(2))
    assert eng.X.shape == (R + 1, 6) and np.all(eng.X[~np.isnan(eng.X[:, 4]), 4] > 0)      # column 5 stays on the original scale
    with pytest.raises(ValueError):
        eng.incorporate(R + 1, {5: -1.0})
    for s, o in enumerate(oracles):
        assert eng._ch.rngs[s].random_sample() == o.rng.random_sample()
        assert eng._ch.diagnostics[s]['iterations'] == o.iterations
    # simulate (last: it consumes the chains' RandomStates) returns positive values for lognormal columns
    smp = eng.simulate(-1, [90, 5], N=200)
    a = np.array([[d[90], d[5]] for d in smp[0]])
    assert np.all(a > 0) and np.isfinite(a).all()
