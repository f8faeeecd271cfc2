"""CPU: the oracle restatement is pinned against (1) vectors generated from the
reference itself (tests/golden/make_golden.py) and (2) the known-answer values
held by the reference's own tests."""
import math
from collections import OrderedDict

import numpy as np
import pytest

import crosscat_oracle as co
from helpers import load_golden, golden_table, snapshot_oracle, assert_snapshot_equal


@pytest.mark.parametrize('fixture', ['trajectory_40x6.json', 'trajectory_lognormal_40x6.json'])
def test_oracle_reproduces_reference_trajectory_bit_for_bit(fixture):
    g = load_golden(fixture)
    X, cct, da = golden_table(g)
    seeds = np.random.RandomState(g['engine_seed']).randint(low=1, high=2 ** 32 - 1, size=g['num_states'])
    states = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd)) for sd in seeds]
    for s, o in enumerate(states):
        assert_snapshot_equal(snapshot_oracle(o), g['init'][s], ('init', s))
        assert o.logpdf_score() == g['init'][s]['logpdf_score']
    for it in range(g['iterations']):
        for s, o in enumerate(states):
            o.transition(N=1, kernels=g['kernels'])
            assert_snapshot_equal(snapshot_oracle(o), g['after'][it][s], (it, s))
            assert o.logpdf_score() == g['after'][it][s]['logpdf_score']
    for q in g['queries']:
        t = {int(k): v for k, v in q['targets'].items()}
        c = {int(k): v for k, v in q['constraints'].items()}
        for s, o in enumerate(states):
            assert o.logpdf(q['rowid'], t, c) == q['logpdf'][s]
    for s, o in enumerate(states):
        assert o.rng.random_sample() == g['stream_check'][s]


def test_oracle_reproduces_reference_c1_trajectory():
    """BASELINE.json configs[0] (1 000 x 10 mixed, 4 chains, reference-faithful CRP-prior initialisation with
    up to ~500 clusters per view, kernels rows/columns/view_alphas/alpha/column_hypers): the first iterations of
    the trajectory generated from the reference (tests/golden/make_golden_c1.py), bit for bit.  The GPU test
    (test_gpu_golden.py) follows all of them."""
    from helpers import snapshot_digest
    g = load_golden('trajectory_c1_1000x10.json.gz')
    X, cct, da = golden_table(g)
    seeds = np.random.RandomState(g['engine_seed']).randint(low=1, high=2 ** 32 - 1, size=g['num_states'])
    chains = (0, 3)                  # the cheapest and the largest initial state (K up to 504); all four on the GPU
    states = {s: co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seeds[s])) for s in chains}
    for s, o in states.items():
        assert_snapshot_equal(snapshot_oracle(o), g['full']['init'][s], ('init', s))
    for it in range(2):
        for s, o in states.items():
            o.transition(N=1, kernels=g['kernels'])
            assert snapshot_digest(snapshot_oracle(o)) == g['digests'][it][s], (it, s)
            if it == 0:
                assert_snapshot_equal(snapshot_oracle(o), g['full']['0'][s], (it, s))


def test_primitive_known_answers():
    k = load_golden('primitive_kats.json')
    for args, ref in k['normal_predictive']:
        assert co.normal_predictive(*args) == ref
    for args, ref in k['normal_marginal']:
        assert co.normal_marginal(*args) == ref
    for args, ref in k['normal_posterior']:
        assert list(co.normal_posterior(*args)) == ref
    for (x, N, counts, a), ref in k['categorical_predictive']:
        assert co.categorical_predictive(x, N, np.array(counts, float), a) == ref
    for (N, counts, a), ref in k['categorical_marginal']:
        assert float(co.categorical_marginal(N, np.array(counts, float), a)) == ref
    counts = OrderedDict([(0, 2), (2, 3), (6, 1)])
    for (x, N, a), ref in k['crp_predictive']:
        assert co.crp_predictive(x, N, counts, a) == ref
    for (N, a), ref in k['crp_marginal']:
        assert float(co.crp_marginal(N, counts, a)) == ref
    rng = np.random.RandomState(42)
    assert [int(co.log_pflip([-1, -2, -.5], array=[5, 7, 9], rng=rng)) for _ in range(8)] == k['log_pflip']
    for args, ref in k['logsumexp']:
        assert co.logsumexp(args) == ref
    grids = co.normal_grids([1.0, 2.5, -0.5, 4.0])
    for n, v in k['normal_grids'].items():
        assert [float(x) for x in grids[n]] == v
    assert [float(x) for x in co.crp_grids(7)] == k['crp_grid']


def relerr(expected, actual):
    return abs((actual - expected) / expected)


def test_reference_test_logsumexp_known_answers():
    """/root/reference/tests/test_logsumexp.py:26-60."""
    inf, nan = float('inf'), float('nan')
    assert co.logsumexp([]) == -inf
    assert co.logsumexp([-1000.]) == -1000.
    assert co.logsumexp([-1000., -1000.]) == -1000. + math.log(2.)
    assert relerr(math.log(2.), co.logsumexp([0., 0.])) < 1e-15
    assert co.logsumexp([-inf, 1]) == 1
    assert co.logsumexp([-inf, -inf]) == -inf
    assert co.logsumexp([+inf, +inf]) == +inf
    assert math.isnan(co.logsumexp([-inf, +inf]))
    assert math.isnan(co.logsumexp([nan, inf]))
    assert math.isnan(co.logsumexp([nan, -3]))
    assert relerr(999.4586751453871, co.logsumexp(range(1000))) < 1e-15
    assert co.logmeanexp([]) == -inf
    assert relerr(992.550919866405, co.logmeanexp(range(1000))) < 1e-15
    assert co.logmeanexp([-1000., -1000.]) == -1000.
    assert relerr(math.log(0.5 * (1 + math.exp(-1.))), co.logmeanexp([0., -1.])) < 1e-15
    assert relerr(math.log(0.5), co.logmeanexp([0., -1000.])) < 1e-15
    assert relerr(-3 - math.log(2.), co.logmeanexp([-inf, -3])) < 1e-15
    assert relerr(-3 - math.log(2.), co.logmeanexp([-3, -inf])) < 1e-15
    assert co.logmeanexp([+inf, -3]) == +inf
    assert co.logmeanexp([-inf, 0, +inf]) == +inf
    assert math.isnan(co.logmeanexp([nan, -3]))
    assert math.isnan(co.logmeanexp([nan]))


def test_reference_test_crp_gibbs_tables():
    """/root/reference/tests/test_crp.py:117-177: counts {0:2, 2:3, 6:1}, alpha 1.5."""
    counts = OrderedDict([(0, 2), (2, 3), (6, 1)])
    K, lp = co.crp_gibbs_logps(counts, 0, 1.5)
    assert K == [0, 2, 6, 7] and np.allclose(np.exp(lp), [1, 3, 1, 1.5])
    K, lp = co.crp_gibbs_logps(counts, 2, 1.5)
    assert K == [0, 2, 6, 7] and np.allclose(np.exp(lp), [2, 2, 1, 1.5])
    K, lp = co.crp_gibbs_logps(counts, 6, 1.5)                # the singleton: no auxiliary table
    assert K == [0, 2, 6] and np.allclose(np.exp(lp), [2, 3, 1.5])
    # marginal = sum of sequential predictives (tests/test_crp.py:180-212)
    data, seen, N, acc = [0, 0, 2, 2, 2, 6], OrderedDict(), 0, 0.0
    for x in data:
        acc += co.crp_predictive(x, N, seen, 1.5)
        seen[x] = seen.get(x, 0) + 1
        N += 1
    assert abs(acc - co.crp_marginal(6, counts, 1.5)) < 1e-12


def test_reference_view_logpdf_cluster_identities():
    """/root/reference/tests/test_view_logpdf_cluster.py:25-92: 5x3 view with NaNs,
    alpha = 2, Zr = [0,0,0,1,1]: CRP prior 3/7, 2/7, 2/7 and sum_k p(k | x) = 1."""
    X = np.array([[1, np.nan, 2, -1, np.nan], [1, 3, 2, -1, -5], [18, -7, -2, 11, -12]], dtype=float).T
    o = co.OState(X, cctypes=['normal'] * 3, Zv={0: 0, 1: 0, 2: 0}, Zrv={0: [0, 0, 0, 1, 1]},
                  view_alphas={0: 2.}, alpha=1., rng=np.random.RandomState(0))
    view = o.views[0]
    K = sorted(view.Nk) + [max(view.Nk) + 1]
    prior = [math.exp(co.crp_predictive(k, 5, view.Nk, 2.)) for k in K]
    assert np.allclose(prior, [3. / 7, 2. / 7, 2. / 7])
    # Bayes rule over clusters for a hypothetical row with evidence on columns 0 and 2
    ev = {0: 1., 2: 3.}
    joint = [co.crp_predictive(k, 5, view.Nk, 2.) + co.OState._row_logpdf(view, ev, k) for k in K]
    post = np.exp(np.asarray(joint) - co.logsumexp(joint))
    assert abs(post.sum() - 1) < 1e-12
    assert abs(o.logpdf(-1, ev) - co.logsumexp(joint)) < 1e-12
    # chain rule p(x0, x2) = p(x0) p(x2 | x0)
    assert abs(o.logpdf(-1, ev) - (o.logpdf(-1, {0: 1.}) + o.logpdf(-1, {2: 3.}, {0: 1.}))) < 1e-10


def test_normal_predictive_agrees_with_student_t_like_the_reference_test():
    """The cross-derivation of the reference's tests/test_teh_murphy.py:69-129 (its hyperparameter sets, its two
    datasets): the Normal-Inverse-Gamma posterior predictive `Z(N+1) - Z(N) - log(2 pi)/2` (normal.py:165-173) is a
    Student-t with nu_n degrees of freedom, location m_n and scale^2 = s_n (r_n + 1) / (nu_n r_n) (Murphy eq. 100).
    The reference allows atol 1e-2; here the tolerance is 1e-9, and the same values are compared with the one-log form
    the CUDA kernels evaluate:
        G(N) - log(s_n)/2 - ((nu_n + 1)/2) log1p(r_n (x - m_n)^2 / ((r_n + 1) s_n)),
        G(N) = lgamma((nu_n + 1)/2) - lgamma(nu_n/2) + log(r_n / (r_n + 1))/2 - log(pi)/2."""
    from scipy.stats import t as student_t
    hypers = [(1., 2., 2., 4.), (7., 18., 6., .6), (.43, 3., 15., 14.), (1.2, 11., 55., 8.)]
    rng = np.random.RandomState(0)
    datasets = [rng.normal(10, 3, size=100), rng.normal(-3, 7, size=100)]
    for (m, r, s, nu) in hypers:
        for x in datasets:
            for n in (0, 1, 17, 100):
                xs = x[:n]
                N, sx, sxx = float(n), float(np.sum(xs)), float(np.sum(xs * xs))
                mn, rn, sn, nun = co.normal_posterior(N, sx, sxx, m, r, s, nu)
                G = math.lgamma(.5 * (nun + 1.)) - math.lgamma(.5 * nun) + .5 * math.log(rn / (rn + 1.)) - .5 * co.LOGPI
                for xt in np.linspace(-30.5, 80.8, 15):
                    ref = co.normal_predictive(float(xt), N, sx, sxx, m, r, s, nu)
                    via_t = student_t.logpdf(xt, nun, loc=mn, scale=math.sqrt(sn * (rn + 1.) / (nun * rn)))
                    kernel_form = G - .5 * math.log(sn) - .5 * (nun + 1.) * math.log1p(rn * (xt - mn) ** 2 / ((rn + 1.) * sn))
                    assert abs(ref - via_t) <= 1e-9 * max(1., abs(ref)), (m, r, s, nu, n, xt, ref, via_t)
                    assert abs(ref - kernel_form) <= 1e-9 * max(1., abs(ref)), (m, r, s, nu, n, xt, ref, kernel_form)


def test_chain_rule_marginal_equals_sum_of_sequential_predictives():
    """The identity behind every collapsed kernel, which the reference pins for the CRP (tests/test_crp.py:180-212:
    marginal = sum of predictives) and, through Bayes / chain-rule identities, for a view
    (tests/test_view_logpdf_cluster.py:45-92): for each primitive on the path, the marginal likelihood of a sequence
    equals the sum of the predictive log-densities of its items taken one after the other, with the sufficient
    statistics updated by the reference's += in between (normal.py:64-70, categorical.py:53-61, crp.py:45-52, ...)."""
    from scipy.special import gammaln
    rng = np.random.RandomState(11)
    tol = lambda ref: 1e-9 * max(1., abs(ref))
    # normal
    for (m, r, s, nu) in [(0., 1., 1., 1.), (.5, 2., 3., 4.), (-2., .1, 20., .7)]:
        xs = rng.normal(1.5, 2.0, size=40)
        N = sx = sxx = 0.
        acc = 0.
        for x in xs:
            acc += co.normal_predictive(float(x), N, sx, sxx, m, r, s, nu)
            N += 1; sx += x; sxx += x * x
        ref = co.normal_marginal(N, sx, sxx, m, r, s, nu)
        assert abs(acc - ref) <= tol(ref), ('normal', acc, ref)
    # categorical
    for k, alpha in [(2, 1.), (4, .7), (16, 3.)]:
        xs = rng.randint(k, size=60)
        counts = np.zeros(k)
        N, acc = 0, 0.
        for x in xs:
            acc += co.categorical_predictive(int(x), N, counts, alpha)
            counts[x] += 1; N += 1
        ref = co.categorical_marginal(N, counts, alpha)
        assert abs(acc - ref) <= tol(ref), ('categorical', acc, ref)
    # bernoulli, exponential, geometric, poisson
    xs = rng.randint(2, size=50)
    N = sx = 0; acc = 0.
    for x in xs:
        acc += co.bernoulli_predictive(int(x), N, sx, 1.3, .4); N += 1; sx += x
    ref = co.bernoulli_marginal(N, sx, 1.3, .4)
    assert abs(acc - ref) <= tol(ref), ('bernoulli', acc, ref)
    xs = rng.exponential(2.5, size=50)
    N = 0; sx = 0.; acc = 0.
    for x in xs:
        acc += co.exponential_predictive(float(x), N, sx, 1.5, .8); N += 1; sx += x
    ref = co.exponential_marginal(N, sx, 1.5, .8)
    assert abs(acc - ref) <= tol(ref), ('exponential', acc, ref)
    xs = rng.geometric(.3, size=50) - 1
    N = sx = 0; acc = 0.
    for x in xs:
        acc += co.geometric_predictive(int(x), N, sx, 2., 1.5); N += 1; sx += x
    ref = co.geometric_marginal(N, sx, 2., 1.5)
    assert abs(acc - ref) <= tol(ref), ('geometric', acc, ref)
    xs = rng.poisson(4., size=50)
    N = sx = 0; slf = 0.; acc = 0.
    for x in xs:
        acc += co.poisson_predictive(int(x), N, sx, 2., .5); N += 1; sx += x; slf += gammaln(x + 1)
    ref = co.poisson_marginal(N, sx, slf, 2., .5)
    assert abs(acc - ref) <= tol(ref), ('poisson', acc, ref)
    # CRP: seat the customers of a partition one by one
    for alpha in (.3, 1.5, 20.):
        z = [0, 0, 1, 0, 2, 1, 1, 3, 0, 2, 4, 4, 0]
        counts, N, acc = OrderedDict(), 0, 0.
        for t in z:
            acc += co.crp_predictive(t, N, counts, alpha)
            counts[t] = counts.get(t, 0) + 1; N += 1
        ref = co.crp_marginal(N, counts, alpha)
        assert abs(acc - ref) <= tol(ref), ('crp', alpha, acc, ref)
