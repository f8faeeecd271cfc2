"""Posterior summaries of the device-generator (philox) engine agree statistically with
the reference's (fixture generated from the reference by tests/golden/make_posterior_golden.py):
number of views, clusters per view, logpdf_score, dependence probabilities."""
import numpy as np
import pytest
from scipy import stats

from helpers import load_golden

pytestmark = pytest.mark.gpu


def test_posterior_summaries_match_reference_fixture():
    from cgpm_b200 import synth
    from cgpm_b200.engine import Engine, gen_rng
    g = load_golden('posterior_summary.json')
    da = [d if d else None for d in g['distargs']]
    X, Zv, _ = synth.gen_table(g['R'], g['cctypes'], da, n_views=2, clusters_per_view=3, separation=0.9,
                               seed=g['table_seed'])
    assert [int(z) for z in Zv] == g['true_Zv']
    S = 256
    eng = Engine(X, num_states=S, rng=gen_rng(23), cctypes=g['cctypes'], distargs=da, stream='philox')
    eng.transition(N=g['N'], kernels=g['kernels'], progress=False)
    ch = eng._ch
    nv = ch.dev.t['nv'].cpu().numpy(); nclu = ch.dev.t['nclu'].cpu().numpy()
    n_views = (nv > 0).sum(axis=1)
    n_clusters = np.array([nclu[s][nv[s] > 0].mean() for s in range(S)])
    scores = np.array(eng.logpdf_score())
    ref_v, ref_k, ref_s = np.array(g['n_views']), np.array(g['n_clusters']), np.array(g['logpdf_score'])
    # two-sample tests (reference: 64 chains; ours: 256 chains)
    for name, a, b in (('views', n_views, ref_v), ('clusters', n_clusters, ref_k), ('score', scores, ref_s)):
        se = np.sqrt(a.var() / len(a) + b.var() / len(b))
        assert abs(a.mean() - b.mean()) <= 4 * se + 1e-9, (name, a.mean(), b.mean(), se)
    assert stats.ks_2samp(scores, ref_s).pvalue > 1e-3
    D = np.mean(eng.dependence_probability_pairwise(), axis=0)
    Dref = np.array(g['dependence'])
    se = np.sqrt(Dref * (1 - Dref) / 64 + D * (1 - D) / S) + 0.02
    assert np.all(np.abs(D - Dref) <= 4 * se), (D, Dref)
    lp = np.mean(eng.logpdf(-1, {0: float(np.nanmean(X[:, 0])), 3: float(np.nanmean(X[:, 3]))}))
    assert abs(lp - g['probe_logpdf_mean'][0]) < 0.25
