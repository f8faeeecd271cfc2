"""One small invocation of the hot path on cuda:0, checked against the CPU oracle
(used by __graft_entry__.smoke()).  The oracle is the checker, never the product."""
import numpy as np

import crosscat_oracle as co
from helpers import small_table, oracle_suffstats, device_suffstats, assert_suffstats_equal


def run():
    import torch
    assert torch.cuda.is_available(), 'smoke() needs a CUDA device'
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = small_table(50, 77)
    S = 2
    eng = Engine(X, num_states=S, rng=gen_rng(5), cctypes=cct, distargs=da, stream='numpy')
    seeds = gen_rng(5).randint(low=1, high=2 ** 32 - 1, size=S)
    oracles = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd)) for sd in seeds]
    kernels = ['rows', 'columns', 'view_alphas', 'alpha', 'column_hypers']
    eng.transition(N=2, kernels=kernels, progress=False)
    for o in oracles:
        o.transition(N=2, kernels=kernels)
    scores = eng.logpdf_score()
    for s, o in enumerate(oracles):
        got, dl = device_suffstats(eng._ch.dev, s, eng._ch.ctype)
        assert_suffstats_equal(oracle_suffstats(o), got)                 # bit-exact
        for j, (vid, view) in enumerate(o.views.items()):
            assert np.array_equal(dl['cid_by_slot'][j][dl['zr'][j]], np.array([view.Zr[i] for i in range(50)]))
        ref = o.logpdf_score()
        assert abs(scores[s] - ref) <= 1e-10 * abs(ref), (scores[s], ref)
    lp = eng.logpdf(-1, {0: 1.0, 1: 2})
    for s, o in enumerate(oracles):
        ref = o.logpdf(-1, {0: 1.0, 1: 2})
        assert abs(lp[s] - ref) <= 1e-9 * max(1.0, abs(ref))
    # the throughput (device-generator) path on the same engine class
    eng2 = Engine(X, num_states=4, rng=gen_rng(6), cctypes=cct, distargs=da, stream='philox')
    eng2.transition(N=3, kernels=['rows', 'columns'], progress=False)
    assert all(np.isfinite(eng2.logpdf_score()))
    return True
