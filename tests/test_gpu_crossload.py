"""GPU <-> reference cross-loading of the serialised state: metadata written by the GPU engine
loads into the REFERENCE Engine (oracle/_ref, the py3-shimmed /root/reference -- it travels to the GPU box) and
gives the same scores and densities there; a pickle written by the reference loads onto the GPU likewise.
src/crosscat/state.py:1189-1284, src/crosscat/engine.py:394-435."""
import io
import os
import pickle
import sys

import numpy as np
import pytest

from helpers import small_table

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'cgpm', '.built')),
                                 reason='oracle/_ref not built (reference tree absent)')]

QUERIES = [(-1, {0: 1.2}, {}), (-1, {0: 1.2, 1: 2}, {2: 0.3}), (-1, {3: 1, 5: -0.5}, {0: 2., 4: 1., 1: 0}), (3, {0: 0.5}, {})]


def table(n, seed):
    X, cct, da = small_table(n, seed)
    X[3, 0] = np.nan                 # the observed-row query above asks for this cell
    return X, cct, da


def ref_modules():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from cgpm.crosscat.engine import Engine as REngine
    from cgpm.crosscat.state import State as RState
    from cgpm.utils import general as gu
    return REngine, RState, gu


def stats_of(md_state):
    out = []
    for col in md_state['suffstats']:
        out.append(sorted((int(k), sorted((f, np.asarray(v, dtype=float).tolist()) for f, v in st.items())) for k, st in col))
    return out


def assert_stats_close(a, b, rtol=1e-12):
    """Same clusters and fields; sums equal up to the order of summation (a state loaded from metadata re-sums
    every cluster in row order, state.py:1246-1267, while a transitioned state carries incrementally updated sums)."""
    assert len(a) == len(b)
    for ca, cb in zip(a, b):
        assert [k for k, _ in ca] == [k for k, _ in cb]
        for (_, fa), (_, fb) in zip(ca, cb):
            assert [f for f, _ in fa] == [f for f, _ in fb]
            for (_, va), (_, vb) in zip(fa, fb):
                assert np.allclose(va, vb, rtol=rtol, atol=0.0), (va, vb)


def test_gpu_metadata_loads_into_the_reference():
    from cgpm_b200.engine import Engine, gen_rng
    REngine, RState, gu = ref_modules()
    X, cct, da = table(60, 31)
    eng = Engine(X, num_states=3, rng=gen_rng(4), cctypes=cct, distargs=da, stream='numpy')
    eng.transition(N=3, progress=False)
    md = eng.to_metadata()
    ref = REngine.from_metadata(pickle.loads(pickle.dumps(md)), rng=gu.gen_rng(2), multiprocess=0)
    assert ref.num_states() == 3
    got, want = eng.logpdf_score(), ref.logpdf_score(multiprocess=0)
    for s in range(3):
        assert abs(got[s] - want[s]) <= 1e-10 * abs(want[s])
    # the reference rebuilds the statistics from (X, Zv, Zrv) in row order: bit for bit what the GPU engine rebuilds
    # from the same metadata, and equal to the transitioned engine's own sums up to the order of summation
    rmd = ref.to_metadata()
    gmd = Engine.from_metadata(md, rng=gen_rng(9), stream='numpy').to_metadata()
    for s in range(3):
        assert sorted(md['states'][s]['Zv']) == sorted(rmd['states'][s]['Zv'])
        assert [(v, list(z)) for v, z in md['states'][s]['Zrv']] == [(v, list(z)) for v, z in rmd['states'][s]['Zrv']]
        assert stats_of(gmd['states'][s]) == stats_of(rmd['states'][s])
        assert_stats_close(stats_of(md['states'][s]), stats_of(rmd['states'][s]))
        assert [dict(h) for h in md['states'][s]['hypers']] == [dict(h) for h in rmd['states'][s]['hypers']]
    for rowid, t, c in QUERIES:
        a = eng.logpdf(rowid, t, c)
        b = ref.logpdf(rowid, t, c, multiprocess=0)
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12)
    # and the reference keeps running from it
    ref.transition(N=1, progress=False, multiprocess=0)
    # single states too (State.to_metadata carries X)
    st = eng.get_state(1)
    rst = RState.from_metadata(st.to_metadata(), rng=gu.gen_rng(0))
    assert abs(rst.logpdf_score() - st.logpdf_score()) <= 1e-10 * abs(rst.logpdf_score())


def test_reference_pickle_loads_onto_the_gpu(tmp_path):
    from cgpm_b200.engine import Engine, State, gen_rng
    REngine, RState, gu = ref_modules()
    X, cct, da = table(60, 32)
    ref = REngine(X, num_states=3, rng=gu.gen_rng(7), cctypes=cct, distargs=da, multiprocess=0)
    ref.transition(N=3, progress=False, multiprocess=0)
    buf = io.BytesIO()
    pickle.dump(ref.to_metadata(), buf)            # engine.py:429-431 (binary mode: the reference's text-mode file is a py2-ism)
    buf.seek(0)
    eng = Engine.from_pickle(buf, rng=gen_rng(1), stream='numpy')
    got, want = eng.logpdf_score(), ref.logpdf_score(multiprocess=0)
    for s in range(3):
        assert abs(got[s] - want[s]) <= 1e-10 * abs(want[s])
    md, rmd = eng.to_metadata(), ref.to_metadata()
    buf.seek(0)
    rmd2 = REngine.from_metadata(pickle.load(buf), rng=gu.gen_rng(3), multiprocess=0).to_metadata()
    for s in range(3):
        assert stats_of(md['states'][s]) == stats_of(rmd2['states'][s])         # both re-summed in row order
        assert_stats_close(stats_of(md['states'][s]), stats_of(rmd['states'][s]))
        assert md['states'][s]['diagnostics']['iterations'] == rmd['states'][s]['diagnostics']['iterations']
    for rowid, t, c in QUERIES:
        assert np.allclose(eng.logpdf(rowid, t, c), ref.logpdf(rowid, t, c, multiprocess=0), rtol=1e-9, atol=1e-12)
    # a single reference State through a pickle file on disk
    path = str(tmp_path / 'state.pkl')
    with open(path, 'wb') as f:
        pickle.dump(ref.get_state(2).to_metadata(), f)
    st = State.from_pickle(path, rng=gen_rng(0))
    assert abs(st.logpdf_score() - want[2]) <= 1e-10 * abs(want[2])
    eng.transition(N=1, progress=False)


def test_constrained_column_crp_scores_like_the_reference():
    """State(Cd=...) scores the column partition with the constrained CRP (state.py:370-379) and draws its
    initial partition from it; State(Ci=...) refuses to score."""
    from cgpm_b200.engine import State, gen_rng
    REngine, RState, gu = ref_modules()
    X, cct, da = small_table(40, 33)
    Cd = [[0, 2], [3, 4, 5]]
    st = State(X, cctypes=cct, distargs=da, Cd=Cd, rng=gen_rng(5))
    rst = RState(X, cctypes=cct, distargs=da, Cd=Cd, rng=gu.gen_rng(5))
    assert dict(st.Zv()) == dict(rst.Zv())
    assert abs(st.logpdf_score() - rst.logpdf_score()) <= 1e-10 * abs(rst.logpdf_score())
    st.transition(N=2, kernels=['rows', 'alpha', 'view_alphas', 'column_hypers'], progress=False)
    rst.transition(N=2, kernels=['rows', 'alpha', 'view_alphas', 'column_hypers'], progress=False)
    assert abs(st.logpdf_score() - rst.logpdf_score()) <= 1e-10 * abs(rst.logpdf_score())
    Ci = [(0, 1), (2, 3)]
    st = State(X, cctypes=cct, distargs=da, Ci=Ci, rng=gen_rng(6))
    rst = RState(X, cctypes=cct, distargs=da, Ci=Ci, rng=gu.gen_rng(6))
    assert dict(st.Zv()) == dict(rst.Zv())
    with pytest.raises(ValueError):
        st.logpdf_score()
    with pytest.raises(ValueError):
        rst.logpdf_score()
    with pytest.raises(ValueError):
        State(X, cctypes=cct, distargs=da, Ci=Ci, Zv={c: 0 for c in range(6)}, rng=gen_rng(6))
