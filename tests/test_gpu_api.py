"""Reference API behaviour of Engine.transition and friends (tests/test_iter_counter.py,
test_diagnostics.py, test_engine_dimensions.py, test_dependence_constraints.py of the reference)."""
import time

import numpy as np
import pytest

from helpers import small_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def table():
    return small_table(50, 61)


def make(table, S=3, **kw):
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = table
    return Engine(X, num_states=S, rng=gen_rng(7), cctypes=cct, distargs=da, **kw)


def test_iteration_counters_per_kernel(table):
    """tests/test_iter_counter.py:28-54."""
    eng = make(table)
    eng.transition(N=4, kernels=['alpha', 'rows'], progress=False)
    eng.transition(N=2, kernels=['columns', 'column_hypers', 'view_alphas', 'column_params'], progress=False)
    for s in range(3):
        it = eng.get_state(s).diagnostics['iterations']
        assert it == {'alpha': 4, 'rows': 4, 'columns': 2, 'column_hypers': 2, 'view_alphas': 2, 'column_params': 2}
    eng.transition(N=1, progress=False)                       # all six kernels
    assert eng.get_state(0).diagnostics['iterations']['rows'] == 5


def test_checkpoint_diagnostics_and_timed_runs(table):
    """tests/test_diagnostics.py:39-75: checkpoint=k appends every k iterations; S= stops on time."""
    eng = make(table, S=2)
    eng.transition(N=6, checkpoint=2, progress=False)
    d = eng.get_state(1).diagnostics
    assert len(d['logscore']) == 3 and len(d['column_crp_alpha']) == 3 and len(d['column_partition']) == 3
    assert all(len(p) == 6 for p in d['column_partition'])
    assert abs(d['logscore'][-1] - eng.logpdf_score()[1]) < 1e-9 * abs(d['logscore'][-1])
    t0 = time.time()
    eng.transition(S=0.5, progress=False)
    assert 0.4 < time.time() - t0 < 5.0
    assert eng.get_state(0).diagnostics['iterations']['rows'] > 6


def test_statenos_subset_and_argument_errors(table):
    eng = make(table, S=4)
    before = eng.to_metadata()['states']
    eng.transition(N=3, statenos=[1, 3], progress=False)
    after = eng.to_metadata()['states']
    for s in (0, 2):
        assert before[s]['Zrv'] == after[s]['Zrv'] and before[s]['Zv'] == after[s]['Zv']
        assert after[s]['diagnostics']['iterations'] == {}
    assert after[1]['diagnostics']['iterations']['rows'] == 3
    assert eng.logpdf_score(statenos=[1, 3]) == [eng.logpdf_score()[1], eng.logpdf_score()[3]]
    with pytest.raises(KeyError):
        eng.transition(N=1, kernels=['nope'], progress=False)
    with pytest.raises(ValueError):
        eng.transition(N=1, cols=[99], progress=False)         # state.py:733-734
    with pytest.raises(ValueError):
        make(table, Rd={0: [(1, 2)]})
    with pytest.raises(ValueError):
        make(table, inputs=[3])


def test_dependence_constraint_blocks_the_columns_kernel(table):
    """state.py:1056-1059: Cd + 'columns' raises, other kernels run."""
    eng = make(table, S=1, Cd=[(0, 2)], Zv={0: 0, 1: 1, 2: 0, 3: 1, 4: 1, 5: 1})
    eng.transition(N=1, kernels=['rows', 'column_hypers'], progress=False)
    with pytest.raises(ValueError):
        eng.transition(N=1, kernels=['columns'], progress=False)


def test_result_shapes_like_the_reference(table):
    """tests/test_engine_dimensions.py:74-210."""
    eng = make(table, S=3)
    eng.transition(N=2, progress=False)
    lp = eng.logpdf(-1, {0: 1.0})
    assert len(lp) == 3 and all(isinstance(x, float) for x in lp)
    lpb = eng.logpdf_bulk([-1, -1], [{0: 1.0}, {2: 0.5, 1: 1}], statenos=[0, 2])
    assert len(lpb) == 2 and len(lpb[0]) == 2
    sm = eng.simulate(-1, [0, 1], N=4, statenos=[1])
    assert len(sm) == 1 and len(sm[0]) == 4 and set(sm[0][0]) == {0, 1}
    smb = eng.simulate_bulk([-1, 3], [[0], [2, 4]], Ns=[2, 3])
    assert len(smb) == 3 and [len(x) for x in smb[0]] == [2, 3]
    D = eng.dependence_probability_pairwise(colnos=[0, 2, 5])
    assert len(D) == 3 and D[0].shape == (3, 3)
    assert len(eng.row_similarity_pairwise(statenos=[0])) == 1
    st = eng.get_state(2)
    assert st.n_rows() == 50 and st.n_cols() == 6 and set(st.Zv().keys()) == set(range(6))
    assert abs(st.logpdf_score() - eng.logpdf_score()[2]) <= 1e-12 * abs(st.logpdf_score())


def test_device_stream_keys_are_not_reused_after_add_and_drop_state():
    """The Philox stream is keyed by (seed, call counter, purpose, chain, ...): chains re-created by drop_state /
    add_state keep the engine's seed, so the counter must carry over (and keeps growing) or later sweeps would
    replay uniforms already consumed."""
    from cgpm_b200.engine import Engine, gen_rng
    from helpers import small_table
    X, cct, da = small_table(40, 3)
    eng = Engine(X, num_states=3, rng=gen_rng(2), cctypes=cct, distargs=da, stream='philox')
    eng.transition(N=2, progress=False)
    c0, seed = eng._ch.counter, eng._ch.seed
    assert c0 > 1
    eng.drop_state(0)
    assert eng._ch.seed == seed and eng._ch.counter == c0
    eng.transition(N=1, progress=False)
    c1 = eng._ch.counter
    assert c1 > c0
    eng.add_state(count=2)
    assert eng._ch.seed == seed and eng._ch.counter == c1 and eng.num_states() == 4
    eng.transition(N=1, progress=False)
    assert eng._ch.counter > c1
    assert all(np.isfinite(eng.logpdf_score()))
