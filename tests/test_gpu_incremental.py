"""incremental edits on a live GPU engine follow the oracle
(itself bit-identical to the reference for these operations, see
tests/test_oracle_vs_reference.py::test_incremental_edits)."""
import numpy as np
import pytest

import crosscat_oracle as co
from helpers import small_table
from test_gpu_transition import compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('host_edits', [False, True])
def test_incorporate_unincorporate_force_cell_dims_follow_oracle(host_edits, monkeypatch):
    """host_edits=False: row edits run on the device (structure_dev.py: device-to-device re-sizing, statistics updated
    in place); True: the general path through the host (CGPM_B200_HOST_EDITS).  Both must follow the oracle bit for bit."""
    from cgpm_b200.engine import Engine, gen_rng
    if host_edits:
        monkeypatch.setenv('CGPM_B200_HOST_EDITS', '1')
    X, cct, da = small_table(40, 3)
    S = 2
    # view 0 must exist when incorporate_dim(v=None) runs: the reference pairs its per-view
    # scores (insertion order) with the sorted table list (state.py:1065-1075), which only agree
    # when views were created in increasing id order; re-creating view 0 later breaks that in the
    # reference (hard part 9) -- the GPU path always uses the sorted pairing.
    Zv = {0: 0, 1: 0, 2: 1, 3: 1, 4: 0, 5: 1}
    eng = Engine(X, num_states=S, rng=gen_rng(31), cctypes=cct, distargs=da, Zv=Zv, stream='numpy')
    seeds = gen_rng(31).randint(low=1, high=2 ** 32 - 1, size=S)
    oracles = [co.OState(X, cctypes=cct, distargs=da, Zv=Zv, rng=np.random.RandomState(sd)) for sd in seeds]

    def both(tag, fe, fo):
        fe(eng)
        for o in oracles:
            fo(o)
        compare(eng, oracles, tag)

    T0 = list(np.random.RandomState(8).normal(size=40))
    both('dim+0', lambda e: e.incorporate_dim(T0, [70], cctype='normal'), lambda o: o.incorporate_dim(T0, [70], cctype='normal'))
    both('t0', lambda e: e.transition(N=1, kernels=['rows', 'columns'], progress=False),
         lambda o: o.transition(N=1, kernels=['rows', 'columns']))
    both('inc1', lambda e: e.incorporate(40, {0: 1.5, 1: 2, 3: 0}), lambda o: o.incorporate(40, {0: 1.5, 1: 2, 3: 0}))
    both('inc2', lambda e: e.incorporate(41, {2: -0.5, 5: 3.0, 4: 0.1}), lambda o: o.incorporate(41, {2: -0.5, 5: 3.0, 4: 0.1}))
    both('force', lambda e: e.force_cell(41, {0: 0.25}), lambda o: o.force_cell(41, {0: 0.25}))
    both('t1', lambda e: e.transition(N=1, kernels=['rows', 'column_hypers', 'view_alphas'], progress=False),
         lambda o: o.transition(N=1, kernels=['rows', 'column_hypers', 'view_alphas']))
    both('uninc', lambda e: e.unincorporate(41), lambda o: o.unincorporate(41))
    Tc = [float(i % 3) for i in range(41)]
    both('dim+v', lambda e: e.incorporate_dim(Tc, [78], cctype='categorical', distargs={'k': 3}, v=9),
         lambda o: o.incorporate_dim(Tc, [78], cctype='categorical', distargs={'k': 3}, v=9))
    ks = ['rows', 'columns', 'alpha', 'view_alphas', 'column_hypers']
    both('t2', lambda e: e.transition(N=2, kernels=ks, progress=False), lambda o: o.transition(N=2, kernels=ks))
    both('dim-', lambda e: e.unincorporate_dim(2), lambda o: o.unincorporate_dim(2))
    both('t3', lambda e: e.transition(N=1, kernels=['rows', 'columns'], progress=False),
         lambda o: o.transition(N=1, kernels=['rows', 'columns']))
    for s, o in enumerate(oracles):
        assert eng._ch.rngs[s].random_sample() == o.rng.random_sample()
    assert eng.X.shape == (41, 7)
    assert (getattr(eng._ch, 'device_edits', 0) >= 4) == (not host_edits)      # inc1, inc2, force, uninc ran on the device
    with pytest.raises(ValueError):
        eng.incorporate(50, {0: 1.0})
    with pytest.raises(ValueError):
        eng.unincorporate(3)
    with pytest.raises(ValueError):
        eng.force_cell(0, {0: 1.0})
    with pytest.raises(ValueError):
        eng.incorporate_dim([0.0] * 41, [0], cctype='normal')
    # given cluster token: no Gibbs step, the row sits where it was told
    st = eng.get_state(0)
    vid = list(st.views.keys())[0]
    k = list(st.views[vid].Nk().keys())[0]
    st.incorporate(41, {0: 0.1, 10 ** 7 + vid: k})
    assert st.views[vid].Zr(41) == k
