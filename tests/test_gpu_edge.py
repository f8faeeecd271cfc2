"""GPU edge cases against the oracle: global-memory table mode and the mid-sweep
migration out of shared memory, row / column subsets, Ci constraints, NaN-heavy
and categorical-only tables, tiny tables."""
import numpy as np
import pytest

import crosscat_oracle as co
from helpers import oracle_suffstats, device_suffstats, assert_suffstats_equal
from test_gpu_transition import compare

pytestmark = pytest.mark.gpu


def pair(X, cct, da, seed, S=2, **kw):
    from cgpm_b200.engine import Engine, gen_rng
    okw = {k: v for k, v in kw.items() if k in ('Zv', 'Zrv', 'alpha', 'view_alphas', 'Ci')}
    eng = Engine(X, num_states=S, rng=gen_rng(seed), cctypes=cct, distargs=da, stream='numpy', **kw)
    seeds = gen_rng(seed).randint(low=1, high=2 ** 32 - 1, size=S)
    oracles = [co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(sd), **okw) for sd in seeds]
    return eng, oracles


def test_global_table_mode_and_mid_sweep_migration(monkeypatch):
    """The per-(cluster, column) tables live in shared memory when they fit.  With the
    shared tables capped (env override used only here), (1) a 60-column view with 30
    clusters starts in global-table mode (mode 1), (2) a 4-column view that grows from 2 to
    many clusters migrates from shared to global memory mid-sweep (mode 2), (3) so does a
    20-column view.  All must follow the oracle exactly (partitions, bit-identical statistics)."""
    from cgpm_b200 import synth
    monkeypatch.setenv('CGPM_B200_ROWS_PROF', '1')
    modes = set()
    cases = (('global', 60, 64, np.arange(90) % 30, 5.0), ('migrate', 4, 1, np.arange(90) % 2, 150.0),
             ('migrate_wide', 20, 5, np.arange(90) % 2, 1e40))
    for name, C, kb, Zr, alpha in cases:
        monkeypatch.setenv('CGPM_B200_ROWS_SMEM_KB', str(kb))
        cct = ['normal'] * C
        X, _, _ = synth.gen_table(90, cct, [None] * C, n_views=1, clusters_per_view=3, seed=2)
        Zv = {c: 0 for c in range(C)}
        eng, oracles = pair(X, cct, None, 3, S=2, Zv=Zv, Zrv={0: [int(z) for z in Zr]}, view_alphas={0: alpha}, alpha=1.0)
        compare(eng, oracles, (name, 'init'))
        for it in range(3):
            eng.transition(N=1, kernels=['rows'], progress=False)
            for o in oracles:
                o.transition(N=1, kernels=['rows'])
            compare(eng, oracles, (name, it))
            seq = eng._ch.dev.t['seq'].cpu().numpy()
            modes |= set(int(seq[s, 0, 4]) for s in range(2))
    assert {1, 2} <= modes, modes       # the global start and the mid-sweep migrations ran


@pytest.mark.parametrize('case', ['handover_mid_sweep', 'handover_at_start', 'grow_kcap', 'grow_in_large_k_kernel'])
def test_large_k_kernel_and_capacity_growth(case, monkeypatch):
    """Sweeps that outgrow the shared-memory candidate arrays continue in the large-K kernel (everything per
    cluster in global memory), and sweeps that run out of cluster slots altogether stop in front of the move,
    the host grows Kcap and they resume with the same row: both hand-overs go through rows_prog and must leave
    the trajectory of the strictly sequential scan untouched (oracle: partitions, member order, bit-identical
    statistics, stream position)."""
    from cgpm_b200 import synth
    monkeypatch.setenv('CGPM_B200_ROWS_PROF', '1')
    C, R = 3, 120
    cct = ['normal'] * C
    X, _, _ = synth.gen_table(R, cct, [None] * C, n_views=1, clusters_per_view=3, seed=6)
    Zv = {c: 0 for c in range(C)}
    kw = {}
    if case == 'handover_mid_sweep':           # 2 clusters -> dozens (alpha = 150), candidate arrays of 8 slots
        monkeypatch.setenv('CGPM_B200_ROWS_KSM', '8')
        Zr, alpha = np.arange(R) % 2, 150.0
        kw['Kcap'] = 128
    elif case == 'handover_at_start':          # 40 clusters from the start
        monkeypatch.setenv('CGPM_B200_ROWS_KSM', '8')
        Zr, alpha = np.arange(R) % 40, 3.0
        kw['Kcap'] = 128
    elif case == 'grow_kcap':                  # 4 slots per view to begin with
        Zr, alpha = np.arange(R) % 2, 150.0
        kw['Kcap'] = 4
    else:                                      # growth requested by the large-K kernel
        monkeypatch.setenv('CGPM_B200_ROWS_KSM', '6')
        Zr, alpha = np.arange(R) % 3, 150.0
        kw['Kcap'] = 12
    eng, oracles = pair(X, cct, None, 3, S=2, Zv=Zv, Zrv={0: [int(z) for z in Zr]}, view_alphas={0: alpha}, alpha=1.0, **kw)
    compare(eng, oracles, (case, 'init'))
    modes = set()
    for it in range(3):
        eng.transition(N=1, kernels=['rows'], progress=False)
        for o in oracles:
            o.transition(N=1, kernels=['rows'])
        compare(eng, oracles, (case, it))
        seq = eng._ch.dev.t['seq'].cpu().numpy()
        modes |= set(int(seq[s, 0, 4]) for s in range(2))
    dev = eng._ch.dev
    if case.startswith('handover') or case == 'grow_in_large_k_kernel':
        assert 9 in modes, modes               # the large-K kernel finished at least one sweep
    if case.startswith('grow'):
        assert dev.kcap_grown > 0 and dev.Kcap > kw['Kcap']
    eng._ch.check_partitions()


def test_row_and_column_subsets_and_views_argument():
    from helpers import small_table
    X, cct, da = small_table(70, 9)
    eng, oracles = pair(X, cct, da, 9, S=2)
    rows = [3, 5, 8, 13, 21, 34, 55]
    eng.transition(N=1, kernels=['rows'], rowids=rows, progress=False)
    for o in oracles:
        if o.n_rows != 1:
            for v in o.views:
                o.views[v].transition_rows(rows=rows)
            o._count('rows')
    compare(eng, oracles, 'rows subset')
    cols = [4, 1, 2]
    eng.transition(N=1, kernels=['columns', 'column_hypers'], cols=cols, progress=False)
    for o in oracles:
        o.transition_dims(cols=cols)
        o.transition_dim_hypers(cols=cols)
    compare(eng, oracles, 'cols subset')


def test_independence_constraints_are_respected():
    from helpers import small_table
    X, cct, da = small_table(50, 4)
    Ci = [(0, 2), (4, 5)]
    Zv = {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 3}
    eng, oracles = pair(X, cct, da, 4, S=3, Zv=Zv, Ci=Ci)
    for it in range(4):
        eng.transition(N=1, kernels=['columns', 'rows'], progress=False)
        for o in oracles:
            o.transition(N=1, kernels=['columns', 'rows'])
        compare(eng, oracles, ('ci', it))
        for s in range(3):
            zv = dict(eng._ch.Zv_items(s))
            assert zv[0] != zv[2] and zv[4] != zv[5]


def test_nan_heavy_categorical_only_and_tiny_tables():
    from cgpm_b200 import synth
    cct = ['categorical'] * 4
    da = [{'k': 2}, {'k': 3}, {'k': 5}, {'k': 2}]
    X, _, _ = synth.gen_table(40, cct, da, n_views=2, clusters_per_view=2, seed=6, nan_frac=0.3)
    eng, oracles = pair(X, cct, da, 6, S=2)
    for it in range(3):
        eng.transition(N=1, progress=False)
        for o in oracles:
            o.transition(N=1)
        compare(eng, oracles, ('cat', it))
    # two rows, one column
    X2 = np.array([[0.5], [1.5]])
    eng, oracles = pair(X2, ['normal'], None, 7, S=1)
    eng.transition(N=3, progress=False)
    for o in oracles:
        o.transition(N=3)
    compare(eng, oracles, 'tiny')
    # all-NaN column next to a regular one
    X3 = np.column_stack([np.linspace(-1, 1, 12), np.full(12, np.nan)])
    X3[0, 1] = 0.25
    eng, oracles = pair(X3, ['normal', 'normal'], None, 8, S=1)
    eng.transition(N=2, progress=False)
    for o in oracles:
        o.transition(N=2)
    compare(eng, oracles, 'nan column')


def test_cluster_capacity_grows_on_the_device_stream():
    """Kcap = 4 and a CRP that wants dozens of clusters: the engine grows the cluster slots instead of failing
    (round 1 raised a capacity error here) and the partitions stay consistent."""
    from cgpm_b200 import synth
    from cgpm_b200.engine import Engine, gen_rng
    X, _, _ = synth.gen_table(64, ['normal'] * 3, [None] * 3, n_views=1, clusters_per_view=2, seed=1)
    eng = Engine(X, num_states=2, rng=gen_rng(1), cctypes=['normal'] * 3, stream='philox', Kcap=4,
                 Zv={0: 0, 1: 0, 2: 0}, Zrv={0: [0] * 64}, view_alphas={0: 50.0}, alpha=1.0)
    eng.transition(N=3, kernels=['rows'], progress=False)
    dev = eng._ch.dev
    assert dev.kcap_grown > 0 and dev.Kcap > 4
    eng._ch.check_partitions()


def test_view_capacity_grows_during_a_column_sweep():
    """Two view slots per chain to begin with and a column CRP that wants many views (alpha = 50): the column
    kernel's births re-allocate the per-view arrays (DeviceChains.grow_vcap) mid-sweep and the trajectory still
    follows the oracle exactly."""
    from helpers import small_table
    X, cct, da = small_table(60, 13)
    Zv = {c: c % 2 for c in range(6)}
    eng, oracles = pair(X, cct, da, 13, S=2, Zv=Zv, alpha=50.0, Vcap=2)
    compare(eng, oracles, 'vcap init')
    for it in range(3):
        eng.transition(N=1, kernels=['columns', 'rows'], progress=False)
        for o in oracles:
            o.transition(N=1, kernels=['columns', 'rows'])
        compare(eng, oracles, ('vcap', it))
    dev = eng._ch.dev
    assert getattr(dev, 'vcap_grown', 0) > 0 and dev.Vcap > 2
    eng._ch.check_partitions()


def test_very_wide_views_use_the_wide_mode():
    """More than 1024 columns in a view: per-column hypers / types and the tables stay in
    global memory (rows kernel 'wide' mode); results still follow the oracle exactly."""
    from cgpm_b200 import synth
    C, R = 1100, 14
    cct = ['normal'] * C
    X, _, _ = synth.gen_table(R, cct, [None] * C, n_views=2, clusters_per_view=2, seed=4)
    Zv = {c: (0 if c < 1060 else 1) for c in range(C)}
    eng, oracles = pair(X, cct, None, 5, S=1, Zv=Zv, Vcap=8)
    compare(eng, oracles, 'wide init')
    for it in range(2):
        eng.transition(N=1, kernels=['rows'], progress=False)
        for o in oracles:
            o.transition(N=1, kernels=['rows'])
        compare(eng, oracles, ('wide rows', it))
    eng.transition(N=1, kernels=['columns'], cols=list(range(0, C, 37)), progress=False)
    for o in oracles:
        o.transition_dims(cols=list(range(0, C, 37)))
    compare(eng, oracles, 'wide columns')


def test_debug_invariants_like_gpmccdebug(monkeypatch):
    """GPMCCDEBUG (src/utils/config.py:112-114) turns on the partition invariants after every
    kernel; a corrupted partition is reported."""
    import torch
    from helpers import small_table
    from cgpm_b200.engine import Engine, gen_rng
    monkeypatch.setenv('GPMCCDEBUG', '1')
    X, cct, da = small_table(60, 12)
    eng = Engine(X, num_states=3, rng=gen_rng(12), cctypes=cct, distargs=da, stream='philox')
    eng.transition(N=3, progress=False)                      # checks run after each kernel, no violation
    eng._ch.check_partitions()
    ch = eng._ch
    nv = ch.dev.t['nv'].cpu().numpy()
    v = int(np.nonzero(nv[1] > 0)[0][0])
    saved = int(ch.dev.t['zr'][1, v, 5].item())
    ch.dev.t['zr'][1, v, 5] = ch.dev.Kcap - 2                # a dead slot
    with pytest.raises(AssertionError, match='chain 1'):
        ch.check_partitions()
    ch.dev.t['zr'][1, v, 5] = saved
    ch.check_partitions()
    ch.dev.t['stat'][2, 0, 0, :] += 1.0                      # corrupt N of cluster slot 0 in every column
    with pytest.raises(AssertionError, match='chain 2'):
        ch.check_partitions()


def test_host_philox_mirror_matches_the_device_generator():
    """core.aux_alpha_index (numpy Philox4x32-10) predicts the alpha the device draws for every
    auxiliary view: it is used to schedule expensive (chain, column) pairs first."""
    from helpers import small_table
    from cgpm_b200.core import aux_alpha_index
    from cgpm_b200.engine import Engine, gen_rng
    X, cct, da = small_table(40, 5)
    eng = Engine(X, num_states=3, rng=gen_rng(5), cctypes=cct, distargs=da, stream='philox')
    ch = eng._ch
    counter = ch.counter + 1                          # the counter the next kernel call will use
    eng.transition(N=1, kernels=['columns'], progress=False)
    aux_alpha = ch.dev.t['aux_alpha'].cpu().numpy()
    chains = np.repeat(np.arange(3), 6); cols = np.tile(np.arange(6), 3)
    idx = aux_alpha_index(ch.seed, counter, chains, cols)
    assert np.array_equal(aux_alpha[chains, cols], ch.dev.row_alpha_grid_h[idx])
