"""GPU parity of the rows kernel and the table rebuild against the oracle."""
import numpy as np
import pytest
import torch

import crosscat_oracle as co
from helpers import (small_table, ctype_arrays, oracle_spec, oracle_suffstats, rel_tol,
                     device_suffstats, assert_suffstats_equal)

pytestmark = pytest.mark.gpu


def make_pair(n_rows, seed, S=1, **kw):
    from cgpm_b200.device import DeviceChains
    from cgpm_b200 import codec
    X, cct, da = small_table(n_rows, seed)
    ctype, ncat = ctype_arrays(cct, da)
    dev = DeviceChains(X, ctype, ncat, S, **kw)
    oracles = []
    for s in range(S):
        o = co.OState(X, cctypes=cct, distargs=da, rng=np.random.RandomState(seed * 100 + s))
        dev.upload_chain(s, codec.pack_chain(oracle_spec(o), X.shape[0], X.shape[1]))
        oracles.append(o)
    dev.rebuild()
    dev.sync()
    return X, ctype, dev, oracles


@pytest.mark.parametrize('n_rows,seed', [(40, 1), (150, 2)])
def test_rebuild_and_score_match_oracle(n_rows, seed):
    X, ctype, dev, oracles = make_pair(n_rows, seed, S=2)
    scores = dev.logpdf_score().cpu().numpy()
    for s, o in enumerate(oracles):
        got, _ = device_suffstats(dev, s, ctype)
        assert_suffstats_equal(oracle_suffstats(o), got)        # bit-exact
        ref = o.logpdf_score()
        assert abs(scores[s] - ref) <= 1e-10 * abs(ref), (scores[s], ref)


def rows_sweep_numpy_stream(dev, oracles, views_order, rngs, trace=False, hints=False):
    """One 'rows' kernel call with the host numpy stream injected, mirroring the
    draw order of View.transition_rows (one permutation + one uniform per row).
    hints: pass the view widths / types (all-normal tables only), which routes views of
    one or two columns to the single-warp kernel."""
    R = dev.R
    TS = 3 + dev.Kcap
    work, perms, us = [], [], []
    off = 0
    for s, o in enumerate(oracles):
        for v in views_order[s]:
            perms.append(rngs[s].permutation(R).astype(np.int32))
            us.append(rngs[s].random_sample(R))
            work.append(dict(chain=s, view=v, n=R, perm_is_index=1, perm_off=off, u_off=off,
                             trace_off=(len(work) * R * TS if trace else -1)))
            if hints:
                nc = len(list(o.views.values())[v].dims)
                work[-1].update(n_cols=nc, n_normal=nc)
            off += R
    perm = torch.from_numpy(np.concatenate(perms)).to(dev.device)
    u = torch.from_numpy(np.concatenate(us)).to(dev.device)
    tr = torch.zeros(len(work) * R * TS, dtype=torch.float64, device=dev.device) if trace else None
    dev.transition_rows(work, perm=perm, u=u, trace=tr, trace_stride=TS)
    dev.sync()
    dev.check_errors()
    return work, (tr.cpu().numpy().reshape(len(work), R, TS) if trace else None)


@pytest.mark.parametrize('n_rows,seed,kw', [(40, 1, {}), (150, 2, {}), (300, 3, {})])
def test_rows_trajectory_matches_oracle(n_rows, seed, kw):
    """Injected shared stream: identical partitions, bit-identical sufficient
    statistics, conditional log-probabilities within 1e-5 relative."""
    X, ctype, dev, oracles = make_pair(n_rows, seed, S=2, **kw)
    # the oracle draws from its own RandomState; the device gets the same stream
    rngs = [np.random.RandomState(7 + s) for s in range(len(oracles))]
    for s, o in enumerate(oracles):
        o.rng.set_state(rngs[s].get_state())
        for view in o.views.values():
            view.rng = o.rng
    views_order = [list(range(len(o.views))) for o in oracles]
    for sweep in range(4):
        otr = [dict() for _ in oracles]
        for s, o in enumerate(oracles):
            o.transition_view_rows(trace=otr[s])
        work, tr = rows_sweep_numpy_stream(dev, oracles, views_order, rngs, trace=True)
        for s, o in enumerate(oracles):
            assert rngs[s].random_sample() == o.rng.random_sample()
            got, dl = device_suffstats(dev, s, ctype)
            # partitions and member order
            for j, (vid, view) in enumerate(o.views.items()):
                zr_dev = dl['cid_by_slot'][j][dl['zr'][j]]
                zr_ref = np.array([view.Zr[i] for i in range(n_rows)])
                assert np.array_equal(zr_dev, zr_ref), (sweep, s, vid)
                assert np.array_equal(dl['order'][j], np.array(list(view.Zr.keys())))
            assert_suffstats_equal(oracle_suffstats(o), got)
        # per-step conditionals
        wi = 0
        for s, o in enumerate(oracles):
            for vid in o.views:
                steps = otr[s][vid]
                for i, (rowid, K, logp, zb) in enumerate(steps):
                    rec = tr[wi, i]
                    assert int(rec[0]) == rowid and int(rec[1]) == len(K) and int(rec[2]) == zb
                    got = rec[3:3 + len(K)]
                    ref = np.asarray(logp)
                    fin = np.isfinite(ref)
                    assert np.array_equal(np.isfinite(got), fin)
                    assert np.all(np.abs(got[fin] - ref[fin]) <= rel_tol(ref[fin]))
                wi += 1
    scores = dev.logpdf_score().cpu().numpy()
    for s, o in enumerate(oracles):
        ref = o.logpdf_score()
        assert abs(scores[s] - ref) <= 1e-10 * abs(ref)


def check_against_oracle(dev, oracles, ctype, n_rows, otr, tr, rngs, tag):
    for s, o in enumerate(oracles):
        assert rngs[s].random_sample() == o.rng.random_sample()
        got, dl = device_suffstats(dev, s, ctype)
        for j, (vid, view) in enumerate(o.views.items()):
            zr_dev = dl['cid_by_slot'][j][dl['zr'][j]]
            zr_ref = np.array([view.Zr[i] for i in range(n_rows)])
            assert np.array_equal(zr_dev, zr_ref), (tag, s, vid)
            assert np.array_equal(dl['order'][j], np.array(list(view.Zr.keys()))), (tag, s, vid)
        assert_suffstats_equal(oracle_suffstats(o), got)
    wi = 0
    for s, o in enumerate(oracles):
        for vid in o.views:
            for i, (rowid, K, logp, zb) in enumerate(otr[s][vid]):
                rec = tr[wi, i]
                assert int(rec[0]) == rowid and int(rec[1]) == len(K) and int(rec[2]) == zb, (tag, s, vid, i)
                got = rec[3:3 + len(K)]
                ref = np.asarray(logp)
                assert np.all(np.abs(got - ref) <= rel_tol(ref)), (tag, s, vid, i)
            wi += 1


@pytest.mark.parametrize('n_rows,seed,n_init,lvl', [(60, 5, 3, None), (300, 6, 3, None), (200, 7, 45, None),
                                                   (300, 6, 3, 0), (300, 6, 3, 2), (300, 6, 3, 8), (300, 6, 3, 9),
                                                   (200, 7, 45, 0), (200, 7, 45, 9), (1000, 6, 3, None)])
def test_rows_trajectory_narrow_views(n_rows, seed, n_init, lvl, monkeypatch):
    """All-normal table whose views hold one or two columns (rows move at every other step): the
    speculative rounds -- adaptive window (lvl=None) or pinned with CGPM_B200_ROWS_LVL to the most rows in
    flight (0), a middle setting, a single row group per round (8) and one warp sweeping on its own (9; the adaptive
    setting reaches it by itself on the 1 000-row table) -- must follow the oracle step by step
    (n_init = 45 starts from more than 32 clusters per view: one row at a time with the multi-pass
    draw, then speculative rounds once the view has shrunk below 32 clusters)."""
    from cgpm_b200.device import DeviceChains
    from cgpm_b200 import codec
    from cgpm_b200.synth import gen_table
    if lvl is not None:
        monkeypatch.setenv('CGPM_B200_ROWS_LVL', str(lvl))
    monkeypatch.setenv('CGPM_B200_ROWS_PROF', '1')
    cct = ['normal'] * 5
    da = [None] * 5
    X, _, _ = gen_table(n_rows, cct, da, n_views=2, clusters_per_view=3, seed=seed)
    rng0 = np.random.RandomState(seed)
    for _ in range(n_rows // 10):
        X[rng0.randint(n_rows), rng0.randint(5)] = np.nan
    ctype, ncat = ctype_arrays(cct, da)
    S = 2
    dev = DeviceChains(X, ctype, ncat, S, Kcap=256)
    oracles = []
    Zv = {0: 0, 1: 1, 2: 1, 3: 2, 4: 2}
    for s in range(S):
        Zrv = {v: [int(z) for z in rng0.randint(n_init, size=n_rows)] for v in range(3)}
        for v in range(3):                       # cluster ids must be 0..K-1 in order of appearance
            _, inv = np.unique(Zrv[v], return_inverse=True)
            Zrv[v] = [int(z) for z in inv]
        o = co.OState(X, cctypes=cct, distargs=da, Zv=Zv, Zrv=Zrv, rng=np.random.RandomState(seed * 100 + s))
        dev.upload_chain(s, codec.pack_chain(oracle_spec(o), n_rows, 5))
        oracles.append(o)
    dev.rebuild()
    dev.sync()
    rngs = [np.random.RandomState(17 + s) for s in range(S)]
    for s, o in enumerate(oracles):
        o.rng.set_state(rngs[s].get_state())
        for view in o.views.values():
            view.rng = o.rng
    views_order = [list(range(len(o.views))) for o in oracles]
    for sweep in range(3):
        otr = [dict() for _ in oracles]
        for s, o in enumerate(oracles):
            o.transition_view_rows(trace=otr[s])
        work, tr = rows_sweep_numpy_stream(dev, oracles, views_order, rngs, trace=True, hints=True)
        modes = dev.t['seq'].cpu().numpy()[:, :3, 4]
        assert np.all(modes == 0), modes        # shared-memory tables, first kernel
        check_against_oracle(dev, oracles, ctype, n_rows, otr, tr, rngs, (sweep, lvl))


@pytest.mark.parametrize('n_rows,seed,mixed,env', [(150, 11, False, {}), (120, 12, True, {}),
                                                  (150, 11, False, {'CGPM_B200_ROWS_LVL': '0'}),
                                                  (120, 12, True, {'CGPM_B200_ROWS_LVL': '0'}),
                                                  (150, 11, False, {'CGPM_B200_ROWS_DL': '8'}),
                                                  (120, 12, True, {'CGPM_B200_ROWS_NO_BULK': '1'}),
                                                  (120, 12, True, {'CGPM_B200_ROWS_DL': '8'}),
                                                  (600, 13, 40, {}), (600, 14, -40, {}),
                                                  (200, 13, 40, {'CGPM_B200_ROWS_LVL': '2'}),
                                                  (200, 13, 40, {'CGPM_B200_ROWS_DL': '8'}),
                                                  (200, 13, 40, {'CGPM_B200_ROWS_SMEM_KB': '8'}),
                                                  (200, 13, 40, {'CGPM_B200_ROWS_SMEM_KB': '8', 'CGPM_B200_ROWS_DL': '8'}),
                                                  (120, 12, True, {'CGPM_B200_ROWS_SMEM_KB': '4', 'CGPM_B200_ROWS_DL': '8'}),
                                                  (200, 14, -40, {'CGPM_B200_ROWS_SMEM_KB': '8', 'CGPM_B200_ROWS_DL': '8'})])
def test_rows_trajectory_wide_view(n_rows, seed, mixed, env, monkeypatch):
    """A view of 20 columns next to one of 3 (all-normal and mixed with categorical columns), swept by
    speculative rounds (adaptive, or pinned to the most rows in flight where every mover invalidates a long
    window), one row at a time (CGPM_B200_ROWS_DL=8 puts the 20-column view beyond the width limit) and with
    per-cell gathers instead of bulk row copies: all must follow the oracle step by step -- partitions,
    member order, bit-identical statistics and per-step conditionals.
    mixed = +-40: a 40-column view (configs[2]'s kind: every other column categorical when positive, all normal when
    negative) started from a random 5-cluster partition, so that most rows move in the first sweep and few in the third:
    the adaptive window goes through the whole-CTA-on-one-row setting (600 rows: an epoch of 16 rounds, then stretches
    of 256 rows) and back to speculative rounds; pinned variants cover the by-type scoring of the speculative rounds,
    of the one-row path and the global-table fallback at that width.  CGPM_B200_ROWS_SMEM_KB + CGPM_B200_ROWS_DL
    together put the wide view on global-memory tables AND one row at a time -- how configs[2]'s 130-180-column views
    with a dozen clusters run (by-type scoring against the global tables, score_candidate_g); the variants check that
    the tables really left shared memory (mode 1)."""
    from cgpm_b200.device import DeviceChains
    from cgpm_b200 import codec
    from cgpm_b200.synth import gen_table
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    monkeypatch.setenv('CGPM_B200_ROWS_PROF', '1')
    widths = (40, 3) if abs(int(mixed)) == 40 else (20, 3)
    ncol = sum(widths)
    cct = ['normal'] * ncol
    da = [None] * ncol
    if mixed is True:
        for c in (2, 7, 13, 21):
            cct[c] = 'categorical'; da[c] = {'k': 3 + (c % 3)}
    elif mixed == 40:
        for c in range(1, ncol, 2):
            cct[c] = 'categorical'; da[c] = {'k': [2, 4, 8, 16][(c // 2) % 4]}
    X, _, _ = gen_table(n_rows, cct, da, n_views=2, clusters_per_view=3, seed=seed)
    rng0 = np.random.RandomState(seed)
    for _ in range(n_rows // 5):
        X[rng0.randint(n_rows), rng0.randint(ncol)] = np.nan
    ctype, ncat = ctype_arrays(cct, da)
    S = 2
    dev = DeviceChains(X, ctype, ncat, S, Kcap=64)
    Zv = {}
    for v, wd in enumerate(widths):
        for _ in range(wd):
            Zv[len(Zv)] = v
    oracles = []
    for s in range(S):
        Zrv = {}
        for v in range(len(widths)):
            _, inv = np.unique(rng0.randint(5, size=n_rows), return_inverse=True)
            Zrv[v] = [int(z) for z in inv]
        o = co.OState(X, cctypes=cct, distargs=da, Zv=Zv, Zrv=Zrv, rng=np.random.RandomState(seed * 100 + s))
        dev.upload_chain(s, codec.pack_chain(oracle_spec(o), n_rows, ncol))
        oracles.append(o)
    dev.rebuild()
    dev.sync()
    rngs = [np.random.RandomState(27 + s) for s in range(S)]
    for s, o in enumerate(oracles):
        o.rng.set_state(rngs[s].get_state())
        for view in o.views.values():
            view.rng = o.rng
    views_order = [list(range(len(o.views))) for o in oracles]
    for sweep in range(3):
        otr = [dict() for _ in oracles]
        for s, o in enumerate(oracles):
            o.transition_view_rows(trace=otr[s])
        R = dev.R
        TS = 3 + dev.Kcap
        work, perms, us, off = [], [], [], 0
        for s, o in enumerate(oracles):
            for v in views_order[s]:
                perms.append(rngs[s].permutation(R).astype(np.int32))
                us.append(rngs[s].random_sample(R))
                nc = len(list(o.views.values())[v].dims)
                work.append(dict(chain=s, view=v, n=R, perm_is_index=1, perm_off=off, u_off=off,
                                 trace_off=len(work) * R * TS, n_cols=nc, n_normal=0))
                off += R
        perm = torch.from_numpy(np.concatenate(perms)).to(dev.device)
        u = torch.from_numpy(np.concatenate(us)).to(dev.device)
        tr = torch.zeros(len(work) * R * TS, dtype=torch.float64, device=dev.device)
        dev.transition_rows(work, perm=perm, u=u, trace=tr, trace_stride=TS)
        dev.sync()
        dev.check_errors()
        trh = tr.cpu().numpy().reshape(len(work), R, TS)
        if 'CGPM_B200_ROWS_SMEM_KB' not in env:
            assert np.all(dev.t['seq'].cpu().numpy()[:, 0, 4] == 0)      # shared-memory tables, first kernel
        elif 'CGPM_B200_ROWS_DL' in env:
            assert np.all(dev.t['seq'].cpu().numpy()[:, 0, 4] == 1)      # the wide view's tables are in global memory
        check_against_oracle(dev, oracles, ctype, n_rows, otr, trh, rngs, (sweep, mixed, env))
