"""CPU: host-side logic (codec, exact host maths, synthetic data, chain sharding)."""
import os
from collections import OrderedDict

import numpy as np
import pytest

import crosscat_oracle as co
from cgpm_b200 import codec, hostmath as hm, synth
from cgpm_b200.parallel import shard_range, shard_sizes


def test_codec_pack_unpack_round_trip():
    R, C = 12, 4
    views = OrderedDict()
    views[3] = {'alpha': 0.5, 'Zr': np.array([0, 0, 7, 7, 2, 2, 2, 0, 7, 9, 9, 0]), 'order': None}
    views[0] = {'alpha': 2.0, 'Zr': np.zeros(R, dtype=int), 'order': np.arange(R)[::-1].copy()}
    spec = {'Zv': [3, 0, 3, 0], 'views': views, 'alpha': 1.5, 'hyp': np.arange(16.).reshape(4, 4)}
    ch = codec.pack_chain(spec, R, C)
    assert ch['view_id'] == [3, 0] and ch['nv'] == [2, 2]
    assert list(ch['cid'][0]) == [0, 2, 7, 9] and list(ch['nk'][0]) == [4, 3, 3, 2]
    dl = dict(ch)
    dl['live'] = [np.arange(len(c)) for c in ch['cid']]
    dl['cid_by_slot'] = ch['cid']
    back = codec.unpack_chain(dl, views_order=[0, 1])
    assert back['Zv'] == spec['Zv']
    for v in views:
        assert np.array_equal(back['views'][v]['Zr'], views[v]['Zr'])
    assert list(back['views'].keys()) == [3, 0]
    h = codec.hypers_from_array(np.arange(8.).reshape(2, 4), [0, 1])
    assert list(h[0].keys()) == ['m', 'r', 's', 'nu'] and list(h[1].keys()) == ['alpha']
    assert np.array_equal(codec.hypers_to_array(h, [0, 1])[0], [0, 1, 2, 3])


def test_hostmath_matches_oracle_and_reference_kats():
    rng = np.random.RandomState(42)
    assert [int(hm.log_pflip([-1, -2, -.5], array=[5, 7, 9], rng=rng)) for _ in range(8)] == [7, 9, 9, 9, 5, 5, 5, 9]
    assert hm.logsumexp(range(1000)) == co.logsumexp(range(1000))
    assert hm.logmeanexp([0., -1.]) == co.logmeanexp([0., -1.])
    # exact CRP prior simulation consumes the stream like the oracle's view constructor
    a, b = np.random.RandomState(9), np.random.RandomState(9)
    z = hm.crp_simulate_exact(60, 1.3, a)
    crp = co.OCrp(60, 1.3, b)
    ref = []
    for i in range(60):
        x = crp.simulate_fresh(b)
        crp.incorporate(i, x)
        ref.append(int(x))
    assert list(z) == ref and a.random_sample() == b.random_sample()
    zf = hm.crp_simulate_fast(5000, 2.0, np.random.RandomState(1))
    assert zf[0] == 0 and set(np.unique(zf)) == set(range(zf.max() + 1))
    K = zf.max() + 1
    assert 5 < K < 40                      # E[K] ~ alpha log(1 + n/alpha) ~ 15.6


def test_hyper_grids_match_oracle():
    from cgpm_b200.device import hyper_grids, crp_grid
    x = np.array([1.0, np.nan, 2.5, -0.5, 4.0])
    g = hyper_grids(x, 0)
    ref = co.normal_grids([1.0, 2.5, -0.5, 4.0])
    for i, n in enumerate(['m', 'r', 's', 'nu']):
        assert np.array_equal(g[i], ref[n])
    assert np.array_equal(hyper_grids(np.array([0., 1., 2., np.nan]), 1)[0], co.categorical_grids([0., 1., 2.])['alpha'])
    assert np.array_equal(crp_grid(7), co.crp_grids(7))


def test_synth_table_shapes_and_types():
    cct = ['normal', 'categorical', 'normal']
    X, Zv, Zrv = synth.gen_table(200, cct, [None, {'k': 5}, None], n_views=2, clusters_per_view=3, seed=3,
                                 quantize_bits=12, nan_frac=0.05)
    assert X.shape == (200, 3) and len(Zv) == 3 and len(Zrv) == 2
    col = X[:, 1][~np.isnan(X[:, 1])]
    assert np.all(col % 1 == 0) and col.min() >= 0 and col.max() <= 4
    q = X[:, 0][~np.isnan(X[:, 0])] * 4096
    assert np.all(q == np.round(q))
    assert 0 < np.isnan(X).mean() < 0.15


def test_shard_ranges_partition_all_chains():
    for n in (1, 7, 64, 512):
        for w in (1, 2, 3, 8):
            got = [shard_range(n, w, r) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from cgpm_b200.parallel import all_gather_chains, all_reduce_sum, shard_range
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n = 7
    lo, hi = shard_range(n, world, rank)
    local = torch.arange(lo, hi, dtype=torch.float64) * 1.5
    full = all_gather_chains(local, n)
    mat = torch.stack([torch.full((3,), float(i)) for i in range(lo, hi)])
    full2 = all_gather_chains(mat, n)
    acc = all_reduce_sum(torch.ones(2, 2, dtype=torch.float64) * (rank + 1))
    from cgpm_b200.parallel import all_gather_objects, weighted_integrate
    objs = all_gather_objects([{'chain': i} for i in range(lo, hi)])
    lw = weighted_integrate(full.tolist(), [0.5 * x for x in full.tolist()])
    # the array forms of ShardedEngine's bulk queries over a stand-in for the local Engine (no device here):
    # chain c answers c + 0.001 q for query q, and samples [c, q, n, t]
    import types
    import numpy as np
    from cgpm_b200.parallel import ShardedEngine
    sh = ShardedEngine.__new__(ShardedEngine)
    sh.world, sh.rank, sh.num_states_total, sh.lo, sh.hi = world, rank, n, lo, hi
    chains = np.arange(lo, hi, dtype=np.float64)
    sh.local = types.SimpleNamespace(
        _ch=types.SimpleNamespace(dev=types.SimpleNamespace(device=torch.device('cpu'))),
        logpdf_bulk_arrays=lambda tc, tv, cc=None, cv=None: chains[:, None] + 0.001 * np.arange(len(tv))[None, :],
        simulate_bulk_arrays=lambda Q, tc, cc=None, cv=None, N=1: (
            chains[:, None, None, None] + np.zeros((1, Q, N, len(tc) + 1)))[..., :len(tc)])      # a non-contiguous view
    la = sh.logpdf_bulk_arrays([0, 1], np.zeros((5, 2)))
    sa = sh.simulate_bulk_arrays(4, [0, 1], N=3)
    q.put((rank, full.tolist(), full2[:, 0].tolist(), acc[0, 0].item(), [o['chain'] for o in objs], lw,
           la.tolist(), list(sa.shape), sa[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


def test_gather_over_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from cgpm_b200 import hostmath as hm
    for rank, full, full2, acc, objs, lw, la, sa_shape, sa_first in res:
        assert la == [[c + 0.001 * j for j in range(5)] for c in range(7)]           # [num_states, Q] in global chain order
        assert sa_shape == [7, 4, 3, 2] and sa_first == [float(c) for c in range(7)]
        assert full == [i * 1.5 for i in range(7)]
        assert full2 == [float(i) for i in range(7)]
        assert acc == 3.0
        assert objs == list(range(7))                       # python objects gathered in global chain order
        # engine.py:361-371: chains weighted by the density they give the constraints
        assert lw == hm.logmeanexp_weighted([i * 1.5 for i in range(7)], [i * 0.75 for i in range(7)])


def test_constrained_crp_helpers_match_the_reference():
    """cgpm_b200.hostmath restates gu.simulate_crp_constrained[_dependent] / logp_crp_constrained_dependent
    (src/utils/general.py:248-355): same partitions from the same stream, same log-probability."""
    import os
    import sys
    import pytest
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
    if not os.path.exists(os.path.join(ref, 'cgpm', '.built')):
        pytest.skip('oracle/_ref not built')
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from cgpm.utils import general as gu
    from cgpm_b200 import hostmath as hm
    Cd = [[0, 2], [3, 4, 5]]
    Ci = [(0, 1), (2, 3), (6, 7)]
    for seed in range(5):
        a = gu.simulate_crp_constrained(9, 1.7, Cd, Ci, {}, {}, rng=np.random.RandomState(seed))
        r = np.random.RandomState(seed)
        b = hm.simulate_crp_constrained(9, 1.7, Cd, Ci, r)
        assert list(a) == list(b)
        a = gu.simulate_crp_constrained_dependent(9, 0.8, Cd, rng=np.random.RandomState(seed))
        b = hm.simulate_crp_constrained_dependent(9, 0.8, Cd, np.random.RandomState(seed))
        assert list(a) == list(b)
        Z = {i: z for i, z in enumerate(b)}
        assert hm.logp_crp_constrained_dependent(Z, 0.8, Cd) == gu.logp_crp_constrained_dependent(Z, 0.8, Cd)
    assert hm.logp_crp_constrained_dependent({0: 0, 1: 0, 2: 1}, 1.0, [[0, 2]]) == -float('inf')
    with pytest.raises(ValueError):
        hm.validate_dependency_constraints(4, [[0, 1]], [(0, 1)])


def _query_stub(R=50, names=(3, 'b', 7, 11, 'e', 20), lognormal=()):
    """What query._encode reads of a Chains object, without a device."""
    import types
    import torch
    C = len(names)
    ch = types.SimpleNamespace()
    ch.R = R
    ch.col_index = {c: i for i, c in enumerate(names)}
    ch.lognormal = np.array([c in lognormal for c in names], dtype=bool)
    ch.X_raw = np.random.RandomState(0).normal(size=(R, C))
    ch.X_raw[::3, 1] = np.nan
    ch.dev = types.SimpleNamespace(device=torch.device('cpu'))
    return ch


def test_vectorised_query_encoder_equals_the_per_query_loop(monkeypatch):
    """query._encode_fast (hypothetical rows, no lognormal column: one numpy pass) must produce exactly the CSR
    arrays of the per-query loop that mirrors State._validate_cgpm_query (state.py:447-476), in the reference's
    dict order, and must step aside (return None) for every batch the loop treats differently -- observed rows,
    lognormal columns, and queries the reference rejects."""
    from cgpm_b200 import query
    ch = _query_stub()
    names = list(ch.col_index)
    rng = np.random.RandomState(3)
    Q = 200

    def batch(simulate):
        rowids, tl, cl = [], [], []
        for q in range(Q):
            perm = [names[i] for i in rng.permutation(len(names))]
            a, b = rng.randint(0, 4), rng.randint(0, 3)
            t = perm[:a]
            c = perm[a:a + b]
            tl.append(list(t) if simulate else {k: float(rng.normal()) for k in t})
            cl.append(None if (b == 0 and q % 2) else {k: float(rng.normal()) for k in c})
            rowids.append([None, -1, ch.R, ch.R + 17][q % 4])
        return rowids, tl, cl

    def slow(*a):
        monkeypatch.setattr(query, 'FAST_ENCODE_MIN', 10 ** 9)
        try:
            return query._encode(*a)
        finally:
            monkeypatch.setattr(query, 'FAST_ENCODE_MIN', 32)

    for simulate in (False, True):
        rowids, tl, cl = batch(simulate)
        fast = query._encode_fast(ch, rowids, tl, [c if c is not None else {} for c in cl], simulate)
        ref = slow(ch, rowids, tl, cl, simulate)
        assert fast is not None
        for k in ('rowid', 'off', 'col', 'val', 'role'):
            assert fast[k].dtype == ref[k].dtype and np.array_equal(fast[k].numpy(), ref[k].numpy()), (simulate, k)
        assert fast['Q'] == ref['Q'] and not fast['dead'].any() and not fast['jac'].any()
        via = query._encode(ch, rowids, tl, cl, simulate)              # the public path takes the fast encoder
        assert np.array_equal(via['val'].numpy(), ref['val'].numpy())

    # batches the vectorised encoder must leave to the loop
    rowids, tl, cl = batch(False)
    cl = [c or {} for c in cl]
    obs = list(rowids); obs[5] = 4
    assert query._encode_fast(ch, obs, tl, cl, False) is None                       # an observed row
    assert query._encode_fast(_query_stub(lognormal=('b',)), rowids, tl, cl, False) is None
    bad = [dict(t) for t in tl]; bad[7] = {'nope': 1.0}
    assert query._encode_fast(ch, rowids, bad, cl, False) is None                    # unknown column
    with pytest.raises(ValueError, match='Unknown column'):
        query._encode(ch, rowids, bad, cl, False)
    both = [dict(t) for t in tl]; both[9] = {3: 0.5}
    cl2 = [dict(c) for c in cl]; cl2[9] = {3: 0.25}
    assert query._encode_fast(ch, rowids, both, cl2, False) is None                  # target also constrained
    with pytest.raises(ValueError, match='disjoint'):
        query._encode(ch, rowids, both, cl2, False)
    rowids, tl, cl = batch(True)
    cl = [c or {} for c in cl]
    tl[11] = [7, 7]; cl[11] = {}
    assert query._encode_fast(ch, rowids, tl, cl, True) is None                      # repeated target
    with pytest.raises(ValueError, match='unique'):
        query._encode(ch, rowids, tl, cl, True)
    assert query._encode_fast(ch, [None] * 3, [{}, {}, {}], [{}, {}, {}], False) is None    # nothing to encode


def test_sample_decoding_types_and_layout():
    """query._decode_samples: the kernel's [chain][sample][target] array becomes [chain][query] lists of N dicts in
    the query's own target order; integer-valued types as int, lognormal columns exponentiated, the rest as float."""
    import math
    from cgpm_b200 import query, _lib as L
    ch = _query_stub(names=('n', 'c', 'l', 'p'), lognormal=('l',))
    ch.ctype = [L.NORMAL, L.CATEGORICAL, L.NORMAL, L.POISSON]
    targets = [['n', 'l'], ['c'], ['p', 'n', 'c'], ['n']]
    Ns = [2, 1, 3, 2]
    soff = np.concatenate([[0], np.cumsum(Ns)]).astype(np.int32)
    rng = np.random.RandomState(5)
    vals = np.round(rng.normal(2.0, 1.0, size=(3, int(soff[-1]), 3)) * 4)          # [chains][samples][maxt]
    got = query._decode_samples(ch, targets, soff, vals)
    assert len(got) == 3 and all(len(g) == 4 for g in got)
    for i in range(3):
        for q, t in enumerate(targets):
            assert len(got[i][q]) == Ns[q]
            for n_, smp in enumerate(got[i][q]):
                r = soff[q] + n_
                assert list(smp) == t
                for j, c in enumerate(t):
                    v = vals[i, r, j]
                    if c in ('c', 'p'):
                        assert isinstance(smp[c], int) and smp[c] == int(v)
                    elif c == 'l':
                        assert smp[c] == math.exp(v)
                    else:
                        assert isinstance(smp[c], float) and smp[c] == float(v)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: rank 0 prints one JSON line that carries the contract's keys for the reference arm
    (impl, the metric of BASELINE.json, a cpu_baseline describing this very run, an e2e block with no copies), other
    ranks print nothing and exit 0.  Runs the reference's own Engine on a tiny sample (no GPU involved)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, 'oracle', '_ref', 'cgpm', '.built')):
        pytest.skip('oracle/_ref not built here')
    cmd = [sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--burnin', '0', '--ref-rows', '40']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'crosscat_gibbs_cell_updates_per_sec'      # BASELINE.json's metric
    assert line['unit'] == 'cell-updates/s'
    assert line['higher_is_better'] is True and line['n_gpus'] == 1 and line['steps'] == 1 and line['value'] > 0
    cb = line['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == line['value'] and '40 rows' in cb['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    other = subprocess.run(cmd + ['--gpus', '2'], capture_output=True, text=True, timeout=60, env=env)
    assert other.returncode == 0 and other.stdout.strip() == ''
