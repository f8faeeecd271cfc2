"""Shared test helpers (oracle <-> device plumbing)."""
from collections import OrderedDict

import numpy as np

import crosscat_oracle as co
from cgpm_b200 import _lib as L
from cgpm_b200 import codec

CT = {'normal': L.NORMAL, 'categorical': L.CATEGORICAL, 'lognormal': L.NORMAL, 'bernoulli': L.BERNOULLI,
      'poisson': L.POISSON, 'exponential': L.EXPONENTIAL, 'geometric': L.GEOMETRIC}


def small_table(n_rows=60, seed=0, nan=True):
    from cgpm_b200.synth import gen_table
    cct = ['normal', 'categorical', 'normal', 'categorical', 'normal', 'normal']
    da = [None, {'k': 4}, None, {'k': 3}, None, None]
    X, _, _ = gen_table(n_rows, cct, da, n_views=2, clusters_per_view=3, seed=seed)
    if nan:
        rng = np.random.RandomState(seed + 1)
        for _ in range(max(3, n_rows // 15)):
            X[rng.randint(n_rows), rng.randint(len(cct))] = np.nan
    return X, cct, da


def count_table(n_rows=50, seed=0, nan=True):
    """A table with every collapsed primitive of SURVEY 8(f) rank 3: normal, bernoulli, poisson, exponential, geometric,
    categorical, normal, poisson -- two latent views of three clusters each."""
    rng = np.random.RandomState(seed)
    cct = ['normal', 'bernoulli', 'poisson', 'exponential', 'geometric', 'categorical', 'normal', 'poisson']
    da = [None, None, None, None, None, {'k': 3}, None, None]
    za = rng.randint(3, size=n_rows); zb = rng.randint(3, size=n_rows)
    X = np.empty((n_rows, 8))
    X[:, 0] = rng.normal(4.0 * za, 1.0)
    X[:, 1] = (rng.random_sample(n_rows) < np.array([.1, .5, .9])[za]).astype(float)
    X[:, 2] = rng.poisson(np.array([1., 6., 20.])[za])
    X[:, 3] = rng.exponential(np.array([.3, 2., 9.])[zb])
    X[:, 4] = rng.geometric(np.array([.7, .3, .08])[zb]) - 1
    X[:, 5] = np.where(rng.random_sample(n_rows) < .8, zb, rng.randint(3, size=n_rows))
    X[:, 6] = rng.normal(-3.0 * zb, .7)
    X[:, 7] = rng.poisson(np.array([40., 3., 11.])[zb])
    if nan:
        for _ in range(max(3, n_rows // 12)):
            X[rng.randint(n_rows), rng.randint(8)] = np.nan
    return X, cct, da


def ctype_arrays(cct, da):
    ctype = np.array([CT[c] for c in cct], dtype=np.int32)
    ncat = np.array([int(d['k']) if d and 'k' in d else 0 for d in da], dtype=np.int32)
    return ctype, ncat


def oracle_spec(o):
    """OState -> codec spec (views in the oracle's insertion order)."""
    ctype = np.array([CT[o.dim_for(c).cctype] for c in o.outputs], dtype=np.int32)
    views = OrderedDict()
    for v, view in o.views.items():
        views[v] = {'alpha': float(view.crp.alpha),
                    'Zr': np.array([view.Zr[i] for i in range(o.n_rows)]),
                    'order': np.array(list(view.Zr.keys()))}
    hyp = codec.hypers_to_array([o.dim_for(c).hypers for c in o.outputs], ctype)
    return {'Zv': [o.Zv[c] for c in o.outputs], 'views': views,
            'alpha': float(o.crp.alpha), 'hyp': hyp}


def oracle_suffstats(o):
    """{(col index, cluster id): stats} from the oracle."""
    out = {}
    for ci, c in enumerate(o.outputs):
        dim = o.dim_for(c)
        for k, st in dim.clusters.items():
            out[(ci, int(k))] = st
    return out


def device_suffstats(dev, s, ctype):
    """{(col index, cluster id): stats} from the device tables of chain s."""
    dl = dev.download_chain(s, with_stats=True)
    out = {}
    for j, v in enumerate(dl['view_slots']):
        cols = np.nonzero(dl['zv'] == v)[0]
        for pos, k in enumerate(dl['live'][j]):
            cid = int(dl['cid'][j][pos])
            for c in cols:
                N = dl['stat'][L.F_N, k, c]
                if ctype[c] != L.CATEGORICAL:
                    out[(int(c), cid)] = [N, dl['stat'][L.F_SX, k, c], dl['stat'][L.F_SXX, k, c]]
                else:
                    o0 = dev.catoff_h[c]
                    out[(int(c), cid)] = [N, dl['cnt'][k, o0:o0 + dev.ncat_h[c]].copy()]
    return out, dl


def assert_suffstats_equal(a, b):
    assert set(a.keys()) == set(b.keys()), (sorted(a.keys())[:10], sorted(b.keys())[:10])
    for key in a:
        sa, sb = a[key], b[key]
        assert float(sa[0]) == float(sb[0]), (key, sa, sb)
        if len(sa) == 3:
            assert float(sa[1]) == float(sb[1]) and float(sa[2]) == float(sb[2]), (key, sa, sb)
        else:
            assert np.array_equal(np.asarray(sa[1], float), np.asarray(sb[1], float)), (key, sa, sb)


# ---------------------------------------------------------------------------- golden snapshots
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    if name.endswith('.gz'):
        import gzip
        with gzip.open(os.path.join(GOLDEN, name), 'rt') as f:
            return json.load(f)
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def snapshot_digest(snap):
    """sha256 of the canonical snapshot (tests/golden/make_golden_c1.py: digest)."""
    import hashlib
    body = {k: snap[k] for k in ('Zv', 'Zrv', 'row_order', 'alpha', 'view_alphas', 'hypers', 'suffstats')}
    return hashlib.sha256(json.dumps(body, sort_keys=True).encode()).hexdigest()


def golden_table(g):
    X = np.array([[np.nan if v is None else v for v in row] for row in g['X']], dtype=float)
    da = [d if d else None for d in g['distargs']]
    return X, g['cctypes'], da


def hx(x):
    return float(x).hex()


def _pack_stats(st):
    if len(st) == 3:
        return {'N': hx(st[0]), 'sum_x': hx(st[1]), 'sum_x_sq': hx(st[2])}
    return {'N': hx(st[0]), 'counts': [hx(c) for c in st[1]]}


def snapshot_oracle(o):
    md = o.to_metadata()
    out = {
        'Zv': sorted([int(c), int(v)] for c, v in md['Zv']),
        'Zrv': [[int(v), [int(z) for z in zr]] for v, zr in md['Zrv']],
        'row_order': [[int(v), [int(r) for r in order]] for v, order in md['row_order']],
        'alpha': hx(md['alpha']),
        'view_alphas': [[int(v), hx(a)] for v, a in md['view_alphas']],
        'hypers': [{k: hx(v) for k, v in h.items()} for h in md['hypers']],
        'suffstats': [],
    }
    for c in o.outputs:
        dim = o.dim_for(c)
        items = [[int(k), _pack_stats(st)] for k, st in dim.clusters.items()]
        items.append([max(dim.clusters) + 1, _pack_stats(dim._empty())])
        out['suffstats'].append(sorted(items))
    return out


def snapshot_engine(eng, s):
    """Same structure from the GPU engine (chain s)."""
    from cgpm_b200.engine import _chain_metadata
    ch = eng._ch
    md = _chain_metadata(ch, s)
    spec, _ = ch.chain_spec(s)
    out = {
        'Zv': sorted([int(c), int(v)] for c, v in md['Zv']),
        'Zrv': [[int(v), [int(z) for z in zr]] for v, zr in md['Zrv']],
        'row_order': [[int(v), [int(r) for r in vw['order']]] for v, vw in spec['views'].items()],
        'alpha': hx(md['alpha']),
        'view_alphas': [[int(v), hx(a)] for v, a in md['view_alphas']],
        'hypers': [{k: hx(v) for k, v in h.items()} for h in md['hypers']],
        'suffstats': [],
    }
    for col in md['suffstats']:
        items = []
        for k, st in col:
            if 'counts' in st:
                items.append([int(k), {'N': hx(st['N']), 'counts': [hx(c) for c in st['counts']]}])
            elif 'sum_log_x' in st:        # lognormal: the same slots hold the sums of log x
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_log_x']), 'sum_x_sq': hx(st['sum_log_x_sq'])}])
            elif 'x_sum' in st:            # bernoulli
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['x_sum']), 'sum_x_sq': hx(0)}])
            elif 'sum_x_sq' not in st:     # poisson (sum_log_fact_x in the third slot), exponential, geometric
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_x']), 'sum_x_sq': hx(st.get('sum_log_fact_x', 0))}])
            else:
                items.append([int(k), {'N': hx(st['N']), 'sum_x': hx(st['sum_x']), 'sum_x_sq': hx(st['sum_x_sq'])}])
        out['suffstats'].append(sorted(items))
    return out


def assert_snapshot_equal(got, ref, tag=''):
    for key in ('Zv', 'Zrv', 'row_order', 'alpha', 'view_alphas', 'hypers', 'suffstats'):
        assert got[key] == ref[key], (tag, key, got[key], ref[key])


def rel_tol(ref):
    """north_star: conditional log-probabilities within 1e-5 RELATIVE of the reference's.  Purely relative down to
    |ref| = 1e-5; below that the bound stays at 1e-10 absolute -- a conditional is a sum of O(1) fp64 terms, so
    asking for better than 1e-10 absolute near a zero crossing would be asking for more than the reference's own
    arithmetic carries."""
    return 1e-5 * np.maximum(1e-5, np.abs(np.asarray(ref, dtype=float)))
