"""Incremental edits of a live GPU engine: incorporate / unincorporate rows,
force_cell, incorporate_dim / unincorporate_dim (src/crosscat/state.py:187-320,
src/mixtures/view.py:149-191, src/crosscat/engine.py:119-174).

These change the shape of the table, so the device arrays are re-created; the
chains are carried across *with their sufficient statistics* (updated on the host
with the reference's own `+=` / `-=`, which are the same IEEE operations), the
hyper grids keep the values they had when each column / CRP was created
(dim.py:176-184 is only run at creation), and the device only recomputes the
cached predictive constants (`cc_refresh_tables`) -- or, for a new column, its
statistics under the view's member order (`cc_rebuild` with a column mask).
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from . import codec
from . import hostmath as hm
from .core import Chains, LOGNORMAL_HYPERS, lognormal_grids
from .device import hyper_grids, crp_grid

CRP_ID_VIEW = 10 ** 7          # state.py:142: cluster-token output id of view v is 10**7 + v


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def full_state(ch, s):
    """Everything that defines chain s, on the host."""
    spec, dl = ch.chain_spec(s, with_stats=True)
    stats = {}
    for vid, vw in spec['views'].items():
        cols = [i for i, v in enumerate(spec['Zv']) if v == vid]
        for pos, k in enumerate(vw['live']):
            cid = int(vw['cid'][pos])
            for i in cols:
                N = float(dl['stat'][L.F_N, k, i])
                if ch.ctype[i] != L.CATEGORICAL:
                    stats[(i, cid)] = [N, float(dl['stat'][L.F_SX, k, i]), float(dl['stat'][L.F_SXX, k, i])]
                else:
                    o0 = ch.dev.catoff_h[i]
                    stats[(i, cid)] = [N, dl['cnt'][k, o0:o0 + ch.ncat[i]].copy()]
    views = OrderedDict()
    for vid, vw in spec['views'].items():
        views[vid] = {'alpha': vw['alpha'], 'Zr': [int(z) for z in vw['Zr']], 'order': [int(r) for r in vw['order']],
                      'grid': np.asarray(vw['grid'], dtype=np.float64)}
    return {'Zv': list(spec['Zv']), 'views': views, 'alpha': spec['alpha'], 'hyp': np.array(spec['hyp']),
            'stats': stats, 'rng': ch.rngs[s], 'diag': ch.diagnostics[s], 'hyper_order': ch.hyper_order[s]}


def _empty(ch_ctype, ncat):
    return [0.0, 0.0, 0.0] if ch_ctype != L.CATEGORICAL else [0.0, np.zeros(ncat)]


def _second(x, ctype):
    """What the column type adds to its second statistic: x*x (normal.py:69), gammaln(x+1) (poisson.py:56), nothing."""
    if ctype == L.NORMAL:
        return x * x
    if ctype == L.POISSON:
        from scipy.special import gammaln
        return gammaln(x + 1)
    return 0.0


def _add(st, x, ctype):
    """normal.py:64-70 / categorical.py:53-61 / bernoulli.py:46-53 / poisson.py:49-57 / exponential.py:47-53 /
    geometric.py:47-53."""
    if ctype != L.CATEGORICAL:
        st[0] += 1.
        st[1] += x
        st[2] += _second(x, ctype)
    else:
        st[0] += 1
        st[1][int(x)] += 1


def _sub(st, x, ctype):
    """normal.py:72-76 / categorical.py:63-66 and the unincorporate of the other primitives."""
    if ctype != L.CATEGORICAL:
        st[0] -= 1
        st[1] -= x
        st[2] -= _second(x, ctype)
    else:
        st[0] -= 1
        st[1][int(x)] -= 1


def check_cell(ch, i, x):
    """The ValueError the column's primitive raises for a value it cannot incorporate."""
    from .device import COUNT_NAMES, count_valid
    if math.isnan(x):
        return
    if ch.ctype[i] == L.CATEGORICAL and not (x % 1 == 0 and 0 <= x < ch.ncat[i]):
        raise ValueError('Invalid Categorical(%d): %s' % (ch.ncat[i], x))
    if ch.ctype[i] in COUNT_NAMES and not bool(count_valid(ch.ctype[i], [x])[0]):
        raise ValueError('Invalid %s: %s' % (COUNT_NAMES[ch.ctype[i]], x))


def build(old, X, outputs, cctypes, distargs, grids, states, rebuild_cols=()):
    """New Chains over table X holding `states`; statistics are uploaded, not recomputed,
    except for the columns listed in rebuild_cols (new columns)."""
    S = len(states)
    ch = Chains(X, outputs=outputs, cctypes=cctypes, distargs=distargs, Cd=old.Cd, Ci=old.Ci, S=S,
                stream=old.stream, device=old.dev.device, Vcap=old.dev.Vcap,
                Kcap=max(old.dev.Kcap, min(X.shape[0] + 1, 1024)), seed=old.seed, grids=grids,
                row_alpha_grid=crp_grid(X.shape[0]), col_alpha_grid=old.dev.col_alpha_grid_h)
    ch.counter = old.counter
    ch.device_edits = getattr(old, 'device_edits', 0)
    dev = ch.dev
    for s, fs in enumerate(states):
        spec = {'Zv': fs['Zv'], 'alpha': fs['alpha'], 'hyp': fs['hyp'],
                'views': OrderedDict((vid, {'alpha': vw['alpha'], 'Zr': np.asarray(vw['Zr'], dtype=np.int64),
                                            'order': np.asarray(vw['order'], dtype=np.int32), 'grid': vw['grid']})
                                     for vid, vw in fs['views'].items())}
        packed = codec.pack_chain(spec, ch.R, ch.C)
        dev.upload_chain(s, packed)
        stat = np.zeros((L.NFIELD, dev.Kcap, ch.C))
        cnt = np.zeros((dev.Kcap, dev.CT))
        for j, vid in enumerate(fs['views']):
            slot_of = {int(cid): p for p, cid in enumerate(packed['cid'][j])}
            for i, v in enumerate(fs['Zv']):
                if v != vid:
                    continue
                for cid, p in slot_of.items():
                    st = fs['stats'].get((i, cid))
                    if st is None:
                        continue
                    stat[L.F_N, p, i] = st[0]
                    if ch.ctype[i] != L.CATEGORICAL:
                        stat[L.F_SX, p, i] = st[1]
                        stat[L.F_SXX, p, i] = st[2]
                    else:
                        o0 = dev.catoff_h[i]
                        cnt[p, o0:o0 + ch.ncat[i]] = st[1]
        dev.t['stat'][s].copy_(torch.from_numpy(stat))
        dev.t['cnt'][s].copy_(torch.from_numpy(cnt))
        ch.rngs[s] = fs['rng']
        ch.views_order[s] = list(range(len(fs['views'])))
        ch.hyper_order[s] = fs['hyper_order']
        ch.diagnostics[s] = fs['diag']
    pairs = [(s, j) for s, fs in enumerate(states) for j in range(len(fs['views']))]
    chs = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
    vws = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
    L.check(dev.lib.cc_refresh_tables(C.byref(dev.st), len(pairs), _p(chs), _p(vws), dev.stream_ptr()), 'cc_refresh_tables')
    if rebuild_cols:
        mask = torch.zeros((S, ch.C), dtype=torch.uint8, device=dev.device)
        mask[:, list(rebuild_cols)] = 1
        dev.rebuild(pairs, mask=mask)
    dev.sync()
    ch._mirror = None
    return ch


def _schema(ch):
    grids = [ch.dev.grids_h[i] for i in range(ch.C)]
    return ch.X_raw.copy(), list(ch.outputs), list(ch.cctypes), [dict(d) for d in ch.distargs], grids      # X on the original scale


def _dev_value(ch, i, x):
    """The value of a cell as the device table and the sufficient statistics hold it: log x for a lognormal
    column (lognormal.py:56-64; a non-positive value is invalid), x otherwise."""
    if i < len(ch.lognormal) and ch.lognormal[i] and not math.isnan(x):
        if x <= 0:
            raise ValueError('Invalid Lognormal: %s' % str(x))
        return math.log(x)
    return x


def _on_device(ch, name, *args):
    """Row edits run on the device (structure_dev.py) unless CGPM_B200_HOST_EDITS is set or the edit changes a
    view's cluster list (then the general path below takes the chains through the host)."""
    import os
    from . import structure_dev
    if os.environ.get('CGPM_B200_HOST_EDITS'):
        return False
    if getattr(structure_dev, name)(ch, *args):
        ch.device_edits = getattr(ch, 'device_edits', 0) + 1
        return True
    return False


def incorporate(ch, rowid, observation, chains=None):
    """State.incorporate (state.py:255-283) for every chain; returns the new Chains."""
    if _on_device(ch, 'incorporate', rowid, observation):
        return ch
    X, outputs, cctypes, distargs, grids = _schema(ch)
    R = X.shape[0]
    if rowid != R:
        raise ValueError('Only contiguous rowids supported: %d' % (rowid,))
    states = [full_state(ch, s) for s in range(ch.S)]
    for fs in states:
        tokens = set(CRP_ID_VIEW + v for v in fs['views'])
        if not all(q in ch.col_index or q in tokens for q in observation):
            raise ValueError('Invalid observation: %s' % observation)
    if any(math.isnan(v) for v in observation.values()):
        raise ValueError('Cannot incorporate nan: %s.' % observation)
    row = np.array([float(observation.get(c, float('nan'))) for c in outputs])
    for i in range(ch.C):
        check_cell(ch, i, row[i])
    X = np.vstack([X, row[None, :]])
    todo = []           # (chain, view id) pairs whose new row gets a Gibbs step
    for s, fs in enumerate(states):
        for vid, vw in fs['views'].items():
            given = observation.get(CRP_ID_VIEW + vid)
            k = 0 if given is None else int(given)
            fresh = k not in vw['Zr']
            vw['Zr'].append(k)
            vw['order'].append(rowid)
            for i, v in enumerate(fs['Zv']):
                if v != vid:
                    continue
                if fresh:
                    fs['stats'][(i, k)] = _empty(ch.ctype[i], ch.ncat[i])       # dim.py:99-101
                if not math.isnan(row[i]):
                    _add(fs['stats'][(i, k)], _dev_value(ch, i, row[i]), ch.ctype[i])
            if given is None:
                todo.append((s, vid))
    new = build(ch, X, outputs, cctypes, distargs, grids, states)
    # view.py:171-172: one Gibbs step for the new row where the cluster was not given
    if todo:
        for s in range(new.S):
            vids = [vid for (ss, vid) in todo if ss == s]
            if vids:
                new.transition_view_rows([s], views=vids, rows=[rowid])
                new.diagnostics[s]['iterations']['rows'] -= 1          # not a kernel iteration of State.transition
                if new.diagnostics[s]['iterations']['rows'] == 0:
                    del new.diagnostics[s]['iterations']['rows']
        new.dev.sync()
        new.dev.check_errors()
    return new


def unincorporate(ch, rowid):
    """State.unincorporate (state.py:285-299)."""
    if _on_device(ch, 'unincorporate', rowid):
        return ch
    X, outputs, cctypes, distargs, grids = _schema(ch)
    R = X.shape[0]
    if rowid != R - 1:
        raise ValueError('Only last rowid may be unincorporated.')
    if R == 1:
        raise ValueError('Cannot unincorporate last rowid.')
    states = [full_state(ch, s) for s in range(ch.S)]
    row = X[rowid]
    for fs in states:
        for vid, vw in fs['views'].items():
            k = vw['Zr'][rowid]
            cols = [i for i, v in enumerate(fs['Zv']) if v == vid]
            for i in cols:
                if not math.isnan(row[i]):
                    _sub(fs['stats'][(i, k)], _dev_value(ch, i, row[i]), ch.ctype[i])
            vw['Zr'].pop()
            vw['order'].remove(rowid)
            if k not in vw['Zr']:                                                # view.py:181-183
                for i in cols:
                    del fs['stats'][(i, k)]
    return build(ch, X[:-1], outputs, cctypes, distargs, grids, states)


def force_cell(ch, rowid, observation):
    """State.force_cell (state.py:302-315)."""
    if _on_device(ch, 'force_cell', rowid, observation):
        return ch
    X, outputs, cctypes, distargs, grids = _schema(ch)
    if not 0 <= rowid < X.shape[0]:
        raise ValueError('Force observation requires existing rowid.')
    if not all(np.isnan(X[rowid, ch.col_index[c]]) for c in observation):
        raise ValueError('Force observations requires NaN cells.')
    states = [full_state(ch, s) for s in range(ch.S)]
    for c, value in observation.items():
        i = ch.col_index[c]
        check_cell(ch, i, float(value))
        X[rowid, i] = float(value)
        for fs in states:
            k = fs['views'][fs['Zv'][i]]['Zr'][rowid]
            _add(fs['stats'][(i, k)], _dev_value(ch, i, float(value)), ch.ctype[i])
    return build(ch, X, outputs, cctypes, distargs, grids, states)


def incorporate_dim(ch, T, outputs_new, inputs=None, cctype=None, distargs=None, v=None):
    """State.incorporate_dim (state.py:187-236)."""
    X, outputs, cctypes, dargs, grids = _schema(ch)
    R = X.shape[0]
    if len(T) != R:
        raise ValueError('%d rows are required, received: %d.' % (R, len(T)))
    if len(outputs_new) != 1:
        raise ValueError('Cannot incorporate multivariate outputs: %s.' % outputs_new)
    if outputs_new[0] in outputs:
        raise ValueError('Specified outputs already exist: %s, %s.' % (outputs_new, outputs))
    if inputs:
        raise ValueError('Cannot incorporate dim with inputs: %s.' % inputs)
    cctype = cctype or 'normal'
    from .core import CCTYPES
    if cctype not in CCTYPES:
        raise ValueError('Only normal, lognormal, categorical, bernoulli, poisson, exponential and geometric cgpms '
                         'supported by cgpm_b200: %s' % (cctype,))
    col = outputs_new[0]
    Tn = np.asarray(T, dtype=np.float64)
    code = CCTYPES[cctype]
    canon = list(codec.HYPER_NAMES[code])
    if cctype == 'lognormal':
        if np.any(Tn[~np.isnan(Tn)] <= 0):
            raise ValueError('Invalid Lognormal: %s' % (Tn[~np.isnan(Tn)][Tn[~np.isnan(Tn)] <= 0][0],))
        g_new = lognormal_grids(Tn)
        names = list(LOGNORMAL_HYPERS)             # the order in which the reference's Dim draws them
    else:
        g_new = hyper_grids(Tn, code)
        names = list(canon)
    states = [full_state(ch, s) for s in range(ch.S)]
    v_add = 0 if v is None else v
    exact = ch.stream == 'numpy'
    for fs in states:
        rng = fs['rng']
        if v_add not in fs['views']:
            # View(self.X, outputs=[crp_id_view + v_add], rng) (view.py:90-103)
            grid = crp_grid(R)
            alpha = rng.choice(grid)
            zr = (hm.crp_simulate_exact if exact else hm.crp_simulate_fast)(R, alpha, rng)
            fs['views'][v_add] = {'alpha': float(alpha), 'Zr': [int(z) for z in zr], 'order': list(range(R)),
                                  'grid': grid}
        hyp_new = np.zeros(4)
        for nm in names:                                                  # dim.py:180-183
            gq = g_new[canon.index(nm)]
            hyp_new[canon.index(nm)] = rng.choice(gq[~np.isnan(gq)])          # (a grid shorter than 30 is NaN-padded)
        fs['hyp'] = np.vstack([fs['hyp'], hyp_new[None, :]])
        fs['Zv'] = fs['Zv'] + [v_add]
        fs['hyper_order'] = list(fs['hyper_order']) + [names]
    X = np.column_stack([X, Tn])
    new = build(ch, X, outputs + [col], cctypes + [cctype], dargs + [dict(distargs) if distargs else {}],
                grids + [g_new], states, rebuild_cols=[len(outputs)])
    chains = list(range(new.S))
    if v is None:
        new.transition_dims(chains, cols=[col])
    else:
        new._count(chains, 'columns')        # transition_dims(cols=[]) still counts an iteration (state.py:804-810)
    new.transition_dim_hypers(chains, cols=[col])
    new.dev.sync()
    new.dev.check_errors()
    return new


def unincorporate_dim(ch, col):
    """State.unincorporate_dim (state.py:238-253)."""
    X, outputs, cctypes, dargs, grids = _schema(ch)
    if len(outputs) == 1:
        raise ValueError('State has only one dim, cannot unincorporate.')
    if col not in outputs:
        raise ValueError('col does not exist: %s, %s.' % (col, outputs))
    i0 = ch.col_index[col]
    states = [full_state(ch, s) for s in range(ch.S)]
    keep = [i for i in range(len(outputs)) if i != i0]
    for fs in states:
        v_del = fs['Zv'][i0]
        if sum(1 for v in fs['Zv'] if v == v_del) == 1:
            del fs['views'][v_del]
        fs['stats'] = {((i if i < i0 else i - 1), k): st for (i, k), st in fs['stats'].items() if i != i0}
        fs['Zv'] = [fs['Zv'][i] for i in keep]
        fs['hyp'] = fs['hyp'][keep]
        fs['hyper_order'] = [fs['hyper_order'][i] for i in keep]
    return build(ch, X[:, keep], [outputs[i] for i in keep], [cctypes[i] for i in keep], [dargs[i] for i in keep],
                 [grids[i] for i in keep], states)
