"""Row edits of a live GPU engine done ON THE DEVICE: State.incorporate / unincorporate / force_cell
(src/crosscat/state.py:255-315, src/mixtures/view.py:149-191) without taking the chains through the host.

The chain state never leaves the device: the arrays with a row dimension are re-sized by device-to-device
copies, the affected sufficient statistics are updated in place with the reference's own `+= x` / `-= x`
(single IEEE additions, so they stay bit-identical), the cached predictive constants are refreshed by
`cc_refresh_tables`, and the Gibbs step of a newly incorporated row runs through the rows kernel.  The
two cases that change the cluster lists themselves -- a new row sent to a cluster id that does not exist yet, and a
removed row that empties its cluster -- return False and are handled by the general host path of structure.py.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from .device import crp_grid

CRP_ID_VIEW = 10 ** 7
ROW_ARRAYS = ('zr', 'order', 'seq', 'moved', 'movedflag')


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _dev_value(ch, i, x):
    if ch.lognormal[i] and not math.isnan(x):
        if x <= 0:
            raise ValueError('Invalid Lognormal: %s' % str(x))
        return math.log(x)
    return x


def _resize_rows(ch, newR):
    """Re-allocate every array with a row dimension for newR rows (device-to-device copies of the common part)."""
    dev = ch.dev
    t, R = dev.t, dev.R
    keep = min(R, newR)
    for name in ROW_ARRAYS:
        a = torch.zeros(t[name].shape[:2] + (newR,), dtype=t[name].dtype, device=dev.device)
        a[:, :, :keep] = t[name][:, :, :keep]
        t[name] = a
    a = torch.full((newR, dev.Cp), float('nan'), dtype=torch.float64, device=dev.device)
    a[:keep] = t['Xrm'][:keep]
    t['Xrm'] = a
    a = torch.full((dev.C, newR), float('nan'), dtype=torch.float64, device=dev.device)
    a[:, :keep] = t['Xcm'][:, :keep]
    t['Xcm'] = a
    if dev.aux_cols:
        a = torch.full((len(dev.aux_cols), newR), float('nan'), dtype=torch.float64, device=dev.device)
        a[:, :keep] = t['Xaux'][:, :keep]
        t['Xaux'] = a
        dev.st.Xaux = a.data_ptr()
    t['aux_zr'] = torch.zeros((dev.S, newR), dtype=torch.int32, device=dev.device)
    dev.aux_zr_all = None               # scratch of the column kernel: re-allocated lazily for the new row count
    dev.arena = None
    dev.st.aux_zr_all = None
    dev.st.arena = None
    dev.st.arena_bytes = 0
    dev.R = newR
    dev.st.R = newR
    for name in ROW_ARRAYS + ('Xrm', 'Xcm', 'aux_zr'):
        setattr(dev.st, name, t[name].data_ptr())
    dev.row_alpha_grid_h = crp_grid(newR)                           # the grid new views will get (view.py:98-99)
    t['row_alpha_grid'].copy_(torch.from_numpy(dev.row_alpha_grid_h))
    dev.X_h = None
    dev._view_cost = {}
    ch.R = newR
    ch._aux_pre = None
    ch._mirror = None


def _live_pairs(ch):
    m = ch.mirror()
    return [(s, int(v)) for s in range(ch.S) for v in np.nonzero(m['nv'][s] > 0)[0]]


def _refresh(ch, pairs):
    dev = ch.dev
    chs = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
    vws = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
    L.check(dev.lib.cc_refresh_tables(C.byref(dev.st), len(pairs), _p(chs), _p(vws), dev.stream_ptr()), 'cc_refresh_tables')


def _update_stats(ch, slot_sc, vals, sign):
    """stat / cnt of cluster slot slot_sc[s, c] += sign * (the cell vals[c]) for every chain s and column c whose
    value is not NaN: N, sum_x, sum_x_sq (x*x rounded first, normal.py:64-76) or the category count
    (categorical.py:53-66).  One IEEE addition per statistic, as in the reference."""
    dev = ch.dev
    t = dev.t
    S, Cn = ch.S, ch.C
    vals = np.asarray(vals, dtype=np.float64)
    ok = ~np.isnan(vals)
    if not ok.any():
        return
    cols = np.nonzero(ok)[0]
    sidx = torch.arange(S, device=dev.device).unsqueeze(1).expand(S, len(cols)).reshape(-1)
    cidx = torch.from_numpy(cols.astype(np.int64)).to(dev.device).unsqueeze(0).expand(S, len(cols)).reshape(-1)
    pidx = slot_sc[:, cidx.view(S, -1)[0]].reshape(-1).long()
    x = torch.from_numpy(vals[cols]).to(dev.device).unsqueeze(0).expand(S, len(cols)).reshape(-1)
    sg = float(sign)
    fN = torch.full_like(x, L.F_N, dtype=torch.long)
    t['stat'].index_put_((sidx, fN, pidx, cidx), torch.full_like(x, sg), accumulate=True)
    isn = torch.from_numpy((ch.ctype[cols] != L.CATEGORICAL)).to(dev.device).unsqueeze(0).expand(S, len(cols)).reshape(-1)
    if bool(isn.any()):
        # the sum types: sum_x, and the second sum -- x*x (normal.py:69), gammaln(x+1) as the host computes it
        # (poisson.py:56), nothing for the others
        from .device import log_fact
        second = np.where(ch.ctype[cols] == L.NORMAL, vals[cols] * vals[cols],
                          np.where(ch.ctype[cols] == L.POISSON, log_fact(vals[cols]), 0.0))
        x2 = torch.from_numpy(second).to(dev.device).unsqueeze(0).expand(S, len(cols)).reshape(-1)
        xs = x[isn]
        t['stat'].index_put_((sidx[isn], torch.full_like(sidx[isn], L.F_SX), pidx[isn], cidx[isn]), sg * xs, accumulate=True)
        t['stat'].index_put_((sidx[isn], torch.full_like(sidx[isn], L.F_SXX), pidx[isn], cidx[isn]), sg * x2[isn], accumulate=True)
    isc = ~isn
    if bool(isc.any()):
        off = torch.from_numpy(dev.catoff_h[cols].astype(np.int64)).to(dev.device).unsqueeze(0).expand(S, len(cols)).reshape(-1)
        kidx = off[isc] + x[isc].long()
        t['cnt'].index_put_((sidx[isc], pidx[isc], kidx), torch.full_like(x[isc], sg), accumulate=True)


def incorporate(ch, rowid, observation):
    """State.incorporate for every chain, on the device.  Returns False (nothing touched) when some view would
    need a cluster that does not exist yet."""
    dev = ch.dev
    R = ch.R
    if rowid != R:
        raise ValueError('Only contiguous rowids supported: %d' % (rowid,))
    m = ch.mirror()
    for s in range(ch.S):
        tokens = set(CRP_ID_VIEW + int(m['view_id'][s][v]) for v in np.nonzero(m['nv'][s] > 0)[0])
        if not all(q in ch.col_index or q in tokens for q in observation):
            raise ValueError('Invalid observation: %s' % observation)
    if any(math.isnan(v) for v in observation.values()):
        raise ValueError('Cannot incorporate nan: %s.' % observation)
    row = np.array([float(observation.get(c, float('nan'))) for c in ch.outputs])
    from .structure import check_cell
    for i in range(ch.C):
        check_cell(ch, i, row[i])
    row_dev = np.array([_dev_value(ch, i, row[i]) for i in range(ch.C)])
    # target cluster of the new row in every view: the given one, else cluster id 0 (view.py:149-172)
    V = dev.Vcap
    kid = np.zeros((ch.S, V), np.int32)
    given = np.zeros((ch.S, V), bool)
    for s in range(ch.S):
        for v in np.nonzero(m['nv'][s] > 0)[0]:
            tok = CRP_ID_VIEW + int(m['view_id'][s][v])
            if tok in observation:
                kid[s, v] = int(observation[tok]); given[s, v] = True
    live = torch.from_numpy(m['nv'] > 0).to(dev.device)
    match = (dev.t['cid'] == torch.from_numpy(kid).to(dev.device).unsqueeze(2)) & (dev.t['nk'] > 0)
    if not bool((match.any(dim=2) | ~live).all()):
        return False
    slot = match.int().argmax(dim=2).to(torch.int32)                      # [S, V]
    _resize_rows(ch, R + 1)
    t = dev.t
    t['Xrm'][R, :ch.C] = torch.from_numpy(row_dev).to(dev.device)
    t['Xcm'][:, R] = torch.from_numpy(row_dev).to(dev.device)
    if dev.aux_cols:
        from .device import log_fact
        t['Xaux'][:, R] = torch.from_numpy(log_fact(row_dev[dev.aux_cols])).to(dev.device)
    t['zr'][:, :, R] = slot
    t['order'][:, :, R] = R
    t['nk'].scatter_add_(2, slot.long().unsqueeze(2), live.to(torch.int32).unsqueeze(2))
    slot_sc = slot.gather(1, t['zv'].long())                              # [S, C]: slot of the row's cluster in each column's view
    _update_stats(ch, slot_sc, row_dev, +1.0)
    ch.X_raw = np.vstack([ch.X_raw, row[None, :]])
    for i in np.nonzero(ch.lognormal)[0]:
        if not math.isnan(row_dev[i]):
            ch._ln_const -= float(row_dev[i])
    _refresh(ch, _live_pairs(ch))
    # view.py:171-172: one Gibbs step for the new row where the cluster was not given
    for s in range(ch.S):
        vids = [int(m['view_id'][s][v]) for v in ch.views_order[s] if not given[s, v]]
        if vids:
            ch.transition_view_rows([s], views=vids, rows=[rowid])
            it = ch.diagnostics[s]['iterations']
            it['rows'] -= 1                                       # not a kernel iteration of State.transition
            if it['rows'] == 0:
                del it['rows']
    dev.sync()
    dev.check_errors()
    return True


def unincorporate(ch, rowid):
    """State.unincorporate (state.py:285-299) on the device; False when the row is the last member of a cluster."""
    dev = ch.dev
    R = ch.R
    if rowid != R - 1:
        raise ValueError('Only last rowid may be unincorporated.')
    if R == 1:
        raise ValueError('Cannot unincorporate last rowid.')
    m = ch.mirror()
    t = dev.t
    live = torch.from_numpy(m['nv'] > 0).to(dev.device)
    slot = t['zr'][:, :, rowid].clone()                                   # [S, V]
    nk_at = t['nk'].gather(2, slot.long().clamp(min=0).unsqueeze(2)).squeeze(2)
    if bool(((nk_at <= 1) & live).any()):
        return False
    row_dev = t['Xrm'][rowid, :ch.C].cpu().numpy()
    slot_sc = slot.gather(1, t['zv'].long())
    _update_stats(ch, slot_sc, row_dev, -1.0)
    t['nk'].scatter_add_(2, slot.long().clamp(min=0).unsqueeze(2), -live.to(torch.int32).unsqueeze(2))
    # the member order loses the row (an OrderedDict pop, crp.py:52): stable compaction
    ar = torch.arange(R, dtype=torch.int32, device=dev.device)
    order = torch.where(live.unsqueeze(2), t['order'], ar.view(1, 1, R))
    t['order'][:, :, :R - 1] = order[order != rowid].view(ch.S, dev.Vcap, R - 1)
    _resize_rows(ch, R - 1)
    for i in np.nonzero(ch.lognormal)[0]:
        if not math.isnan(row_dev[i]):
            ch._ln_const += float(row_dev[i])
    ch.X_raw = ch.X_raw[:-1].copy()
    _refresh(ch, _live_pairs(ch))
    dev.sync()
    return True


def force_cell(ch, rowid, observation):
    """State.force_cell (state.py:302-315): in place, nothing is re-sized."""
    dev = ch.dev
    if not 0 <= rowid < ch.R:
        raise ValueError('Force observation requires existing rowid.')
    if not all(np.isnan(ch.X_raw[rowid, ch.col_index[c]]) for c in observation):
        raise ValueError('Force observations requires NaN cells.')
    vals = np.full(ch.C, np.nan)
    for c, value in observation.items():
        i = ch.col_index[c]
        from .structure import check_cell
        check_cell(ch, i, float(value))
        vals[i] = _dev_value(ch, i, float(value))
    t = dev.t
    for c, value in observation.items():
        i = ch.col_index[c]
        ch.X_raw[rowid, i] = float(value)
        t['Xrm'][rowid, i] = float(vals[i])
        t['Xcm'][i, rowid] = float(vals[i])
        if dev.auxcol_h[i] >= 0:
            from .device import log_fact
            t['Xaux'][int(dev.auxcol_h[i]), rowid] = float(log_fact([vals[i]])[0])
        if ch.lognormal[i]:
            ch._ln_const -= float(vals[i])
    slot_sc = t['zr'][:, :, rowid].gather(1, t['zv'].long())
    _update_stats(ch, slot_sc, vals, +1.0)
    _refresh(ch, _live_pairs(ch))
    dev.X_h = None
    dev.sync()
    return True
