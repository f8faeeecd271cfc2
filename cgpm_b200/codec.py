"""Chain codec: reference-schema chain description <-> dense slot arrays.

The reference keeps a chain as a graph of Python objects whose identifiers are
sparse ints (view ids, cluster ids: src/primitives/crp.py:58-59,142;
src/crosscat/state.py:1113-1115,1143-1154).  The device keeps dense "slots" and
carries the reference ids beside them (`view_id`, `cid`), so that
`State.to_metadata()` (state.py:1189-1240) can be rebuilt exactly.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

HYPER_NAMES = {0: ['m', 'r', 's', 'nu'], 1: ['alpha'],                          # normal, categorical
               2: ['alpha', 'beta'], 3: ['a', 'b'], 4: ['a', 'b'], 5: ['a', 'b']}       # bernoulli, poisson, exponential, geometric
# suffstat names of State.to_metadata (get_suffstats of each primitive): (N, first sum, second sum or None)
SUFFSTAT_NAMES = {0: ('sum_x', 'sum_x_sq'), 2: ('x_sum', None), 3: ('sum_x', 'sum_log_fact_x'), 4: ('sum_x', None), 5: ('sum_x', None)}


def pack_chain(spec, R, C):
    """spec: dict(Zv=[view id per column index], views=OrderedDict(vid ->
    dict(alpha, Zr=[cluster id per row], order=None|[rowids])), alpha, hyp[C][4]).
    View slots are assigned in the insertion order of spec['views']."""
    vids = list(spec['views'].keys())
    slot_of = {vid: j for j, vid in enumerate(vids)}
    zv = np.asarray([slot_of[v] for v in spec['Zv']], dtype=np.int32)
    out = {
        'zv': zv, 'hyp': np.ascontiguousarray(spec['hyp'], dtype=np.float64),
        'col_alpha': float(spec['alpha']),
        'view_slots': list(range(len(vids))), 'view_id': [int(v) for v in vids],
        'nv': [int(np.sum(zv == j)) for j in range(len(vids))],
        'view_alpha': [], 'cid': [], 'nk': [], 'zr': [], 'order': [], 'view_grid': [],
    }
    for vid in vids:
        vw = spec['views'][vid]
        ids = np.asarray(vw['Zr'], dtype=np.int64)
        assert ids.shape == (R,)
        uniq, inv, counts = np.unique(ids, return_inverse=True, return_counts=True)
        out['view_alpha'].append(float(vw['alpha']))
        out['view_grid'].append(vw.get('grid'))
        out['cid'].append(uniq.astype(np.int32))
        out['nk'].append(counts.astype(np.int32))
        out['zr'].append(inv.astype(np.int32))
        order = vw.get('order')
        out['order'].append(np.arange(R, dtype=np.int32) if order is None
                            else np.asarray(order, dtype=np.int32))
    return out


def unpack_chain(dl, views_order=None):
    """Inverse of pack_chain from DeviceChains.download_chain output.
    views_order: view slots in the reference's insertion order (default:
    ascending view id)."""
    slots = dl['view_slots']
    idx = {v: j for j, v in enumerate(slots)}
    if views_order is None:
        views_order = [slots[j] for j in np.argsort(dl['view_id'], kind='stable')]
    vid_of_slot = {v: dl['view_id'][idx[v]] for v in slots}
    spec = {
        'Zv': [int(vid_of_slot[int(z)]) for z in dl['zv']],
        'alpha': dl['col_alpha'], 'hyp': dl['hyp'], 'views': OrderedDict(),
    }
    for v in views_order:
        j = idx[v]
        ids = dl['cid_by_slot'][j][dl['zr'][j]]
        spec['views'][int(vid_of_slot[v])] = {
            'alpha': dl['view_alpha'][j], 'Zr': ids.astype(np.int64),
            'order': dl['order'][j].copy(), 'slot': v, 'grid': (dl['view_grid'][j] if 'view_grid' in dl else None),
            'cid': dl['cid'][j], 'nk': dl['nk'][j], 'live': dl['live'][j],
        }
    return spec


def hypers_to_array(hypers, ctype):
    """list of reference hyper dicts -> [C][4]."""
    C = len(ctype)
    out = np.zeros((C, 4))
    for c in range(C):
        names = HYPER_NAMES[int(ctype[c])]
        for i, n in enumerate(names):
            out[c, i] = float(hypers[c][n])
    return out


def hypers_from_array(hyp, ctype):
    out = []
    for c in range(len(ctype)):
        names = HYPER_NAMES[int(ctype[c])]
        out.append(OrderedDict((n, float(hyp[c, i])) for i, n in enumerate(names)))
    return out
