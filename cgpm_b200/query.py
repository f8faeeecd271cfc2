"""Host side of the batched queries: encode the reference's dict-style queries
(State.logpdf_bulk / simulate_bulk, src/crosscat/state.py:479-523) as CSR
arrays, run the CUDA kernels, decode.  Validation follows
State._validate_cgpm_query (state.py:447-476)."""
from __future__ import annotations

import ctypes as C
import itertools
import math

import numpy as np
import torch

from . import _lib as L


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _validate(ch, rowid, targets, constraints, simulate):
    fresh = rowid is None or not (0 <= rowid < ch.R)
    tcols = list(targets)
    if simulate and len(set(tcols)) != len(tcols):
        raise ValueError('Columns in targets must be unique.')
    for c in list(tcols) + list(constraints):
        if c not in ch.col_index:
            raise ValueError('Unknown column: %s' % (c,))
    X = ch.X_raw
    if not fresh and not simulate and any(not np.isnan(X[rowid, ch.col_index[q]]) for q in tcols):
        raise ValueError('Cannot constrain observed cell.')
    if constraints:
        if set(tcols) & set(constraints):
            raise ValueError('Targets and constraints must be disjoint.')
        if not fresh:
            for e, val in constraints.items():
                x = X[rowid, ch.col_index[e]]
                if not (np.isnan(x) or np.allclose(x, val)):
                    raise ValueError('Cannot use observed cell in constraints.')
    return fresh


FAST_ENCODE_MIN = 32       # batches of at least this many queries try the vectorised encoder first


def _encode_fast(ch, rowids, targets_list, constraints_list, simulate):
    """The arrays of _encode for a batch of hypothetical rows (every rowid None or outside the table:
    state.py:447-476 has nothing to check against the data then) on a table without lognormal columns, built
    with numpy in one piece instead of a Python loop per query.  Returns None whenever the batch needs the general
    path: an observed row, a lognormal column, or any query the loop would reject (it then raises the reference's
    error for the first offending query)."""
    Q = len(rowids)
    if Q == 0 or ch.lognormal.any():
        return None
    try:
        rid = np.fromiter((-1 if r is None else r for r in rowids), dtype=np.int64, count=Q)
    except (TypeError, ValueError):
        return None
    if np.any((rid >= 0) & (rid < ch.R)):
        return None
    cl = [c or {} for c in constraints_list]
    ci = ch.col_index
    try:
        lt = np.fromiter(map(len, targets_list), dtype=np.int64, count=Q)
        lc = np.fromiter(map(len, cl), dtype=np.int64, count=Q)
        nt, nc = int(lt.sum()), int(lc.sum())
        flat = itertools.chain.from_iterable          # iterating a dict yields its keys, in insertion order
        tcol = np.fromiter(map(ci.__getitem__, flat(targets_list)), dtype=np.int64, count=nt)
        ccol = np.fromiter(map(ci.__getitem__, flat(cl)), dtype=np.int64, count=nc)
        if simulate:
            tval = np.zeros(nt)
        else:
            tval = np.fromiter(flat(map(dict.values, targets_list)), dtype=np.float64, count=nt)
        cval = np.fromiter(flat(map(dict.values, cl)), dtype=np.float64, count=nc)
    except (KeyError, TypeError, ValueError, AttributeError):
        return None              # unknown column, unhashable key, a value that is not a number: the loop reports it
    if nt + nc == 0:
        return None
    Cn = len(ci) + 1
    qt = np.repeat(np.arange(Q, dtype=np.int64), lt)
    qc = np.repeat(np.arange(Q, dtype=np.int64), lc)
    tkey, ckey = qt * Cn + tcol, qc * Cn + ccol
    if Q * Cn <= (1 << 26):      # a (query, column) bitmap; sorting the keys otherwise
        seen = np.zeros(Q * Cn, dtype=bool)
        seen[tkey] = True
        repeated = simulate and int(np.count_nonzero(seen)) != nt
        overlap = nc and bool(seen[ckey].any())
    else:
        repeated = simulate and np.unique(tkey).size != nt
        overlap = nc and np.intersect1d(tkey, ckey).size > 0
    if repeated:
        return None              # 'Columns in targets must be unique.'
    if overlap:
        return None              # 'Targets and constraints must be disjoint.'
    per = lt + lc
    off = np.zeros(Q + 1, np.int64)
    np.cumsum(per, out=off[1:])
    if off[-1] >= 2 ** 31:
        return None
    t0 = np.zeros(Q, np.int64); c0 = np.zeros(Q, np.int64)
    np.cumsum(lt[:-1], out=t0[1:]); np.cumsum(lc[:-1], out=c0[1:])
    pos_t = np.repeat(off[:-1] - t0, lt) + np.arange(nt)              # a query's targets first, in dict order ...
    pos_c = np.repeat(off[:-1] + lt - c0, lc) + np.arange(nc)         # ... then its constraints
    n = int(off[-1])
    cols = np.empty(n, np.int32); vals = np.empty(n, np.float64); roles = np.zeros(n, np.uint8)
    cols[pos_t] = tcol; vals[pos_t] = tval
    cols[pos_c] = ccol; vals[pos_c] = cval; roles[pos_c] = 1
    dev = ch.dev.device
    return {
        'Q': Q, 'rowid': torch.full((Q,), -1, dtype=torch.int32, device=dev),
        'off': torch.from_numpy(off.astype(np.int32)).to(dev),
        'col': torch.from_numpy(cols).to(dev), 'val': torch.from_numpy(vals).to(dev), 'role': torch.from_numpy(roles).to(dev),
        'constraints': constraints_list, 'jac': np.zeros(Q), 'dead': np.zeros(Q, dtype=bool),
    }


def _encode(ch, rowids, targets_list, constraints_list, simulate):
    Q = len(rowids)
    if constraints_list is None:
        constraints_list = [{} for _ in range(Q)]
    assert len(targets_list) == Q and len(constraints_list) == Q
    if Q >= FAST_ENCODE_MIN:
        enc = _encode_fast(ch, rowids, targets_list, constraints_list, simulate)
        if enc is not None:
            return enc
    rid = np.full(Q, -1, np.int32)
    off = np.zeros(Q + 1, np.int32)
    cols, vals, roles = [], [], []
    jac = np.zeros(Q)                       # lognormal targets: - log x (lognormal.py:72-80)
    dead = np.zeros(Q, dtype=bool)          # a lognormal target <= 0 has density 0

    def on_device(c, x, q, target):
        i = ch.col_index[c]
        x = float(x)
        if not ch.lognormal[i]:
            return x
        if x <= 0:
            if not target:
                raise ValueError('Zero density constraints: %s' % ({c: x},))
            dead[q] = True
            return 0.0
        if target:
            jac[q] -= math.log(x)
        return math.log(x)

    for q in range(Q):
        t = targets_list[q]
        cns = constraints_list[q] or {}
        fresh = _validate(ch, rowids[q], t, cns, simulate)
        rid[q] = -1 if fresh else int(rowids[q])
        if simulate:
            for c in t:
                cols.append(ch.col_index[c]); vals.append(0.0); roles.append(0)
        else:
            for c, x in t.items():
                cols.append(ch.col_index[c]); vals.append(on_device(c, x, q, True)); roles.append(0)
        for c, x in cns.items():
            cols.append(ch.col_index[c]); vals.append(on_device(c, x, q, False)); roles.append(1)
        off[q + 1] = len(cols)
    dev = ch.dev.device
    enc = {
        'Q': Q, 'rowid': torch.from_numpy(rid).to(dev), 'off': torch.from_numpy(off).to(dev),
        'col': torch.from_numpy(np.asarray(cols or [0], np.int32)).to(dev),
        'val': torch.from_numpy(np.asarray(vals or [0.0], np.float64)).to(dev),
        'role': torch.from_numpy(np.asarray(roles or [0], np.uint8)).to(dev),
        'constraints': constraints_list, 'jac': jac, 'dead': dead,
    }
    return enc


def logpdf_bulk(ch, chains, rowids, targets_list, constraints_list=None):
    """Returns [len(chains)][Q] floats (Engine.logpdf_bulk, engine.py:202-210)."""
    for t in targets_list:
        assert isinstance(t, dict)
    enc = _encode(ch, rowids, targets_list, constraints_list, simulate=False)
    return logpdf_bulk_encoded(ch, chains, enc).tolist()


def logpdf_bulk_encoded(ch, chains, enc):
    dev = ch.dev
    n, Q = len(chains), enc['Q']
    out = torch.empty((n, Q), dtype=torch.float64, device=dev.device)
    flag = torch.zeros((n, Q), dtype=torch.int32, device=dev.device)
    chs = np.ascontiguousarray(chains, dtype=np.int32)
    L.check(dev.lib.cc_logpdf_bulk(C.byref(dev.st), n, chs.ctypes.data_as(C.c_void_p), Q, _dp(enc['rowid']),
                                   _dp(enc['off']), _dp(enc['col']), _dp(enc['val']), _dp(enc['role']),
                                   _dp(out), _dp(flag), dev.stream_ptr()), 'cc_logpdf_bulk')
    res = out.cpu().numpy()
    bad = flag.cpu().numpy()
    if bad.any():
        q = int(np.nonzero(bad.any(axis=0))[0][0])
        raise ValueError('Zero density constraints: %s' % (enc['constraints'][q],))   # sampling.py:78-79
    res = res + enc['jac'][None, :]
    res[:, enc['dead']] = -np.inf
    return res


def _decode_samples(ch, targets_list, soff, vals):
    """[chain][query] lists of sample dicts {column: value} from the kernel's [chain][sample][target] array
    (samples soff[q]..soff[q+1] belong to query q): integer-valued types come back as int, lognormal columns on
    their original scale (the shapes of State.simulate_bulk, state.py:505-523)."""
    Q = len(targets_list)
    integer = (L.CATEGORICAL, L.BERNOULLI, L.POISSON, L.GEOMETRIC)
    plan = []
    for q in range(Q):
        tcols = list(targets_list[q])
        kinds = [1 if ch.ctype[ch.col_index[c]] in integer else 2 if ch.lognormal[ch.col_index[c]] else 0 for c in tcols]
        plan.append((tcols, kinds, range(int(soff[q]), int(soff[q + 1]))))
    res = []
    for i in range(vals.shape[0]):
        rows = vals[i].tolist()                  # Python floats in one pass
        per_chain = []
        for tcols, kinds, span in plan:
            if not any(kinds):
                per_chain.append([dict(zip(tcols, rows[r])) for r in span])
            else:
                per_chain.append([{c: (int(rows[r][j]) if kinds[j] == 1 else math.exp(rows[r][j]) if kinds[j] == 2
                                       else rows[r][j]) for j, c in enumerate(tcols)} for r in span])
        res.append(per_chain)
    return res


def simulate_bulk(ch, chains, rowids, targets_list, constraints_list=None, Ns=None):
    """Returns [len(chains)][Q] lists of N sample dicts (Engine.simulate_bulk,
    engine.py:241-251)."""
    Q = len(rowids)
    if Ns is None:
        Ns = [1] * Q
    for t in targets_list:
        assert isinstance(t, (list, tuple))
    enc = _encode(ch, rowids, targets_list, constraints_list, simulate=True)
    soff = np.zeros(Q + 1, np.int32)
    soff[1:] = np.cumsum([int(n if n is not None else 1) for n in Ns])
    total = int(soff[-1])
    maxt = max(1, max(len(t) for t in targets_list)) if Q else 1
    dev = ch.dev
    n = len(chains)
    out = torch.zeros((2, n, total, maxt), dtype=torch.float64, device=dev.device)
    flag = torch.zeros((n, Q), dtype=torch.int32, device=dev.device)
    chs = np.ascontiguousarray(chains, dtype=np.int32)
    soff_d = torch.from_numpy(soff).to(dev.device)
    # fold the chains' RandomStates into the device key so that Engine.simulate never
    # repeats itself (engine.py:232,352-355 reseeds every state first)
    salt = 0
    for s in chains:
        if ch.rngs[s] is not None:
            salt = (salt * 1000003 + int(ch.rngs[s].randint(0, 2 ** 31 - 1))) & 0xFFFFFFFFFFFFFFFF
    L.check(dev.lib.cc_simulate_bulk(C.byref(dev.st), n, chs.ctypes.data_as(C.c_void_p), Q, _dp(enc['rowid']),
                                     _dp(enc['off']), _dp(enc['col']), _dp(enc['val']), _dp(enc['role']),
                                     _dp(soff_d), total, maxt, _dp(out), _dp(flag),
                                     C.c_uint64(ch.seed ^ salt), C.c_uint64(ch._next_counter()), dev.stream_ptr()),
            'cc_simulate_bulk')
    vals = out[0].cpu().numpy()
    bad = flag.cpu().numpy()
    if bad.any():
        q = int(np.nonzero(bad.any(axis=0))[0][0])
        raise ValueError('Zero density constraints: %s' % (enc['constraints'][q],))
    return _decode_samples(ch, targets_list, soff, vals)


# ---------------------------------------------------------------------------------------------
# Array form of the bulk queries: Q hypothetical rows (rowid = -1) that all ask for the same target columns
# given the same constraint columns -- the shape of a scoring / imputation batch (BASELINE.json configs[4]).
# The dict form above costs a Python loop per query; here the CSR arrays are built with numpy in one piece.
def _encode_arrays(ch, target_cols, target_vals, constraint_cols, constraint_vals, simulate):
    target_cols = list(target_cols)
    constraint_cols = list(constraint_cols or [])
    for c in target_cols + constraint_cols:
        if c not in ch.col_index:
            raise ValueError('Unknown column: %s' % (c,))
    if len(set(target_cols)) != len(target_cols):
        raise ValueError('Columns in targets must be unique.')
    if set(target_cols) & set(constraint_cols):
        raise ValueError('Targets and constraints must be disjoint.')
    nt, nc = len(target_cols), len(constraint_cols)
    cv = np.asarray(constraint_vals, dtype=np.float64).reshape(-1, nc) if nc else None
    if simulate:
        Q = int(target_vals)                 # number of queries
        tv = np.zeros((Q, nt))
    else:
        tv = np.asarray(target_vals, dtype=np.float64).reshape(-1, nt)
        Q = tv.shape[0]
    if cv is not None and cv.shape[0] != Q:
        raise ValueError('constraint values must have one row per query')
    jac = np.zeros(Q)
    dead = np.zeros(Q, dtype=bool)
    tv = tv.copy()
    for j, c in enumerate(target_cols):
        if ch.lognormal[ch.col_index[c]] and not simulate:
            bad = tv[:, j] <= 0
            dead |= bad
            tv[bad, j] = 1.0
            tv[:, j] = np.log(tv[:, j])
            jac -= np.where(bad, 0.0, tv[:, j])
    if cv is not None:
        cv = cv.copy()
        for j, c in enumerate(constraint_cols):
            if ch.lognormal[ch.col_index[c]]:
                if np.any(cv[:, j] <= 0):
                    raise ValueError('Zero density constraints: %s' % ({c: float(cv[cv[:, j] <= 0, j][0])},))
                cv[:, j] = np.log(cv[:, j])
    per = nt + nc
    col_row = np.asarray([ch.col_index[c] for c in target_cols + constraint_cols], np.int32)
    role_row = np.asarray([0] * nt + [1] * nc, np.uint8)
    vals = tv if cv is None else np.concatenate([tv, cv], axis=1)
    dev = ch.dev.device
    return {
        'Q': Q, 'rowid': torch.full((Q,), -1, dtype=torch.int32, device=dev),
        'off': torch.arange(0, (Q + 1) * per, per, dtype=torch.int32, device=dev),
        'col': torch.from_numpy(np.tile(col_row, Q)).to(dev),
        'val': torch.from_numpy(np.ascontiguousarray(vals.reshape(-1))).to(dev),
        'role': torch.from_numpy(np.tile(role_row, Q)).to(dev),
        'constraints': None, 'jac': jac, 'dead': dead, 'constraint_cols': constraint_cols, 'cv': cv,
    }


def logpdf_bulk_arrays(ch, chains, target_cols, target_vals, constraint_cols=None, constraint_vals=None):
    """[len(chains), Q] array of State.logpdf(-1, {target_cols[j]: target_vals[q, j]}, {constraint ...})."""
    enc = _encode_arrays(ch, target_cols, target_vals, constraint_cols, constraint_vals, simulate=False)
    if enc['constraints'] is None:
        enc['constraints'] = _LazyConstraints(enc)
    return logpdf_bulk_encoded(ch, chains, enc)


class _LazyConstraints(object):
    """constraints_list[q] for error messages, built on demand."""

    def __init__(self, enc):
        self.enc = enc

    def __getitem__(self, q):
        cv = self.enc['cv']
        return {} if cv is None else {c: float(cv[q, j]) for j, c in enumerate(self.enc['constraint_cols'])}


def simulate_bulk_arrays(ch, chains, Q, target_cols, constraint_cols=None, constraint_vals=None, N=1):
    """[len(chains), Q, N, len(target_cols)] array of predictive samples of hypothetical rows (categorical
    columns as integer-valued floats, lognormal columns on their original scale)."""
    enc = _encode_arrays(ch, target_cols, Q, constraint_cols, constraint_vals, simulate=True)
    enc['constraints'] = _LazyConstraints(enc)
    N = int(N)
    soff = np.arange(0, (Q + 1) * N, N, dtype=np.int32)
    total = int(Q * N)
    maxt = max(1, len(target_cols))
    dev = ch.dev
    n = len(chains)
    out = torch.zeros((2, n, total, maxt), dtype=torch.float64, device=dev.device)
    flag = torch.zeros((n, Q), dtype=torch.int32, device=dev.device)
    chs = np.ascontiguousarray(chains, dtype=np.int32)
    soff_d = torch.from_numpy(soff).to(dev.device)
    salt = 0
    for s in chains:
        if ch.rngs[s] is not None:
            salt = (salt * 1000003 + int(ch.rngs[s].randint(0, 2 ** 31 - 1))) & 0xFFFFFFFFFFFFFFFF
    L.check(dev.lib.cc_simulate_bulk(C.byref(dev.st), n, chs.ctypes.data_as(C.c_void_p), Q, _dp(enc['rowid']),
                                     _dp(enc['off']), _dp(enc['col']), _dp(enc['val']), _dp(enc['role']),
                                     _dp(soff_d), total, maxt, _dp(out), _dp(flag),
                                     C.c_uint64(ch.seed ^ salt), C.c_uint64(ch._next_counter()), dev.stream_ptr()),
            'cc_simulate_bulk')
    vals = out[0].cpu().numpy().reshape(n, Q, N, maxt)
    bad = flag.cpu().numpy()
    if bad.any():
        q = int(np.nonzero(bad.any(axis=0))[0][0])
        raise ValueError('Zero density constraints: %s' % (enc['constraints'][q],))
    for j, c in enumerate(target_cols):
        if ch.lognormal[ch.col_index[c]]:
            vals[..., j] = np.exp(vals[..., j])
    return vals[..., :len(target_cols)]
