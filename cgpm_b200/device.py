"""Device-resident chain state: torch tensors behind the `cc_state` C struct.

PyTorch is used for device memory and streams only; every computation on the
tensors is done by the CUDA kernels of libcgpm_b200.so through the C ABI
(include/cgpm_b200.h).  Layouts are documented there.
"""
from __future__ import annotations

import collections
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

MAXCAT = 256


def log_linspace(a, b, n):
    """src/utils/general.py:203-205."""
    return np.exp(np.linspace(math.log(a), math.log(b), n))


def hyper_grids(xcol, ctype, n_grid=L.NGRID):
    """Empirical-Bayes hyper grids of one column from its non-NaN values.
    normal: src/primitives/normal.py:128-139; categorical:
    src/primitives/categorical.py:110-114.  Returns float64 [4][n_grid]."""
    X = xcol[~np.isnan(xcol)]
    g = np.zeros((4, n_grid))
    if ctype == L.NORMAL:
        N = len(X) + 1.
        ssqdev = np.var(X) * len(X) + 1.
        g[0] = np.linspace(min(X), max(X) + 5, n_grid)
        g[1] = log_linspace(1. / N, N, n_grid)
        g[2] = log_linspace(ssqdev / 100., ssqdev, n_grid)
        g[3] = log_linspace(1., N, n_grid)
    elif ctype == L.CATEGORICAL:
        g[0] = log_linspace(1., float(len(X)), n_grid)
    elif ctype == L.BERNOULLI:                                   # bernoulli.py:107-112
        g[0] = log_linspace(1., float(len(X)), n_grid)
        g[1] = log_linspace(1., float(len(X)), n_grid)
    elif ctype == L.POISSON:                                     # poisson.py:119-126: integers only, possibly fewer than n_grid
        a = np.unique(np.round(np.linspace(1, len(X), n_grid)))
        g[0] = np.nan
        g[0, :len(a)] = a
        g[1] = log_linspace(.1, float(len(X)), n_grid)
    elif ctype == L.EXPONENTIAL:                                 # exponential.py:109-114
        g[0] = log_linspace(.5, float(len(X)), n_grid)
        g[1] = log_linspace(.5, float(len(X)), n_grid)
    elif ctype == L.GEOMETRIC:                                   # geometric.py:105-110
        g[0] = log_linspace(1, float(len(X)) / 2., n_grid)
        g[1] = log_linspace(.1, float(len(X)) / 2., n_grid)
    else:
        raise ValueError('unknown column type %s' % (ctype,))
    return g


COUNT_NAMES = {L.BERNOULLI: 'Bernoulli', L.POISSON: 'Poisson', L.EXPONENTIAL: 'Exponential', L.GEOMETRIC: 'Geometric'}


def count_valid(ctype, col):
    """Values the primitive's incorporate accepts (bernoulli.py:49, poisson.py:52, exponential.py:50, geometric.py:50);
    NaN cells are never incorporated."""
    col = np.asarray(col, dtype=np.float64)
    nan = np.isnan(col)
    if ctype == L.BERNOULLI:
        return nan | (col == 0) | (col == 1)
    if ctype == L.EXPONENTIAL:
        return nan | (col >= 0)
    return nan | ((col % 1 == 0) & (col >= 0))


def log_fact(col):
    """gammaln(x + 1) of every cell as the reference adds it to sum_log_fact_x (poisson.py:56; scipy.special.gammaln)."""
    from scipy.special import gammaln
    col = np.asarray(col, dtype=np.float64)
    out = np.full(col.shape, np.nan)
    ok = ~np.isnan(col)
    out[ok] = gammaln(col[ok] + 1)
    return out


def crp_grid(n, n_grid=L.NGRID):
    """src/primitives/crp.py:148-152 with n members."""
    return log_linspace(1. / n, n, n_grid)


# CUDA kernels launched per C-ABI call (for the bench's gpu_launches count)
# kernels per C-ABI call (a lower bound: cc_transition_rows launches one kernel per view class present, up to three)
KERNELS_PER_CALL = {'cc_rebuild': 2, 'cc_refresh_tables': 1, 'cc_logpdf_score': 1, 'cc_transition_rows': 1, 'cc_col_aux': 1,
                    'cc_col_marginals': 2, 'cc_col_scan': 1, 'cc_col_install': 1,
                    'cc_transition_col_hypers': 1, 'cc_transition_alphas': 1, 'cc_logpdf_bulk': 1,
                    'cc_simulate_bulk': 1}


class _CountingLib(object):
    """Forwards to the ctypes library and counts the kernels each call launches."""

    def __init__(self, lib, counts):
        self._lib, self._counts = lib, counts

    def __getattr__(self, name):
        f = getattr(self._lib, name)
        n = KERNELS_PER_CALL.get(name)
        if n is None:
            return f
        counts = self._counts

        def call(*args):
            counts[name] += n
            return f(*args)
        return call


class DeviceChains(object):
    """S CrossCat chains over one shared table, resident on one GPU."""

    def __init__(self, X, ctype, ncat, S, Vcap=None, Kcap=None, device=None, grids=None,
                 row_alpha_grid=None, col_alpha_grid=None):
        L.load()          # fail loudly when the CUDA extension is missing
        if not torch.cuda.is_available():
            raise L.CgpmB200Error('cgpm_b200 needs a CUDA device (no CPU fallback).')
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        R, Cn = X.shape
        ctype = np.asarray(ctype, dtype=np.int32)
        ncat = np.asarray(ncat, dtype=np.int32)
        if np.any(ncat > MAXCAT):
            raise ValueError('categorical(k) supports k <= %d' % MAXCAT)
        for c in range(Cn):
            if ctype[c] == L.CATEGORICAL:
                col = X[:, c]
                ok = np.isnan(col) | ((col % 1 == 0) & (col >= 0) & (col < ncat[c]))
                if not np.all(ok):      # categorical.py:56-57
                    bad = col[~ok][0]
                    raise ValueError('Invalid Categorical(%d): %s' % (ncat[c], bad))
            elif ctype[c] in COUNT_NAMES:
                ok = count_valid(ctype[c], X[:, c])
                if not np.all(ok):
                    raise ValueError('Invalid %s: %s' % (COUNT_NAMES[ctype[c]], X[:, c][~ok][0]))
        self.device = torch.device('cuda', torch.cuda.current_device()) \
            if device is None else torch.device(device)
        self.R, self.C, self.S = R, Cn, int(S)
        self.Cp = (Cn + 1) & ~1
        self.Vcap = int(Vcap) if Vcap else int(min(Cn, 16))
        # cluster slots per view: one more than the clusters a view can hold (the last slot keeps the empty-cluster
        # record).  R + 1 always suffices; smaller values grow on demand (grow_kcap), so this is only a first guess.
        self.Kcap = int(Kcap) if Kcap else int(min(max(R + 1, 4), 1024))
        if self.Kcap < 4:
            raise ValueError('Kcap must be at least 4')
        self.Kcap = min(self.Kcap, max(R + 1, 4))
        self.ctype_h, self.ncat_h = ctype, ncat
        catoff = np.zeros(Cn, dtype=np.int32)
        catoff[1:] = np.cumsum(ncat)[:-1]
        self.catoff_h = catoff
        self.CT = int(max(int(ncat.sum()), 1))
        self.X_h = X
        # grids are fixed when a column / CRP is created (dim.py:176-184); callers that edit the
        # table afterwards (incorporate rows / dims) pass the original ones
        self.grids_h = np.stack([grids[c] if grids is not None and grids[c] is not None else hyper_grids(X[:, c], ctype[c])
                                 for c in range(Cn)])
        self.row_alpha_grid_h = crp_grid(R) if row_alpha_grid is None else np.asarray(row_alpha_grid, dtype=np.float64)
        self.col_alpha_grid_h = crp_grid(Cn) if col_alpha_grid is None else np.asarray(col_alpha_grid, dtype=np.float64)
        dev = self.device
        f64, i32 = torch.float64, torch.int32
        Xp = np.full((R, self.Cp), np.nan)
        Xp[:, :Cn] = X
        t = {}
        t['Xrm'] = torch.from_numpy(Xp).to(dev)
        t['Xcm'] = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)
        t['ctype'] = torch.from_numpy(ctype).to(dev)
        t['ncat'] = torch.from_numpy(ncat).to(dev)
        t['catoff'] = torch.from_numpy(catoff).to(dev)
        t['grids'] = torch.from_numpy(self.grids_h).to(dev)
        # per-cell constants of the poisson columns: gammaln(x + 1) exactly as the host computes it
        self.aux_cols = [c for c in range(Cn) if ctype[c] == L.POISSON]
        auxcol = np.full(Cn, -1, np.int32)
        auxcol[self.aux_cols] = np.arange(len(self.aux_cols))
        self.auxcol_h = auxcol
        if self.aux_cols:
            t['Xaux'] = torch.from_numpy(np.ascontiguousarray(np.stack([log_fact(X[:, c]) for c in self.aux_cols]))).to(dev)
            t['auxcol'] = torch.from_numpy(auxcol).to(dev)
        t['row_alpha_grid'] = torch.from_numpy(self.row_alpha_grid_h).to(dev)
        t['col_alpha_grid'] = torch.from_numpy(self.col_alpha_grid_h).to(dev)
        S_, V, K = self.S, self.Vcap, self.Kcap
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        t['hyp'] = z((S_, Cn, 4), f64)
        t['zv'] = z((S_, Cn), i32)
        t['nv'] = z((S_, V), i32)
        t['view_id'] = torch.full((S_, V), -1, dtype=i32, device=dev)
        t['col_alpha'] = torch.ones((S_,), dtype=f64, device=dev)
        t['view_alpha'] = torch.ones((S_, V), dtype=f64, device=dev)
        t['view_grid'] = torch.from_numpy(np.tile(self.row_alpha_grid_h, (S_, V, 1))).to(dev)
        t['nclu'] = z((S_, V), i32)
        t['live'] = torch.full((S_, V, K), -1, dtype=i32, device=dev)
        t['cid'] = z((S_, V, K), i32)
        t['nk'] = z((S_, V, K), i32)
        t['zr'] = z((S_, V, R), i32)
        t['order'] = z((S_, V, R), i32)
        t['stat'] = z((S_, L.NFIELD, K, Cn), f64)
        t['cnt'] = z((S_, K, self.CT), f64)
        t['err'] = z((S_,), i32)
        t['rows_prog'] = z((S_, V, 4), i32)
        t['seq'] = z((S_, V, R), i32)
        t['moved'] = z((S_, V, R), i32)
        t['movedflag'] = z((S_, V, R), torch.uint8)
        t['colL'] = z((S_, Cn, V + 1), f64)
        t['aux_zr'] = z((S_, R), i32)
        t['aux_K'] = z((S_,), i32)
        t['aux_alpha'] = z((S_, Cn), f64)
        t['aux_K_all'] = z((S_, Cn), i32)
        self.t = t
        self.arena = None
        st = L.cc_state()
        st.R, st.C, st.Cp, st.S = R, Cn, self.Cp, S_
        st.Vcap, st.Kcap, st.CT = V, K, self.CT
        for name in ('Xrm', 'Xcm', 'ctype', 'ncat', 'catoff', 'grids', 'row_alpha_grid',
                     'col_alpha_grid', 'hyp', 'zv', 'nv', 'view_id', 'col_alpha', 'view_alpha', 'view_grid',
                     'nclu', 'live', 'cid', 'nk', 'zr', 'order', 'stat', 'cnt', 'err', 'rows_prog', 'seq',
                     'moved', 'movedflag', 'colL', 'aux_zr', 'aux_K', 'aux_alpha', 'aux_K_all'):
            setattr(st, name, t[name].data_ptr())
        st.Xaux = t['Xaux'].data_ptr() if self.aux_cols else None
        st.auxcol = t['auxcol'].data_ptr() if self.aux_cols else None
        st.arena = None
        st.arena_bytes = 0
        st.aux_zr_all = None
        self.aux_zr_all = None
        self.st = st
        self.lib_calls = collections.Counter()
        self.lib = _CountingLib(L.load(), self.lib_calls)
        self._rows_pending = None
        self.kcap_grown = 0
        self._view_cost = {}          # (chain, view slot) -> kilocycles of the view's last rows sweep (launch order)

    # ------------------------------------------------------------------ helpers
    def stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        torch.cuda.synchronize(self.device)

    def check_errors(self):
        err = self.t['err'].cpu().numpy()
        if np.any(err != 0):
            bad = np.nonzero(err)[0]
            code = int(err[bad[0]])
            self.t['err'].zero_()
            if code == L.ERR_CAPACITY:
                raise L.CgpmB200Error(
                    'cluster/view capacity exceeded in chains %s (Kcap=%d, Vcap=%d): '
                    'recreate the engine with larger capacities' % (bad.tolist(), self.Kcap, self.Vcap))
            raise L.CgpmB200Error('numerical failure (NaN or zero-mass conditional) in chains %s'
                                  % (bad.tolist(),))

    # ------------------------------------------------------- chain load / read
    def upload_chain(self, s, ch):
        """ch: dict from codec.pack_chain (numpy arrays, slot form)."""
        t, dev = self.t, self.device
        V, K, R = self.Vcap, self.Kcap, self.R
        nviews = len(ch['view_slots'])
        if nviews > V or any(v >= V for v in ch['view_slots']):
            self.grow_vcap(max(nviews, max(ch['view_slots']) + 1))
            V = self.Vcap
        kmax = max([len(c_) for c_ in ch['cid']] or [0])
        if kmax > K - 1:
            self.grow_kcap(kmax + 1 + max(16, kmax // 8))
            K = self.Kcap
        t['hyp'][s].copy_(torch.from_numpy(ch['hyp']))
        t['zv'][s].copy_(torch.from_numpy(ch['zv']))
        nv = np.zeros(V, np.int32); view_id = np.full(V, -1, np.int32)
        valpha = np.ones(V); nclu = np.zeros(V, np.int32)
        live = np.full((V, K), -1, np.int32); cid = np.zeros((V, K), np.int32)
        nk = np.zeros((V, K), np.int32)
        for j, v in enumerate(ch['view_slots']):
            Kv = len(ch['cid'][j])
            if Kv > K - 1:
                raise L.CgpmB200Error('view needs %d clusters > Kcap-1=%d' % (Kv, K - 1))
            nv[v] = ch['nv'][j]; view_id[v] = ch['view_id'][j]; valpha[v] = ch['view_alpha'][j]
            nclu[v] = Kv
            live[v, :Kv] = np.arange(Kv); cid[v, :Kv] = ch['cid'][j]; nk[v, :Kv] = ch['nk'][j]
            if ch.get('view_grid') is not None and ch['view_grid'][j] is not None:
                t['view_grid'][s, v].copy_(torch.from_numpy(np.asarray(ch['view_grid'][j], dtype=np.float64)))
            else:
                t['view_grid'][s, v].copy_(torch.from_numpy(self.row_alpha_grid_h))
            t['zr'][s, v].copy_(torch.from_numpy(ch['zr'][j]))
            t['order'][s, v].copy_(torch.from_numpy(ch['order'][j]))
        t['nv'][s].copy_(torch.from_numpy(nv)); t['view_id'][s].copy_(torch.from_numpy(view_id))
        t['view_alpha'][s].copy_(torch.from_numpy(valpha)); t['nclu'][s].copy_(torch.from_numpy(nclu))
        t['live'][s].copy_(torch.from_numpy(live)); t['cid'][s].copy_(torch.from_numpy(cid))
        t['nk'][s].copy_(torch.from_numpy(nk))
        t['col_alpha'][s] = float(ch['col_alpha'])

    def download_chain(self, s, with_stats=False):
        t = self.t
        out = {}
        nv = t['nv'][s].cpu().numpy()
        slots = [int(v) for v in np.nonzero(nv > 0)[0]]
        out['view_slots'] = slots
        out['nv'] = [int(nv[v]) for v in slots]
        view_id = t['view_id'][s].cpu().numpy()
        out['view_id'] = [int(view_id[v]) for v in slots]
        valpha = t['view_alpha'][s].cpu().numpy()
        out['view_alpha'] = [float(valpha[v]) for v in slots]
        out['col_alpha'] = float(t['col_alpha'][s].item())
        vg = t['view_grid'][s].cpu().numpy()
        out['view_grid'] = [vg[v].copy() for v in slots]
        out['zv'] = t['zv'][s].cpu().numpy()
        out['hyp'] = t['hyp'][s].cpu().numpy()
        nclu = t['nclu'][s].cpu().numpy()
        live = t['live'][s].cpu().numpy(); cid = t['cid'][s].cpu().numpy(); nk = t['nk'][s].cpu().numpy()
        out['live'] = [live[v, :nclu[v]].copy() for v in slots]
        out['cid'] = [cid[v][live[v, :nclu[v]]].copy() for v in slots]
        out['nk'] = [nk[v][live[v, :nclu[v]]].copy() for v in slots]
        out['zr'] = [t['zr'][s, v].cpu().numpy() for v in slots]         # cluster slots
        out['cid_by_slot'] = [cid[v].copy() for v in slots]
        out['order'] = [t['order'][s, v].cpu().numpy() for v in slots]
        if with_stats:
            out['stat'] = t['stat'][s].cpu().numpy()
            out['cnt'] = t['cnt'][s].cpu().numpy()
        return out

    # ------------------------------------------------------------------ kernels
    def rebuild(self, pairs=None, mask=None):
        """Recompute tables for (chain, view slot) pairs (None = everything);
        mask: optional uint8 device tensor [S][C] restricting the columns."""
        mp = C.c_void_p(mask.data_ptr()) if mask is not None else None
        if pairs is None:
            rc = self.lib.cc_rebuild(C.byref(self.st), 0, None, None, mp, self.stream_ptr())
        else:
            if len(pairs) == 0:
                return
            ch = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
            vw = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
            rc = self.lib.cc_rebuild(C.byref(self.st), len(pairs), ch.ctypes.data_as(C.c_void_p),
                                     vw.ctypes.data_as(C.c_void_p), mp, self.stream_ptr())
        L.check(rc, 'cc_rebuild')

    def logpdf_score(self):
        out = torch.empty((self.S,), dtype=torch.float64, device=self.device)
        L.check(self.lib.cc_logpdf_score(C.byref(self.st), C.c_void_p(out.data_ptr()), self.stream_ptr()),
                'cc_logpdf_score')
        return out

    def transition_rows(self, work, perm=None, u=None, trace=None, trace_stride=0, seed=0, counter=0, defer=False):
        """work: list of dicts(chain, view, n, perm_is_index, perm_off, u_off, trace_off).
        A sweep that runs out of cluster slots stops in front of the move that needs one (rows_prog status
        CC_ROWS_GROW); finish_rows() then grows Kcap and resumes it with the same row -- nothing is lost.  With
        defer=True the caller promises to call finish_rows() itself (after enqueueing work of its own that should
        run beside the sweep); otherwise it is called here."""
        n = len(work)
        if n == 0:
            return
        self._rows_pending = dict(work=work, perm=perm, u=u, trace=trace, trace_stride=int(trace_stride), seed=seed,
                                  counter=counter)
        self._launch_rows(resume=0)
        if not defer:
            self.finish_rows()

    def _launch_rows(self, resume):
        pd = self._rows_pending
        work = pd['work']
        n = len(work)
        arr = (L.cc_row_work * n)()
        for i, w in enumerate(work):
            arr[i].chain = w['chain']; arr[i].view = w['view']; arr[i].n = w['n']
            arr[i].perm_is_index = w.get('perm_is_index', 1)
            arr[i].perm_off = w.get('perm_off', -1)
            arr[i].u_off = w.get('u_off', -1)
            arr[i].trace_off = w.get('trace_off', -1)
            arr[i].n_cols = w.get('n_cols', 0)
            arr[i].n_normal = w.get('n_normal', 0)
            arr[i].resume = resume
            arr[i].cost_hint = self._view_cost.get((w['chain'], w['view']), 0)
        perm, u, trace = pd['perm'], pd['u'], pd['trace']
        rc = self.lib.cc_transition_rows(
            C.byref(self.st), n, arr,
            C.c_void_p(perm.data_ptr()) if perm is not None else None,
            C.c_void_p(u.data_ptr()) if u is not None else None,
            C.c_void_p(trace.data_ptr()) if trace is not None else None,
            pd['trace_stride'], C.c_uint64(pd['seed']), C.c_uint64(pd['counter']), self.stream_ptr())
        L.check(rc, 'cc_transition_rows')

    def finish_rows(self):
        """Complete the rows sweep launched last: while some view stopped for lack of cluster slots, grow Kcap and
        resume.  Needs one small device-to-host read per sweep whenever Kcap < R + 1."""
        if self._rows_pending is None:
            return
        try:
            while self.Kcap < self.R + 1:
                prog = self.t['rows_prog'].cpu().numpy()                      # synchronises the stream
                worst = max(int(prog[w['chain'], w['view'], 2]) for w in self._rows_pending['work'])
                for w in self._rows_pending['work']:
                    self._view_cost[(w['chain'], w['view'])] = int(prog[w['chain'], w['view'], 3])
                if worst != L.ROWS_GROW:
                    if worst != L.ROWS_DONE:
                        raise L.CgpmB200Error('rows sweep left in state %d' % worst)
                    break
                self.grow_kcap(2 * self.Kcap)
                self._launch_rows(resume=1)
        finally:
            self._rows_pending = None

    # ------------------------------------------------------------------ capacity
    def grow_callback(self):
        """The cc_grow_fn handed to cc_transition: re-allocates through grow_kcap / grow_vcap, which update the
        cc_state struct in place."""
        if getattr(self, '_grow_ref', None) is None:
            def cb(_user, min_kcap, min_vcap):
                try:
                    if min_kcap > self.Kcap:
                        self.grow_kcap(min_kcap)
                    if min_vcap > self.Vcap:
                        self.grow_vcap(min_vcap)
                    return 0
                except Exception:            # out of memory: cc_transition reports the capacity error
                    return 1
            self._grow_ref = L.GROW_FN(cb)
        return self._grow_ref

    def grow_kcap(self, want):
        """Re-allocate the per-cluster arrays with at least `want` cluster slots per view (at most R + 1, which
        always suffices).  Slots keep their numbers; the empty-cluster record moves to the new last slot."""
        new = int(min(max(want, self.Kcap + 1), self.R + 1))
        old = self.Kcap
        if new <= old:
            return
        t, dev = self.t, self.device
        S_, V, Cn = self.S, self.Vcap, self.C
        # the per-cluster tables are dense over (chain, slot, column): old and new copies exist side by side for a moment
        need = S_ * new * (8 * L.NFIELD * Cn + 8 * self.CT + 12 * V)
        free = torch.cuda.mem_get_info(dev)[0] + torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        if need > free:
            raise L.CgpmB200Error(
                'a view needs %d cluster slots: the per-cluster tables of %d chains x %d columns would take %.1f GB '
                '(%.1f GB free).  Views of this size come from columns whose fixed hyperparameters favour one cluster '
                'per row; transition column_hypers, or run fewer chains per device' % (want, S_, Cn, need / 2**30, free / 2**30))
        for name, fill in (('live', -1), ('cid', 0), ('nk', 0)):
            a = torch.full((S_, V, new), fill, dtype=torch.int32, device=dev)
            a[:, :, :old - 1] = t[name][:, :, :old - 1]
            t[name] = a
        stat = torch.zeros((S_, L.NFIELD, new, Cn), dtype=torch.float64, device=dev)
        stat[:, :, :old - 1] = t['stat'][:, :, :old - 1]
        stat[:, :, new - 1] = t['stat'][:, :, old - 1]
        t['stat'] = stat
        cnt = torch.zeros((S_, new, self.CT), dtype=torch.float64, device=dev)
        cnt[:, :old - 1] = t['cnt'][:, :old - 1]
        cnt[:, new - 1] = t['cnt'][:, old - 1]
        t['cnt'] = cnt
        self.Kcap = new
        self.st.Kcap = new
        for name in ('live', 'cid', 'nk', 'stat', 'cnt'):
            setattr(self.st, name, t[name].data_ptr())
        self.kcap_grown += 1

    def grow_vcap(self, want):
        """Re-allocate the per-view arrays with at least `want` view slots per chain (at most C, which always
        suffices).  Slots keep their numbers; the auxiliary-view column of colL moves to the new last position."""
        new = int(min(max(want, self.Vcap + 1), max(self.C, self.Vcap)))
        old = self.Vcap
        if new <= old:
            return
        t, dev = self.t, self.device
        S_, R, K, Cn = self.S, self.R, self.Kcap, self.C

        def widen(name, shape_tail, dtype, fill):
            a = torch.full((S_, new) + shape_tail, fill, dtype=dtype, device=dev)
            a[:, :old] = t[name]
            t[name] = a
        i32, f64 = torch.int32, torch.float64
        widen('nv', (), i32, 0)
        widen('view_id', (), i32, -1)
        widen('view_alpha', (), f64, 1.0)
        widen('nclu', (), i32, 0)
        widen('rows_prog', (4,), i32, 0)
        widen('live', (K,), i32, -1)
        widen('cid', (K,), i32, 0)
        widen('nk', (K,), i32, 0)
        vg = torch.from_numpy(np.tile(self.row_alpha_grid_h, (S_, new, 1))).to(dev)
        vg[:, :old] = t['view_grid']
        t['view_grid'] = vg
        for name, dt in (('zr', i32), ('order', i32), ('seq', i32), ('moved', i32), ('movedflag', torch.uint8)):
            widen(name, (R,), dt, 0)
        colL = torch.zeros((S_, Cn, new + 1), dtype=f64, device=dev)
        colL[:, :, :old] = t['colL'][:, :, :old]
        colL[:, :, new] = t['colL'][:, :, old]
        t['colL'] = colL
        self.Vcap = new
        self.st.Vcap = new
        for name in ('nv', 'view_id', 'view_alpha', 'nclu', 'rows_prog', 'live', 'cid', 'nk', 'view_grid', 'zr', 'order',
                     'seq', 'moved', 'movedflag', 'colL'):
            setattr(self.st, name, t[name].data_ptr())
        self._view_cost = {}
        self.vcap_grown = getattr(self, 'vcap_grown', 0) + 1
