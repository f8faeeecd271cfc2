"""Host orchestration of S CrossCat chains on one GPU.

`Chains` owns a DeviceChains (device memory + C ABI) and the small host-side
mirrors the reference keeps implicitly in its object graph: each chain's
numpy RandomState, the insertion order of its views (State.views is an
OrderedDict, src/crosscat/state.py:141-159), the order of hyper names, and the
diagnostics counters (state.py:907-914).  It runs the kernel scheduler of
State.transition / _transition_generic (state.py:727-761,866-905) for all
selected chains in lock-step, one CUDA launch per CrossCat kernel.

Two random-stream modes:
  stream='numpy'   every draw comes from the chain's RandomState in exactly the
                   order the reference consumes it and is
                   injected into the kernels; chains then follow the reference's
                   trajectories for the same seeds.  The column kernel needs a
                   host round-trip per column in this mode (its stream use is
                   data dependent, src/primitives/crp.py:141).
  stream='philox'  counter-based device generator; nothing crosses PCIe inside
                   a sweep.  Same kernels, same conditional distributions.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from . import codec
from . import hostmath as hm
from .device import DeviceChains

# lognormal (src/primitives/lognormal.py) is a normal column on log x: the device sees the transformed column, the
# host keeps the original values, the column's own hyper grids, the -log x Jacobian of densities and scores,
# and the reference's order of the hyperparameter dict (m, r, nu, s: lognormal.py:151-157).
CCTYPES = {'normal': L.NORMAL, 'categorical': L.CATEGORICAL, 'lognormal': L.NORMAL, 'bernoulli': L.BERNOULLI,
           'poisson': L.POISSON, 'exponential': L.EXPONENTIAL, 'geometric': L.GEOMETRIC}
LOGNORMAL_HYPERS = ['m', 'r', 'nu', 's']


def lognormal_grids(xcol, n_grid=L.NGRID):
    """Lognormal.construct_hyper_grids (lognormal.py:151-157) of the column's positive, non-NaN values, laid out in
    the device's canonical hyper order (m, r, s, nu)."""
    clean = [float(x) for x in xcol if not math.isnan(x)]
    g = {'m': hm.log_linspace(1e-4, max(clean), n_grid), 'r': hm.log_linspace(.1, float(len(clean)), n_grid),
         'nu': hm.log_linspace(.1, float(len(clean)), n_grid), 's': hm.log_linspace(.1, float(len(clean)), n_grid)}
    return np.stack([g[n] for n in codec.HYPER_NAMES[L.NORMAL]])
KERNELS = ['alpha', 'view_alphas', 'column_params', 'column_hypers', 'rows', 'columns']


def philox4x32_10(ctr_lo, ctr_hi, seed):
    """numpy mirror of the device generator (csrc/cc_common.cuh: Philox): returns the first
    output word for arrays of 64-bit counters.  Used to order work by its expected cost."""
    ctr_lo = np.asarray(ctr_lo, dtype=np.uint64); ctr_hi = np.asarray(ctr_hi, dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    c0, c1 = ctr_lo & m32, ctr_lo >> np.uint64(32)
    c2, c3 = ctr_hi & m32, ctr_hi >> np.uint64(32)
    a = np.uint64(int(seed) & 0xFFFFFFFF); b = np.uint64((int(seed) >> 32) & 0xFFFFFFFF)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & m32
        hi1, lo1 = p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ a) & m32, lo1, (hi0 ^ c3 ^ b) & m32, lo0
        a = (a + np.uint64(0x9E3779B9)) & m32; b = (b + np.uint64(0xBB67AE85)) & m32
    return c0


def aux_alpha_index(seed, counter, chains, cols):
    """Grid index the device will draw for the auxiliary view of each (chain, column)
    (columns.cu: stream 4): ((x * 30) >> 32)."""
    sid = (np.uint64(4) << np.uint64(56)) | ((np.asarray(chains, dtype=np.uint64) & np.uint64(0xFFFFFF)) << np.uint64(32)) \
        | np.asarray(cols, dtype=np.uint64)
    x = philox4x32_10(np.full(len(sid), counter, dtype=np.uint64), sid, seed)
    return ((x * np.uint64(L.NGRID)) >> np.uint64(32)).astype(np.int64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Chains(object):
    def __init__(self, X, outputs=None, cctypes=None, distargs=None, Cd=None, Ci=None,
                 S=1, stream='philox', device=None, Vcap=None, Kcap=None, seed=0, grids=None,
                 row_alpha_grid=None, col_alpha_grid=None):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError('X must be a 2-D table.')
        R, Cn = X.shape
        self.outputs = list(outputs) if outputs else list(range(Cn))
        if len(self.outputs) != Cn or any(o < 0 for o in self.outputs):
            raise ValueError('outputs must be %d non-negative ids' % Cn)
        cctypes = list(cctypes) if cctypes else ['normal'] * Cn
        distargs = [dict(d) if d else {} for d in (distargs or [None] * Cn)]
        bad = [c for c in cctypes if c not in CCTYPES]
        if bad:                                           # same rule as state.py:817-820
            raise ValueError('Only normal, lognormal, categorical, bernoulli, poisson, exponential and geometric cgpms '
                             'supported by cgpm_b200: %s' % (cctypes,))
        for ct, d in zip(cctypes, distargs):
            if ct == 'categorical' and 'k' not in d:
                raise ValueError('Categorical requires distarg `k`.')
        self.cctypes, self.distargs = cctypes, distargs
        self.ctype = np.array([CCTYPES[c] for c in cctypes], dtype=np.int32)
        self.ncat = np.array([int(d['k']) if ct == 'categorical' else 0
                              for ct, d in zip(cctypes, distargs)], dtype=np.int32)
        self.Cd = list(Cd) if Cd else []
        self.Ci = list(Ci) if Ci else []
        if stream not in ('numpy', 'philox'):
            raise ValueError("stream must be 'numpy' or 'philox'")
        self.stream = stream
        self.col_index = {c: i for i, c in enumerate(self.outputs)}
        if Vcap is None:
            # a chain can hold up to C views (state.py:1125-1154); give every chain C view slots when the per-slot
            # arrays (17 bytes per row) stay within a quarter of the device memory, otherwise what fits (>= 16)
            per_slot = S * (17 * R + 64)
            try:
                free, _total = torch.cuda.mem_get_info(device) if device is not None else torch.cuda.mem_get_info()
            except Exception:
                free = 16 << 30
            Vcap = int(min(Cn, max(16, (free // 4) // max(per_slot, 1))))
        # lognormal columns: log x on the device (math.log, as the reference computes it: lognormal.py:60-63)
        self.lognormal = np.array([c == 'lognormal' for c in cctypes], dtype=bool)
        self.X_raw = X
        self._ln_const = 0.0
        if self.lognormal.any():
            X = X.copy()
            grids = list(grids) if grids is not None else [None] * Cn
            mlog = np.frompyfunc(math.log, 1, 1)
            for i in np.nonzero(self.lognormal)[0]:
                col = self.X_raw[:, i]
                ok = ~np.isnan(col)
                if np.any(col[ok] <= 0):
                    raise ValueError('Invalid Lognormal: %s' % (col[ok][col[ok] <= 0][0],))
                X[ok, i] = mlog(col[ok]).astype(np.float64)
                if grids[i] is None:
                    grids[i] = lognormal_grids(col)
                self._ln_const -= float(np.sum(X[ok, i]))          # sum over rows of -log x (lognormal.py:98-102)
        self.dev = DeviceChains(X, self.ctype, self.ncat, S, Vcap=Vcap, Kcap=Kcap, device=device, grids=grids,
                                row_alpha_grid=row_alpha_grid, col_alpha_grid=col_alpha_grid)
        self.R, self.C, self.S = R, Cn, S
        self.rngs = [None] * S
        self.views_order = [[] for _ in range(S)]
        self.hyper_order = [[self._default_hyper_order(i) for i in range(Cn)] for _ in range(S)]
        self.diagnostics = [self._fresh_diag() for _ in range(S)]
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.counter = 1
        self._aux_pre = None          # auxiliary views launched beside a rows sweep (see _prelaunch_aux)
        self._defer_rows = False
        self._aux_stream = None
        self._ci_dev = None
        self._mirror = None
        self.events = None          # bench instrumentation: list of (kernel name, start event, end event)
        if self.Ci:
            self._build_ci()

    # ------------------------------------------------------------------ set-up
    def _default_hyper_order(self, i):
        """Key order of the column's hyperparameter dict in the reference (= the order of its grid dict, in which
        a fresh Dim draws them: dim.py:180-183)."""
        return list(LOGNORMAL_HYPERS) if self.lognormal[i] else list(codec.HYPER_NAMES[int(self.ctype[i])])

    @staticmethod
    def _fresh_diag():
        return {'iterations': {}, 'logscore': [], 'column_crp_alpha': [], 'column_partition': []}

    def _build_ci(self):
        adj = [[] for _ in range(self.C)]
        for pair in self.Ci:
            for a in pair:
                for b in pair:
                    if a != b:
                        adj[self.col_index[a]].append(self.col_index[b])
        off = np.zeros(self.C + 1, np.int32)
        off[1:] = np.cumsum([len(a) for a in adj])
        flat = np.array([x for a in adj for x in a] or [0], np.int32)
        self._ci_dev = (torch.from_numpy(off).to(self.dev.device), torch.from_numpy(flat).to(self.dev.device))

    def init_chain(self, s, rng, Zv=None, Zrv=None, alpha=None, view_alphas=None, hypers=None,
                   diagnostics=None):
        """State.__init__ (src/crosscat/state.py:47-182) for chain s, consuming
        `rng` in the reference's order: column-CRP alpha, column partition, then
        per view (in `set(Zv.values())` order) its alpha, row partition and the
        hypers of its columns."""
        dev, R, Cn = self.dev, self.R, self.C
        exact = self.stream == 'numpy'
        sim = hm.crp_simulate_exact if exact else hm.crp_simulate_fast
        self.rngs[s] = rng
        if alpha is None:
            alpha = rng.choice(dev.col_alpha_grid_h)
        if Zv is None:
            if self.Ci or self.Cd:                                # state.py:96-113
                if self.outputs != list(range(Cn)):
                    raise ValueError('Use zero-based outputs with constraints.')
                if self.Ci:
                    zv_ids = hm.simulate_crp_constrained(Cn, alpha, self.Cd, self.Ci, rng)
                else:
                    zv_ids = hm.simulate_crp_constrained_dependent(Cn, alpha, self.Cd, rng)
            else:
                zv_ids = sim(Cn, alpha, rng)
            Zv_od = OrderedDict((c, int(z)) for c, z in zip(self.outputs, zv_ids))
        else:
            Zv_od = OrderedDict((c, int(z)) for c, z in dict(Zv).items())
            if set(Zv_od.keys()) != set(self.outputs):
                raise ValueError('Zv must assign every output.')
            if (self.Cd or self.Ci) and not hm.validate_crp_constrained_partition(Zv_od, self.Cd, self.Ci):
                raise ValueError('Zv violates the dependence / independence constraints.')
        view_alphas = dict(view_alphas) if view_alphas else {}
        Zrv = dict(Zrv) if Zrv else {}
        if Zrv and set(Zrv.keys()) != set(Zv_od.values()):
            raise ValueError('Zrv keys must match the views of Zv.')
        hyp = np.zeros((Cn, 4))
        names_all = [None] * Cn
        views = OrderedDict()
        for v in set(Zv_od.values()):
            va = view_alphas.get(v)
            if va is None:
                va = rng.choice(dev.row_alpha_grid_h)
            zr = Zrv.get(v)
            zr = sim(R, va, rng) if zr is None else np.asarray(zr, dtype=np.int64)
            if zr.shape != (R,):
                raise ValueError('Zrv[%s] must have one entry per row.' % (v,))
            views[v] = {'alpha': float(va), 'Zr': zr, 'order': None}
            for c in [o for o in self.outputs if Zv_od[o] == v]:
                i = self.col_index[c]
                names = list(codec.HYPER_NAMES[int(self.ctype[i])])
                given = hypers[i] if hypers else None
                if given:
                    names_all[i] = list(given.keys())
                    for n in names:
                        hyp[i, names.index(n)] = float(given[n])
                else:
                    names_all[i] = self._default_hyper_order(i)
                    for n in names_all[i]:                      # dim.py:180-183
                        g = dev.grids_h[i][names.index(n)]
                        hyp[i, names.index(n)] = rng.choice(g[~np.isnan(g)])     # (a grid shorter than 30 is NaN-padded)
        spec = {'Zv': [Zv_od[c] for c in self.outputs], 'views': views, 'alpha': float(alpha), 'hyp': hyp}
        dev.upload_chain(s, codec.pack_chain(spec, R, Cn))
        self.views_order[s] = list(range(len(views)))
        self.hyper_order[s] = names_all
        self.diagnostics[s] = self._fresh_diag() if diagnostics is None else self._load_diag(diagnostics)
        self._mirror = None

    @staticmethod
    def _load_diag(d):
        out = Chains._fresh_diag()
        for k, v in dict(d).items():
            out[k] = v
        out['iterations'] = dict(out.get('iterations', {}))
        return out

    def finish_init(self, chains=None):
        if chains is None:
            self.dev.rebuild()
        else:
            nv = self.dev.t['nv'].cpu().numpy()
            self.dev.rebuild([(s, int(v)) for s in chains for v in np.nonzero(nv[s] > 0)[0]])
        self._mirror = None

    # ------------------------------------------------------------------ mirror
    def mirror(self):
        """Small per-chain structure read back from the device: zv, nv, view_id."""
        if self._mirror is None:
            t = self.dev.t
            self._mirror = {'zv': t['zv'].cpu().numpy(), 'nv': t['nv'].cpu().numpy(),
                            'view_id': t['view_id'].cpu().numpy()}
        return self._mirror

    def _refresh_views_order(self, chains):
        m = self.mirror()
        for s in chains:
            live = [int(v) for v in np.nonzero(m['nv'][s] > 0)[0]]
            keep = [v for v in self.views_order[s] if v in live]
            new = sorted([v for v in live if v not in keep], key=lambda v: m['view_id'][s][v])
            self.views_order[s] = keep + new

    # ------------------------------------------------------------- transitions
    def transition(self, N=None, S=None, kernels=None, rowids=None, cols=None, views=None,
                   progress=True, checkpoint=None, chains=None):
        """State.transition (state.py:727-761) + _transition_generic (:866-905) for
        the selected chains in lock-step."""
        chains = list(range(self.S)) if chains is None else list(chains)
        if cols and any(c not in self.col_index for c in cols):
            raise ValueError('Only CrossCat columns may be transitioned.')
        if kernels is None:
            kernels = list(KERNELS)
        for k in kernels:
            if k not in KERNELS:
                raise KeyError(k)
        if 'columns' in kernels and self.Cd:
            raise ValueError('Cannot transition columns with dependence constraint, '
                             'use State.transition_lovecat.')
        table = {
            'alpha': lambda: self.transition_crp_alpha(chains),
            'view_alphas': lambda: self.transition_view_alphas(chains, views=views, cols=cols),
            'column_params': lambda: self._count(chains, 'column_params'),
            'column_hypers': lambda: self.transition_dim_hypers(chains, cols=cols),
            'rows': lambda: self.transition_view_rows(chains, views=views, cols=cols, rows=rowids),
            'columns': lambda: self.transition_dims(chains, cols=cols),
        }
        if (self.stream == 'philox' and rowids is None and cols is None and views is None and S is None and N
                and self.events is None and not self.check_env_debug() and not os.environ.get('CGPM_B200_PYTHON_SCHEDULE')):
            return self._transition_native(chains, int(N), kernels, progress, checkpoint)
        funcs = [self._timed(k, table[k]) for k in kernels]
        if self.stream == 'philox' and not os.environ.get('CGPM_B200_NO_AUX_OVERLAP'):
            # rows immediately followed by columns: generate the column sweep's auxiliary views beside the rows sweep
            for j, k in enumerate(kernels):
                if k == 'rows' and j + 1 < len(kernels) and kernels[j + 1] == 'columns':
                    funcs[j] = self._rows_then_aux(funcs[j], chains, cols)
        if N is None and S is None:
            N = 1
        if progress is None:
            progress = True
        iters, start = 0, time.time()

        def done():
            if S is not None:
                self.dev.sync()         # launches are asynchronous: the time budget is checked against finished work
            p_sec = 0 if S is None else (time.time() - start) / S
            p_it = 0 if N is None else float(iters) / N
            return max(p_it, p_sec)
        while funcs:
            broke = False
            for f in funcs:
                p = done()
                if progress:
                    self._progress(p)
                if p >= 1.:
                    broke = True
                    break
                f()
            if broke:
                break
            iters += 1
            if checkpoint and iters % checkpoint == 0:
                self._increment_diagnostics(chains)
        self.dev.sync()
        self._aux_pre = None           # auxiliary views generated ahead are only valid within this call
        self.dev.check_errors()
        if progress:
            print('\rCompleted: %d iterations in %f seconds.' % (iters, time.time() - start))
        return iters

    # kernels the C schedule launches per kernel code and iteration (a lower bound, for the launch counter)
    NATIVE_LAUNCHES = {'alpha': 1, 'view_alphas': 1, 'column_hypers': 1, 'rows': 1, 'columns': 7, 'column_params': 0}

    def _transition_native(self, chains, N, kernels, progress, checkpoint):
        """The whole schedule through cc_transition (csrc/transition.cu): N iterations of the kernel list in one C call
        per checkpoint interval -- no Python between kernels.  Same kernels, same conditional distributions as the
        Python schedule below (which stays for the injected numpy stream, row / column / view subsets, time budgets,
        per-kernel timing and the debug invariants)."""
        dev = self.dev
        start = time.time()
        codes = np.ascontiguousarray([L.KERNEL_CODES[k] for k in kernels if k != 'column_params'], dtype=np.int32)
        ch = np.ascontiguousarray(chains, dtype=np.int32)
        if 'columns' in kernels:
            self._ensure_arena(len(chains) * self.C)
            self._ensure_aux_store()
        ci_off, ci_adj = (self._ci_dev if self._ci_dev else (None, None))
        self._aux_pre = None
        done = 0
        while done < N:
            n = min(N - done, checkpoint) if checkpoint else N - done
            ctr = C.c_uint64(self.counter)
            rc = dev.lib.cc_transition(C.byref(dev.st), len(ch), _p(ch), n, len(codes), _p(codes), _dp(ci_off), _dp(ci_adj),
                                       C.c_uint64(self.seed), C.byref(ctr), dev.grow_callback(), None, dev.stream_ptr())
            self.counter = int(ctr.value)
            self._mirror = None
            L.check(rc, 'cc_transition')
            dev.lib_calls['cc_transition'] += n * sum(self.NATIVE_LAUNCHES[k] for k in kernels)
            for k in kernels:
                for s in chains:
                    it = self.diagnostics[s]['iterations']
                    it[k] = it.get(k, 0) + n
            done += n
            if checkpoint and done % checkpoint == 0:
                self._refresh_views_order(chains)
                self._increment_diagnostics(chains)
        dev.sync()
        self._refresh_views_order(chains)
        dev.check_errors()
        if progress:
            print('\rCompleted: %d iterations in %f seconds.' % (N, time.time() - start))
        return N

    CHECK_CODES = {1: 'nv is not the histogram of zv', 2: 'live view without id', 3: 'column in no view',
                   4: 'non-positive CRP alpha', 5: 'cluster count out of range', 6: 'invalid or repeated live slot',
                   7: 'live cluster without members', 8: 'cluster ids not ascending', 9: 'cluster sizes do not sum to R',
                   10: 'row in a dead cluster', 11: 'cluster size is not the histogram of zr',
                   12: 'member order is not a permutation', 13: 'scratch flags not clear',
                   14: 'N of a (cluster, column) differs from its non-NaN members'}

    def check_partitions(self):
        """Device form of _check_partitions (view.py:528-560, state.py:1162-1184)."""
        out = torch.zeros((self.S,), dtype=torch.int32, device=self.dev.device)
        L.check(self.dev.lib.cc_check_state(C.byref(self.dev.st), _dp(out), self.dev.stream_ptr()), 'cc_check_state')
        codes = out.cpu().numpy()
        bad = np.nonzero(codes)[0]
        if len(bad):
            raise AssertionError('partition invariants violated: ' + ', '.join(
                'chain %d: %s' % (s, self.CHECK_CODES.get(int(codes[s]), codes[s])) for s in bad[:8]))

    @staticmethod
    def check_env_debug():
        """src/utils/config.py:112-114."""
        import os
        return 'GPMCCDEBUG' in os.environ

    def _timed(self, name, f):
        def run():
            if self.check_env_debug():
                f()
                return self.check_partitions()
            if self.events is None:
                return f()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            self.events.append((name, e0, e1))
        return run

    @staticmethod
    def _progress(p):
        width = 30
        k = int(min(max(p, 0), 1) * width)
        sys.stdout.write('\r[%s%s] %.2f%%' % ('#' * k, '-' * (width - k), 100 * min(p, 1)))
        sys.stdout.flush()

    def _count(self, chains, name):
        for s in chains:
            it = self.diagnostics[s]['iterations']
            it[name] = it.get(name, 0) + 1

    def _increment_diagnostics(self, chains):
        """state.py:911-914."""
        scores = self.logpdf_score()
        alphas = self.dev.t['col_alpha'].cpu().numpy()
        for s in chains:
            d = self.diagnostics[s]
            d['logscore'].append(float(scores[s]))
            d['column_crp_alpha'].append(float(alphas[s]))
            d['column_partition'].append(self.Zv_items(s))

    def _next_counter(self):
        self.counter += 1
        return self.counter

    # -- alpha ---------------------------------------------------------------
    def transition_crp_alpha(self, chains):
        ch = np.ascontiguousarray(chains, dtype=np.int32)
        u = None
        if self.stream == 'numpy':
            u = np.ascontiguousarray([self.rngs[s].random_sample() for s in chains])
        L.check(self.dev.lib.cc_transition_alphas(C.byref(self.dev.st), len(ch), _p(ch), None, _p(u), 1,
                                                  C.c_uint64(self.seed), C.c_uint64(self._next_counter()),
                                                  self.dev.stream_ptr()), 'cc_transition_alphas')
        self._count(chains, 'alpha')

    def _views_for(self, s, views, cols):
        """Selection rule of state.py:768-769 / 798-799 (view slots)."""
        m = self.mirror()
        if views is not None:
            ids = list(views)
        elif cols:
            ids = list(set(int(m['view_id'][s][m['zv'][s][self.col_index[c]]]) for c in cols))
        else:
            return list(self.views_order[s])
        slot_of = {int(m['view_id'][s][v]): v for v in range(self.dev.Vcap) if m['nv'][s][v] > 0}
        return [slot_of[i] for i in ids]

    def transition_view_alphas(self, chains, views=None, cols=None):
        ch, vw, us = [], [], []
        for s in chains:
            for v in self._views_for(s, views, cols):
                ch.append(s)
                vw.append(v)
                if self.stream == 'numpy':
                    us.append([self.rngs[s].random_sample(), self.rngs[s].random_sample()])
        ch = np.ascontiguousarray(ch, dtype=np.int32)
        vw = np.ascontiguousarray(vw, dtype=np.int32)
        u = np.ascontiguousarray(us, dtype=np.float64) if self.stream == 'numpy' else None
        L.check(self.dev.lib.cc_transition_alphas(C.byref(self.dev.st), len(ch), _p(ch), _p(vw), _p(u), 2,
                                                  C.c_uint64(self.seed), C.c_uint64(self._next_counter()),
                                                  self.dev.stream_ptr()), 'cc_transition_alphas')
        self._count(chains, 'view_alphas')

    # -- column hypers -------------------------------------------------------
    def transition_dim_hypers(self, chains, cols=None):
        cidx = [self.col_index[c] for c in cols] if cols is not None else list(range(self.C))
        ch, co, hord, us = [], [], [], []
        for s in chains:
            for i in cidx:
                ch.append(s)
                co.append(i)
                if self.stream == 'numpy':
                    names = list(self.hyper_order[s][i])
                    self.rngs[s].shuffle(names)                           # dim.py:154-155
                    canon = codec.HYPER_NAMES[int(self.ctype[i])]
                    o = [canon.index(n) for n in names]
                    hord.append(o + [-1] * (4 - len(o)))
                    uu = [self.rngs[s].random_sample() for _ in names]
                    us.append(uu + [0.0] * (4 - len(uu)))
        ch = np.ascontiguousarray(ch, dtype=np.int32)
        co = np.ascontiguousarray(co, dtype=np.int32)
        ho = np.ascontiguousarray(hord, dtype=np.int32) if self.stream == 'numpy' else None
        u = np.ascontiguousarray(us, dtype=np.float64) if self.stream == 'numpy' else None
        L.check(self.dev.lib.cc_transition_col_hypers(C.byref(self.dev.st), len(ch), _p(ch), _p(co), _p(ho), _p(u),
                                                      C.c_uint64(self.seed), C.c_uint64(self._next_counter()),
                                                      self.dev.stream_ptr()), 'cc_transition_col_hypers')
        self._count(chains, 'column_hypers')

    # -- rows ------------------------------------------------------------------
    def transition_view_rows(self, chains, views=None, cols=None, rows=None, trace=None):
        """State.transition_view_rows (state.py:795-802) -> View.transition_rows."""
        if self.R == 1:
            return
        R, dev = self.R, self.dev
        n = R if rows is None else len(rows)
        work, perms, us = [], [], []
        off = 0
        m = self.mirror()
        is_normal = self.ctype == CCTYPES['normal']
        for s in chains:
            for v in self._views_for(s, views, cols):
                w = dict(chain=s, view=v, n=n, perm_is_index=1 if rows is None else 0,
                         n_cols=int(m['nv'][s][v]), n_normal=int(np.sum(is_normal & (m['zv'][s] == v))))
                if self.stream == 'numpy':
                    rng = self.rngs[s]
                    perms.append(np.asarray(rng.permutation(R if rows is None else rows), dtype=np.int32))
                    us.append(rng.random_sample(n))
                    w['perm_off'] = off
                    w['u_off'] = off
                    off += n
                elif rows is not None:
                    perms.append(np.asarray(np.random.RandomState((self.seed + self.counter + 7919 * s + v) % (2 ** 32))
                                            .permutation(rows), dtype=np.int32))
                    w['perm_off'] = off
                    off += n
                if trace is not None:
                    w['trace_off'] = len(work) * n * trace['stride']
                work.append(w)
        perm = torch.from_numpy(np.concatenate(perms)).to(dev.device) if perms else None
        u = torch.from_numpy(np.concatenate(us)).to(dev.device) if us else None
        tr = None
        if trace is not None:
            tr = torch.zeros(max(len(work) * n * trace['stride'], 1), dtype=torch.float64, device=dev.device)
        dev.transition_rows(work, perm=perm, u=u, trace=tr, trace_stride=trace['stride'] if trace else 0,
                            seed=self.seed, counter=self._next_counter(), defer=self._defer_rows and trace is None)
        if trace is not None:
            dev.sync()
            trace['work'] = work
            trace['data'] = tr.cpu().numpy()[:len(work) * n * trace['stride']].reshape(len(work), n, trace['stride'])
        self._count(chains, 'rows')

    # -- columns ---------------------------------------------------------------
    def _ensure_arena(self, n_pairs):
        dev = self.dev
        R = self.R
        n1 = (R + 31) >> 5
        n2 = (n1 + 31) >> 5
        per = (max(R + (R + n1 + n2 + 32) + (R + 1) + R + R, ((R + 7) & ~7) + 8 * R) + 7) & ~7      # columns.cu: cc_col_aux
        want = min(max(n_pairs, 1), 148 * 32)
        free, _total = torch.cuda.mem_get_info(dev.device)
        budget = max(int(free * 0.5), per * 4)
        workers = max(1, min(want, budget // (per * 4)))
        need = workers * per
        if dev.arena is None or dev.arena.numel() < need:
            dev.arena = None
            dev.arena = torch.empty(need, dtype=torch.int32, device=dev.device)
            dev.st.arena = dev.arena.data_ptr()
            dev.st.arena_bytes = need * 4

    def _ensure_aux_store(self):
        """Room for every (chain, column) auxiliary partition, when it costs less than a
        quarter of the free device memory; otherwise a birth regenerates its partition."""
        dev = self.dev
        if dev.aux_zr_all is not None:
            return True
        need = self.S * self.C * self.R * 4
        free, _total = torch.cuda.mem_get_info(dev.device)
        if need > free // 4:
            return False
        dev.aux_zr_all = torch.empty(self.S * self.C * self.R, dtype=torch.int32, device=dev.device)
        dev.st.aux_zr_all = dev.aux_zr_all.data_ptr()
        return True

    def _rows_then_aux(self, rows_func, chains, cols):
        def run():
            if self.check_env_debug():
                return rows_func()
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev.device))
            self._defer_rows = True            # the capacity check of the sweep (a host read) waits until the
            try:                               # auxiliary views are enqueued beside it
                rows_func()
                self._prelaunch_aux(chains, cols, ev)
            finally:
                self._defer_rows = False
                self.dev.finish_rows()
        return run

    def _column_order(self, chains, cidx, counter):
        """Column visiting order of every chain: rng.permutation(cols) (state.py:807)."""
        perm = np.zeros((len(chains), len(cidx)), np.int32)
        for i, s in enumerate(chains):
            if self.stream == 'numpy':
                perm[i] = np.asarray(self.rngs[s].permutation(cidx), dtype=np.int32)
            else:
                perm[i] = np.asarray(np.random.RandomState((self.seed * 31 + counter * 1009 + s) % (2 ** 32))
                                     .permutation(cidx), dtype=np.int32)
        return perm

    def _launch_aux(self, ch_h, perm, counter, stored, stream_ptr):
        """cc_col_aux for every (chain, column) of a device-stream column sweep."""
        a_ch = np.repeat(ch_h, perm.shape[1]).astype(np.int32)
        a_co = perm.reshape(-1).astype(np.int32)
        # the cost of a pair grows with its alpha (number of tables): hand out the expensive ones first
        cost = np.argsort(-aux_alpha_index(self.seed, counter, a_ch, a_co), kind='stable')
        a_ch = np.ascontiguousarray(a_ch[cost]); a_co = np.ascontiguousarray(a_co[cost])
        L.check(self.dev.lib.cc_col_aux(C.byref(self.dev.st), len(a_ch), _p(a_ch), _p(a_co), None, None, 2 if stored else 0,
                                        C.c_uint64(self.seed), C.c_uint64(counter), stream_ptr), 'cc_col_aux')

    def _prelaunch_aux(self, chains, cols, ev_before_rows):
        """The auxiliary views of the NEXT column sweep depend on the data, the hyperparameters and their own
        random stream only -- not on the row partitions -- so on the device stream they are generated on a side
        stream while the rows kernels (latency-bound, most of the GPU idle) are still running.  transition_dims
        picks the result up if it is called next with the same chains and columns."""
        dev = self.dev
        cidx = [self.col_index[c] for c in cols] if cols is not None else list(range(self.C))
        if not cidx or not chains:
            return
        counter = self.counter + 1                  # the counter transition_dims will draw
        self._ensure_arena(len(chains) * len(cidx))
        stored = self._ensure_aux_store()
        perm = self._column_order(chains, cidx, counter)
        if self._aux_stream is None:
            self._aux_stream = torch.cuda.Stream(device=dev.device)
        side = self._aux_stream
        side.wait_event(ev_before_rows)
        # allocated on the main stream, used on the side stream: tell the caching allocator
        dev.arena.record_stream(side)
        if dev.aux_zr_all is not None:
            dev.aux_zr_all.record_stream(side)
        with torch.cuda.stream(side):
            self._launch_aux(np.ascontiguousarray(chains, dtype=np.int32), perm, counter, stored, dev.stream_ptr())
            done = torch.cuda.Event()
            done.record(side)
        self._aux_pre = {'counter': counter, 'key': (tuple(chains), tuple(cidx)), 'perm': perm, 'done': done}

    def transition_dims(self, chains, cols=None, trace=None):
        """State.transition_dims (state.py:804-810) for the selected chains."""
        dev, lib, st = self.dev, self.dev.lib, C.byref(self.dev.st)
        sp = dev.stream_ptr
        cidx = [self.col_index[c] for c in cols] if cols is not None else list(range(self.C))
        ncols = len(cidx)
        nch = len(chains)
        if ncols == 0 or nch == 0:
            return
        counter = self._next_counter()
        seed = C.c_uint64(self.seed)
        ci_off, ci_adj = (self._ci_dev if self._ci_dev else (None, None))
        # column visiting order: rng.permutation(cols) (state.py:807)
        pre = self._aux_pre
        self._aux_pre = None
        if pre is not None and (pre['counter'] != counter or pre['key'] != (tuple(chains), tuple(cidx))):
            pre['done'].synchronize()          # a prelaunched auxiliary pass nobody will use: let it finish
            pre = None
        if pre is not None:
            perm = pre['perm']
        else:
            perm = self._column_order(chains, cidx, counter)
        ch_h = np.ascontiguousarray(chains, dtype=np.int32)
        ch_d = torch.from_numpy(ch_h).to(dev.device)
        ts = trace['stride'] if trace is not None else 0
        tr_d = torch.zeros(max(nch * ncols * ts, 1), dtype=torch.float64, device=dev.device) if trace is not None else None
        self._mirror = None
        if self.stream == 'philox':
            if pre is not None:
                pre['done'].synchronize()      # the prelaunched pass may still be using the arena: never re-size under it
            self._ensure_arena(nch * ncols)
            stored = self._ensure_aux_store()
            cols_d = torch.from_numpy(perm).to(dev.device)
            pos_d = torch.zeros(nch, dtype=torch.int32, device=dev.device)
            pend_d = torch.full((nch,), -1, dtype=torch.int32, device=dev.device)
            # auxiliary views of every (chain, column) and marginals under every live view
            if pre is not None:
                torch.cuda.current_stream(dev.device).wait_event(pre['done'])     # launched beside the rows sweep
            else:
                self._launch_aux(ch_h, perm, counter, stored, sp())
            m = self.mirror()
            w_ch = np.ascontiguousarray([s for s in chains for v in np.nonzero(m['nv'][s] > 0)[0]], dtype=np.int32)
            w_vw = np.ascontiguousarray([v for s in chains for v in np.nonzero(m['nv'][s] > 0)[0]], dtype=np.int32)
            if cols is None:
                L.check(lib.cc_col_marginals(st, len(w_ch), _p(w_ch), _p(w_vw), None, None, None, self.C, sp()), 'cc_col_marginals')
            else:
                row_of = {s: i for i, s in enumerate(chains)}
                off = np.ascontiguousarray([row_of[s] * ncols for s in w_ch], dtype=np.int32)
                cnt = np.full(len(w_ch), ncols, np.int32)
                L.check(lib.cc_col_marginals(st, len(w_ch), _p(w_ch), _p(w_vw), _dp(cols_d), _p(off), _p(cnt), ncols, sp()), 'cc_col_marginals')
            self._mirror = None
            rounds = 0
            while True:
                rounds += 1
                L.check(lib.cc_col_scan(st, nch, _dp(ch_d), _dp(cols_d), ncols, _dp(pos_d), _dp(pend_d), None,
                                        seed, C.c_uint64(counter), _dp(ci_off), _dp(ci_adj), _dp(tr_d), ts, sp()), 'cc_col_scan')
                pend = pend_d.cpu().numpy()
                idx = np.nonzero(pend >= 0)[0]
                if len(idx) == 0:
                    self.last_col_rounds = rounds
                    break
                b_ch = np.ascontiguousarray(ch_h[idx], dtype=np.int32)
                b_co = np.ascontiguousarray(perm[idx, pend[idx]], dtype=np.int32)
                if not stored:
                    # regenerate the same auxiliary partitions (same key) and keep them
                    L.check(lib.cc_col_aux(st, len(b_ch), _p(b_ch), _p(b_co), None, None, 1, seed, C.c_uint64(counter), sp()), 'cc_col_aux')
                nv_before = self._make_room_for_births(b_ch, b_co, stored)
                st = C.byref(dev.st)
                L.check(lib.cc_col_install(st, len(b_ch), _p(b_ch), _p(b_co), 2 if stored else 1, sp()), 'cc_col_install')
                new_slots = np.array([int(np.nonzero(nv_before[s] == 0)[0][0]) if np.any(nv_before[s] == 0) else 0
                                      for s in b_ch], dtype=np.int32)
                rem_off = np.ascontiguousarray([i * ncols + pend[i] + 1 for i in idx], dtype=np.int32)
                rem_cnt = np.ascontiguousarray([ncols - pend[i] - 1 for i in idx], dtype=np.int32)
                keep = rem_cnt > 0
                if np.any(keep):
                    L.check(lib.cc_col_marginals(st, int(keep.sum()), _p(np.ascontiguousarray(b_ch[keep])),
                                                 _p(np.ascontiguousarray(new_slots[keep])), _dp(cols_d),
                                                 _p(np.ascontiguousarray(rem_off[keep])), _p(np.ascontiguousarray(rem_cnt[keep])),
                                                 int(rem_cnt.max()), sp()), 'cc_col_marginals')
                pos_d += (pend_d >= 0).to(torch.int32)
                dev.sync()
                dev.check_errors()
        else:
            self._ensure_arena(nch)
            R = self.R
            for j in range(ncols):
                m = self.mirror()
                a_ch, a_co, a_ai, a_u, draws = [], [], [], [], np.zeros((nch, 1))
                for i, s in enumerate(chains):
                    c = int(perm[i, j])
                    own = m['zv'][s][c]
                    singleton = m['nv'][s][own] == 1
                    nviews = int(np.sum(m['nv'][s] > 0))
                    rng = self.rngs[s]
                    if not singleton:
                        a_ch.append(s); a_co.append(c)
                        a_ai.append(int(rng.randint(0, L.NGRID)))          # rng.choice(grid), dim.py:183
                        a_u.append(rng.random_sample(R - 1) if R > 1 else np.zeros(0))
                    if nviews + (0 if singleton else 1) > 1:
                        draws[i, 0] = rng.random_sample()                   # state.py:1106
                a_ch = np.ascontiguousarray(a_ch, dtype=np.int32); a_co = np.ascontiguousarray(a_co, dtype=np.int32)
                a_ai = np.ascontiguousarray(a_ai, dtype=np.int32)
                if len(a_ch):
                    u_d = torch.from_numpy(np.ascontiguousarray(np.concatenate(a_u) if R > 1 else np.zeros(1))).to(dev.device)
                    L.check(lib.cc_col_aux(st, len(a_ch), _p(a_ch), _p(a_co), _p(a_ai), _dp(u_d), 1, seed, C.c_uint64(counter), sp()), 'cc_col_aux')
                w_ch = np.ascontiguousarray([s for s in chains for v in np.nonzero(m['nv'][s] > 0)[0]], dtype=np.int32)
                w_vw = np.ascontiguousarray([v for s in chains for v in np.nonzero(m['nv'][s] > 0)[0]], dtype=np.int32)
                row_of = {s: i for i, s in enumerate(chains)}
                colj_d = torch.from_numpy(np.ascontiguousarray(perm[:, j])).to(dev.device)
                off = np.ascontiguousarray([row_of[s] for s in w_ch], dtype=np.int32)
                cnt = np.ones(len(w_ch), np.int32)
                L.check(lib.cc_col_marginals(st, len(w_ch), _p(w_ch), _p(w_vw), _dp(colj_d), _p(off), _p(cnt), 1, sp()), 'cc_col_marginals')
                pos_d = torch.zeros(nch, dtype=torch.int32, device=dev.device)
                pend_d = torch.full((nch,), -1, dtype=torch.int32, device=dev.device)
                u_draw = torch.from_numpy(draws).to(dev.device)
                trj = None
                if tr_d is not None:
                    trj = torch.zeros(nch * ts, dtype=torch.float64, device=dev.device)
                L.check(lib.cc_col_scan(st, nch, _dp(ch_d), _dp(colj_d), 1, _dp(pos_d), _dp(pend_d), _dp(u_draw),
                                        seed, C.c_uint64(counter), _dp(ci_off), _dp(ci_adj), _dp(trj), ts, sp()), 'cc_col_scan')
                if trj is not None:
                    tr_d.view(nch, ncols, ts)[:, j, :] = trj.view(nch, ts)
                pend = pend_d.cpu().numpy()
                idx = np.nonzero(pend >= 0)[0]
                if len(idx):
                    b_ch = np.ascontiguousarray(ch_h[idx], dtype=np.int32)
                    b_co = np.ascontiguousarray(perm[idx, j], dtype=np.int32)
                    self._make_room_for_births(b_ch, b_co, False)
                    st = C.byref(dev.st)
                    L.check(lib.cc_col_install(st, len(b_ch), _p(b_ch), _p(b_co), 1, sp()), 'cc_col_install')
                self._mirror = None
                self._refresh_views_order(chains)
            dev.sync()
            dev.check_errors()
        # re-incorporate every visited column under its final view (state.py:1110-1121)
        mask = torch.zeros((self.S, self.C), dtype=torch.uint8, device=dev.device)
        mask[ch_d.long().unsqueeze(1), torch.from_numpy(np.ascontiguousarray(cidx, dtype=np.int64)).to(dev.device).unsqueeze(0)] = 1
        self._mirror = None
        m = self.mirror()
        pairs = [(s, int(v)) for s in chains for v in np.nonzero(m['nv'][s] > 0)[0]]
        dev.rebuild(pairs, mask=mask)
        self._refresh_views_order(chains)
        if trace is not None:
            dev.sync()
            trace['perm'] = perm
            trace['data'] = tr_d.cpu().numpy()[:nch * ncols * ts].reshape(nch, ncols, ts)
        self._count(chains, 'columns')

    def _make_room_for_births(self, b_ch, b_co, stored):
        """Before auxiliary views are installed as new views: grow the cluster slots to the largest partition about
        to be installed and the view slots of chains that have none free (a chain can hold up to C views).  Returns
        nv as it is before the install."""
        dev = self.dev
        if stored:
            ak = dev.t['aux_K_all'][torch.from_numpy(b_ch.astype(np.int64)).to(dev.device),
                                    torch.from_numpy(b_co.astype(np.int64)).to(dev.device)]
        else:
            ak = dev.t['aux_K'][torch.from_numpy(b_ch.astype(np.int64)).to(dev.device)]
        kmax = int(ak.max().item())
        if kmax > dev.Kcap - 1:
            dev.grow_kcap(kmax + 1 + max(16, kmax // 8))
        nv_before = dev.t['nv'].cpu().numpy()
        if any(np.all(nv_before[s] > 0) for s in b_ch):
            dev.grow_vcap(max(dev.Vcap + 8, (3 * dev.Vcap) // 2))
            nv_before = dev.t['nv'].cpu().numpy()
        return nv_before

    # ------------------------------------------------------------------ queries
    def logpdf_score(self):
        """State.logpdf_score (state.py:370-387).  With dependence constraints the column-CRP term is the
        constrained one (gu.logp_crp_constrained_dependent); with independence constraints the reference refuses."""
        if self.Ci:
            raise ValueError('Cannot compute crp score with independences.')
        out = self.dev.logpdf_score().cpu().numpy() + self._ln_const
        if self.Cd:
            m = self.mirror()
            alphas = self.dev.t['col_alpha'].cpu().numpy()
            for s in range(self.S):
                nvs = [int(x) for x in m['nv'][s] if x > 0]
                if not nvs:
                    continue
                Z = {c: int(m['zv'][s][i]) for i, c in enumerate(self.outputs)}
                out[s] += hm.logp_crp_constrained_dependent(Z, float(alphas[s]), self.Cd) \
                    - hm.logp_crp(self.C, nvs, float(alphas[s]))
        return out

    def Zv_items(self, s):
        m = self.mirror()
        return [(self.outputs[i], int(m['view_id'][s][m['zv'][s][i]])) for i in range(self.C)]

    def chain_spec(self, s, with_stats=False):
        dl = self.dev.download_chain(s, with_stats=with_stats)
        return codec.unpack_chain(dl, self.views_order[s]), dl

    def logpdf_bulk(self, chains, rowids, targets_list, constraints_list=None):
        from . import query
        return query.logpdf_bulk(self, chains, rowids, targets_list, constraints_list)

    def simulate_bulk(self, chains, rowids, targets_list, constraints_list=None, Ns=None):
        from . import query
        return query.simulate_bulk(self, chains, rowids, targets_list, constraints_list, Ns)

    # ------------------------------------------------- dense host <-> device state
    SMALL = ('hyp', 'zv', 'nv', 'view_id', 'col_alpha', 'view_alpha', 'nclu', 'live', 'cid', 'nk')

    def export_dense(self, pinned=None):
        """Device -> host copy of the chain state in dense slot form (the
        write-back half of the lovecat-style bridge, lovecat.py:226-290).
        Returns pinned CPU tensors; zr/order only for live (chain, view) slots."""
        t = self.dev.t
        out = pinned if pinned is not None else {}
        for k in self.SMALL:
            if k not in out:
                out[k] = torch.empty(t[k].shape, dtype=t[k].dtype).pin_memory()
            out[k].copy_(t[k], non_blocking=True)
        self.dev.sync()
        live = torch.nonzero(out['nv'] > 0)
        idx = (live[:, 0] * self.dev.Vcap + live[:, 1]).to(self.dev.device)
        n = int(idx.numel())
        for k in ('zr', 'order'):
            if k not in out or out[k].shape[0] != n:
                out[k] = torch.empty((n, self.R), dtype=torch.int32).pin_memory()
            out[k].copy_(t[k].view(-1, self.R).index_select(0, idx), non_blocking=True)
        out['live_index'] = live
        self.dev.sync()
        return out

    def import_dense(self, d, X=None):
        """Host -> device copy of a dense chain state (+ optionally the table),
        then the ordered rebuild of all tables: the export half of the bridge
        (lovecat.py:73-223)."""
        t = self.dev.t
        if X is not None:
            t['Xrm'][:, :self.C].copy_(X, non_blocking=True)
            t['Xcm'].copy_(t['Xrm'][:, :self.C].t())
        for k in self.SMALL:
            t[k].copy_(d[k], non_blocking=True)
        live = d['live_index']
        idx = (live[:, 0] * self.dev.Vcap + live[:, 1]).to(self.dev.device)
        for k in ('zr', 'order'):
            t[k].view(-1, self.R).index_copy_(0, idx, d[k].to(self.dev.device, non_blocking=True))
        self._mirror = None
        self.dev.rebuild()
