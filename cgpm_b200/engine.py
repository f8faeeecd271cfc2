"""`Engine` and `State`: the reference's public API for the CrossCat path, GPU-backed.

Mirrors cgpm.crosscat.engine.Engine (src/crosscat/engine.py:61-435) and
cgpm.crosscat.state.State (src/crosscat/state.py:44-1284) for the inference hot
path: construction ("initialize"), transition(N, S, kernels, rowids, cols,
views, progress, checkpoint, statenos), logpdf_score, dependence probability,
logpdf / simulate (bulk), and the metadata / pickle round trip ("serialize").
Same names, argument meaning and error behaviour; the collapsed primitives are accepted as
column types (normal, lognormal, categorical, bernoulli, poisson, exponential, geometric); conditional and foreign
cgpms are not (the rule lovecat applies more narrowly, state.py:817-822).
The reference's process fan-out (utils/parallel_map.py) is replaced by one CUDA
launch over all chains; `multiprocess` is accepted and ignored.

Extra keyword arguments (not in the reference): stream ('philox' | 'numpy',
see core.py), device, Vcap, Kcap.
"""
from __future__ import annotations

import pickle
from collections import OrderedDict

import math

import numpy as np

from . import _lib as L
from . import codec
from . import hostmath as hm
from .core import Chains, KERNELS

STATE_KW = ('outputs', 'inputs', 'cctypes', 'distargs', 'Zv', 'Zrv', 'alpha', 'view_alphas',
            'hypers', 'Cd', 'Ci', 'Rd', 'Ri', 'diagnostics', 'loom_path')
OPT_KW = ('stream', 'device', 'Vcap', 'Kcap')


def gen_rng(seed=None):
    """src/utils/general.py:37-40."""
    if seed is None:
        seed = np.random.randint(low=1, high=2 ** 31)
    return np.random.RandomState(seed)


def _check_state_kwargs(kw):
    if kw.get('inputs'):
        raise ValueError('State does not accept inputs.')
    if kw.get('Rd') is not None:
        raise ValueError('Row dependence constraints not implemented.')
    if kw.get('Ri') is not None:
        raise ValueError('Row independence constraints not implemented.')
    unknown = [k for k in kw if k not in STATE_KW]
    if unknown:
        raise TypeError('unexpected keyword arguments: %s' % (unknown,))


def _chain_metadata(ch, s, X_list=None):
    """State.to_metadata (src/crosscat/state.py:1189-1240) for chain s."""
    spec, dl = ch.chain_spec(s, with_stats=True)
    md = dict()
    if X_list is not None:
        md['X'] = X_list
    md['outputs'] = list(ch.outputs)
    md['alpha'] = float(spec['alpha'])
    md['Zv'] = [(ch.outputs[i], int(v)) for i, v in enumerate(spec['Zv'])]
    md['cctypes'] = list(ch.cctypes)
    hyp = spec['hyp']
    md['hypers'] = []
    for i in range(ch.C):
        canon = codec.HYPER_NAMES[int(ch.ctype[i])]
        md['hypers'].append(OrderedDict((n, float(hyp[i, canon.index(n)])) for n in ch.hyper_order[s][i]))
    md['distargs'] = [dict(d) for d in ch.distargs]
    # suffstats: list per column of (cluster id, stats) + the empty aux model at max+1 (dim.py:212-217)
    md['suffstats'] = []
    view_of_col = {i: v for i, v in enumerate(spec['Zv'])}
    for i in range(ch.C):
        vw = spec['views'][view_of_col[i]]
        items = []
        for pos, k in enumerate(vw['live']):
            cid = int(vw['cid'][pos])
            N = float(dl['stat'][L.F_N, k, i])
            if ch.ctype[i] != L.CATEGORICAL:
                sx, sxx = (('sum_log_x', 'sum_log_x_sq') if ch.lognormal[i] else codec.SUFFSTAT_NAMES[int(ch.ctype[i])])
                d = {'N': N, sx: float(dl['stat'][L.F_SX, k, i])}
                if sxx is not None:
                    d[sxx] = float(dl['stat'][L.F_SXX, k, i])
                items.append((cid, d))
            else:
                o0 = ch.dev.catoff_h[i]
                items.append((cid, {'N': int(N), 'counts': [float(x) for x in dl['cnt'][k, o0:o0 + ch.ncat[i]]]}))
        top = max(int(c) for c in vw['cid']) + 1
        if ch.ctype[i] != L.CATEGORICAL:
            sx, sxx = (('sum_log_x', 'sum_log_x_sq') if ch.lognormal[i] else codec.SUFFSTAT_NAMES[int(ch.ctype[i])])
            d = {'N': 0, sx: 0}
            if sxx is not None:
                d[sxx] = 0
            items.append((top, d))
        else:
            items.append((top, {'N': 0, 'counts': [0.0] * int(ch.ncat[i])}))
        md['suffstats'].append(items)
    md['Cd'] = ch.Cd
    md['Ci'] = ch.Ci
    md['Zrv'] = [(int(v), [int(z) for z in vw['Zr']]) for v, vw in spec['views'].items()]
    md['view_alphas'] = [(int(v), float(vw['alpha'])) for v, vw in spec['views'].items()]
    md['diagnostics'] = ch.diagnostics[s]
    md['hooked_cgpms'] = dict()
    md['loom_path'] = None
    md['factory'] = ('cgpm.crosscat.state', 'State')
    return md


class _ViewInfo(object):
    """Read-only stand-in for cgpm.mixtures.view.View (Zr / Nk / alpha / dims)."""

    def __init__(self, vid, vw, cols):
        self.vid = vid
        self._zr = vw['Zr']
        self._alpha = vw['alpha']
        self._order = vw['order']
        self.dims = OrderedDict((c, None) for c in cols)
        self.outputs = [10 ** 7 + vid] + list(cols)

    def Zr(self, rowid=None):
        if rowid is not None:
            return int(self._zr[rowid])
        return OrderedDict((int(r), int(self._zr[r])) for r in self._order)

    def Nk(self, k=None):
        ids, counts = np.unique(self._zr, return_counts=True)
        d = OrderedDict((int(i), int(c)) for i, c in zip(ids, counts))
        return d[k] if k is not None else d

    def alpha(self):
        return self._alpha


class _Api(object):
    """Behaviour shared by Engine (S chains) and State (1 chain)."""

    _ch = None

    # ---- per-chain accessors -------------------------------------------------
    def _Zv(self, s):
        return OrderedDict(self._ch.Zv_items(s))

    def _views(self, s):
        spec, _ = self._ch.chain_spec(s)
        out = OrderedDict()
        for vid, vw in spec['views'].items():
            cols = [self._ch.outputs[i] for i, v in enumerate(spec['Zv']) if v == vid]
            out[vid] = _ViewInfo(vid, vw, cols)
        return out

    def _dependence_probability(self, s, col0, col1):
        zv = self._Zv(s)
        if col0 not in zv or col1 not in zv:
            raise ValueError('Unknown columns: %s, %s' % (col0, col1))
        return float(zv[col0] == zv[col1])

    def _dependence_probability_pairwise(self, s, colnos=None):
        """state.py:546-554."""
        if colnos is None:
            colnos = self._ch.outputs
        zv = self._Zv(s)
        z = np.array([zv[c] for c in colnos])
        return (z[:, None] == z[None, :]).astype(float)


    # ---- ensemble queries built on the query kernels -----------------------
    def _row_similarity(self, s, row0, row1, cols=None):
        """state.py:580-584: fraction of the views of `cols` in which the rows share a cluster."""
        views = self._views(s)
        zv = self._Zv(s)
        vids = set(zv[c] for c in (cols if cols is not None else self._ch.outputs))
        return float(np.mean([views[v].Zr(row0) == views[v].Zr(row1) for v in vids]))

    def _row_similarity_pairwise(self, s, cols=None):
        """state.py:586-594, evaluated on the device partitions."""
        import torch
        ch = self._ch
        m = ch.mirror()
        cidx = [ch.col_index[c] for c in cols] if cols is not None else list(range(ch.C))
        slots = sorted(set(int(m['zv'][s][i]) for i in cidx))
        acc = torch.zeros((ch.R, ch.R), dtype=torch.float64, device=ch.dev.device)
        for v in slots:
            z = ch.dev.t['zr'][s, v]
            acc += (z[:, None] == z[None, :]).to(torch.float64)
        return (acc / len(slots)).cpu().numpy()

    def _relevance_probability(self, s, rowid_target, rowid_query, col, hypotheticals=None):
        """state.py:599-629.  Hypothetical rows are incorporated into the column's view with one Gibbs step each
        (view.py:149-172), consuming the chain's RandomState like the reference; this is done on a temporary copy
        of the chain, so -- unlike the reference, whose incorporate / unincorporate pair may change the view's sums
        in the last bit -- the live chain's statistics are untouched."""
        if col not in self._ch.col_index:
            raise ValueError('Unknown column: %s' % (col,))
        vid = self._Zv(s)[col]
        view = self._views(s)[vid]
        hyps = []
        for h in (hypotheticals or []):
            q = {d: float(h[d]) for d in view.dims if d in h and not math.isnan(h[d])}
            if q:
                hyps.append(q)
        rowid_query = list(rowid_query)
        if hyps:
            from .structure import CRP_ID_VIEW
            md = _chain_metadata(self._ch, s, X_list=self._ch.X_raw.tolist())
            opts = dict(getattr(self, '_opts', {}) or {})
            opts['stream'] = self._ch.stream
            tmp = State.from_metadata(md, rng=self._ch.rngs[s], **opts)
            R = tmp._ch.R
            # the other views keep the row out of the way: seated at an existing table, no Gibbs step, no cells
            tokens = {CRP_ID_VIEW + w: int(vw.Zr(0)) for w, vw in tmp.views.items() if w != vid}
            for j, q in enumerate(hyps):
                tmp.incorporate(R + j, dict(q, **tokens))
            view = tmp.views[vid]
            rowid_query = rowid_query + list(range(R, R + len(hyps)))
        return int(all(view.Zr(rowid_target) == view.Zr(rq) for rq in rowid_query)) if rowid_query else 0

    def _mutual_information(self, s, col0, col1, constraints=None, T=None, N=None):
        """state.py:638-722 for non-composite states: the connected components are the views."""
        from . import query
        constraints = dict(constraints or {})
        if any(c in constraints for c in col0 + col1):
            raise ValueError('Target and constraints columns must be disjoint.')
        if any(c in col1 for c in col0) and set(col0) != set(col1):
            raise ValueError('Targets must match exactly or be disjoint.')
        zv = self._Zv(s)
        blocks = {}
        for c in col0:
            blocks.setdefault(zv[c], ([], [], {}))[0].append(c)
        for c in col1:
            blocks.setdefault(zv[c], ([], [], {}))[1].append(c)
        for c, x in constraints.items():
            blocks.setdefault(zv[c], ([], [], {}))[2][c] = x
        N = N or 100
        T = T or 100
        ch = self._ch

        def lp(targets_list, const):
            n = len(targets_list)
            return np.asarray(query.logpdf_bulk(ch, [s], [-1] * n, targets_list, [const] * n)[0])

        def estimator(c0, c1, const):
            if set(c0) != set(c1):
                samples = query.simulate_bulk(ch, [s], [-1], [c0 + c1], [const], [N])[0][0]
                pxy = lp(samples, const)
                px = lp([{c: smp[c] for c in c0} for smp in samples], const)
                py = lp([{c: smp[c] for c in c1} for smp in samples], const)
                return float((pxy.sum() - px.sum() - py.sum()) / N)
            samples = query.simulate_bulk(ch, [s], [-1], [c0], [const], [N])[0][0]
            return float(-lp([{c: smp[c] for c in c0} for smp in samples], const).sum() / N)

        total = 0.0
        for c0, c1, const in blocks.values():
            if not (c0 and c1):
                continue
            e_const = {e: x for e, x in const.items() if x is not None}
            m_const = [e for e, x in const.items() if x is None]
            if not m_const:
                total += estimator(c0, c1, const)
                continue
            m_samples = query.simulate_bulk(ch, [s], [-1], [m_const], [{}], [T])[0][0]
            total += sum(estimator(c0, c1, dict(e_const, **smp)) for smp in m_samples) / float(T)
        return total

    # ---- incremental edits (state.py:187-320, engine.py:119-174) ---------------------------
    def _after_edit(self, new):
        self._ch = new

    def incorporate(self, rowid, observation, inputs=None, multiprocess=1):
        if inputs:
            raise ValueError('Cannot incorporate with inputs: %s' % inputs)
        from . import structure
        self._after_edit(structure.incorporate(self._ch, rowid, observation))

    def incorporate_bulk(self, rowids, observations, inputs=None, multiprocess=1):
        for rowid, observation in zip(rowids, observations):
            self.incorporate(rowid, observation, inputs)

    def unincorporate(self, rowid, multiprocess=1):
        from . import structure
        self._after_edit(structure.unincorporate(self._ch, rowid))

    def force_cell(self, rowid, observation, multiprocess=1):
        from . import structure
        self._after_edit(structure.force_cell(self._ch, rowid, observation))

    def force_cell_bulk(self, rowids, queries, multiprocess=1):
        for rowid, query in zip(rowids, queries):
            self.force_cell(rowid, query)

    def incorporate_dim(self, T, outputs, inputs=None, cctype=None, distargs=None, v=None, multiprocess=1):
        from . import structure
        self._after_edit(structure.incorporate_dim(self._ch, T, outputs, inputs, cctype, distargs, v))

    def unincorporate_dim(self, col, multiprocess=1):
        from . import structure
        self._after_edit(structure.unincorporate_dim(self._ch, col))


class State(_Api):
    """cgpm.crosscat.state.State for one chain (src/crosscat/state.py:44-1284)."""

    def __init__(self, X, outputs=None, inputs=None, cctypes=None, distargs=None, Zv=None, Zrv=None,
                 alpha=None, view_alphas=None, hypers=None, Cd=None, Ci=None, Rd=None, Ri=None,
                 diagnostics=None, loom_path=None, rng=None, **opts):
        _check_state_kwargs(dict(inputs=inputs, Rd=Rd, Ri=Ri))
        bad = [k for k in opts if k not in OPT_KW]
        if bad:
            raise TypeError('unexpected keyword arguments: %s' % (bad,))
        self.rng = gen_rng() if rng is None else rng
        X = np.asarray(X)
        opts.setdefault('stream', 'numpy')
        self._ch = Chains(X, outputs=outputs, cctypes=cctypes, distargs=distargs, Cd=Cd, Ci=Ci, S=1,
                          seed=int(self.rng.get_state()[1][0]), **opts)
        self._ch.init_chain(0, self.rng, Zv=Zv, Zrv=Zrv, alpha=alpha, view_alphas=view_alphas,
                            hypers=hypers, diagnostics=diagnostics)
        self._ch.finish_init()
        self.outputs = list(self._ch.outputs)
        self.inputs = []
        self.Cd, self.Ci = self._ch.Cd, self._ch.Ci

    def _after_edit(self, new):
        self._ch = new
        self.outputs = list(new.outputs)

    # ---- structure ---------------------------------------------------------------
    @property
    def diagnostics(self):
        return self._ch.diagnostics[0]

    @property
    def views(self):
        return self._views(0)

    def n_rows(self):
        return self._ch.R

    def n_cols(self):
        return self._ch.C

    def cctypes(self):
        return list(self._ch.cctypes)

    def distargs(self):
        return [dict(d) for d in self._ch.distargs]

    def data_array(self):
        return self._ch.X_raw.copy()

    def alpha(self):
        return float(self._ch.dev.t['col_alpha'][0].item())

    def Zv(self, col=None):
        zv = self._Zv(0)
        return zv[col] if col is not None else zv

    def Nv(self, v=None):
        zv = self._Zv(0)
        out = OrderedDict()
        for c, w in zv.items():
            out[w] = out.get(w, 0) + 1
        return out[v] if v is not None else out

    def Zrv(self):
        return OrderedDict((vid, [vw.Zr(i) for i in range(self.n_rows())]) for vid, vw in self.views.items())

    def hypers(self):
        return _chain_metadata(self._ch, 0)['hypers']

    def is_composite(self):
        return False

    def hypothetical(self, rowid):
        return rowid is None or not 0 <= rowid < self.n_rows()

    # ---- inference -------------------------------------------------------------
    def transition(self, N=None, S=None, kernels=None, rowids=None, cols=None, views=None,
                   progress=True, checkpoint=None):
        self._ch.transition(N=N, S=S, kernels=kernels, rowids=rowids, cols=cols, views=views,
                            progress=progress, checkpoint=checkpoint, chains=[0])

    def transition_crp_alpha(self):
        self._ch.transition_crp_alpha([0])

    def transition_view_alphas(self, views=None, cols=None):
        self._ch.transition_view_alphas([0], views=views, cols=cols)

    def transition_dim_hypers(self, cols=None):
        self._ch.transition_dim_hypers([0], cols=cols)

    def transition_view_rows(self, views=None, rows=None, cols=None):
        self._ch.transition_view_rows([0], views=views, cols=cols, rows=rows)

    def transition_dims(self, cols=None):
        self._ch.transition_dims([0], cols=cols)
        self._ch.dev.sync()
        self._ch.dev.check_errors()

    # ---- queries -----------------------------------------------------------------
    def logpdf_score(self):
        return float(self._ch.logpdf_score()[0])

    def dependence_probability(self, col0, col1):
        return self._dependence_probability(0, col0, col1)

    def dependence_probability_pairwise(self, colnos=None):
        return self._dependence_probability_pairwise(0, colnos)

    def row_similarity(self, row0, row1, cols=None):
        return self._row_similarity(0, row0, row1, cols)

    def row_similarity_pairwise(self, cols=None):
        return self._row_similarity_pairwise(0, cols)

    def relevance_probability(self, rowid_target, rowid_query, col, hypotheticals=None):
        return self._relevance_probability(0, rowid_target, rowid_query, col, hypotheticals)

    def mutual_information(self, col0, col1, constraints=None, T=None, N=None, progress=None):
        return self._mutual_information(0, col0, col1, constraints, T, N)

    def logpdf(self, rowid, targets, constraints=None, inputs=None, accuracy=None):
        return self._ch.logpdf_bulk([0], [rowid], [targets], [constraints or {}])[0][0]

    def logpdf_bulk(self, rowids, targets_list, constraints_list=None, inputs_list=None):
        return self._ch.logpdf_bulk([0], rowids, targets_list, constraints_list)[0]

    def simulate(self, rowid, targets, constraints=None, inputs=None, N=None, accuracy=None):
        out = self._ch.simulate_bulk([0], [rowid], [targets], [constraints or {}], [N if N is not None else 1])[0][0]
        return out if N is not None else out[0]

    def simulate_bulk(self, rowids, targets_list, constraints_list=None, inputs_list=None, Ns=None):
        return self._ch.simulate_bulk([0], rowids, targets_list, constraints_list, Ns)[0]

    # ---- serialise ---------------------------------------------------------------
    def to_metadata(self):
        return _chain_metadata(self._ch, 0, X_list=self._ch.X_raw.tolist())

    def to_pickle(self, fileptr):
        pickle.dump(self.to_metadata(), fileptr)

    @classmethod
    def from_metadata(cls, metadata, rng=None, **opts):
        """state.py:1246-1267: suffstats are ignored and rebuilt from X, Zv, Zrv."""
        if rng is None:
            rng = gen_rng(0)
        if metadata.get('hooked_cgpms'):
            raise ValueError('cgpm_b200: composite states (hooked cgpms) are out of scope.')
        to_dict = lambda val: None if val is None else OrderedDict(val)
        return cls(
            np.asarray(metadata['X']), outputs=metadata.get('outputs'), cctypes=metadata.get('cctypes'),
            distargs=metadata.get('distargs'), alpha=metadata.get('alpha'), Zv=to_dict(metadata.get('Zv')),
            Zrv=to_dict(metadata.get('Zrv')), view_alphas=to_dict(metadata.get('view_alphas')),
            hypers=metadata.get('hypers'), Cd=metadata.get('Cd'), Ci=metadata.get('Ci'),
            diagnostics=metadata.get('diagnostics'), loom_path=metadata.get('loom_path'), rng=rng, **opts)

    @classmethod
    def from_pickle(cls, fileptr, rng=None, **opts):
        if isinstance(fileptr, str):
            with open(fileptr, 'rb') as f:
                metadata = pickle.load(f)
        else:
            metadata = pickle.load(fileptr)
        return cls.from_metadata(metadata, rng=rng, **opts)


class Engine(_Api):
    """cgpm.crosscat.engine.Engine (src/crosscat/engine.py:61-435)."""

    def __init__(self, X, num_states=1, rng=None, multiprocess=1, **kwargs):
        self.rng = gen_rng(1) if rng is None else rng
        self._opts = {k: kwargs.pop(k) for k in OPT_KW if k in kwargs}
        self._opts.setdefault('stream', 'philox')
        _check_state_kwargs(kwargs)
        self.X = np.asarray(X, dtype=float)
        self._state_kwargs = kwargs
        seeds = self._get_seeds(num_states)
        self._make_chains(num_states, seeds[0] if num_states else 0)
        for s, seed in enumerate(seeds):                       # engine.py:36-38
            self._init_from_kwargs(s, gen_rng(seed), kwargs)
        if num_states:
            self._ch.finish_init()

    def _after_edit(self, new):
        self._ch = new
        self.X = new.X_raw.copy()
        self._state_kwargs = dict(self._state_kwargs, outputs=list(new.outputs), cctypes=list(new.cctypes),
                                  distargs=[dict(d) for d in new.distargs])

    def _make_chains(self, S, seed):
        kw = self._state_kwargs
        old = getattr(self, '_ch', None)
        self._ch = Chains(self.X, outputs=kw.get('outputs'), cctypes=kw.get('cctypes'), distargs=kw.get('distargs'),
                          Cd=kw.get('Cd'), Ci=kw.get('Ci'), S=S, seed=int(seed), **self._opts)
        if old is not None:
            # re-created chains (drop_state / add_state) continue the call counter of the device generator: its keys are
            # (seed, counter, purpose, chain, ...), so no key a chain -- or the chain that inherits its index -- has
            # consumed is ever drawn again
            self._ch.counter = old.counter

    def _init_from_kwargs(self, s, rng, kw):
        self._ch.init_chain(s, rng, Zv=kw.get('Zv'), Zrv=kw.get('Zrv'), alpha=kw.get('alpha'),
                            view_alphas=kw.get('view_alphas'), hypers=kw.get('hypers'),
                            diagnostics=kw.get('diagnostics'))

    # ---- inference -------------------------------------------------------------
    def transition(self, N=None, S=None, kernels=None, rowids=None, cols=None, views=None,
                   progress=True, checkpoint=None, statenos=None, multiprocess=1):
        statenos = list(statenos) if statenos else list(range(self.num_states()))
        self._ch.transition(N=N, S=S, kernels=kernels, rowids=rowids, cols=cols, views=views,
                            progress=progress, checkpoint=checkpoint, chains=statenos)

    # ---- queries -----------------------------------------------------------------
    def logpdf_score(self, statenos=None, multiprocess=1):
        statenos = statenos or range(self.num_states())
        scores = self._ch.logpdf_score()
        return [float(scores[s]) for s in statenos]

    def logpdf(self, rowid, targets, constraints=None, inputs=None, accuracy=None, statenos=None,
               multiprocess=1):
        statenos = list(statenos or range(self.num_states()))
        out = self._ch.logpdf_bulk(statenos, [rowid], [targets], [constraints or {}])
        return [o[0] for o in out]

    def logpdf_bulk(self, rowids, targets_list, constraints_list=None, inputs_list=None, statenos=None,
                    multiprocess=1):
        statenos = list(statenos or range(self.num_states()))
        return self._ch.logpdf_bulk(statenos, rowids, targets_list, constraints_list)

    def simulate(self, rowid, targets, constraints=None, inputs=None, N=None, accuracy=None,
                 statenos=None, multiprocess=1):
        self._seed_states()
        statenos = list(statenos or range(self.num_states()))
        out = self._ch.simulate_bulk(statenos, [rowid], [targets], [constraints or {}], [N if N is not None else 1])
        return [o[0] if N is not None else o[0][0] for o in out]

    def simulate_bulk(self, rowids, targets_list, constraints_list=None, inputs_list=None, Ns=None,
                      statenos=None, multiprocess=1):
        self._seed_states()
        statenos = list(statenos or range(self.num_states()))
        return self._ch.simulate_bulk(statenos, rowids, targets_list, constraints_list, Ns)

    # ---- array form of the bulk queries (not in the reference: its dict-per-query form costs a Python loop per
    #      query; a batch of hypothetical rows over fixed target / constraint columns is encoded with numpy) ----
    def logpdf_bulk_arrays(self, target_cols, target_vals, constraint_cols=None, constraint_vals=None, statenos=None):
        """[n_states, Q] array: logpdf(-1, {target_cols[j]: target_vals[q, j]}, {constraint_cols[j]: constraint_vals[q, j]})."""
        from . import query
        statenos = list(statenos or range(self.num_states()))
        return query.logpdf_bulk_arrays(self._ch, statenos, target_cols, target_vals, constraint_cols, constraint_vals)

    def simulate_bulk_arrays(self, Q, target_cols, constraint_cols=None, constraint_vals=None, N=1, statenos=None):
        """[n_states, Q, N, len(target_cols)] array of predictive samples for Q hypothetical rows."""
        from . import query
        self._seed_states()
        statenos = list(statenos or range(self.num_states()))
        return query.simulate_bulk_arrays(self._ch, statenos, Q, target_cols, constraint_cols, constraint_vals, N)

    # ---- dense state bridge and instrumentation (the lovecat-style export / write-back of INTEGRATION.md) ----
    def to_dense(self, pinned=None):
        """Device -> pinned host copy of every chain's state in dense slot form (the write-back half of
        lovecat._update_state, src/crosscat/lovecat.py:226-290): dict of CPU tensors, reusable as `pinned`."""
        return self._ch.export_dense(pinned=pinned)

    def load_dense(self, dense, X=None):
        """Host -> device copy of a dense state (and optionally the table, a pinned [R, C] float64 tensor), followed
        by the ordered rebuild of all sufficient statistics (the export half of the bridge, lovecat.py:73-223)."""
        self._ch.import_dense(dense, X=X)

    def start_kernel_timing(self):
        """Record a CUDA event pair around every CrossCat kernel call of transition()."""
        self._ch.events = []

    def stop_kernel_timing(self):
        """{kernel name: [milliseconds per call]} since start_kernel_timing()."""
        import torch
        torch.cuda.synchronize(self._ch.dev.device)
        events, self._ch.events = self._ch.events or [], None
        out = {}
        for name, a, b in events:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out

    def native_launches(self, reset=False):
        """Kernels of libcgpm_b200.so launched so far (a lower bound: counted per C-ABI call)."""
        n = sum(self._ch.dev.lib_calls.values())
        if reset:
            self._ch.dev.lib_calls.clear()
        return n

    def dependence_probability(self, col0, col1, statenos=None, multiprocess=1):
        statenos = statenos or range(self.num_states())
        return [self._dependence_probability(s, col0, col1) for s in statenos]

    def dependence_probability_pairwise(self, colnos=None, statenos=None, multiprocess=1):
        statenos = statenos or range(self.num_states())
        return [self._dependence_probability_pairwise(s, colnos) for s in statenos]

    def row_similarity(self, row0, row1, cols=None, statenos=None, multiprocess=1):
        statenos = statenos or range(self.num_states())
        return [self._row_similarity(s, row0, row1, cols) for s in statenos]

    def row_similarity_pairwise(self, cols=None, statenos=None, multiprocess=1):
        statenos = statenos or range(self.num_states())
        return [self._row_similarity_pairwise(s, cols) for s in statenos]

    def relevance_probability(self, rowid_target, rowid_query, col, hypotheticals=None, statenos=None,
                              multiprocess=1):
        statenos = statenos or range(self.num_states())
        return [self._relevance_probability(s, rowid_target, rowid_query, col, hypotheticals) for s in statenos]

    def mutual_information(self, col0, col1, constraints=None, T=None, N=None, progress=None, statenos=None,
                           multiprocess=1):
        """engine.py:253-263: one estimate per state."""
        self._seed_states()
        statenos = statenos or range(self.num_states())
        return [self._mutual_information(s, col0, col1, constraints, T, N) for s in statenos]

    def _likelihood_weighted_integrate(self, logpdfs, rowid, constraints=None, inputs=None, statenos=None,
                                       multiprocess=1):
        """engine.py:361-371."""
        if constraints:
            weights = self.logpdf(rowid, constraints, inputs, statenos=statenos)
            return hm.logmeanexp_weighted(logpdfs, weights)
        return hm.logmeanexp(logpdfs)

    # ---- states ------------------------------------------------------------------
    def num_states(self):
        return self._ch.S if self._ch is not None else 0

    def get_state(self, index):
        md = _chain_metadata(self._ch, index, X_list=self.X.tolist())
        st = State.from_metadata(md, rng=self._ch.rngs[index], **self._opts_for_state())
        return st

    def _opts_for_state(self):
        o = dict(self._opts)
        return o

    @property
    def states(self):
        return [self.get_state(i) for i in range(self.num_states())]

    def _all_metadata(self):
        return [_chain_metadata(self._ch, s) for s in range(self.num_states())]

    def _reload(self, mds, rngs):
        kw = self._state_kwargs
        if mds:
            kw = dict(kw, outputs=mds[0]['outputs'], cctypes=mds[0]['cctypes'], distargs=mds[0]['distargs'],
                      Cd=mds[0].get('Cd'), Ci=mds[0].get('Ci'))
            self._state_kwargs = {k: v for k, v in kw.items() if k in ('outputs', 'cctypes', 'distargs', 'Cd', 'Ci')}
        self._make_chains(len(mds), self._ch.seed if self._ch is not None else 0)
        for s, (md, rng) in enumerate(zip(mds, rngs)):
            self._ch.init_chain(s, rng, Zv=OrderedDict(md['Zv']), Zrv=OrderedDict(md['Zrv']), alpha=md['alpha'],
                                view_alphas=OrderedDict(md['view_alphas']), hypers=md['hypers'],
                                diagnostics=md.get('diagnostics'))
        if mds:
            self._ch.finish_init()

    def drop_state(self, index):
        mds, rngs = self._all_metadata(), list(self._ch.rngs)
        del mds[index]
        del rngs[index]
        self._reload(mds, rngs)

    def add_state(self, count=1, multiprocess=1, **kwargs):
        """engine.py:331-346."""
        forbidden = ['X', 'outputs', 'cctypes', 'distargs']
        if [f for f in forbidden if f in kwargs]:
            raise ValueError('Cannot specify arguments for: %s.' % (forbidden,))
        _check_state_kwargs(kwargs)
        mds, rngs = self._all_metadata(), list(self._ch.rngs)
        old = len(mds)
        seeds = self._get_seeds(count)
        tmp_kw = dict(self._state_kwargs)
        self._state_kwargs = dict(outputs=self._ch.outputs, cctypes=self._ch.cctypes, distargs=self._ch.distargs,
                                  Cd=kwargs.get('Cd', self._ch.Cd), Ci=kwargs.get('Ci', self._ch.Ci))
        self._make_chains(old + count, self._ch.seed)
        for s, (md, rng) in enumerate(zip(mds, rngs)):
            self._ch.init_chain(s, rng, Zv=OrderedDict(md['Zv']), Zrv=OrderedDict(md['Zrv']), alpha=md['alpha'],
                                view_alphas=OrderedDict(md['view_alphas']), hypers=md['hypers'],
                                diagnostics=md.get('diagnostics'))
        for j, seed in enumerate(seeds):
            self._init_from_kwargs(old + j, gen_rng(seed), kwargs)
        self._ch.finish_init()
        del tmp_kw

    # ---- internal ------------------------------------------------------------------
    def _seed_states(self):
        """engine.py:352-355."""
        seeds = self._get_seeds()
        for seed, rng in zip(seeds, self._ch.rngs):
            rng.seed(seed)

    def _get_seeds(self, N=None):
        """engine.py:357-359."""
        num_draws = N if N is not None else self.num_states()
        return self.rng.randint(low=1, high=2 ** 32 - 1, size=num_draws)

    # ---- serialise -------------------------------------------------------------------
    def to_metadata(self):
        """engine.py:394-401."""
        return {'X': self.X.tolist(), 'states': self._all_metadata(),
                'factory': ('cgpm.crosscat.engine', 'Engine')}

    @classmethod
    def from_metadata(cls, metadata, rng=None, multiprocess=1, **opts):
        """engine.py:403-424."""
        if rng is None:
            rng = gen_rng(0)
        states = metadata['states']
        if any(m.get('hooked_cgpms') for m in states):
            raise ValueError('cgpm_b200: composite states (hooked cgpms) are out of scope.')
        eng = cls.__new__(cls)
        eng.rng = rng
        eng._opts = {k: opts[k] for k in OPT_KW if k in opts}
        eng._opts.setdefault('stream', 'philox')
        eng.X = np.asarray(metadata['X'], dtype=float)
        eng._state_kwargs = {}
        eng._ch = None
        seeds = eng.rng.randint(low=1, high=2 ** 32 - 1, size=len(states))
        eng._ch = None
        if states:
            eng._state_kwargs = dict(outputs=states[0]['outputs'], cctypes=states[0]['cctypes'],
                                     distargs=states[0]['distargs'], Cd=states[0].get('Cd'), Ci=states[0].get('Ci'))
        eng._make_chains(len(states), seeds[0] if len(states) else 0)
        for s, (md, seed) in enumerate(zip(states, seeds)):
            eng._ch.init_chain(s, gen_rng(seed), Zv=OrderedDict(md['Zv']), Zrv=OrderedDict(md['Zrv']),
                               alpha=md['alpha'], view_alphas=OrderedDict(md['view_alphas']),
                               hypers=md['hypers'], diagnostics=md.get('diagnostics'))
        if states:
            eng._ch.finish_init()
        return eng

    def to_pickle(self, fileptr):
        pickle.dump(self.to_metadata(), fileptr)

    @classmethod
    def from_pickle(cls, fileptr, rng=None, **opts):
        if isinstance(fileptr, str):
            with open(fileptr, 'rb') as f:
                metadata = pickle.load(f)
        else:
            metadata = pickle.load(fileptr)
        return cls.from_metadata(metadata, rng=rng, **opts)
