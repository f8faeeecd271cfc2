"""CPU restatement of the CrossCat Gibbs hot path of probcomp/cgpm.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; nothing under
cgpm_b200/ does.  It is a from-scratch, single-file restatement (plain Python
containers + math/scipy scalars) of the algorithm in the reference files
listed below; it is NOT the product and is not fast.

PARITY PINNED: tests/test_oracle_vs_reference.py replays whole transition
trajectories of this module against the reference itself (the py3-shimmed
build under oracle/_ref, made by oracle/build_ref.py from /root/reference/src)
with the same numpy RandomState seeds and requires bit-identical partitions,
hyperparameters, sufficient statistics and scores; tests/golden/*.json holds
vectors generated from the reference by tests/golden/make_golden.py and is
checked on any machine (no /root/reference needed); the known-answer values of
the reference's own tests (test_logsumexp.py:31,47, test_crp.py:117-177) are
checked in tests/test_oracle_kat.py.

Reference map (all paths relative to /root/reference/src):
  rows kernel      mixtures/view.py:247-252,393-429   OView.transition_rows
  row CRP          primitives/crp.py:45-59,111-146    crp_gibbs_tables/_logps
  column kernel    crosscat/state.py:804-810,1030-1154  OState.transition_dims
  bulk rebuild     mixtures/view.py:483-494           ODim.bulk_incorporate
  dim routing/NaN  mixtures/dim.py:95-132,152-184     ODim
  normal           primitives/normal.py:64-76,128-139,165-200
  categorical      primitives/categorical.py:53-66,110-114,144-155
  lognormal        primitives/lognormal.py:56-102,151-157
  bernoulli        primitives/bernoulli.py:46-61,107-112,141-155
  poisson          primitives/poisson.py:49-66,119-126,147-172
  exponential      primitives/exponential.py:47-58,109-114,140-166
  geometric        primitives/geometric.py:47-58,105-110,140-164
  sampling         utils/general.py:80-86,123-157,203-205
  scheduler        crosscat/state.py:727-810,866-914
  queries          crosscat/sampling.py:37-116, crosscat/state.py:384-387,538-554

Every floating-point expression keeps the reference's operand order so that the
results are bit-identical on the same libm / numpy / scipy.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from math import lgamma, log

import numpy as np
from scipy.special import betaln, gammaln

LOG2 = log(2)
LOGPI = log(math.pi)
LOG2PI = LOG2 + LOGPI
NEG_INF = float('-inf')


# ----------------------------------------------------------------------------
# utils/general.py:80-86,123-157,203-205

def logsumexp(array):
    """general.py:141-157 (sequential left-to-right sum of exp(a - max))."""
    if len(array) == 0:
        return NEG_INF
    m = max(array)
    if math.isinf(m) and min(array) != -m and \
            all(not math.isnan(a) for a in array):
        return m
    return m + math.log(sum(math.exp(a - m) for a in array))


def logmeanexp(array):
    """general.py:159-182."""
    if len(array) == 0:
        return NEG_INF
    noninfs = [a for a in array if not a == NEG_INF]
    return logsumexp(noninfs) - math.log(len(array))


def log_normalize(logp):
    return np.subtract(logp, logsumexp(logp))


def pflip(p, array=None, size=None, rng=None):
    """general.py:128-139.  One random_sample() per draw unless len(p)==1."""
    if array is None:
        array = list(range(len(p)))
    if len(p) == 1:
        return array[0] if size is None else [array[0]] * size
    p = np.asarray(p, dtype=float) / sum(p)
    return rng.choice(array, size=size, p=p)


def log_pflip(logp, array=None, size=None, rng=None):
    """general.py:123-126."""
    p = np.exp(log_normalize(logp))
    return pflip(p, array=array, size=size, rng=rng)


def log_linspace(a, b, n):
    return np.exp(np.linspace(log(a), log(b), n))


# ----------------------------------------------------------------------------
# primitives/normal.py

def normal_posterior(N, sx, sxx, m, r, s, nu):
    """normal.py:183-191."""
    rn = r + float(N)
    nun = nu + float(N)
    mn = (r*m + sx)/rn
    sn = s + sxx + r*m*m - rn*mn*mn
    if sn == 0:
        sn = s
    return mn, rn, sn, nun


def normal_log_Z(r, s, nu):
    """normal.py:193-200."""
    return (
        ((nu + 1.) / 2.) * LOG2
        + .5 * LOGPI
        - .5 * log(r)
        - (nu/2.) * log(s)
        + lgamma(nu/2.))


def normal_predictive(x, N, sx, sxx, m, r, s, nu):
    """normal.py:165-173."""
    _mn, rn, sn, nun = normal_posterior(N, sx, sxx, m, r, s, nu)
    _mm, rm, sm, num = normal_posterior(N+1, sx+x, sxx+x*x, m, r, s, nu)
    ZN = normal_log_Z(rn, sn, nun)
    ZM = normal_log_Z(rm, sm, num)
    return -.5 * LOG2PI + ZM - ZN


def normal_marginal(N, sx, sxx, m, r, s, nu):
    """normal.py:175-181."""
    _mn, rn, sn, nun = normal_posterior(N, sx, sxx, m, r, s, nu)
    Z0 = normal_log_Z(r, s, nu)
    ZN = normal_log_Z(rn, sn, nun)
    return -(N/2.) * LOG2PI + ZN - Z0


def normal_grids(X, n_grid=30):
    """normal.py:128-139 (X = the column's non-NaN values)."""
    N = len(X) + 1.
    ssqdev = np.var(X) * len(X) + 1.
    return OrderedDict([
        ('m', np.linspace(min(X), max(X) + 5, n_grid)),
        ('r', log_linspace(1. / N, N, n_grid)),
        ('s', log_linspace(ssqdev / 100., ssqdev, n_grid)),
        ('nu', log_linspace(1., N, n_grid)),
    ])


def lognormal_grids(X, n_grid=30):
    """lognormal.py:151-157 (X = the column's non-NaN values, on the original scale)."""
    return OrderedDict([
        ('m', log_linspace(1e-4, max(X), n_grid)),
        ('r', log_linspace(.1, float(len(X)), n_grid)),
        ('nu', log_linspace(.1, float(len(X)), n_grid)),
        ('s', log_linspace(.1, float(len(X)), n_grid)),
    ])


# ----------------------------------------------------------------------------
# primitives/categorical.py

def categorical_valid(x, k):
    return x % 1 == 0 and 0 <= x < k


def categorical_predictive(x, N, counts, alpha):
    """categorical.py:144-148."""
    numer = log(alpha + counts[x])
    denom = log(np.sum(counts) + alpha * len(counts))
    return numer - denom


def categorical_marginal(N, counts, alpha):
    """categorical.py:150-155."""
    K = len(counts)
    A = K * alpha
    lg = sum(gammaln(counts[k] + alpha) for k in range(K))
    return gammaln(A) - gammaln(A+N) + lg - K * gammaln(alpha)


def categorical_grids(X, n_grid=30):
    """categorical.py:110-114."""
    return OrderedDict([('alpha', log_linspace(1., float(len(X)), n_grid))])


# ----------------------------------------------------------------------------
# primitives/bernoulli.py, poisson.py, exponential.py, geometric.py: statistics [N, sum_x, sum_log_fact_x]
# (the third slot is used by poisson only)

def bernoulli_predictive(x, N, x_sum, alpha, beta):
    """bernoulli.py:141-151."""
    log_denom = log(N + alpha + beta)
    if x == 1:
        return log(x_sum + alpha) - log_denom
    return log(N - x_sum + beta) - log_denom


def bernoulli_marginal(N, x_sum, alpha, beta):
    """bernoulli.py:153-155."""
    return betaln(x_sum + alpha, N - x_sum + beta) - betaln(alpha, beta)


def gamma_log_Z(a, b):
    """poisson.py:169-172, exponential.py:163-166."""
    return gammaln(a) - a*log(b)


def poisson_predictive(x, N, sum_x, a, b):
    """poisson.py:147-153."""
    an, bn = a + sum_x, b + N
    am, bm = a + (sum_x + x), b + (N + 1)
    ZN = gamma_log_Z(an, bn)
    ZM = gamma_log_Z(am, bm)
    return ZM - ZN - gammaln(x+1)


def poisson_marginal(N, sum_x, sum_log_fact_x, a, b):
    """poisson.py:155-160."""
    an, bn = a + sum_x, b + N
    Z0 = gamma_log_Z(a, b)
    ZN = gamma_log_Z(an, bn)
    return ZN - Z0 - sum_log_fact_x


def exponential_predictive(x, N, sum_x, a, b):
    """exponential.py:140-146."""
    an, bn = a + N, b + sum_x
    am, bm = a + (N + 1), b + (sum_x + x)
    ZN = gamma_log_Z(an, bn)
    ZM = gamma_log_Z(am, bm)
    return ZM - ZN


def exponential_marginal(N, sum_x, a, b):
    """exponential.py:148-153."""
    an, bn = a + N, b + sum_x
    return gamma_log_Z(an, bn) - gamma_log_Z(a, b)


def geometric_predictive(x, N, sum_x, a, b):
    """geometric.py:140-146."""
    an, bn = a + N, b + sum_x
    am, bm = a + (N + 1), b + (sum_x + x)
    return betaln(am, bm) - betaln(an, bn)


def geometric_marginal(N, sum_x, a, b):
    """geometric.py:148-153."""
    an, bn = a + N, b + sum_x
    return betaln(an, bn) - betaln(a, b)


def count_grids(cctype, X, n_grid=30):
    """construct_hyper_grids of bernoulli.py:107-112, poisson.py:119-126, exponential.py:109-114, geometric.py:105-110."""
    n = float(len(X))
    if cctype == 'bernoulli':
        return OrderedDict([('alpha', log_linspace(1., n, n_grid)), ('beta', log_linspace(1., n, n_grid))])
    if cctype == 'poisson':
        return OrderedDict([('a', np.unique(np.round(np.linspace(1, len(X), n_grid)))), ('b', log_linspace(.1, n, n_grid))])
    if cctype == 'exponential':
        return OrderedDict([('a', log_linspace(.5, n, n_grid)), ('b', log_linspace(.5, n, n_grid))])
    return OrderedDict([('a', log_linspace(1, n / 2., n_grid)), ('b', log_linspace(.1, n / 2., n_grid))])


COUNT_TYPES = ('bernoulli', 'poisson', 'exponential', 'geometric')


def count_valid(cctype, x):
    """The values each primitive's incorporate accepts / its logpdf gives mass to."""
    if cctype == 'bernoulli':
        return x in [0, 1]
    if cctype == 'exponential':
        return x >= 0
    return x % 1 == 0 and x >= 0


# ----------------------------------------------------------------------------
# primitives/crp.py

def crp_grids(n, n_grid=30):
    """crp.py:148-152 with X = [1]*n."""
    return log_linspace(1./n, n, n_grid)


def crp_predictive(x, N, counts, alpha):
    """crp.py:178-182."""
    return log(counts.get(x, alpha)) - log(N + alpha)


def crp_marginal(N, counts, alpha):
    """crp.py:184-188."""
    return len(counts) * log(alpha) + sum(gammaln(list(counts.values()))) \
        + gammaln(alpha) - gammaln(N + alpha)


def crp_gibbs_tables(counts, own, singleton):
    """crp.py:126-143 (m = 1)."""
    K = sorted(counts)
    return K if singleton else K + [max(counts) + 1]


def crp_gibbs_logps(counts, own, alpha):
    """crp.py:111-124 (m = 1)."""
    singleton = counts[own] == 1
    p_row = alpha if singleton else counts[own] - 1
    tables = crp_gibbs_tables(counts, own, singleton)

    def weight(t):
        if t == own:
            return p_row
        if t not in counts:
            return alpha
        return counts[t]
    return tables, [log(weight(t)) for t in tables]


class OCrp(object):
    """Open-set CRP with the reference's container semantics: `data` keeps the
    insertion order of members (a re-seated member moves to the end,
    crp.py:45-59) and `counts` the order in which tables were opened."""

    def __init__(self, n_grid_items, alpha, rng):
        self.grid = crp_grids(n_grid_items)
        self.alpha = rng.choice(self.grid) if alpha is None else alpha
        self.N = 0
        self.data = OrderedDict()
        self.counts = OrderedDict()

    def incorporate(self, item, x):
        x = int(x)
        self.N += 1
        if x not in self.counts:
            self.counts[x] = 0
        self.counts[x] += 1
        self.data[item] = x

    def unincorporate(self, item):
        x = self.data.pop(item)
        self.N -= 1
        self.counts[x] -= 1
        if self.counts[x] == 0:
            del self.counts[x]

    def simulate_fresh(self, rng):
        """crp.py:71-81 for a member that is not seated."""
        K = sorted(self.counts) + [max(self.counts) + 1] if self.counts else [0]
        logps = [crp_predictive(x, self.N, self.counts, self.alpha) for x in K]
        return log_pflip(logps, array=K, rng=rng)

    def logpdf_score(self):
        return crp_marginal(self.N, self.counts, self.alpha)

    def transition_alpha(self, rng):
        """Dim.transition_hypers (dim.py:152-174) specialised to the one
        hyper of a CRP: shuffling a 1-list draws nothing."""
        names = ['alpha']
        rng.shuffle(names)
        logps = []
        for g in self.grid:
            logp_k = 0
            logp_k += crp_marginal(self.N, self.counts, g)
            logps.append(logp_k)
        index = log_pflip(logps, rng=rng)
        self.alpha = self.grid[index]


# ----------------------------------------------------------------------------
# mixtures/dim.py

class ODim(object):
    """One column: cluster id -> sufficient statistics (dim.py:26-255)."""

    def __init__(self, col, cctype, distargs, hypers, xcol, rng):
        self.col = col
        self.cctype = cctype
        self.distargs = dict(distargs) if distargs else {}
        if cctype not in ('normal', 'categorical', 'lognormal') + COUNT_TYPES:
            raise ValueError('oracle does not cover cctype %s' % cctype)
        self.k = int(self.distargs['k']) if cctype == 'categorical' else 0
        self.hypers = OrderedDict(hypers) if hypers else OrderedDict()
        clean = [x for x in xcol if not math.isnan(x)]
        self.grids = normal_grids(clean) if cctype == 'normal' \
            else lognormal_grids(clean) if cctype == 'lognormal' \
            else count_grids(cctype, clean) if cctype in COUNT_TYPES \
            else categorical_grids(clean)
        if not self.hypers:                                   # dim.py:180-183
            for h in self.grids:
                self.hypers[h] = rng.choice(self.grids[h])
        self.clusters = {}

    # sufficient statistics ---------------------------------------------------
    def _empty(self):
        if self.cctype in ('normal', 'lognormal') + COUNT_TYPES:
            return [0, 0, 0]                                  # N, sum_x, sum_x_sq (of log x for lognormal; sum_log_fact_x for poisson)
        return [0, np.zeros(self.k)]                          # N, counts

    def _add(self, st, x):
        if self.cctype == 'normal':                           # normal.py:64-70
            st[0] += 1.
            st[1] += x
            st[2] += x*x
        elif self.cctype == 'lognormal':                      # lognormal.py:56-64
            if x <= 0:
                raise ValueError('Invalid Lognormal: %s' % str(x))
            st[0] += 1
            st[1] += math.log(x)
            st[2] += math.log(x) * math.log(x)
        elif self.cctype in COUNT_TYPES:                      # bernoulli.py:46-53, poisson.py:49-57, exponential.py:47-53, geometric.py:47-53
            if not count_valid(self.cctype, x):
                raise ValueError('Invalid %s: %s' % (self.cctype.capitalize(), str(x)))
            st[0] += 1
            st[1] += x
            if self.cctype == 'poisson':
                st[2] += gammaln(x+1)
        else:                                                 # categorical.py:53-61
            if not categorical_valid(x, self.k):
                raise ValueError('Invalid Categorical(%d): %s' % (self.k, x))
            st[0] += 1
            st[1][int(x)] += 1

    def _sub(self, st, x):
        if self.cctype == 'normal':                           # normal.py:72-76
            st[0] -= 1
            st[1] -= x
            st[2] -= x*x
        elif self.cctype == 'lognormal':                      # lognormal.py:66-70
            st[0] -= 1
            st[1] -= math.log(x)
            st[2] -= math.log(x) * math.log(x)
        elif self.cctype in COUNT_TYPES:
            st[0] -= 1
            st[1] -= x
            if self.cctype == 'poisson':
                st[2] -= gammaln(x+1)
        else:
            st[0] -= 1
            st[1][int(x)] -= 1

    def incorporate(self, x, k):
        """dim.py:95-106: a NaN cell still opens the cluster."""
        if k not in self.clusters:
            self.clusters[k] = self._empty()
        if not math.isnan(x):
            self._add(self.clusters[k], x)

    def unincorporate(self, x, k):
        if not math.isnan(x):
            self._sub(self.clusters[k], x)

    def bulk_incorporate(self, X, Zr):
        """view.py:483-494: rebuild in the view's member order."""
        self.clusters = {}
        xcol = X[self.col]
        for rowid, k in Zr.items():
            self.incorporate(xcol[rowid], k)

    # densities ---------------------------------------------------------------
    def _predictive(self, st, x):
        h = self.hypers
        if self.cctype == 'normal':
            return normal_predictive(x, st[0], st[1], st[2],
                                     h['m'], h['r'], h['s'], h['nu'])
        if self.cctype == 'lognormal':                        # lognormal.py:72-80
            if x <= 0:
                return NEG_INF
            return - math.log(x) + normal_predictive(math.log(x), st[0], st[1], st[2],
                                                     h['m'], h['r'], h['s'], h['nu'])
        if self.cctype in COUNT_TYPES:
            if not count_valid(self.cctype, x):
                return NEG_INF
            if self.cctype == 'bernoulli':
                return bernoulli_predictive(x, st[0], st[1], h['alpha'], h['beta'])
            f = {'poisson': poisson_predictive, 'exponential': exponential_predictive, 'geometric': geometric_predictive}[self.cctype]
            return f(x, st[0], st[1], h['a'], h['b'])
        if not categorical_valid(x, self.k):
            return NEG_INF
        return categorical_predictive(int(x), st[0], st[1], h['alpha'])

    def _marginal(self, st):
        h = self.hypers
        if self.cctype == 'normal':
            return normal_marginal(st[0], st[1], st[2],
                                   h['m'], h['r'], h['s'], h['nu'])
        if self.cctype == 'lognormal':                        # lognormal.py:98-102
            return -st[1] + normal_marginal(st[0], st[1], st[2],
                                            h['m'], h['r'], h['s'], h['nu'])
        if self.cctype == 'bernoulli':
            return bernoulli_marginal(st[0], st[1], h['alpha'], h['beta'])
        if self.cctype == 'poisson':
            return poisson_marginal(st[0], st[1], st[2], h['a'], h['b'])
        if self.cctype == 'exponential':
            return exponential_marginal(st[0], st[1], h['a'], h['b'])
        if self.cctype == 'geometric':
            return geometric_marginal(st[0], st[1], h['a'], h['b'])
        return categorical_marginal(st[0], st[1], h['alpha'])

    def logpdf(self, x, k):
        """dim.py:127-132: NaN -> 0, unknown cluster -> empty model."""
        if math.isnan(x):
            return 0
        st = self.clusters.get(k)
        if st is None:
            st = self._empty()
        return self._predictive(st, x)

    def logpdf_score(self):
        return sum(self._marginal(self.clusters[k]) for k in self.clusters)

    def transition_hypers(self, rng):
        """dim.py:152-174."""
        names = list(self.hypers.keys())
        rng.shuffle(names)
        for name in names:
            logps = []
            for g in self.grids[name]:
                self.hypers[name] = g
                logp_k = 0
                for k in self.clusters:
                    logp_k += self._marginal(self.clusters[k])
                logps.append(logp_k)
            index = log_pflip(logps, rng=rng)
            self.hypers[name] = self.grids[name][index]

    def suffstats(self):
        """dim.py:212-217 (+ the empty aux model at max+1)."""
        def pack(st):
            if self.cctype == 'normal':
                return {'N': st[0], 'sum_x': st[1], 'sum_x_sq': st[2]}
            if self.cctype == 'lognormal':                    # lognormal.py:131-133
                return {'N': st[0], 'sum_log_x': st[1], 'sum_log_x_sq': st[2]}
            if self.cctype == 'bernoulli':                    # bernoulli.py:101-102
                return {'N': st[0], 'x_sum': st[1]}
            if self.cctype == 'poisson':                      # poisson.py:112-114
                return {'N': st[0], 'sum_x': st[1], 'sum_log_fact_x': st[2]}
            if self.cctype in ('exponential', 'geometric'):
                return {'N': st[0], 'sum_x': st[1]}
            return {'N': st[0], 'counts': list(st[1])}
        if not self.clusters:
            return [(0, pack(self._empty()))]
        out = [(k, pack(self.clusters[k])) for k in self.clusters]
        out.append((max(self.clusters) + 1, pack(self._empty())))
        return out


# ----------------------------------------------------------------------------
# mixtures/view.py

class OView(object):
    def __init__(self, X, n_rows, alpha, Zr, rng):
        """view.py:90-106 (row CRP part of the constructor)."""
        self.X = X
        self.rng = rng
        self.crp = OCrp(n_rows, alpha, rng)
        if Zr is None:
            for i in range(n_rows):
                self.crp.incorporate(i, self.crp.simulate_fresh(rng))
        else:
            for i, z in enumerate(Zr):
                self.crp.incorporate(i, z)
        self.dims = {}

    @property
    def Zr(self):
        return self.crp.data

    @property
    def Nk(self):
        return self.crp.counts

    def incorporate_dim(self, dim, reassign=True):
        """view.py:134-141."""
        if reassign:
            dim.bulk_incorporate(self.X, self.Zr)
        self.dims[dim.col] = dim
        return dim.logpdf_score()

    def unincorporate_dim(self, dim):
        del self.dims[dim.col]
        return dim.logpdf_score()

    # rows kernel ---------------------------------------------------------------
    def row_conditional(self, rowid):
        """view.py:393-401: candidate tables, CRP weights, data terms.
        Scoring the member's own cluster replays the reference's
        unincorporate / logpdf / incorporate sequence (view.py:412-422), so
        the sufficient statistics drift exactly as they do there."""
        z = self.Zr[rowid]
        K, logp_crp = crp_gibbs_logps(self.Nk, z, self.crp.alpha)
        logp_data = []
        for k in K:
            terms = []
            for dim in self.dims.values():
                x = self.X[dim.col][rowid]
                if z == k:
                    dim.unincorporate(x, k)
                    lp = dim.logpdf(x, k)
                    dim.incorporate(x, k)
                else:
                    lp = dim.logpdf(x, k)
                terms.append(lp)
            logp_data.append(sum(terms))
        return K, logp_crp, logp_data

    def gibbs_row(self, rowid, trace=None):
        K, logp_crp, logp_data = self.row_conditional(rowid)
        p_cluster = np.add(logp_data, logp_crp)
        z_b = log_pflip(p_cluster, array=K, rng=self.rng)
        if trace is not None:
            trace.append((int(rowid), list(K), [float(v) for v in p_cluster], int(z_b)))
        if self.Zr[rowid] != z_b:
            self.migrate_row(rowid, z_b)

    def migrate_row(self, rowid, k_new):
        """view.py:424-429,174-183,149-172."""
        k_old = self.Zr[rowid]
        for dim in self.dims.values():
            dim.unincorporate(self.X[dim.col][rowid], k_old)
        self.crp.unincorporate(rowid)
        if k_old not in self.Nk:
            for dim in self.dims.values():
                del dim.clusters[k_old]
        self.crp.incorporate(rowid, k_new)
        for dim in self.dims.values():
            dim.incorporate(self.X[dim.col][rowid], k_new)

    # incremental observation (view.py:149-191) ---------------------------------------
    def incorporate(self, rowid, values, cluster=None):
        """view.py:149-172: seat the row (table 0 unless given), add its cells, then one
        Gibbs step for it when the table was not given."""
        k = 0 if cluster is None else cluster
        self.crp.incorporate(rowid, k)
        for dim in self.dims.values():
            dim.incorporate(values[dim.col], k)
        if cluster is None:
            self.transition_rows(rows=[rowid])

    def unincorporate(self, rowid):
        """view.py:174-183."""
        k = self.Zr[rowid]
        for dim in self.dims.values():
            dim.unincorporate(self.X[dim.col][rowid] if rowid < len(self.X[dim.col]) else self._popped[dim.col], k)
        self.crp.unincorporate(rowid)
        if k not in self.Nk:
            for dim in self.dims.values():
                del dim.clusters[k]

    def force_cell(self, rowid, observation):
        """view.py:186-191 (the cell was NaN: nothing to remove)."""
        k = self.Zr[rowid]
        for d, x in observation.items():
            self.dims[d].incorporate(x, k)

    def transition_rows(self, rows=None, trace=None):
        """view.py:247-252."""
        if rows is None:
            rows = list(self.Zr.keys())
        rows = self.rng.permutation(rows)
        for rowid in rows:
            self.gibbs_row(rowid, trace)

    def logpdf_score(self):
        """view.py:257-270."""
        lp_prior = self.crp.logpdf_score()
        lp_likelihood = sum([d.logpdf_score() for d in self.dims.values()])
        return lp_prior + lp_likelihood


# ----------------------------------------------------------------------------
# crosscat/state.py

class OState(object):
    """CrossCat state restricted to normal / categorical columns."""

    KERNELS = ('alpha', 'view_alphas', 'column_params', 'column_hypers',
               'rows', 'columns')

    def __init__(self, X, outputs=None, cctypes=None, distargs=None, Zv=None,
                 Zrv=None, alpha=None, view_alphas=None, hypers=None, Ci=None,
                 rng=None):
        """state.py:47-182."""
        self.rng = np.random.RandomState() if rng is None else rng
        X = np.asarray(X, dtype=float)
        self.n_rows, self.n_cols = X.shape
        self.outputs = list(outputs) if outputs else list(range(self.n_cols))
        self.X = OrderedDict(
            (c, X[:, i].tolist()) for i, c in enumerate(self.outputs))
        self.Ci = [] if Ci is None else Ci
        self.crp = OCrp(self.n_cols, alpha, self.rng)
        if Zv is None:
            for c in self.outputs:
                self.crp.incorporate(c, self.crp.simulate_fresh(self.rng))
        else:
            for c, z in Zv.items():
                self.crp.incorporate(c, z)
        cctypes = cctypes or ['normal'] * self.n_cols
        distargs = distargs or [None] * self.n_cols
        hypers = hypers or [None] * self.n_cols
        view_alphas = view_alphas or {}
        Zrv = Zrv or {}
        self.views = OrderedDict()
        for v in set(self.Zv.values()):
            view = OView(self.X, self.n_rows, view_alphas.get(v), Zrv.get(v),
                         self.rng)
            for c in [o for o in self.outputs if self.Zv[o] == v]:
                i = self.outputs.index(c)
                dim = ODim(c, cctypes[i], distargs[i], hypers[i], self.X[c],
                           self.rng)
                view.incorporate_dim(dim)
            self.views[v] = view
        self.iterations = {}

    @property
    def Zv(self):
        return self.crp.data

    @property
    def Nv(self):
        return self.crp.counts

    def dim_for(self, c):
        return self.views[self.Zv[c]].dims[c]

    # kernels -------------------------------------------------------------------
    def _count(self, name):
        self.iterations[name] = self.iterations.get(name, 0) + 1

    def transition_crp_alpha(self):
        self.crp.transition_alpha(self.rng)
        self._count('alpha')

    def transition_view_alphas(self):
        """state.py:767-772; view.py:231-233 transitions each alpha twice."""
        for v in self.views:
            self.views[v].crp.transition_alpha(self.rng)
            self.views[v].crp.transition_alpha(self.rng)
        self._count('view_alphas')

    def transition_dim_hypers(self, cols=None):
        for c in (cols if cols is not None else self.outputs):
            self.dim_for(c).transition_hypers(self.rng)
        self._count('column_hypers')

    def transition_view_rows(self, trace=None):
        if self.n_rows == 1:
            return
        for v in self.views:
            self.views[v].transition_rows(
                trace=None if trace is None else trace.setdefault(v, []))
        self._count('rows')

    def transition_dims(self, cols=None, trace=None):
        """state.py:804-810."""
        if cols is None:
            cols = self.outputs
        cols = self.rng.permutation(cols)
        for c in cols:
            self.gibbs_dim(c, trace)
        self._count('columns')

    def gibbs_dim(self, col, trace=None):
        """state.py:1050-1123 with m = 1 auxiliary view."""
        dim = self.dim_for(col)
        logp_data = []
        for v in self.views:                                   # :1065-1072
            view = self.views[v]
            logp_data.append(view.incorporate_dim(dim))
            view.unincorporate_dim(dim)
        own = self.Zv[col]
        singleton = self.Nv[own] == 1
        tables = crp_gibbs_tables(self.Nv, own, singleton)     # :1075
        t_aux = tables[len(self.views):]
        aux_views = [OView(self.X, self.n_rows, None, None, self.rng)
                     for _t in t_aux]                          # :1078-1081
        for view in aux_views:
            logp_data.append(view.incorporate_dim(dim))
            view.unincorporate_dim(dim)
        _tables, logp_crp = crp_gibbs_logps(self.Nv, own, self.crp.alpha)
        logp_views = np.add(logp_data, logp_crp)
        avoid = [a for p in self.Ci if col in p for a in p if a != col]
        for a in avoid:                                        # :1098-1102
            index = list(self.views.keys()).index(self.Zv[a])
            logp_views[index] = NEG_INF
        draw = log_pflip(logp_views, rng=self.rng)
        v_sampled = tables[draw]
        if trace is not None:
            trace.append((int(col), list(tables),
                          [float(v) for v in logp_views], int(v_sampled)))
        if own != v_sampled:
            if v_sampled > max(self.views):
                self.views[v_sampled] = aux_views[draw - len(self.views)]
            delete = self.Nv[own] == 1                         # :1125-1145
            self.views[v_sampled].incorporate_dim(dim)
            self.crp.unincorporate(col)
            self.crp.incorporate(col, v_sampled)
            if delete:
                del self.views[own]
        else:
            self.views[own].incorporate_dim(dim)

    # incremental structure changes (state.py:187-320) ---------------------------------
    CRP_ID_VIEW = 10 ** 7

    def incorporate(self, rowid, observation):
        """state.py:255-283."""
        if rowid != self.n_rows:
            raise ValueError('Only contiguous rowids supported: %d' % (rowid,))
        tokens = {self.CRP_ID_VIEW + v: v for v in self.views}
        if not all(q in self.outputs or q in tokens for q in observation):
            raise ValueError('Invalid observation: %s' % observation)
        if any(math.isnan(v) for v in observation.values()):
            raise ValueError('Cannot incorporate nan: %s.' % observation)
        for c in self.outputs:
            self.X[c].append(observation.get(c, float('nan')))
        self.n_rows += 1
        for v in self.views:
            view = self.views[v]
            values = {d: self.X[d][rowid] for d in view.dims}
            view.incorporate(rowid, values, observation.get(self.CRP_ID_VIEW + v))

    def unincorporate(self, rowid):
        """state.py:285-299."""
        if rowid != self.n_rows - 1:
            raise ValueError('Only last rowid may be unincorporated.')
        if self.n_rows == 1:
            raise ValueError('Cannot unincorporate last rowid.')
        popped = {c: self.X[c].pop() for c in self.outputs}
        self.n_rows -= 1
        for v in self.views:
            self.views[v]._popped = popped
            self.views[v].unincorporate(rowid)

    def force_cell(self, rowid, observation):
        """state.py:302-315."""
        if not 0 <= rowid < self.n_rows:
            raise ValueError('Force observation requires existing rowid.')
        if not all(math.isnan(self.X[c][rowid]) for c in observation):
            raise ValueError('Force observations requires NaN cells.')
        for col, value in observation.items():
            self.X[col][rowid] = value
        for col, value in observation.items():
            self.views[self.Zv[col]].force_cell(rowid, {col: value})

    def incorporate_dim(self, T, outputs, cctype=None, distargs=None, v=None):
        """state.py:187-236."""
        if len(T) != self.n_rows:
            raise ValueError('%d rows are required, received: %d.' % (self.n_rows, len(T)))
        col = outputs[0]
        if col in self.outputs:
            raise ValueError('Specified outputs already exist: %s, %s.' % (outputs, self.outputs))
        self.X[col] = [float(x) for x in T]
        self.outputs = self.outputs + [col]
        self.n_cols += 1
        transition = [col] if v is None else []
        v_add = 0 if v is None else v
        if v_add in self.views:
            view = self.views[v_add]
        else:
            view = OView(self.X, self.n_rows, None, None, self.rng)
            self.views[v_add] = view
        dim = ODim(col, cctype, distargs, None, self.X[col], self.rng)
        view.incorporate_dim(dim)
        self.crp.incorporate(col, v_add)
        self.transition_dims(cols=transition)
        self.transition_dim_hypers(cols=[col])

    def unincorporate_dim(self, col):
        """state.py:238-253."""
        if self.n_cols == 1:
            raise ValueError('State has only one dim, cannot unincorporate.')
        if col not in self.outputs:
            raise ValueError('col does not exist: %s' % (col,))
        v_del = self.Zv[col]
        delete = self.Nv[v_del] == 1
        del self.views[v_del].dims[col]
        self.crp.unincorporate(col)
        if delete:
            del self.views[v_del]
        del self.X[col]
        self.outputs = [i for i in self.outputs if i != col]
        self.n_cols -= 1

    def transition(self, N=1, kernels=None, trace=None):
        """state.py:727-761,866-905 (N iterations, kernels in caller order)."""
        if kernels is None:
            kernels = list(self.KERNELS)
        table = {
            'alpha': self.transition_crp_alpha,
            'view_alphas': self.transition_view_alphas,
            'column_params': lambda: self._count('column_params'),
            'column_hypers': self.transition_dim_hypers,
            'rows': lambda: self.transition_view_rows(
                None if trace is None else trace.setdefault('rows', {})),
            'columns': lambda: self.transition_dims(
                trace=None if trace is None else trace.setdefault('columns', [])),
        }
        for _ in range(N):
            for k in kernels:
                table[k]()

    # queries -------------------------------------------------------------------
    def logpdf_score(self):
        """state.py:384-387."""
        logp_crp = self.crp.logpdf_score()
        logp_views = sum(v.logpdf_score() for v in self.views.values())
        return logp_crp + logp_views

    def logpdf(self, rowid, targets, constraints=None):
        """sampling.py:37-49,70-82,102-107."""
        constraints = constraints or {}
        by_view_t, by_view_c = OrderedDict(), {}
        for c in targets:
            by_view_t.setdefault(self.Zv[c], {})[c] = targets[c]
        for c in constraints:
            by_view_c.setdefault(self.Zv[c], {})[c] = constraints[c]
        total = []
        for v in by_view_t:
            total.append(self._view_logpdf(
                self.views[v], rowid, by_view_t[v], by_view_c.get(v, {})))
        return sum(total)

    @staticmethod
    def _row_logpdf(view, cells, k):
        return sum(view.dims[c].logpdf(x, k) for c, x in cells.items())

    def _view_logpdf(self, view, rowid, targets, constraints):
        if rowid is not None and 0 <= rowid < self.n_rows:
            return self._row_logpdf(view, targets, view.Zr[rowid])
        Nk = view.Nk
        K = sorted(Nk) + [max(Nk) + 1]
        lp_crp = [crp_predictive(k, self.n_rows, Nk, view.crp.alpha) for k in K]
        lp_con = [self._row_logpdf(view, constraints, k) for k in K]
        if all(np.isinf(lp_con)):
            raise ValueError('Zero density constraints: %s' % (constraints,))
        lp_cluster = log_normalize(np.add(lp_crp, lp_con))
        lp_tgt = [self._row_logpdf(view, targets, k) for k in K]
        return logsumexp(np.add(lp_cluster, lp_tgt))

    def relevance_probability(self, rowid_target, rowid_query, col, hypotheticals=None):
        """state.py:599-629: hypothetical rows are incorporated into the column's view (one Gibbs step each),
        the answer is read off the partition, and they are removed again."""
        view = self.views[self.Zv[col]]
        hyps = []
        for h in (hypotheticals or []):
            q = {d: h.get(d, float('nan')) for d in view.dims}
            if not all(math.isnan(v) for v in q.values()):
                hyps.append(q)
        rowid_hyp = list(range(self.n_rows, self.n_rows + len(hyps)))
        for rowid, query in zip(rowid_hyp, hyps):
            for d in view.dims:
                self.X[d].append(query[d])
            view.incorporate(rowid, query)
        rowid_all = list(rowid_query) + rowid_hyp
        relevance = all(view.Zr[rowid_target] == view.Zr[rq] for rq in rowid_all) if rowid_all else 0
        for rowid in reversed(rowid_hyp):
            view._popped = {d: self.X[d].pop() for d in view.dims}
            view.unincorporate(rowid)
        return int(relevance)

    def dependence_probability(self, c0, c1):
        return float(self.Zv[c0] == self.Zv[c1])

    # serialisation -------------------------------------------------------------
    def to_metadata(self):
        """Subset of state.py:1189-1240 that defines the chain."""
        dims = [self.dim_for(c) for c in self.outputs]
        return {
            'outputs': list(self.outputs),
            'alpha': float(self.crp.alpha),
            'Zv': [(int(c), int(v)) for c, v in self.Zv.items()],
            'cctypes': [d.cctype for d in dims],
            'distargs': [dict(d.distargs) for d in dims],
            'hypers': [{h: float(x) for h, x in d.hypers.items()} for d in dims],
            'suffstats': [d.suffstats() for d in dims],
            'Zrv': [(int(v), [int(view.Zr[i]) for i in sorted(view.Zr)])
                    for v, view in self.views.items()],
            'view_alphas': [(int(v), float(view.crp.alpha))
                            for v, view in self.views.items()],
            'row_order': [(int(v), [int(i) for i in view.Zr])
                          for v, view in self.views.items()],
            'iterations': dict(self.iterations),
        }
