"""Host-side scalar math with the reference's exact semantics.

Used only where the host must consume a chain's numpy RandomState exactly like
the reference does (state initialisation, src/crosscat/state.py:84-159): the
draw-to-index mapping of `gu.log_pflip` (src/utils/general.py:123-139) and the
sequential CRP prior simulation (src/primitives/crp.py:71-81).
"""
from __future__ import annotations

import math

import numpy as np


def logsumexp(array):
    """src/utils/general.py:141-157."""
    if len(array) == 0:
        return float('-inf')
    m = max(array)
    if math.isinf(m) and min(array) != -m and all(not math.isnan(a) for a in array):
        return m
    return m + math.log(sum(math.exp(a - m) for a in array))


def log_linspace(a, b, n):
    """src/utils/general.py:203-205."""
    return np.exp(np.linspace(math.log(a), math.log(b), n))


def logmeanexp(array):
    """src/utils/general.py:159-182."""
    if len(array) == 0:
        return float('-inf')
    noninfs = [a for a in array if not a == float('-inf')]
    return logsumexp(noninfs) - math.log(len(array))


def logmeanexp_weighted(log_A, log_W):
    """src/utils/general.py:184-201."""
    assert len(log_W) == len(log_A)
    return logsumexp([w + a for w, a in zip(log_W, log_A)]) - logsumexp(log_W)


def log_pflip(logp, array=None, size=None, rng=None):
    """src/utils/general.py:123-139: one random_sample() per draw unless the
    candidate list has length 1."""
    p = np.exp(np.subtract(logp, logsumexp(logp)))
    if array is None:
        array = list(range(len(p)))
    if len(p) == 1:
        return array[0] if size is None else [array[0]] * size
    p = np.asarray(p, dtype=float) / sum(p)
    return rng.choice(array, size=size, p=p)


def crp_simulate_exact(n, alpha, rng):
    """Sequential CRP(alpha) prior draw of n members with the reference's
    arithmetic and stream use (crp.py:71-81): member 0 draws nothing."""
    counts = {}
    N = 0
    out = []
    for _ in range(n):
        if counts:
            K = sorted(counts) + [max(counts) + 1]
        else:
            K = [0]
        logps = [math.log(counts.get(x, alpha)) - math.log(N + alpha) for x in K]
        x = int(log_pflip(logps, array=K, rng=rng))
        counts[x] = counts.get(x, 0) + 1
        N += 1
        out.append(x)
    return np.asarray(out, dtype=np.int64)


def crp_simulate_fast(n, alpha, rng):
    """A CRP(alpha) draw by the equivalent 'join a uniformly chosen earlier
    member' construction (vectorised); same law, different stream use."""
    i = np.arange(n)
    new = rng.random_sample(n) < alpha / (i + alpha)
    new[0] = True
    parent = (rng.random_sample(n) * np.maximum(i, 1)).astype(np.int64)
    parent = np.minimum(parent, np.maximum(i - 1, 0))
    parent[new] = i[new]
    while True:
        nxt = parent[parent]
        if np.array_equal(nxt, parent):
            break
        parent = nxt
    # roots are first members, so sorted roots are already in order of appearance
    _, inv = np.unique(parent, return_inverse=True)
    return inv.astype(np.int64)
