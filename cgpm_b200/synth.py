"""Synthetic CrossCat tables and partitions (vectorised).

Same distributions as the reference's generator `cgpm.utils.test.gen_data_table`
(src/utils/test.py:30-105): columns are split into views, every view has its
own row partition; a normal column is N(5 * separation * cluster, 1)
(test.py:135-145), a categorical column draws per-cluster theta from
Dirichlet((1 - separation) * 1_k) and x ~ Categorical(theta) (test.py:260-277).
The reference generator is a per-cell Python loop; this one is numpy so the
benchmark shapes (1e5 .. 1e6 rows) are generated in seconds.
"""
from __future__ import annotations

import numpy as np


def gen_table(n_rows, cctypes, distargs, n_views=2, clusters_per_view=3, separation=0.9,
              seed=0, quantize_bits=None, nan_frac=0.0):
    """Returns X [n_rows, n_cols] float64, Zv [n_cols], Zrv [n_views][n_rows]."""
    rng = np.random.RandomState(seed)
    C = len(cctypes)
    Zv = np.sort(rng.randint(0, n_views, size=C))
    Zv[:min(n_views, C)] = np.arange(min(n_views, C))          # every view non-empty
    Zv = np.sort(Zv)
    Zrv = []
    for _ in range(n_views):
        w = np.arange(1, clusters_per_view + 1, dtype=float)
        Zrv.append(rng.choice(clusters_per_view, size=n_rows, p=w / w.sum()))
    X = np.empty((n_rows, C))
    for c in range(C):
        z = Zrv[Zv[c]]
        if cctypes[c] == 'normal':
            x = rng.normal(loc=5.0 * separation * z, scale=1.0)
            if quantize_bits is not None:
                q = float(1 << quantize_bits)
                x = np.round(x * q) / q
            X[:, c] = x
        elif cctypes[c] == 'categorical':
            k = int(distargs[c]['k'])
            theta = rng.dirichlet(np.ones(k) * max(1.0 - separation, 1e-3), size=clusters_per_view)
            cdf = np.cumsum(theta, axis=1)
            u = rng.random_sample(n_rows)
            X[:, c] = np.minimum((u[:, None] > cdf[z]).sum(axis=1), k - 1)
        else:
            raise ValueError('unsupported cctype %s' % cctypes[c])
    if nan_frac > 0:
        X[rng.random_sample(X.shape) < nan_frac] = np.nan
    return X, Zv, Zrv


def crp_partition(n, alpha, rng):
    """A draw from CRP(alpha) over n items, labelled in order of appearance.
    Uses the equivalent 'join a uniformly chosen earlier customer' construction
    (vectorised pointer jumping), not the reference's sequential inverse-CDF."""
    i = np.arange(n)
    new = rng.random_sample(n) < alpha / (i + alpha)
    new[0] = True
    parent = np.where(new, i, (rng.random_sample(n) * np.maximum(i, 1)).astype(np.int64))
    parent = np.minimum(parent, np.maximum(i - 1, 0))
    parent[new] = i[new]
    while True:
        nxt = parent[parent]
        if np.array_equal(nxt, parent):
            break
        parent = nxt
    _, labels = np.unique(parent, return_inverse=True)
    return labels.astype(np.int64)
