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


# ---------------------------------------------------------------------------------------------
# Column-CRP dependence / independence constraints (State(Cd=..., Ci=...), src/crosscat/state.py:96-113,370-379).
def pflip(p, rng):
    """src/utils/general.py:127-139 with array = range(len(p)): no draw for a single candidate."""
    if len(p) == 1:
        return 0
    p = np.asarray(p, dtype=float) / sum(p)
    return int(rng.choice(list(range(len(p))), p=p))


def validate_dependency_constraints(N, Cd, Ci):
    """src/utils/validation.py:34-66."""
    counts = {}
    for block in Cd:
        if len(block) == 1:
            raise ValueError('Single customer in dependency constraint.')
        for col in block:
            if N is not None and N <= col:
                raise ValueError('Dependence customer out of range.')
            counts[col] = counts.get(col, 0) + 1
            if counts[col] > 1:
                raise ValueError('Multiple customer dependencies.')
        for pair in Ci:
            if pair[0] in block and pair[1] in block:
                raise ValueError('Contradictory customer independence.')
    for pair in Ci:
        if len(pair) != 2:
            raise ValueError('Independencies require two customers.')
        if N is not None and (N <= pair[0] or N <= pair[1]):
            raise ValueError('Independence customer of out range.')
        if pair[0] == pair[1]:
            raise ValueError('Independency specified for same customer.')
    return True


def _compatible(Ci, a, b):
    """src/utils/validation.py:80-89 without row constraints (Rd / Ri are refused by State)."""
    return not ((a, b) in Ci or (b, a) in Ci)


def validate_crp_constrained_partition(Zv, Cd, Ci):
    """src/utils/validation.py:22-32: Zv maps customer -> table."""
    import itertools
    Ci = [tuple(p) for p in Ci]
    valid = True
    for block in Cd:
        valid = valid and all(Zv[block[0]] == Zv[b] for b in block)
        for a, b in itertools.combinations(block, 2):
            valid = valid and _compatible(Ci, a, b)
    for a, b in Ci:
        valid = valid and not Zv[a] == Zv[b]
    return valid


def simulate_crp(N, alpha, rng):
    """src/utils/general.py:207-246 (note its normaliser i - 1 + alpha; pflip renormalises)."""
    alpha = float(alpha)
    partition = [0] * N
    Nk = [1]
    for i in range(1, N):
        K = len(Nk)
        ps = np.zeros(K + 1)
        for k in range(K):
            ps[k] = float(Nk[k])
        ps[K] = alpha
        ps /= (float(i) - 1 + alpha)
        a = pflip(ps, rng)
        if a == K:
            Nk.append(1)
        else:
            Nk[a] += 1
        partition[i] = a
    return partition


def simulate_crp_constrained(N, alpha, Cd, Ci, rng):
    """src/utils/general.py:248-291: friends are seated together, enemies never share a table."""
    validate_dependency_constraints(N, Cd, Ci)
    Ci = [tuple(p) for p in Ci]
    Z = [-1] * N
    friends = {col: block for block in Cd for col in block}
    for cust in range(N):
        if Z[cust] > -1:
            continue
        prob_table = [0] * (max(Z) + 1)
        for t in range(max(Z) + 1):
            t_custs = [i for i, z in enumerate(Z) if z == t]
            prob_table[t] = len(t_custs)
            for tc in t_custs:
                for f in friends.get(cust, [cust]):
                    if not _compatible(Ci, f, tc):
                        prob_table[t] = 0
                        break
        prob_table.append(alpha)
        a = pflip(prob_table, rng)
        for f in friends.get(cust, [cust]):
            Z[f] = a
    assert validate_crp_constrained_partition(Z, Cd, Ci)
    return Z


def crp_constrained_num_effective(N, Cd):
    """src/utils/general.py:358-370."""
    return (N - sum(len(block) for block in Cd)) + len(Cd)


def simulate_crp_constrained_dependent(N, alpha, Cd, rng):
    """src/utils/general.py:293-332: every clique of friends is one customer of a plain CRP."""
    validate_dependency_constraints(N, Cd, [])
    crp_block = simulate_crp(crp_constrained_num_effective(N, Cd), alpha, rng)
    partition = [-1] * N
    for i, block in enumerate(Cd):
        for customer in block:
            partition[customer] = crp_block[i]
    unused = iter(crp_block[len(Cd):])
    for customer, a in enumerate(partition):
        if a == -1:
            partition[customer] = next(unused)
    return partition


def logp_crp(N, Nk, alpha):
    """src/utils/general.py:88-94."""
    return len(Nk) * math.log(alpha) + sum(math.lgamma(c) for c in Nk) + math.lgamma(alpha) - math.lgamma(N + alpha)


def logp_crp_constrained_dependent(Z, alpha, Cd):
    """src/utils/general.py:335-355: Z maps customer -> table."""
    if not validate_crp_constrained_partition(Z, Cd, []):
        return -float('inf')
    counts = {}
    seen = set()
    for block in Cd:
        seen.update(block)
        counts[Z[block[0]]] = counts.get(Z[block[0]], 0) + 1
    for customer in Z:
        if customer not in seen:
            counts[Z[customer]] = counts.get(Z[customer], 0) + 1
    return logp_crp(crp_constrained_num_effective(len(Z), Cd), list(counts.values()), alpha)
