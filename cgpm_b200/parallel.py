"""Chain sharding across GPUs (one process per GPU, torch.distributed).

Chains of an Engine never interact during inference (src/crosscat/engine.py:78-85:
each State is transitioned in isolation), so `num_states` chains are block-
partitioned over the ranks, the table is replicated, and no collective runs
inside `transition`.  Collectives (NCCL on GPUs, gloo in the CPU tests) only
gather per-chain results: scores and diagnostics (all_gather), ensemble
log-densities (all_gather + logmeanexp, engine.py:361-371) and pairwise
dependence probabilities (all_reduce of co-assignment counts, state.py:546-554).
This replaces the reference's fork + pipe fan-out (src/utils/parallel_map.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import hostmath as hm


def shard_range(num_states, world, rank):
    """Contiguous block of chains owned by `rank`: sizes differ by at most one."""
    base, extra = divmod(int(num_states), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(num_states, world):
    return [shard_range(num_states, world, r)[1] - shard_range(num_states, world, r)[0] for r in range(world)]


def all_gather_chains(local, num_states, group=None):
    """local: tensor [n_local, ...] of per-chain values in shard order.  Returns the
    [num_states, ...] tensor on every rank (chains in global order)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(num_states, world)
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def all_reduce_sum(t, group=None):
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_gather_objects(local_list, group=None):
    """Python objects (one list per rank, in shard order) -> the concatenated list on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return list(local_list)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, list(local_list), group=group)
    return [x for part in parts for x in part]


def weighted_integrate(logpdfs, weights=None):
    """Engine._likelihood_weighted_integrate (src/crosscat/engine.py:361-371) over the chains of all ranks:
    logmeanexp of the chains' log-densities, weighted by the log-density of the constraints when given."""
    logpdfs = [float(x) for x in logpdfs]
    if weights is None:
        return hm.logmeanexp(logpdfs)
    return hm.logmeanexp_weighted(logpdfs, [float(w) for w in weights])


class ShardedEngine(object):
    """Engine whose chains are spread over the ranks of the default process group.
    Every rank constructs it with the same arguments; rank r holds chains
    shard_range(num_states, world, r).  Chain seeds are the ones the single-process
    Engine would draw (engine.py:357-359), so sharding does not change the chains."""

    def __init__(self, X, num_states=1, rng=None, **kwargs):
        from .engine import Engine, gen_rng
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.num_states_total = int(num_states)
        rng = gen_rng(1) if rng is None else rng
        seeds = rng.randint(low=1, high=2 ** 32 - 1, size=num_states)
        lo, hi = shard_range(num_states, self.world, self.rank)
        self.lo, self.hi = lo, hi

        class _Seeded(object):               # hands the local Engine exactly its shard of the seeds
            def __init__(self, vals):
                self.vals = list(vals)

            def randint(self, low=None, high=None, size=None):
                out, self.vals = self.vals[:size], self.vals[size:]
                if len(out) < size:
                    out = out + list(np.random.RandomState(len(out) + 1).randint(low, high, size=size - len(out)))
                return np.asarray(out)
        self.local = Engine(X, num_states=hi - lo, rng=_Seeded(seeds[lo:hi]), **kwargs)

    def transition(self, **kw):
        """No data-path collective: every rank advances its own chains."""
        self.local.transition(**kw)

    def _device(self):
        return self.local._ch.dev.device

    def logpdf_score(self):
        loc = torch.as_tensor(self.local.logpdf_score(), dtype=torch.float64, device=self._device())
        return all_gather_chains(loc, self.num_states_total).cpu().numpy().tolist()

    def logpdf_bulk(self, rowids, targets_list, constraints_list=None):
        loc = torch.as_tensor(np.asarray(self.local.logpdf_bulk(rowids, targets_list, constraints_list)).reshape(
            self.hi - self.lo, len(rowids)), dtype=torch.float64, device=self._device())
        return all_gather_chains(loc, self.num_states_total).cpu().numpy()

    def logpdf_ensemble(self, rowid, targets):
        """Engine._likelihood_weighted_integrate without constraints: logmeanexp over all chains."""
        lp = self.logpdf_bulk([rowid], [targets])[:, 0]
        return hm.logmeanexp([float(x) for x in lp])

    def dependence_probability_pairwise(self):
        C = self.local._ch.C
        acc = torch.zeros((C, C), dtype=torch.float64, device=self._device())
        for D in self.local.dependence_probability_pairwise():
            acc += torch.as_tensor(D, dtype=torch.float64, device=self._device())
        all_reduce_sum(acc)
        return (acc / self.num_states_total).cpu().numpy()

    # ---- the rest of the Engine surface (src/crosscat/engine.py:192-281,361-435) over all ranks ----------------
    def num_states(self):
        return self.num_states_total

    def logpdf(self, rowid, targets, constraints=None):
        loc = torch.as_tensor(self.local.logpdf(rowid, targets, constraints), dtype=torch.float64, device=self._device())
        return all_gather_chains(loc, self.num_states_total).cpu().numpy().tolist()

    def logpdf_likelihood_weighted(self, rowid, targets, constraints=None):
        """The ensemble estimate of engine.py:361-371: chains weighted by the density they give the constraints."""
        lp = self.logpdf(rowid, targets, constraints)
        if constraints:
            return weighted_integrate(lp, self.logpdf(rowid, constraints))
        return weighted_integrate(lp)

    def simulate(self, rowid, targets, constraints=None, N=None):
        return all_gather_objects(self.local.simulate(rowid, targets, constraints, N=N))

    def simulate_bulk(self, rowids, targets_list, constraints_list=None, Ns=None):
        """[num_states][Q] lists of sample dicts, chains in global order (engine.py:241-251)."""
        return all_gather_objects(self.local.simulate_bulk(rowids, targets_list, constraints_list, Ns=Ns))

    def logpdf_bulk_arrays(self, target_cols, target_vals, constraint_cols=None, constraint_vals=None):
        """[num_states, Q] array, chains in global order: Engine.logpdf_bulk_arrays of every rank's shard."""
        loc = self.local.logpdf_bulk_arrays(target_cols, target_vals, constraint_cols, constraint_vals)
        loc = torch.as_tensor(np.ascontiguousarray(loc), dtype=torch.float64, device=self._device())
        return all_gather_chains(loc, self.num_states_total).cpu().numpy()

    def simulate_bulk_arrays(self, Q, target_cols, constraint_cols=None, constraint_vals=None, N=1):
        """[num_states, Q, N, len(target_cols)] array of predictive samples, chains in global order."""
        loc = self.local.simulate_bulk_arrays(Q, target_cols, constraint_cols, constraint_vals, N=N)
        loc = torch.as_tensor(np.ascontiguousarray(loc), dtype=torch.float64, device=self._device())
        return all_gather_chains(loc, self.num_states_total).cpu().numpy()

    def dependence_probability(self, col0, col1):
        loc = torch.as_tensor(self.local.dependence_probability(col0, col1), dtype=torch.float64, device=self._device())
        return all_gather_chains(loc, self.num_states_total).cpu().numpy().tolist()

    def to_metadata(self):
        """engine.py:394-401 with the states of every rank, in global order (the same dict on every rank)."""
        md = self.local.to_metadata()
        md['states'] = all_gather_objects(md['states'])
        return md

    def get_state(self, index):
        """The State object of chain `index`, rebuilt on every rank from its owner's metadata."""
        from .engine import State, gen_rng, _chain_metadata
        box = [None]
        owner = [r for r in range(self.world) if shard_range(self.num_states_total, self.world, r)[0] <= index
                 < shard_range(self.num_states_total, self.world, r)[1]][0]
        if self.rank == owner:
            box[0] = _chain_metadata(self.local._ch, index - self.lo, X_list=self.local.X.tolist())
        if dist.is_available() and dist.is_initialized():
            dist.broadcast_object_list(box, src=owner)
        return State.from_metadata(box[0], rng=gen_rng(index), **self.local._opts_for_state())
