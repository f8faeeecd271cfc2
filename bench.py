#!/usr/bin/env python
"""Benchmark of the CrossCat Gibbs hot path (BASELINE.json metric: cell-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one `Engine.transition(N=1, kernels=['rows','columns'])` over all chains, driven through the
public API only (cgpm_b200.engine.Engine / cgpm_b200.parallel.ShardedEngine).

  N = 1   BASELINE.json configs[1]: 100k rows x 50 normal columns, 64 chains on one B200.
  N > 1   BASELINE.json configs[2]: 1M rows x 200 mixed normal / categorical columns, 64 chains per GPU
          (weak scaling; chains never interact, engine.py:78-85), ShardedEngine over NCCL -- the only
          collective gathers the per-chain scores.

Chains start from supplied partitions Zv ~ CRP(1), Zrv ~ CRP(1), alpha = view_alphas = 1 and
run a stated number of burn-in sweeps before the W warm-up and K timed steps.  One JSON line is printed by
rank 0; besides the contract's keys it carries `roofline` (HBM), `roofline_compute` (FP64 pipe), the
reference's CPU Engine timed beside it (`cpu_baseline`) and, at N = 1, an `extra` block with short runs of the
other named configs (configs[2] shard, configs[3] column stress, configs[4] query batch, a K sweep).

--impl reference times the reference's own CPU Engine (oracle/_ref, the py3-shimmed probcomp/cgpm;
multiprocess fan-out over the host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KERNELS = ['rows', 'columns']
C1 = dict(name='configs[1]', rows=100000, cols=50, chains=64, mixed=False, views=4, clusters=4, kcap=256, vcap=24)
C2 = dict(name='configs[2]', rows=1000000, cols=200, chains=64, mixed=True, views=5, clusters=4, kcap=128, vcap=16,
          hypers=True, ref_rows=200)
REF_ROWS = 1000
FP64_PEAK_FLOPS = 148 * 64 * 2 * 1.965e9          # 64 DFMA lanes per SM and clock (B200 vector FP64, 37 TFLOP/s)
FLOPS_PER_CELL_EVAL = 2 * 60                     # one Student-t term: ~60 FP64 instructions (log kernel 35, the rest 25)


def column_types(C, mixed):
    """configs[1]: all normal.  configs[2]/[3] mix: every other column categorical, k cycling through 2, 4, 8, 16."""
    if not mixed:
        return ['normal'] * C, [None] * C
    cct, da = [], []
    for c in range(C):
        if c % 2:
            cct.append('categorical'); da.append({'k': [2, 4, 8, 16][(c // 2) % 4]})
        else:
            cct.append('normal'); da.append(None)
    return cct, da


def make_table(R, C, views, clusters, mixed=False, seed=0):
    from cgpm_b200 import synth
    cct, da = column_types(C, mixed)
    X, _, _ = synth.gen_table(R, cct, da, n_views=views, clusters_per_view=clusters, seed=seed)
    return X, cct, da


def supplied_hypers(X, cct):
    """Column hyperparameters for workloads that do not run the column_hypers kernel (configs[2]): weakly informative
    and on the scale of the data -- normal: m = column mean, r = 1, s = column variance, nu = 1; categorical: alpha = 1.
    Without them each chain draws every hyperparameter once from its grid (dim.py:181-183: up to R for r, nu and the
    Dirichlet alpha) and keeps it for the whole run; a column that draws a very tight variance prior then prefers -- and
    accepts -- auxiliary views of the order of R clusters, which no practitioner's run would keep."""
    import numpy as np
    out = []
    for c, t in enumerate(cct):
        if t == 'normal':
            col = X[:, c]
            out.append({'m': float(np.nanmean(col)), 'r': 1.0, 's': float(np.nanvar(col)) or 1.0, 'nu': 1.0})
        else:
            out.append({'alpha': 1.0})
    return out


def initial_partitions(R, C, seed):
    """Zv ~ CRP(1) over columns, Zrv ~ CRP(1) over rows per view, alphas = 1."""
    import numpy as np
    from cgpm_b200 import hostmath as hm
    rng = np.random.RandomState(seed)
    zv = hm.crp_simulate_fast(C, 1.0, rng)
    Zv = {c: int(z) for c, z in enumerate(zv)}
    Zrv = {int(v): hm.crp_simulate_fast(R, 1.0, rng) for v in set(Zv.values())}
    return Zv, Zrv


def src_sha16():
    h = hashlib.sha256()
    for fn in ('rows.cu', 'rows_common.cuh', 'cc_common.cuh'):
        with open(os.path.join(ROOT, 'cgpm_b200', 'csrc', fn), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------- reference arm
def run_reference(cfg, R, chains, steps, warmup, burnin):
    """The reference Engine (multiprocess) on an R-row sample of the workload; returns a dict."""
    import numpy as np
    ref = os.path.join(ROOT, 'oracle', '_ref')
    kind = 'reference'
    cores = os.cpu_count() or 1
    S = max(1, min(chains, cores))
    X, cct, da = make_table(R, cfg['cols'], cfg['views'], cfg['clusters'], cfg['mixed'])
    Zv, Zrv = initial_partitions(R, cfg['cols'], 12345)
    hyp = supplied_hypers(X, cct) if cfg.get('hypers') else None
    if os.path.exists(os.path.join(ref, 'cgpm', '.built')):
        sys.path.insert(0, ref)
        from cgpm.crosscat.engine import Engine
        from cgpm.utils import general as gu
        eng = Engine(X, num_states=S, rng=gu.gen_rng(1), multiprocess=1, cctypes=cct, distargs=da,
                     Zv=Zv, Zrv={v: list(map(int, z)) for v, z in Zrv.items()}, alpha=1.0,
                     view_alphas={v: 1.0 for v in Zrv}, hypers=hyp)
        step = lambda: eng.transition(N=1, kernels=KERNELS, progress=False, multiprocess=1)
    else:
        kind = 'port'
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import crosscat_oracle as co
        S = 1
        states = [co.OState(X, cctypes=cct, distargs=da, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv},
                            hypers=hyp, rng=np.random.RandomState(2 + s)) for s in range(S)]
        step = lambda: [o.transition(N=1, kernels=KERNELS) for o in states]
    for _ in range(burnin + warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = time.time() - t0
    val = R * cfg['cols'] * S * steps / dt
    return dict(value=val, seconds=dt, cores=min(S, cores) if kind == 'reference' else 1, kind=kind,
                sample='%s sample: %d rows x %d %s cols x %d chains, %d burn-in + %d warm-up + %d timed step(s) of %s, '
                       'supplied CRP(1) init%s; host has %d cores'
                       % (cfg['name'], R, cfg['cols'], 'mixed' if cfg['mixed'] else 'normal', S, burnin, warmup, steps,
                          '+'.join(KERNELS), ' and hypers' if hyp else '', cores), chains=S, ms_per_step=1e3 * dt / steps,
                rows=R)


def reference_main(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = C1 if args.gpus <= 1 else C2
    rows = args.ref_rows or cfg.get('ref_rows', REF_ROWS)
    r = run_reference(cfg, rows, cfg['chains'], args.steps, args.warmup, min(args.burnin, 3))
    line = {
        'impl': 'reference', 'metric': 'crosscat_gibbs_cell_updates_per_sec', 'value': r['value'],
        'unit': 'cell-updates/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s sample: %d rows x %d cols, %d chains, kernels=%s (CPU reference Engine, multiprocess)'
                               % (cfg['name'], r['rows'], cfg['cols'], r['chains'], KERNELS)},
        'cpu_baseline': {'value': r['value'], 'unit': 'cell-updates/s', 'cores': r['cores'], 'kind': r['kind'],
                         'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------- extras (N = 1, rank 0)
def view_stats(eng):
    import numpy as np
    dev = eng._ch.dev
    nv = dev.t['nv'].cpu().numpy(); nclu = dev.t['nclu'].cpu().numpy()
    k = nclu[nv > 0]
    return {'views_total': int((nv > 0).sum()), 'clusters_per_view_mean': float(k.mean()), 'clusters_per_view_max': int(k.max())}


def step_stats(ms, steps):
    """min / median / max over the timed steps of one kernel's milliseconds (calls grouped per step in launch order).
    A step is as slow as its slowest (chain, view): one view whose rows move at most steps can set the rows phase for
    as long as it exists, so the spread is reported beside the mean."""
    per = max(1, len(ms) // max(steps, 1))
    tot = sorted(sum(ms[i:i + per]) for i in range(0, len(ms) - per + 1, per))
    if not tot:
        return None
    return {'min': tot[0], 'median': tot[len(tot) // 2], 'max': tot[-1]}


def timed_steps(eng, steps, kernels):
    import torch
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        eng.transition(N=1, kernels=kernels, progress=False)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extra_config2(budget_s):
    """configs[2] shape on one GPU with a reduced chain count (the full 64-chain shard is what --gpus N > 1 runs)."""
    from cgpm_b200.engine import Engine, gen_rng
    t0 = time.time()
    cfg = C2
    S = 8
    X, cct, da = make_table(cfg['rows'], cfg['cols'], cfg['views'], cfg['clusters'], True)
    Zv, Zrv = initial_partitions(cfg['rows'], cfg['cols'], 4242)
    eng = Engine(X, num_states=S, rng=gen_rng(11), cctypes=cct, distargs=da, Zv=Zv, Zrv=Zrv, alpha=1.0,
                 view_alphas={v: 1.0 for v in Zrv}, hypers=supplied_hypers(X, cct), stream='philox', Kcap=cfg['kcap'],
                 Vcap=cfg['vcap'])
    burn = 0
    while burn < 3 and time.time() - t0 < 0.5 * budget_s:
        eng.transition(N=1, kernels=KERNELS, progress=False)
        burn += 1
    eng.start_kernel_timing()
    ms = timed_steps(eng, 2, KERNELS)
    km = eng.stop_kernel_timing()
    out = {'workload': '1M rows x 200 mixed cols (100 normal, 100 categorical k in {2,4,8,16}), %d chains on 1 GPU, kernels=%s'
                       % (S, KERNELS), 'burn_in_sweeps': burn, 'steps': 2, 'ms_per_step': ms,
           'cell_updates_per_s': cfg['rows'] * cfg['cols'] * S / (ms * 1e-3),
           'kernel_ms_per_step': {k: sum(v) / 2 for k, v in km.items()}, 'setup_s': None}
    out.update(view_stats(eng))
    out['wall_s'] = time.time() - t0
    return out


def extra_config3():
    """configs[3]: 10k rows x 5 000 normal columns, State.transition_dims dominant."""
    from cgpm_b200.engine import Engine, gen_rng
    t0 = time.time()
    R, C, S = 10000, 5000, 4
    X, cct, da = make_table(R, C, 50, 4, False, seed=3)
    Zv, Zrv = initial_partitions(R, C, 77)
    eng = Engine(X, num_states=S, rng=gen_rng(12), cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0,
                 view_alphas={v: 1.0 for v in Zrv}, stream='philox', Kcap=128, Vcap=96)
    eng.transition(N=1, kernels=['columns', 'rows'], progress=False)
    ms = timed_steps(eng, 3, ['columns'])
    out = {'workload': '10k rows x 5000 normal cols, %d chains, kernels=[columns] (1 burn-in sweep of columns+rows)' % S,
           'steps': 3, 'ms_per_step': ms, 'column_transitions_per_s': C * S / (ms * 1e-3),
           'cell_updates_per_s': R * C * S / (ms * 1e-3)}
    out.update(view_stats(eng))
    out['wall_s'] = time.time() - t0
    return out


def extra_config4(eng, C):
    """configs[4]: query batch over the converged chains of the main run (hypothetical rows, 2 targets | 2 constraints)."""
    import numpy as np
    S = eng.num_states()
    rng = np.random.RandomState(5)
    out = {'workload': 'hypothetical rows, targets = cols (0, 1), constraints = cols (2, 3), over the %d chains of the main run '
                       '(configs[1] table after the timed steps)' % S}
    Q = 1000000
    tv = rng.normal(2.0, 2.0, size=(Q, 2)); cv = rng.normal(2.0, 2.0, size=(Q, 2))
    t0 = time.time()
    lp = eng.logpdf_bulk_arrays([0, 1], tv, [2, 3], cv)
    dt = time.time() - t0
    out['logpdf_bulk_arrays'] = {'queries': Q, 'chains': S, 'seconds_incl_host_encode_and_copies': dt,
                                 'query_chains_per_s': Q * S / dt, 'finite': bool(np.isfinite(lp).all())}
    Qd = 20000
    tl = [{0: float(a), 1: float(b)} for a, b in tv[:Qd]]
    cl = [{2: float(a), 3: float(b)} for a, b in cv[:Qd]]
    t0 = time.time()
    lpd = eng.logpdf_bulk([-1] * Qd, tl, cl)
    dt = time.time() - t0
    out['logpdf_bulk_dicts'] = {'queries': Qd, 'chains': S, 'seconds_incl_host_encode': dt, 'query_chains_per_s': Qd * S / dt,
                                'max_abs_diff_vs_arrays': float(np.max(np.abs(np.asarray(lpd) - lp[:, :Qd])))}
    Qs = 200000
    t0 = time.time()
    sm = eng.simulate_bulk_arrays(Qs, [0, 1], [2, 3], cv[:Qs], N=1)
    dt = time.time() - t0
    out['simulate_bulk_arrays'] = {'queries': Qs, 'chains': S, 'seconds': dt, 'samples_per_s': Qs * S / dt,
                                   'finite': bool(np.isfinite(sm).all())}
    t0 = time.time()
    for _ in range(10):
        sc = eng.logpdf_score()
    out['logpdf_score_latency_ms'] = 1e2 * (time.time() - t0)
    return out


def extra_ksweep():
    """Rows kernel against supplied partitions of K clusters per view: cluster evaluations per second."""
    import numpy as np
    from cgpm_b200.engine import Engine, gen_rng
    R, C, S = 20000, 16, 16
    out = []
    for K in (4, 32, 256):
        X, cct, da = make_table(R, C, 2, K, False, seed=9)
        Zv = {c: c % 2 for c in range(C)}
        rng = np.random.RandomState(K)
        Zrv = {}
        for v in (0, 1):
            _, inv = np.unique(rng.randint(K, size=R), return_inverse=True)
            Zrv[v] = [int(z) for z in inv]
        eng = Engine(X, num_states=S, rng=gen_rng(13), cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={0: 1.0, 1: 1.0},
                     stream='philox', Kcap=max(64, 2 * K), Vcap=4)
        eng.transition(N=1, kernels=['rows'], progress=False)
        k0 = view_stats(eng)['clusters_per_view_mean']
        ms = timed_steps(eng, 3, ['rows'])
        k1 = view_stats(eng)['clusters_per_view_mean']
        kk = 0.5 * (k0 + k1)
        out.append({'K_supplied': K, 'K_mean_during': kk, 'ms_per_sweep': ms, 'cell_updates_per_s': R * C * S / (ms * 1e-3),
                    'cluster_evaluations_per_s': R * C * S * (kk + 1) / (ms * 1e-3)})
    return {'workload': '20k rows x 16 normal cols (2 views of 8), %d chains, kernels=[rows], random K-cluster partitions' % S,
            'runs': out}


# --------------------------------------------------------------------------- ours
def ours_main(args):
    import numpy as np
    t_begin = time.time()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = dict(C1 if world == 1 else C2)
    if args.rows:
        cfg['rows'] = args.rows
    if args.cols:
        cfg['cols'] = args.cols
    if args.chains:
        cfg['chains'] = args.chains
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference forks worker processes: run it before CUDA is initialised, in a child
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '3',
                                  '--warmup', '1', '--burnin', '0', '--ref-rows', str(args.ref_rows or REF_ROWS)],
                                 capture_output=True, text=True, timeout=900)
            cpu = json.loads(out.stdout.strip().splitlines()[-1])['cpu_baseline']
        except Exception as e:
            cpu = {'value': None, 'unit': 'cell-updates/s', 'cores': 0, 'kind': 'reference', 'sample': 'failed: %s' % (e,)}
    os.environ.setdefault('CGPM_B200_ROWS_PROF', '1')      # per-view cycle counters of the rows kernel (a few stores per CTA)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from cgpm_b200.engine import Engine, gen_rng
    from cgpm_b200.parallel import ShardedEngine
    R, Cn, S = cfg['rows'], cfg['cols'], cfg['chains']
    t_setup = time.time()
    X, cct, da = make_table(R, Cn, cfg['views'], cfg['clusters'], cfg['mixed'])
    Zv, Zrv = initial_partitions(R, Cn, 12345)
    kw = dict(cctypes=cct, distargs=da, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv},
              stream='philox', Kcap=cfg['kcap'], Vcap=cfg['vcap'])
    if cfg.get('hypers'):
        kw['hypers'] = supplied_hypers(X, cct)
    if world > 1:
        sh = ShardedEngine(X, num_states=S * world, rng=gen_rng(1), **kw)
        eng = sh.local
    else:
        sh = None
        eng = Engine(X, num_states=S, rng=gen_rng(1), **kw)
    t_setup = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- burn-in (stated), warm-up, timed steps: device-resident throughput ----------------------
    if world > 1:
        args.burnin = min(args.burnin, 3)          # a configs[2] sweep takes seconds
        args.e2e_steps = 1
    for _ in range(args.burnin):
        eng.transition(N=1, kernels=KERNELS, progress=False)
    t_warm = 0.0
    for _ in range(args.warmup):
        torch.cuda.synchronize()
        t_warm = time.time()
        eng.transition(N=1, kernels=KERNELS, progress=False)
        torch.cuda.synchronize()
        t_warm = time.time() - t_warm
    steps_requested = args.steps
    if world > 1 and args.time_budget > 0 and args.warmup > 0:
        # Safety valve for the configs[2] run only: a sweep of 64 chains x 1M rows x 200 columns takes seconds, and a run
        # the driver has to kill reports nothing.  If the K requested steps (plus the end-to-end steps behind them) would
        # not fit in the time budget at the pace of the last warm-up step, fewer steps are timed -- the same number on
        # every rank -- and the line says so (`steps` = what was timed, `steps_requested` = K).
        left = args.time_budget - (time.time() - t_begin)
        fit = int((left - 2.0 * (args.e2e_steps + 1) * 1.5 * t_warm) / max(t_warm, 1e-3))
        tfit = torch.tensor([max(2, min(args.steps, fit))], dtype=torch.int64, device='cuda')
        dist.all_reduce(tfit, op=dist.ReduceOp.MIN)
        args.steps = int(tfit.item())
    eng.native_launches(reset=True)
    eng.start_kernel_timing()
    eng._ch.dev.t['seq'][:, :, 0].zero_()              # per-view cycle counters: slots of views that no longer exist are stale
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.transition(N=1, kernels=KERNELS, progress=False)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    launches = eng.native_launches()
    kern_ms = eng.stop_kernel_timing()
    tms = torch.tensor([ms], dtype=torch.float64, device='cuda')
    rank_ms = [ms / args.steps]
    if world > 1:
        # every rank's own device time beside the max: ranks hold different chains and never exchange anything inside a
        # step, so the spread between them is the whole weak-scaling loss (a step ends with the slowest view anywhere)
        allms = torch.zeros(world, dtype=torch.float64, device='cuda')
        dist.all_gather_into_tensor(allms, tms)
        rank_ms = [float(x) / args.steps for x in allms.tolist()]
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = R * Cn * S * world * args.steps / (ms * 1e-3)
    vs = view_stats(eng)
    n_views = vs['views_total']

    def critical_cta():
        """Cycles per row of the slowest (chain, view) CTA of the rows sweeps since the counters were last cleared
        (CGPM_B200_ROWS_PROF counters; the timed steps, then each device-resident step)."""
        dv = eng._ch.dev
        c = float(dv.t['seq'][:, :, 0].max().item()) * 1024.0 / R
        dv.t['seq'][:, :, 0].zero_()
        return c
    crit = [critical_cta()]

    # ---- end to end: host table + host chain state in, host state + scores out, every step -------------
    Xpin = torch.from_numpy(X).pin_memory()
    dense = eng.to_dense()
    e2e_ms = []
    h2d = d2h = 0
    small = list(eng._ch.SMALL) + ['zr', 'order']
    for it in range(args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        eng.load_dense(dense, X=Xpin)
        eng.transition(N=1, kernels=KERNELS, progress=False)
        dense = eng.to_dense(pinned=dense)
        scores = eng.logpdf_score()
        barrier()
        if it > 0:
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
        h2d = Xpin.numel() * 8 + sum(dense[k].numel() * dense[k].element_size() for k in small)
        d2h = sum(dense[k].numel() * dense[k].element_size() for k in small) + 8 * len(scores)
    te = torch.tensor([sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device='cuda')
    # resident variant: the chain state stays on the device, only the scores come back every step
    res_ms = []
    critical_cta()                                      # clear: what follows is per step
    for it in range(args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        eng.transition(N=1, kernels=KERNELS, progress=False)
        scores = eng.logpdf_score()
        barrier()
        if it > 0:
            res_ms.append(1e3 * (time.perf_counter() - t0))
        crit.append(critical_cta())
    tr = torch.tensor([sum(res_ms) / len(res_ms)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    e2e_val = R * Cn * S * world / (float(te.item()) * 1e-3)
    res_val = R * Cn * S * world / (float(tr.item()) * 1e-3)

    # ---- gather per-chain scores (the only collective of the path) --------------------
    sc = np.asarray(sh.logpdf_score() if sh is not None else eng.logpdf_score())
    eng._ch.dev.check_errors()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        rows_ms = kern_ms.get('rows', [float('nan')])
        rows_avg = sum(rows_ms) / len(rows_ms)
        # algorithmic bytes of one rows-kernel launch: 8 B per cell visited +
        # 16 B per (row, view) visit (row id, cluster in, cluster out, member order)
        alg_bytes = 8.0 * R * Cn * S + 16.0 * R * n_views
        achieved = alg_bytes / (rows_avg * 1e-3) / 1e9
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'rows_kernel_profile.json')))
        except Exception:
            pass
        sha = src_sha16()
        stale = prof.get('src_sha16') != sha
        traffic = prof.get('dram_bytes_per_launch') if (R, Cn, S) == (100000, 50, 64) else None
        # FP64 work of one rows launch: every cell of a row is scored against each of the view's clusters, the empty
        # cluster and (the row removed) its own cluster once more
        evals = R * Cn * S * (vs['clusters_per_view_mean'] + 2.0)
        line = {
            'metric': 'crosscat_gibbs_cell_updates_per_sec', 'value': value, 'unit': 'cell-updates/s',
            'n_gpus': world, 'steps': args.steps, 'steps_requested': steps_requested, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'ms_per_step_by_rank': rank_ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {
                'workload': '%s: %d rows x %d %s cols, %d chains per GPU, kernels=%s, philox stream, public Engine API%s'
                            % (cfg['name'], R, Cn, 'mixed normal/categorical' if cfg['mixed'] else 'normal', S, KERNELS,
                               ' (ShardedEngine over NCCL)' if world > 1 else ''),
                'init': 'supplied Zv~CRP(1), Zrv~CRP(1) (the same for every chain, as in the reference arm), alpha=1, '
                        'view_alphas=1 (SURVEY 8d)%s; %d burn-in sweeps, then %d warm-up sweeps'
                        % ('; column hypers supplied on the scale of the data (normal: m=mean, r=1, s=var, nu=1; categorical: '
                           'alpha=1) because column_hypers is not among the kernels' if cfg.get('hypers') else '',
                           args.burnin, args.warmup),
                'cache': 'inputs larger than L2: per-step traffic (table %.0f MB + chain state %.0f MB) exceeds the 126 MB L2'
                         % (X.nbytes / 1e6, n_views * R * 8 / 1e6),
                'Kcap': eng._ch.dev.Kcap, 'Vcap': eng._ch.dev.Vcap, 'setup_s': t_setup,
            },
            'e2e': {'value': e2e_val, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': float(te.item()),
                    'what': 'Engine.load_dense(host chain state, host table) [pinned -> device, ordered rebuild of all tables], '
                            'transition(N=1), Engine.to_dense() + logpdf_score() [device -> host], every step: the worst case '
                            'of the lovecat-style bridge (INTEGRATION.md)',
                    'when': 'sweeps %d.. of the same chains, after the K timed steps: a step is as slow as its slowest view, and '
                            'which views exist changes from sweep to sweep (kernel_ms_step_stats), so e2e can be faster or '
                            'slower than `value`' % (args.burnin + args.warmup + args.steps + 1)},
            'e2e_resident': {'value': res_val, 'unit': 'cell-updates/s', 'ms_per_step': float(tr.item()),
                             'd2h_bytes_per_step': 8 * S, 'what': 'chain state stays on the device: transition(N=1) + logpdf_score() to the host'},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'rows_kernel<224, 2, false> (one launch per sweep, one CTA per (chain, view))',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_stale': bool(stale) if traffic is not None else None,
                         'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback',
                         'ms_per_launch': rows_avg, 'alg_bytes_per_launch': alg_bytes,
                         'traffic_source': prof.get('source'),
                         'note': 'latency / FP64-issue bound sequential scan (SURVEY 8d): the HBM fraction is <1% by construction; '
                                 'roofline_compute is the binding one'},
            'roofline_compute': {
                'bound': 'fp64', 'unit': 'GFLOP/s',
                'achieved': evals * FLOPS_PER_CELL_EVAL / (rows_avg * 1e-3) / 1e9, 'peak': FP64_PEAK_FLOPS / 1e9,
                'frac': evals * FLOPS_PER_CELL_EVAL / (rows_avg * 1e-3) / FP64_PEAK_FLOPS,
                'cell_evaluations_per_launch': evals, 'flops_per_cell_evaluation': FLOPS_PER_CELL_EVAL,
                'cluster_evaluations_per_s': evals / (rows_avg * 1e-3),
                'peak_source': '148 SMs x 64 FP64 FMA lanes x 2 x 1.965 GHz (no measured FP64 figure in MEASURED_PEAKS.json)',
                'cycles_per_row_critical_cta': crit, 'ubench_floor_cycles_per_row': [640, 830],
                'fp64_pipe_pct_of_active_cycles': prof.get('fp64_pipe_pct_active'),
                'issue_slots_pct_active': prof.get('issue_active_pct'), 'inst_per_cycle_per_sm': prof.get('inst_per_cycle_active'),
                'stall_barrier_per_issue': prof.get('stall_barrier'), 'profile_stale': bool(stale), 'profile_source': prof.get('source'),
                'note': 'no SFU path in fp64: log / exp are FP64-pipe instruction sequences (north_star asks for the SFU '
                        'fraction; MUFU is used only for reciprocal seeds)'},
            'kernel_ms_per_step': {k: sum(v) / args.steps for k, v in kern_ms.items()},
            'kernel_ms_step_stats': {k: step_stats(v, args.steps) for k, v in kern_ms.items()},
            'cpu_baseline': cpu,
            'logscore_mean': float(sc.mean()),
        }
        line['config'].update(vs)
        if world == 1 and not args.no_extra:
            extra = {}
            t_extra = time.time()
            for name, fn in (('configs[4]_queries', lambda: extra_config4(eng, Cn)), ('k_sweep', extra_ksweep),
                             ('configs[3]_columns', extra_config3),
                             ('configs[2]_shard', lambda: extra_config2(args.extra_budget))):
                if time.time() - t_extra > args.extra_budget:
                    extra[name] = {'skipped': 'time budget of %d s for the extra block used up' % args.extra_budget}
                    continue
                if name != 'configs[4]_queries':
                    eng = None
                try:
                    extra[name] = fn()
                except Exception as e:           # an extra must never take the headline down
                    extra[name] = {'failed': '%s: %s' % (type(e).__name__, e)}
                torch.cuda.empty_cache()
            line['extra'] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--burnin', type=int, default=15)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rows', type=int, default=0)
    ap.add_argument('--cols', type=int, default=0)
    ap.add_argument('--chains', type=int, default=0)
    ap.add_argument('--ref-rows', type=int, default=0, help='rows of the reference arm\'s sample (0: per config)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--extra-budget', type=int, default=150)
    ap.add_argument('--time-budget', type=int, default=600,
                    help='N > 1 only: seconds the whole run may take; the timed steps are cut (and reported) if K would not fit')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_main(args)
    else:
        ours_main(args)


if __name__ == '__main__':
    main()
