#!/usr/bin/env python
"""Benchmark of the CrossCat Gibbs hot path (BASELINE.json metric: cell-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one `Engine.transition(N=1, kernels=['rows','columns'])` over all
chains.  Workload at N=1: BASELINE.json configs[1] -- 100k rows x 50 normal
columns, 64 chains, kernels rows+columns, on one B200 (synthetic table from
cgpm_b200.synth, chains initialised with supplied CRP(1) partitions as
SURVEY.md section 8(d) prescribes).  For N>1 every rank holds 64 chains (weak
scaling; chains never interact, engine.py:78-85) and NCCL only gathers the
per-chain scores.  One JSON line is printed by rank 0.

--impl reference times the reference's own CPU Engine (oracle/_ref, the
py3-shimmed probcomp/cgpm; multiprocess fan-out over the host cores) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KERNELS = ['rows', 'columns']
WORKLOAD = dict(rows=100000, cols=50, chains=64, views=4, clusters=4)
REF_SAMPLE = dict(rows=1000, cols=50)


def make_table(R, C, views, clusters, seed=0):
    from cgpm_b200 import synth
    cct = ['normal'] * C
    da = [None] * C
    X, _, _ = synth.gen_table(R, cct, da, n_views=views, clusters_per_view=clusters, seed=seed)
    return X, cct, da


def initial_partitions(R, C, seed):
    """Zv ~ CRP(1) over columns, Zrv ~ CRP(1) over rows per view, alphas = 1."""
    import numpy as np
    from cgpm_b200 import hostmath as hm
    rng = np.random.RandomState(seed)
    zv = hm.crp_simulate_fast(C, 1.0, rng)
    Zv = {c: int(z) for c, z in enumerate(zv)}
    Zrv = {int(v): hm.crp_simulate_fast(R, 1.0, rng) for v in set(Zv.values())}
    return Zv, Zrv


# --------------------------------------------------------------------------- reference arm
def run_reference(R, C, chains, steps, warmup, views, clusters):
    """The reference Engine (multiprocess) on an R x C sample; returns a dict."""
    import numpy as np
    ref = os.path.join(ROOT, 'oracle', '_ref')
    kind = 'reference'
    cores = os.cpu_count() or 1
    S = max(1, min(chains, cores))
    X, cct, da = make_table(R, C, views, clusters)
    Zv, Zrv = initial_partitions(R, C, 12345)
    if os.path.exists(os.path.join(ref, 'cgpm', '.built')):
        sys.path.insert(0, ref)
        from cgpm.crosscat.engine import Engine
        from cgpm.utils import general as gu
        eng = Engine(X, num_states=S, rng=gu.gen_rng(1), multiprocess=1, cctypes=cct,
                     Zv=Zv, Zrv={v: list(map(int, z)) for v, z in Zrv.items()}, alpha=1.0,
                     view_alphas={v: 1.0 for v in Zrv})
        step = lambda: eng.transition(N=1, kernels=KERNELS, progress=False, multiprocess=1)
    else:
        kind = 'port'
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import crosscat_oracle as co
        S = 1
        states = [co.OState(X, cctypes=cct, Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv},
                            rng=np.random.RandomState(2 + s)) for s in range(S)]
        step = lambda: [o.transition(N=1, kernels=KERNELS) for o in states]
    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = time.time() - t0
    val = R * C * S * steps / dt
    return dict(value=val, seconds=dt, cores=min(S, cores) if kind == 'reference' else 1, kind=kind,
                sample='%d rows x %d normal cols x %d chains, %d step(s) of %s, supplied CRP(1) init; host has %d cores'
                       % (R, C, S, steps, '+'.join(KERNELS), cores), chains=S, ms_per_step=1e3 * dt / steps)


def reference_main(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = run_reference(args.ref_rows, args.cols, args.chains, args.steps, args.warmup, args.views, args.clusters)
    line = {
        'impl': 'reference', 'metric': 'crosscat_gibbs_cell_updates_per_sec', 'value': r['value'],
        'unit': 'cell-updates/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'configs[1] sample: %d rows x %d normal cols, %d chains, kernels=%s (CPU reference Engine, multiprocess)'
                               % (args.ref_rows, args.cols, r['chains'], KERNELS)},
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


# --------------------------------------------------------------------------- ours
def ours_main(args):
    import numpy as np
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference forks worker processes: run it before CUDA is initialised, in a child
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1',
                                  '--warmup', '0', '--ref-rows', str(args.ref_rows), '--cols', str(args.cols)],
                                 capture_output=True, text=True, timeout=900)
            cpu = json.loads(out.stdout.strip().splitlines()[-1])['cpu_baseline']
        except Exception as e:
            cpu = {'value': None, 'unit': 'cell-updates/s', 'cores': 0, 'kind': 'reference', 'sample': 'failed: %s' % (e,)}
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from cgpm_b200.engine import Engine, gen_rng
    from cgpm_b200 import _lib as L
    R, Cn, S = args.rows, args.cols, args.chains
    X, cct, da = make_table(R, Cn, args.views, args.clusters)
    eng = Engine(X, num_states=0, rng=gen_rng(1 + rank), cctypes=cct, stream='philox', Kcap=args.kcap, Vcap=args.vcap)
    eng._make_chains(S, 1000 + rank)
    ch = eng._ch
    for s in range(S):
        Zv, Zrv = initial_partitions(R, Cn, 100 * rank + s)
        ch.init_chain(s, gen_rng(7 + s), Zv=Zv, Zrv=Zrv, alpha=1.0, view_alphas={v: 1.0 for v in Zrv})
    ch.finish_init()
    ch.dev.sync()
    dev = ch.dev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -----------------------------------------------
    for _ in range(args.warmup):
        eng.transition(N=1, kernels=KERNELS, progress=False)
    dev.lib_calls.clear()
    ch.events = []
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
    launches = sum(dev.lib_calls.values())
    events, ch.events = ch.events, None
    kern_ms = {}
    for name, a, b in events:
        kern_ms.setdefault(name, []).append(a.elapsed_time(b))
    tms = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = R * Cn * S * world * args.steps / (ms * 1e-3)
    nv = dev.t['nv'].cpu().numpy(); nclu = dev.t['nclu'].cpu().numpy()
    n_views = int((nv > 0).sum())
    kstat = nclu[nv > 0]

    # ---- end to end: host state in, host state out, every step -----------------------
    Xpin = torch.from_numpy(X).pin_memory()
    dense = ch.export_dense()
    e2e_ms = []
    h2d = d2h = 0
    for it in range(args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        ch.import_dense(dense, X=Xpin)
        eng.transition(N=1, kernels=KERNELS, progress=False)
        dense = ch.export_dense(pinned=dense)
        scores = ch.logpdf_score()
        barrier()
        if it > 0:
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
        h2d = Xpin.numel() * 8 + sum(dense[k].numel() * dense[k].element_size() for k in list(ch.SMALL) + ['zr', 'order'])
        d2h = sum(dense[k].numel() * dense[k].element_size() for k in list(ch.SMALL) + ['zr', 'order']) + scores.nbytes
    te = torch.tensor([sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = R * Cn * S * world / (float(te.item()) * 1e-3)

    # ---- gather per-chain scores (the only collective of the path) --------------------
    sc = dev.logpdf_score()
    if world > 1:
        allsc = [torch.empty_like(sc) for _ in range(world)]
        dist.all_gather(allsc, sc)
        sc = torch.cat(allsc)
    dev.check_errors()

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
        line = {
            'metric': 'crosscat_gibbs_cell_updates_per_sec', 'value': value, 'unit': 'cell-updates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {
                'workload': 'configs[1]: %d rows x %d normal cols, %d chains per GPU, kernels=%s, philox stream'
                            % (R, Cn, S, KERNELS),
                'init': 'supplied Zv~CRP(1), Zrv~CRP(1), alpha=1, view_alphas=1 (SURVEY 8d); %d warm-up sweeps' % args.warmup,
                'cache': 'inputs larger than L2: per-step traffic (table %.0f MB + chain state %.0f MB) exceeds the 126 MB L2'
                         % (X.nbytes / 1e6, n_views * R * 8 / 1e6),
                'views_total': n_views, 'clusters_per_view_mean': float(kstat.mean()), 'clusters_per_view_max': int(kstat.max()),
                'Kcap': dev.Kcap, 'Vcap': dev.Vcap,
            },
            'e2e': {'value': e2e_val, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': float(te.item()),
                    'what': 'host table + dense chain state -> device, table rebuild, transition(N=1), state + scores -> host'},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'rows phase: rows_kernel<224, PIPE> + rows_kernel<128> + rows_narrow_kernel (concurrent)', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': (3.75e9 if (R, Cn, S) == (100000, 50, 64) else None), 'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback',
                         'ms_per_launch': rows_avg, 'alg_bytes_per_launch': alg_bytes,
                         'traffic_source': 'profiles/r01_rows_pipe_kernel_v7_raw.csv + r01_rows_kernel_v7_raw.csv + r01_rows_narrow_kernel_v7_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum of the three concurrent launches of one rows phase at this workload',
                         'note': 'latency-bound sequential scan (SURVEY 8d): HBM fraction is <1% by construction; see DESIGN.md'},
            'kernel_ms_per_step': {k: sum(v) / args.steps for k, v in kern_ms.items()},
            'cpu_baseline': cpu,
            'logscore_mean': float(sc.mean().item()),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rows', type=int, default=WORKLOAD['rows'])
    ap.add_argument('--cols', type=int, default=WORKLOAD['cols'])
    ap.add_argument('--chains', type=int, default=WORKLOAD['chains'])
    ap.add_argument('--views', type=int, default=WORKLOAD['views'])
    ap.add_argument('--clusters', type=int, default=WORKLOAD['clusters'])
    ap.add_argument('--kcap', type=int, default=256)
    ap.add_argument('--vcap', type=int, default=24)
    ap.add_argument('--ref-rows', type=int, default=REF_SAMPLE['rows'])
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_main(args)
    else:
        ours_main(args)


if __name__ == '__main__':
    main()
