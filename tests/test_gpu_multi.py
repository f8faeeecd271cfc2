"""Two GPUs, NCCL: ShardedEngine (cgpm_b200/parallel.py) holds the chains of one Engine spread over the ranks
and answers like the single-process Engine with the same seeds.  Skipped with fewer than two devices."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from cgpm_b200.parallel import ShardedEngine
    from cgpm_b200.engine import gen_rng
    from helpers import small_table
    X, cct, da = small_table(50, 8)
    eng = ShardedEngine(X, num_states=5, rng=gen_rng(3), cctypes=cct, distargs=da, stream='numpy')
    eng.transition(N=2, progress=False)
    out = {
        'scores': eng.logpdf_score(),
        'lp': eng.logpdf(-1, {0: 1.0}, {2: 0.5}),
        'lw': eng.logpdf_likelihood_weighted(-1, {0: 1.0}, {2: 0.5}),
        'dep': eng.dependence_probability_pairwise().tolist(),
        'sim_shape': [len(eng.simulate_bulk([-1, -1], [[0, 1], [2]], None, Ns=[3, 2])),
                      len(eng.simulate(-1, [0, 1], N=4))],
        'n_md': len(eng.to_metadata()['states']),
        'state3': eng.get_state(3).logpdf_score(),
    }
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_engine_over_two_gpus_matches_the_single_process_engine():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    from cgpm_b200.engine import Engine, gen_rng
    from cgpm_b200 import hostmath as hm
    from helpers import small_table
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the same five chains in one process (numpy stream: seeds decide everything)
    X, cct, da = small_table(50, 8)
    eng = Engine(X, num_states=5, rng=gen_rng(3), cctypes=cct, distargs=da, stream='numpy')
    eng.transition(N=2, progress=False)
    want = eng.logpdf_score()
    lp = eng.logpdf(-1, {0: 1.0}, {2: 0.5})
    lw = hm.logmeanexp_weighted(lp, eng.logpdf(-1, {2: 0.5}))
    dep = np.mean(eng.dependence_probability_pairwise(), axis=0)
    for rank in (0, 1):
        out = res[rank]
        assert np.allclose(out['scores'], want, rtol=1e-12)
        assert np.allclose(out['lp'], lp, rtol=1e-12)
        assert abs(out['lw'] - lw) <= 1e-12 * abs(lw)
        assert np.allclose(out['dep'], dep)
        assert out['sim_shape'] == [5, 5] and out['n_md'] == 5
        assert abs(out['state3'] - want[3]) <= 1e-10 * abs(want[3])
