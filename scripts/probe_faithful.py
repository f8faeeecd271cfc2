"""Reference-faithful default initialisation at configs[1] size: Engine(X, num_states=8) with no
supplied partitions, R = 100 000: views start with up to ~7e4 clusters."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cgpm_b200.engine import Engine, gen_rng
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 50
S = int(sys.argv[3]) if len(sys.argv) > 3 else 8
X, cct, da = bench.make_table(R, C, 4, 4)
t0 = time.time()
eng = Engine(X, num_states=S, rng=gen_rng(3), cctypes=cct, stream='philox')
dev = eng._ch.dev
torch.cuda.synchronize()
nv = dev.t['nv'].cpu().numpy(); K = dev.t['nclu'].cpu().numpy()
print('init %.1fs  Kcap %d Vcap %d  views %d  K per view: max %d, >1000: %d, >10000: %d' % (
    time.time() - t0, dev.Kcap, dev.Vcap, int((nv > 0).sum()), K.max(), int((K > 1000).sum()), int((K > 10000).sum())), flush=True)
os.environ['CGPM_B200_ROWS_PROF'] = '1'
for it in range(3):
    t0 = time.time()
    try:
        eng.transition(N=1, kernels=['rows'], progress=False)
    except Exception as e:
        seq = dev.t['seq'].cpu().numpy()
        for s in range(S):
            for v in range(dev.Vcap):
                if nv[s, v] > 0 and seq[s, v, 6] != 0:
                    print('draw failed: chain %d view %d path %d row position %d K %d D %d mode %d' % (s, v, seq[s, v, 6], seq[s, v, 7], seq[s, v, 8], seq[s, v, 9], seq[s, v, 4]))
                    print('first bad candidate position %d slot %d, u*1e9 %d, tot finite %d, z of that row?' % (seq[s, v, 10], seq[s, v, 11], seq[s, v, 12], seq[s, v, 13]))
                    k = int(seq[s, v, 11])
                    if k >= 0:
                        st = dev.t['stat'][s].cpu().numpy(); cols = np.nonzero(dev.t['zv'][s].cpu().numpy() == v)[0]
                        for c in cols[:50]:
                            rec = st[:, k, c]
                            if not np.all(np.isfinite(rec)):
                                print('col', c, 'record', rec.tolist(), 'hyp', dev.t['hyp'][s, c].cpu().numpy().tolist(), 'nk', int(dev.t['nk'][s, v, k]))
                                break
                    vs = v
                    hyp = dev.t['hyp'][s].cpu().numpy(); zv = dev.t['zv'][s].cpu().numpy()
                    print('hypers of its columns', hyp[zv == v][:6].tolist(), 'alpha', float(dev.t['view_alpha'][s, v]))
        raise
    torch.cuda.synchronize(); t1 = time.time()
    eng.transition(N=1, kernels=['columns'], progress=False)
    torch.cuda.synchronize(); t2 = time.time()
    nv = dev.t['nv'].cpu().numpy(); K = dev.t['nclu'].cpu().numpy()
    print('sweep %d: rows %.2fs columns %.2fs  Kcap %d (grown %d)  K max %d mean %.1f' % (
        it, t1 - t0, t2 - t1, dev.Kcap, dev.kcap_grown, K[nv > 0].max(), K[nv > 0].mean()), flush=True)
eng._ch.check_partitions()
print('partitions consistent; score', eng.logpdf_score()[:3])
