import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_gpu_fullsize as T
eng = T.build(S=12, seed=5)
ch = eng._ch
for rnd in range(4):
    eng.transition(N=3, kernels=['rows', 'columns', 'alpha', 'view_alphas', 'column_hypers'] if rnd % 2 else ['rows', 'columns'], progress=False)
    ch.check_partitions()
    stat_inc = ch.dev.t['stat'].clone(); zv = ch.dev.t['zv'].cpu().numpy()
    live = ch.dev.t['nv'].cpu().numpy() > 0
    ch.dev.rebuild(); ch.dev.sync()
    stat_new = ch.dev.t['stat']
    nclu = ch.dev.t['nclu'].cpu().numpy(); livel = ch.dev.t['live'].cpu().numpy()
    bad = 0
    for s in range(ch.S):
        for v in np.nonzero(live[s])[0]:
            slots = torch.as_tensor(livel[s, v, :nclu[s, v]].astype(np.int64), device=ch.dev.device)
            cols = torch.as_tensor(np.nonzero(zv[s] == v)[0].astype(np.int64), device=ch.dev.device)
            for f in (0, 1, 2):
                a = stat_inc[s, f][slots][:, cols]; b = stat_new[s, f][slots][:, cols]
                bad += int(not torch.equal(a, b))
    print('round', rnd, 'views', int(live.sum()), 'K max', int(nclu.max()), 'mismatching (view, field) blocks:', bad, 'score', float(np.mean(ch.logpdf_score())))
    assert bad == 0
print('stress ok')
