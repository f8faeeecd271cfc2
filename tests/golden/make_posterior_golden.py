#!/usr/bin/env python
"""Posterior-summary fixture from the REFERENCE (oracle/_ref): 64 independent chains
of the reference Engine on a small mixed table; per-chain number of views, number of
clusters per view (mean), logpdf_score and the pairwise dependence matrix averaged
over chains.  Used by tests/test_gpu_posterior.py (device-generator stream, so only
distributions can agree).   python tests/golden/make_posterior_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, ROOT)
import build_ref  # noqa: E402
sys.path.insert(0, build_ref.build(quiet=True))
from cgpm.crosscat.engine import Engine   # noqa: E402
from cgpm.utils import general as gu      # noqa: E402
from cgpm_b200 import synth               # noqa: E402

R, S, N = 120, 64, 60
cct = ['normal', 'normal', 'categorical', 'normal', 'categorical', 'normal']
da = [None, None, {'k': 3}, None, {'k': 4}, None]
X, Zv, Zrv = synth.gen_table(R, cct, da, n_views=2, clusters_per_view=3, separation=0.9, seed=5)
kernels = ['rows', 'columns', 'view_alphas', 'alpha', 'column_hypers']
eng = Engine(X, num_states=S, rng=gu.gen_rng(11), cctypes=cct, distargs=da, multiprocess=1)
eng.transition(N=N, kernels=kernels, progress=False, multiprocess=1)
n_views = [len(s.views) for s in eng.states]
n_clusters = [float(np.mean([len(v.Nk()) for v in s.views.values()])) for s in eng.states]
scores = [float(x) for x in eng.logpdf_score()]
D = np.mean(eng.dependence_probability_pairwise(), axis=0)
probe = [float(np.mean(eng.logpdf(-1, {0: float(np.nanmean(X[:, 0])), 3: float(np.nanmean(X[:, 3]))}, multiprocess=1)))]
out = {'R': R, 'S': S, 'N': N, 'cctypes': cct, 'distargs': [d or {} for d in da], 'table_seed': 5,
       'kernels': kernels, 'n_views': n_views, 'n_clusters': n_clusters, 'logpdf_score': scores,
       'dependence': D.tolist(), 'true_Zv': [int(z) for z in Zv], 'probe_logpdf_mean': probe}
with open(os.path.join(HERE, 'posterior_summary.json'), 'w') as f:
    json.dump(out, f)
print('views', np.mean(n_views), 'clusters', np.mean(n_clusters), 'score', np.mean(scores))
