"""ShardedEngine's array-form bulk queries in a single process (world = 1): equal to the local Engine's."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))      # tests/helpers.py imports the checker
import numpy as np
from cgpm_b200.parallel import ShardedEngine
from cgpm_b200.engine import gen_rng
from helpers import small_table
X, cct, da = small_table(50, 8)
sh = ShardedEngine(X, num_states=3, rng=gen_rng(3), cctypes=cct, distargs=da)
sh.transition(N=1, progress=False)
tv = np.random.RandomState(0).normal(size=(40, 1))
cv = np.zeros((40, 1))
a = sh.logpdf_bulk_arrays([0], tv, [2], cv)
b = sh.local.logpdf_bulk_arrays([0], tv, [2], cv)
s = sh.simulate_bulk_arrays(40, [0, 1], [2], cv, N=2)
d = np.asarray(sh.logpdf_bulk([-1] * 40, [{0: float(v)} for v in tv[:, 0]], [{2: 0.0}] * 40))
assert a.shape == (3, 40) and np.array_equal(a, b) and s.shape == (3, 40, 2, 2) and np.isfinite(s).all()
assert np.allclose(a, d, rtol=1e-12, atol=0), float(np.max(np.abs(a - d)))
print('sharded arrays ok', a.shape, s.shape)
