"""A C++ host with no Python in the loop runs inference through the C ABI alone: the driver in
tests/c_driver/driver.cu allocates a cc_state with cudaMalloc, loads an initial state, calls cc_rebuild and ONE
cc_transition (30 iterations of rows, columns, view_alphas, alpha, column_hypers -- the LocalEngine.analyze shape of
src/crosscat/lovecat.py:377-391), then checks invariants, scores and a query batch."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_only_host_runs_inference_through_cc_transition():
    import __graft_entry__ as g
    exe = g.build_c_driver()
    if exe is None:
        pytest.skip('nvcc not available')
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, 'cgpm_b200') + ':' + os.environ.get('LD_LIBRARY_PATH', ''))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0 and 'C driver OK' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
