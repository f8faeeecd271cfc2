#!/usr/bin/env python
"""Self-check of the py3 shim (oracle/_ref): run the reference's OWN hot-path
tests against it.  TEST INFRASTRUCTURE ONLY; needs /root/reference/tests, so it
runs in the build container only (never on the GPU box).

The test files are converted with the same mechanical transform as the sources
and written to a temporary directory; tests/conftest.py (imports matplotlib,
absent here) is replaced by an empty one.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_ref  # noqa: E402

TESTS = [
    'test_crp.py', 'test_logsumexp.py', 'test_view_logpdf_cluster.py',
    'test_logpdf_score.py', 'test_iter_counter.py',
    'test_dependence_probability.py', 'test_add_state.py', 'test_type_check.py',
    'test_engine_seed.py', 'test_state_initialize.py', 'test_incorporate_row.py',
]


def main(ref_tests='/root/reference/tests'):
    root = build_ref.build(quiet=True)
    if root is None or not os.path.isdir(ref_tests):
        print('reference not present; skipped')
        return 0
    tmp = tempfile.mkdtemp(prefix='cgpm_reftests_')
    try:
        for fn in TESTS:
            src = os.path.join(ref_tests, fn)
            if not os.path.exists(src):
                continue
            with open(src) as f:
                text = f.read()
            with open(os.path.join(tmp, fn), 'w') as f:
                f.write(build_ref.convert_source(text, fn))
        open(os.path.join(tmp, 'conftest.py'), 'w').write('')
        open(os.path.join(tmp, 'hacks.py'), 'w').write(
            'import pytest\ndef skip(reason):\n'
            '    raise pytest.skip.Exception(reason, allow_module_level=True)\n')
        open(os.path.join(tmp, 'markers.py'), 'w').write(
            'import pytest\nintegration = pytest.mark.skip(reason="integration")\n')
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + tmp)
        cmd = [sys.executable, '-m', 'pytest', '-q', '-W', 'ignore', '-p', 'no:cacheprovider',
               '-k', 'not __ci_ and not random_forest'] + [t for t in TESTS if os.path.exists(os.path.join(tmp, t))]
        return subprocess.call(cmd, cwd=tmp, env=env)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    sys.exit(main())
