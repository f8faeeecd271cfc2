"""CPU: the C-ABI library loads and exports every function include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'cgpm_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'^\s*(?:const\s+char\s*\*|int)\s+(cc_\w+)\s*\(', text, flags=re.M)
    return sorted(set(names))


def test_header_and_binding_agree():
    from cgpm_b200 import _lib
    assert declared_functions() == sorted(_lib.SYMBOLS)


def test_library_loads_and_exports_all_symbols():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    from cgpm_b200 import _lib
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.cc_abi_version() == _lib.ABI_VERSION == 4
    assert lib.cc_last_error() is not None


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Compile the header with gcc and compare sizeof/offsetof with the ctypes mirror."""
    import shutil
    import subprocess
    from cgpm_b200 import _lib
    if shutil.which('gcc') is None:
        import pytest
        pytest.skip('gcc not available')
    fields = [f[0] for f in _lib.cc_state._fields_]
    wfields = [f[0] for f in _lib.cc_row_work._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "cgpm_b200.h"', 'int main(void){',
           'printf("%zu %zu\\n", sizeof(cc_state), sizeof(cc_row_work));']
    for f in fields:
        src.append('printf("%%zu\\n", offsetof(cc_state, %s));' % f)
    for f in wfields:
        src.append('printf("%%zu\\n", offsetof(cc_row_work, %s));' % f)
    src.append('return 0;}')
    c = tmp_path / 'layout.c'
    c.write_text('\n'.join(src))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(c), '-o', str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == ctypes.sizeof(_lib.cc_state) and int(out[1]) == ctypes.sizeof(_lib.cc_row_work)
    offs = [int(x) for x in out[2:]]
    want = [getattr(_lib.cc_state, f).offset for f in fields] + [getattr(_lib.cc_row_work, f).offset for f in wfields]
    assert offs == want


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from cgpm_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    import pytest
    with pytest.raises(_lib.CgpmB200Error):
        _lib.load()


def test_no_product_import_of_the_oracle():
    """The product package never imports or executes anything under oracle/."""
    pkg = os.path.join(ROOT, 'cgpm_b200')
    for dirpath, _d, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, fn)).read()
                assert 'crosscat_oracle' not in text and 'oracle/' not in text and 'build_ref' not in text, fn
