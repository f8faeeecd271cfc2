"""ctypes binding of libcgpm_b200.so (the C ABI declared in include/cgpm_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There
is no CPU fallback: if the shared object is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CGPM_B200_LIB: another build of the same library (probe builds with phase counters; scripts/ only)
LIB_PATH = os.environ.get('CGPM_B200_LIB') or os.path.join(HERE, 'libcgpm_b200.so')

NGRID = 30
NFIELD = 8
NORMAL, CATEGORICAL, BERNOULLI, POISSON, EXPONENTIAL, GEOMETRIC = range(6)
F_N, F_SX, F_SXX, F_MN, F_A, F_CST, F_GN, F_GM1 = range(8)
OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY = 0, 1, 2, 3
ROWS_DONE, ROWS_HANDOVER, ROWS_GROW = 0, 1, 2
ABI_VERSION = 4


class cc_state(C.Structure):
    _fields_ = [
        ('R', C.c_int32), ('C', C.c_int32), ('Cp', C.c_int32), ('S', C.c_int32),
        ('Vcap', C.c_int32), ('Kcap', C.c_int32), ('CT', C.c_int32), ('_pad', C.c_int32),
        ('Xrm', C.c_void_p), ('Xcm', C.c_void_p), ('ctype', C.c_void_p),
        ('ncat', C.c_void_p), ('catoff', C.c_void_p), ('grids', C.c_void_p), ('Xaux', C.c_void_p), ('auxcol', C.c_void_p),
        ('row_alpha_grid', C.c_void_p), ('col_alpha_grid', C.c_void_p),
        ('hyp', C.c_void_p), ('zv', C.c_void_p), ('nv', C.c_void_p),
        ('view_id', C.c_void_p), ('col_alpha', C.c_void_p), ('view_alpha', C.c_void_p), ('view_grid', C.c_void_p),
        ('nclu', C.c_void_p), ('live', C.c_void_p), ('cid', C.c_void_p), ('nk', C.c_void_p),
        ('zr', C.c_void_p), ('order', C.c_void_p), ('stat', C.c_void_p), ('cnt', C.c_void_p),
        ('err', C.c_void_p), ('rows_prog', C.c_void_p),
        ('seq', C.c_void_p), ('moved', C.c_void_p), ('movedflag', C.c_void_p),
        ('colL', C.c_void_p), ('aux_zr', C.c_void_p), ('aux_K', C.c_void_p),
        ('aux_alpha', C.c_void_p), ('aux_zr_all', C.c_void_p), ('aux_K_all', C.c_void_p),
        ('arena', C.c_void_p), ('arena_bytes', C.c_int64),
    ]


class cc_row_work(C.Structure):
    _fields_ = [
        ('chain', C.c_int32), ('view', C.c_int32), ('n', C.c_int32),
        ('perm_is_index', C.c_int32), ('perm_off', C.c_int64), ('u_off', C.c_int64),
        ('trace_off', C.c_int64), ('n_cols', C.c_int32), ('n_normal', C.c_int32),
        ('resume', C.c_int32), ('cost_hint', C.c_int32),
    ]


GROW_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32)
K_ALPHA, K_VIEW_ALPHAS, K_COLUMN_HYPERS, K_ROWS, K_COLUMNS = range(5)
KERNEL_CODES = {'alpha': K_ALPHA, 'view_alphas': K_VIEW_ALPHAS, 'column_hypers': K_COLUMN_HYPERS, 'rows': K_ROWS,
                'columns': K_COLUMNS}


class CgpmB200Error(RuntimeError):
    pass


_lib = None


def _declare(lib):
    P, I, U64, D = C.c_void_p, C.c_int, C.c_uint64, C.c_double
    lib.cc_last_error.restype = C.c_char_p
    lib.cc_last_error.argtypes = []
    lib.cc_abi_version.restype = I
    lib.cc_abi_version.argtypes = []
    lib.cc_rebuild.restype = I
    lib.cc_rebuild.argtypes = [C.POINTER(cc_state), I, P, P, P, P]
    lib.cc_refresh_tables.restype = I
    lib.cc_refresh_tables.argtypes = [C.POINTER(cc_state), I, P, P, P]
    lib.cc_check_state.restype = I
    lib.cc_check_state.argtypes = [C.POINTER(cc_state), P, P]
    lib.cc_logpdf_score.restype = I
    lib.cc_logpdf_score.argtypes = [C.POINTER(cc_state), P, P]
    lib.cc_transition_rows.restype = I
    lib.cc_transition_rows.argtypes = [C.POINTER(cc_state), I, C.POINTER(cc_row_work),
                                       P, P, P, I, U64, U64, P]
    lib.cc_col_aux.restype = I
    lib.cc_col_aux.argtypes = [C.POINTER(cc_state), I, P, P, P, P, I, U64, U64, P]
    lib.cc_col_marginals.restype = I
    lib.cc_col_marginals.argtypes = [C.POINTER(cc_state), I, P, P, P, P, P, I, P]
    lib.cc_col_scan.restype = I
    lib.cc_col_scan.argtypes = [C.POINTER(cc_state), I, P, P, I, P, P, P, U64, U64, P, P, P, I, P]
    lib.cc_col_install.restype = I
    lib.cc_col_install.argtypes = [C.POINTER(cc_state), I, P, P, I, P]
    lib.cc_transition_col_hypers.restype = I
    lib.cc_transition_col_hypers.argtypes = [C.POINTER(cc_state), I, P, P, P, P, U64, U64, P]
    lib.cc_transition_alphas.restype = I
    lib.cc_transition_alphas.argtypes = [C.POINTER(cc_state), I, P, P, P, I, U64, U64, P]
    lib.cc_transition.restype = I
    lib.cc_transition.argtypes = [C.POINTER(cc_state), I, P, I, I, P, P, P, U64, C.POINTER(C.c_uint64), GROW_FN, P, P]
    lib.cc_logpdf_bulk.restype = I
    lib.cc_logpdf_bulk.argtypes = [C.POINTER(cc_state), I, P, I, P, P, P, P, P, P, P, P]
    lib.cc_simulate_bulk.restype = I
    lib.cc_simulate_bulk.argtypes = [C.POINTER(cc_state), I, P, I, P, P, P, P, P, P, I, I, P, P, U64, U64, P]
    for name, argtypes in OPTIONAL.items():
        if hasattr(lib, name):
            f = getattr(lib, name)
            f.restype = I
            f.argtypes = argtypes


# entry points added after ABI v1 baseline; declared when present
OPTIONAL = {}

# every symbol include/cgpm_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ['cc_last_error', 'cc_abi_version', 'cc_rebuild', 'cc_refresh_tables', 'cc_check_state', 'cc_logpdf_score',
           'cc_transition_rows', 'cc_col_aux', 'cc_col_marginals', 'cc_col_scan', 'cc_col_install',
           'cc_transition_col_hypers', 'cc_transition_alphas', 'cc_transition', 'cc_logpdf_bulk', 'cc_simulate_bulk']


def load():
    """Load the CUDA extension; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CgpmB200Error(
            'CUDA extension %s not found: run `python -c "import __graft_entry__ as g; '
            'g.build()"` at the repo root (there is no CPU fallback).' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    _declare(lib)
    _lib = lib
    return lib


def check(rc, what=''):
    if rc == OK:
        return
    lib = load()
    msg = lib.cc_last_error().decode('utf-8', 'replace')
    if rc == ERR_ARG:
        raise ValueError('%s: %s' % (what, msg))
    raise CgpmB200Error('%s failed (code %d): %s' % (what, rc, msg))
