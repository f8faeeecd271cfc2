#!/usr/bin/env python
"""Materialise a Python-3 runnable build of the reference into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under cgpm_b200/ may import this module or
its output.  The reference (/root/reference, probcomp/cgpm) is Python-2.7-only;
this script applies a mechanical source transform at build time, reading the
sources where they lie and writing the result into oracle/_ref/cgpm/ (listed in
.gitignore: no reference source is ever committed).  The transform is purely
syntactic:

  regex pre-pass   print statements, tuple parameters, cPickle, func_code,
                   implicit relative import, text-mode pickle files
  ast pass         xrange/iteritems/itervalues/iterkeys/unicode renames and
                   list(...) around .keys()/.values()/.items()/map/filter/zip/
                   range so that py2 list semantics are kept
  two hand patches `hypothetical` (None < int comparison) and the serial mapper
                   (`map` is lazy on py3)

Usage:  python oracle/build_ref.py [--src /root/reference/src] [--force]
"""
from __future__ import annotations

import argparse
import ast
import os
import re
import shutil
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SRC = '/root/reference/src'
DEST = os.path.join(HERE, '_ref', 'cgpm')

# Sub-packages that need packages absent from this image are still converted
# (they are imported lazily by the reference), but a failure to convert one of
# them is not fatal.
REQUIRED = ('crosscat', 'mixtures', 'primitives', 'utils', 'network')


def _fix_tuple_params(src: str) -> str:
    """def f((a, b)):  ->  def f(_targ): (a, b) = _targ"""
    pat = re.compile(
        r'^(?P<ind>[ \t]*)def (?P<name>\w+)\(\((?P<args>[^()]*)\)\):[ \t]*$',
        re.M)

    def repl(m):
        ind = m.group('ind')
        return '%sdef %s(_targ):\n%s    (%s) = _targ' % (
            ind, m.group('name'), ind, m.group('args'))
    return pat.sub(repl, src)


def _fix_print(src: str) -> str:
    """print <expr>  ->  print(<expr>)   (handles trailing backslash joins)."""
    lines = src.split('\n')
    out = []
    i = 0
    while i < len(lines):
        line = lines[i]
        m = re.match(r'^([ \t]*)print (?!\()(.*)$', line) or \
            re.match(r'^([ \t]*)print (\(.*\).*[%,].*)$', line)
        if m and not line.lstrip().startswith('#'):
            ind, body = m.group(1), m.group(2)
            while body.rstrip().endswith('\\'):
                i += 1
                body = body.rstrip()[:-1] + ' ' + lines[i].strip()
            depth = body.count('(') - body.count(')')
            while depth > 0:
                i += 1
                body += ' ' + lines[i].strip()
                depth = body.count('(') - body.count(')')
            out.append('%sprint(%s)' % (ind, body))
        else:
            out.append(line)
        i += 1
    return '\n'.join(out)


def _regex_pass(src: str, relpath: str) -> str:
    src = src.replace('import cPickle as pickle', 'import pickle')
    src = src.replace('.func_code', '.__code__')
    src = src.replace('.__func__.__doc__', '.__doc__')
    src = src.replace('from relevance import', 'from cgpm.mixtures.relevance import')
    src = re.sub(r"open\(([^,()]+), 'r'\)", r"open(\1, 'rb')", src) \
        if 'pickle' in src else src
    src = _fix_tuple_params(src)
    src = _fix_print(src)
    if relpath.endswith(os.path.join('crosscat', 'engine.py')):
        # `map` is the serial mapper in the reference; it is lazy on py3.
        src = re.sub(r'parallel_map if multiprocess else map\b',
                     'parallel_map if multiprocess else _serial_map', src)
        src = src.replace(
            '# Multiprocessing functions.',
            'def _serial_map(f, l):\n    return [f(x) for x in l]\n\n'
            '# Multiprocessing functions.', 1)
    # None < int comparisons (py2 orders None before ints).
    src = src.replace(
        'return not (0 <= rowid < len(self.Zr()))',
        'return rowid is None or not (0 <= rowid < len(self.Zr()))')
    src = src.replace(
        'return not 0 <= rowid < self.n_rows()',
        'return rowid is None or not 0 <= rowid < self.n_rows()')
    return src


class _Py2Semantics(ast.NodeTransformer):
    RENAME = {'xrange': 'range', 'unicode': 'str', 'long': 'int',
              'basestring': 'str', 'raw_input': 'input'}
    ITER = {'iteritems': 'items', 'itervalues': 'values', 'iterkeys': 'keys'}
    LISTY_METHODS = {'keys', 'values', 'items'}
    LISTY_FUNCS = {'map', 'filter', 'zip', 'range'}

    def visit_Name(self, node):
        if node.id in self.RENAME:
            node.id = self.RENAME[node.id]
        return node

    def visit_Call(self, node):
        self.generic_visit(node)
        f = node.func
        if isinstance(f, ast.Attribute) and f.attr in self.ITER:
            f.attr = self.ITER[f.attr]
            return node
        wrap = False
        if isinstance(f, ast.Attribute) and f.attr in self.LISTY_METHODS \
                and not node.args and not node.keywords:
            wrap = True
        if isinstance(f, ast.Name) and f.id in self.LISTY_FUNCS:
            wrap = True
        if wrap:
            return ast.copy_location(
                ast.Call(func=ast.Name(id='list', ctx=ast.Load()),
                         args=[node], keywords=[]), node)
        return node


def convert_source(src: str, relpath: str) -> str:
    src = _regex_pass(src, relpath)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', SyntaxWarning)
        tree = ast.parse(src)
    tree = _Py2Semantics().visit(tree)
    ast.fix_missing_locations(tree)
    return ast.unparse(tree) + '\n'


def build(src_root: str = DEFAULT_SRC, force: bool = False, quiet: bool = False) -> str | None:
    """Build oracle/_ref/cgpm from `src_root`.  Returns the path that must be
    put on sys.path (oracle/_ref), or None if the reference tree is absent."""
    if not os.path.isdir(src_root):
        return None
    stamp = os.path.join(DEST, '.built')
    if os.path.exists(stamp) and not force:
        return os.path.dirname(DEST)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    n_ok, failed = 0, []
    for dirpath, _dirnames, filenames in os.walk(src_root):
        rel = os.path.relpath(dirpath, src_root)
        outdir = os.path.normpath(os.path.join(DEST, rel))
        os.makedirs(outdir, exist_ok=True)
        for fn in filenames:
            if not fn.endswith('.py'):
                continue
            relpath = os.path.normpath(os.path.join(rel, fn))
            with open(os.path.join(dirpath, fn), 'r', encoding='utf-8') as f:
                text = f.read()
            try:
                out = convert_source(text, relpath)
            except SyntaxError as e:  # a file outside the hot path
                failed.append((relpath, str(e)))
                if relpath.split(os.sep)[0] in REQUIRED:
                    raise
                continue
            with open(os.path.join(outdir, fn), 'w', encoding='utf-8') as f:
                f.write(out)
            n_ok += 1
    with open(os.path.join(DEST, 'version.py'), 'w') as f:
        f.write("__version__ = '0.0.0+py3shim'\n")
    with open(stamp, 'w') as f:
        f.write('built from %s; %d files converted; %d skipped\n'
                % (src_root, n_ok, len(failed)))
    if not quiet:
        print('oracle/_ref: converted %d files, skipped %d %s'
              % (n_ok, len(failed), [p for p, _ in failed]))
    return os.path.dirname(DEST)


def ref_path() -> str | None:
    """Path to put on sys.path to `import cgpm` (the shimmed reference), or
    None when oracle/_ref has not been built."""
    return os.path.dirname(DEST) if os.path.exists(os.path.join(DEST, '.built')) else None


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default=DEFAULT_SRC)
    ap.add_argument('--force', action='store_true')
    a = ap.parse_args()
    p = build(a.src, a.force)
    if p is None:
        print('reference tree %s not present; nothing built' % a.src)
        sys.exit(0)
    print(p)
