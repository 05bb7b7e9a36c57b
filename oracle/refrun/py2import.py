"""Import hook that loads the reference's Python-2 modules from /root/reference in memory (test infrastructure).

Nothing is written to disk and no reference source enters the repo.  Rewrites applied to the text before compiling:
  * `print x, y` statements -> `print(x, y)`;
  * `a.items() + b.items()` -> list concatenation;
  * every `/` -> Python-2 division (floor for two integers, true division otherwise), by an AST pass;
and `xrange` is defined in the module namespace.
"""
import ast
import importlib.abc
import importlib.util
import os
import re
import sys

REF_ROOT = '/root/reference'

_PRINT = re.compile(r'^(\s*)print[ \t]+(?!\()(.*?)\s*$')
_PRINT_PAREN = re.compile(r'^(\s*)print[ \t]+(\(.*\)\s*[%+*].*)$')     # print ('a' % b) + c  style
_ITEMS = re.compile(r'([A-Za-z_][A-Za-z_0-9]*)\.items\(\)')


def py2div(a, b):
    ints = (int,)
    try:
        import numpy as np
        ints = (int, np.integer)
    except ImportError:
        pass
    if isinstance(a, ints) and isinstance(b, ints) and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    return a / b


class _Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id='__py2div__', ctx=ast.Load()), args=[node.left, node.right],
                                              keywords=[]), node)
        return node


def _depth(s):
    """Net bracket depth of a line (string literals in the reference's print statements hold no brackets that matter)."""
    return sum(s.count(c) for c in '([{') - sum(s.count(c) for c in ')]}')


def convert(text):
    out = []
    lines = text.split('\n')
    i = 0
    while i < len(lines):
        line = lines[i]
        i += 1
        m = _PRINT.match(line)
        if m and not line.lstrip().startswith('#'):
            body = m.group(2)
            while _depth(body) > 0 and i < len(lines):          # a print statement continued over several lines
                body += ' ' + lines[i].strip()
                i += 1
            line = '%sprint(%s)' % (m.group(1), body[:-1] + ", end=' '" if body.endswith(',') else body)
        else:
            m = _PRINT_PAREN.match(line)
            if m:
                line = '%sprint(%s)' % (m.group(1), m.group(2))
        line = _ITEMS.sub(r'list(\1.items())', line)
        out.append(line)
    return '\n'.join(out)


class _Loader(importlib.abc.Loader):
    def __init__(self, path):
        self.path = path

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        if os.path.isdir(self.path):          # package
            init = os.path.join(self.path, '__init__.py')
            text = open(init).read() if os.path.exists(init) else ''
            fname = init
        else:
            text = open(self.path).read()
            fname = self.path
        tree = _Div().visit(ast.parse(convert(text), filename=fname))
        ast.fix_missing_locations(tree)
        module.__dict__['__py2div__'] = py2div
        module.__dict__['xrange'] = range
        module.__file__ = fname
        exec(compile(tree, fname, 'exec'), module.__dict__)


class ReferenceFinder(importlib.abc.MetaPathFinder):
    """Resolves top-level names against /root/reference and /root/reference/models (Python-2 implicit relative imports:
    models/DAE_h.py does `import model_helpers`)."""

    def __init__(self, roots):
        self.roots = roots

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split('.')
        for root in self.roots:
            base = os.path.join(root, *parts)
            if os.path.isdir(base) and os.path.commonpath([os.path.realpath(base), REF_ROOT]) == REF_ROOT \
                    and (os.path.exists(os.path.join(base, '__init__.py'))):
                return importlib.util.spec_from_loader(fullname, _Loader(base), is_package=True)
            if os.path.isfile(base + '.py'):
                return importlib.util.spec_from_loader(fullname, _Loader(base + '.py'))
        return None


def install_stubs():
    """Puts the stand-ins (theano, lasagne, FC_DenseNet) on sys.path.  Idempotent; needs no reference tree."""
    stubs = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')
    if stubs not in sys.path:
        sys.path.insert(0, stubs)


def install():
    """Puts the stubs and the reference finder in place.  Idempotent."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError('%s is not present: the reference can only be executed in the build container' % REF_ROOT)
    install_stubs()
    if not any(isinstance(f, ReferenceFinder) for f in sys.meta_path):
        sys.meta_path.append(ReferenceFinder([REF_ROOT, os.path.join(REF_ROOT, 'models')]))
