"""Run the reference's OWN unmodified ``layers`` / ``models`` packages over the dgl shim.

TEST INFRASTRUCTURE.  Works only where ``/root/reference`` exists (the build container);
the GPU box uses ``oracle/restatement.py`` and the committed ``tests/golden`` fixtures.
"""
import contextlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get('GNNOME_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dgl_shim')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'models', 'full_graph.py'))


def load():
    """Return (dgl_shim_module, reference ``layers`` module, reference ``models`` module)."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    for p in (REFERENCE_ROOT, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in ('dgl', 'layers', 'models'):
        mod = sys.modules.get(name)
        f = getattr(mod, '__file__', '') or ''
        if mod is not None and not (f.startswith(REFERENCE_ROOT) or f.startswith(_SHIM)):
            raise RuntimeError(f'module {name!r} already imported from {f}')
    import dgl
    import layers
    import models
    return dgl, layers, models


def weights_path():
    return os.path.join(REFERENCE_ROOT, 'weights', 'weights.pt')


@contextlib.contextmanager
def quiet():
    """``models/full_graph.py:25`` prints ``x.shape`` on every forward."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_functions(relpath, names, namespace):
    """Execute the *unmodified* source of the named top-level functions of a reference file in ``namespace``.

    For reference modules that cannot be imported whole here (``train.py`` needs ``dgl.data`` and ``wandb`` runs,
    ``utils/data_utils.py`` needs Biopython): the functions' own code still runs, only the module-level imports are
    replaced by what the caller puts in ``namespace`` (torch, F, the dgl shim, ``get_hyperparameters`` ...)."""
    import ast
    path = os.path.join(REFERENCE_ROOT, relpath)
    source = open(path).read()
    tree = ast.parse(source, filename=path)
    found = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, 'exec')
            exec(code, namespace)
            found[node.name] = namespace[node.name]
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f'{relpath}: no top-level function(s) {sorted(missing)}')
    return found
