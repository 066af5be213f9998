"""The C-ABI library loads and exports every symbol that include/gnnome_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'gnnome_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gnb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_entry_points():
    names = _declared()
    assert 'gnb_edge_forward' in names and 'gnb_graph_stage' in names and len(names) >= 12


def test_library_exports_every_declared_symbol():
    from gnnome_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f'{name} missing from {_lib.LIB_PATH}'
    assert set(_declared()) == set(_lib.SIGNATURES), 'ctypes signature table out of sync with the header'
    assert _lib.load().gnb_abi_version() == _lib.ABI_VERSION


def test_argument_errors_do_not_need_a_gpu():
    from gnnome_b200 import _lib
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    assert lib.gnb_graph_stage_workspace(10, 5, ctypes.byref(nbytes)) == 0 and nbytes.value > 0
    assert lib.gnb_graph_stage_workspace(2 ** 31, 5, ctypes.byref(nbytes)) == -1
    assert b'too large' in lib.gnb_last_error()
    assert lib.gnb_edge_chunk(256) > 0 and lib.gnb_edge_chunk(48) == -1


def test_no_cpu_fallback():
    import torch
    import gnnome_b200
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    model = gnnome_b200.models.SymGatedGCNModel(2, 2, 64, 16, 2, 64, 'batch').eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model((torch.tensor([0]), torch.tensor([0]), 1), torch.zeros(1, 2), torch.zeros(1, 2))


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md section 2 maps every C-ABI symbol to the reference lines it replaces."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    assert [n for n in _declared() if f'`{n}`' not in doc] == []
