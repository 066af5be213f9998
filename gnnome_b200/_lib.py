"""ctypes binding of ``libgnnome_b200.so`` (the C ABI declared in ``include/gnnome_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libgnnome_b200.so')

_c_i32p = ctypes.c_void_p
_c_f32p = ctypes.c_void_p


class GnbGraph(ctypes.Structure):
    """Mirror of ``gnb_graph_t``."""
    _fields_ = [
        ('num_nodes', ctypes.c_int64),
        ('num_edges', ctypes.c_int64),
        ('in_ptr', ctypes.c_void_p),
        ('in_src', ctypes.c_void_p),
        ('in_dst', ctypes.c_void_p),
        ('in_eid', ctypes.c_void_p),
        ('out_ptr', ctypes.c_void_p),
        ('out_pos', ctypes.c_void_p),
        ('out_dst', ctypes.c_void_p),
    ]


class GnbWalkGraph(ctypes.Structure):
    """Mirror of ``gnb_walk_graph_t`` (host pointers)."""
    _fields_ = [('num_nodes', ctypes.c_int64), ('succ_ptr', ctypes.c_void_p), ('succ_node', ctypes.c_void_p),
                ('succ_edge', ctypes.c_void_p)]


# name -> (restype, argtypes); must list every symbol of include/gnnome_b200.h
_P, _I, _L, _S = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
SIGNATURES = {
    'gnb_abi_version': (_I, []),
    'gnb_last_error': (ctypes.c_char_p, []),
    'gnb_graph_stage_workspace': (_I, [_L, _L, ctypes.POINTER(_S)]),
    'gnb_graph_stage': (_I, [_P, _P, ctypes.POINTER(GnbGraph), _P, _S, _P]),
    'gnb_encode': (_I, [_P, _P, _L, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'gnb_node_linear': (_I, [_P, _L, _I, _P, _P, _I, _P, _L, _P]),
    'gnb_packed_linear_bytes': (_S, [_I, _I]),
    'gnb_pack_linear_tc': (_I, [_P, _I, _I, _P, _P]),
    'gnb_set_spin_timeout_ms': (None, [ctypes.c_longlong]),
    'gnb_hang_report': (_I, [ctypes.c_char_p, _S]),
    'gnb_edge_chunk': (_I, [_I]),
    'gnb_edge_forward': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _P, _P, _P, _P, _P, _I, _P]),
    'gnb_node_update': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _L, _P, _P, _P, _P]),
    'gnb_reverse_partial': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _L, _L, _P, _P]),
    'gnb_score_forward': (_I, [ctypes.POINTER(GnbGraph), _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'gnb_split16_bytes': (_S, [_L, _I]),
    'gnb_split_rows': (_I, [_P, _P, _L, _I, _P, _P, _P]),
    'gnb_merge_rows': (_I, [_P, _P, _L, _I, _P, _P]),
    'gnb_encode2': (_I, [_P, _P, _L, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'gnb_node_linear_tc2': (_I, [_P, _L, _I, _P, _P, _I, _P, _L, _P, _P]),
    'gnb_edge_tile_tc2': (_I, [_I]),
    'gnb_edge_chunk_tc2': (_I, [_I]),
    'gnb_edge_forward_tc2': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _P, _P, _P, _I, _P]),
    'gnb_debug_edge_timing': (None, [_P]),
    'gnb_debug_store_delay_ns': (None, [_I]),
    'gnb_debug_edge_mode': (None, [_I]),
    'gnb_node_update2': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _L, _P, _P, _P, _P]),
    'gnb_reverse_partial2': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _L, _L, _P, _P]),
    'gnb_score_forward_tc2': (_I, [ctypes.POINTER(GnbGraph), _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'gnb_score_forward2': (_I, [ctypes.POINTER(GnbGraph), _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'gnb_t_gather_add3': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _L, _P, _P, _P]),
    'gnb_t_seg_sum': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _I, _P, _L, _P]),
    'gnb_t_agg_fwd': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _L, _P, _I, _P, _P, _P]),
    'gnb_t_agg_bwd_edge': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _P, _P, _P, _L, _I, _P, _I, _P]),
    'gnb_t_agg_bwd_node': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _P, _P, _I, _P, _L, _P]),
    'gnb_t_gate_fwd': (_I, [_P, _P, _L, _I, _P, _P, _P]),
    'gnb_t_gate_bwd': (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _P]),
    'gnb_t_affine2': (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P]),
    'gnb_t_layer_norm_fwd': (_I, [_P, _P, _P, _L, _I, ctypes.c_float, _P, _P, _P, _P]),
    'gnb_t_layer_norm_bwd': (_I, [_P, _P, _P, _P, _L, _I, _P, _P]),
    'gnb_t_col_stats_workspace': (_S, [_L, _I]),
    'gnb_t_col_stats': (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _P]),
    'gnb_gather_rows': (_I, [_P, _P, _L, _I, _P, _P]),
    'gnb_gather_rows_ld': (_I, [_P, _L, _P, _L, _I, _P, _L, _P]),
    'gnb_scatter_rows': (_I, [_P, _P, _L, _I, _P, _P]),
    'gnb_degree_rows': (_I, [ctypes.POINTER(GnbGraph), _I, _P, _P]),
    'gnb_zscore_workspace': (_S, []),
    'gnb_zscore_cols': (_I, [_P, _L, _I, _I, _P, _P, _P]),
    'gnb_subgraph_workspace': (_I, [_L, _L, ctypes.POINTER(_S)]),
    'gnb_subgraph_count': (_I, [_P, _P, _P, _L, _L, _P, _S, _P, _P]),
    'gnb_subgraph_fill': (_I, [_P, _P, _P, _L, _L, _P, _P, _P, _P, _P, _P]),
    'gnb_greedy_walks': (_I, [ctypes.POINTER(GnbWalkGraph), _P, _P, _L, _P, _P, _I, _P, _L, _P, _P, _P]),
    'gnb_walk_contig_length': (_I, [ctypes.POINTER(GnbWalkGraph), _P, _P, _P, _L, _P]),
    'gnb_walk_jumped_nodes': (_I, [ctypes.POINTER(GnbWalkGraph), ctypes.POINTER(GnbWalkGraph), _P, _L, _P]),
}

ABI_VERSION = 11
GNB_F_SYMMETRIC = 1
GNB_F_RESIDUAL = 2

_lib = None


def load():
    """Load the library once; raise with a build hint if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'or `make -C {os.path.dirname(LIB_PATH)}` (there is no CPU fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    if lib.gnb_abi_version() != ABI_VERSION:
        raise RuntimeError(f'libgnnome_b200 ABI {lib.gnb_abi_version()} != {ABI_VERSION}: rebuild')
    _lib = lib
    return lib


def hang_report():
    """The spin watchdog's record of a device-side wait that timed out in this process ('' if none):
    which kernel / block / warp role waited for which barrier.  Readable after the CUDA context has died."""
    buf = ctypes.create_string_buffer(512)
    return buf.value.decode() if load().gnb_hang_report(buf, len(buf)) else ''


def check(rc, what):
    if rc != 0:
        msg = load().gnb_last_error()
        hang = hang_report()
        raise RuntimeError(f'{what} failed (rc={rc}): {msg.decode() if msg else "?"}' + (f' [{hang}]' if hang else ''))
