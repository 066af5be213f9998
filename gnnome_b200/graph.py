"""GraphIndex: the staged graph (dst-CSR + src-CSR, int32, on the GPU) the kernels walk.

Built once per graph from anything that exposes ``edges()`` / ``num_nodes()`` (a DGLGraph, the
oracle's shim graph) or from ``(src, dst, num_nodes)``; cached on the graph object.  It replaces
the index structures DGL builds lazily behind ``apply_edges`` / ``update_all`` / ``dgl.reverse``
(reference layers/gated_gcn_full.py:99,104,112,125)."""
import ctypes

import torch

from . import _lib


def _cuda_device(device=None):
    if device is not None:
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('gnnome_b200 kernels run on CUDA devices only')
        return device
    if not torch.cuda.is_available():
        raise RuntimeError('gnnome_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def current_stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class GraphIndex:
    def __init__(self, src, dst, num_nodes, device=None):
        lib = _lib.load()
        device = _cuda_device(device if device is not None else
                              (src.device if torch.is_tensor(src) and src.is_cuda else None))
        src = torch.as_tensor(src).to(device=device, dtype=torch.int32).contiguous()
        dst = torch.as_tensor(dst).to(device=device, dtype=torch.int32).contiguous()
        if src.ndim != 1 or src.shape != dst.shape:
            raise ValueError('src/dst must be 1-D tensors of equal length')
        self.device = device
        self.N, self.E = int(num_nodes), int(src.numel())
        if self.E and (int(torch.max(src.max(), dst.max())) >= self.N or int(torch.min(src.min(), dst.min())) < 0):
            raise ValueError('edge endpoint out of range')
        i32 = dict(dtype=torch.int32, device=device)
        self.in_ptr = torch.empty(self.N + 1, **i32)
        self.out_ptr = torch.empty(self.N + 1, **i32)
        self.in_src, self.in_dst, self.in_eid, self.out_pos, self.out_dst = (
            torch.empty(max(self.E, 1), **i32) for _ in range(5))
        self.struct = _lib.GnbGraph(self.N, self.E, *(t.data_ptr() for t in (
            self.in_ptr, self.in_src, self.in_dst, self.in_eid, self.out_ptr, self.out_pos, self.out_dst)))
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.gnb_graph_stage_workspace(self.E, self.N, ctypes.byref(nbytes)), 'gnb_graph_stage_workspace')
        with torch.cuda.device(device):
            ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=device)
            _lib.check(lib.gnb_graph_stage(src.data_ptr(), dst.data_ptr(), ctypes.byref(self.struct),
                                           ws.data_ptr(), nbytes.value, current_stream_ptr(device)),
                       'gnb_graph_stage')
        self._keep = (src, dst)  # original-order endpoints (used by reversed())
        self._in_eid_long = None
        del ws

    @property
    def src(self):
        return self._keep[0]

    @property
    def dst(self):
        return self._keep[1]

    def ref(self):
        return ctypes.byref(self.struct)

    def chunk(self, H, backend='tc2'):
        """Carry granularity (edges per aggregation chunk) of the edge pass of this kernel family
        ('tc2': gnb_edge_forward_tc2, 'ffma': gnb_edge_forward)."""
        lib = _lib.load()
        chunk = lib.gnb_edge_chunk_tc2(H) if backend == 'tc2' else lib.gnb_edge_chunk(H)
        if chunk <= 0:
            raise RuntimeError(f'hidden_features={H} unsupported by the {backend!r} kernels')
        return chunk

    def num_chunks(self, H, backend='tc2'):
        return max(1, -(-self.E // self.chunk(H, backend)))

    def reversed(self):
        """Index of ``dgl.reverse(g)`` (train.py:165): same edge ids, endpoints swapped."""
        return GraphIndex(self._keep[1], self._keep[0], self.N, self.device)

    @staticmethod
    def from_graph(graph, device=None):
        if isinstance(graph, GraphIndex):
            return graph
        if isinstance(graph, (tuple, list)) and len(graph) == 3:
            return GraphIndex(graph[0], graph[1], graph[2], device)
        device = _cuda_device(device)
        cache = getattr(graph, '_gnb_index_cache', None)
        if cache is not None and cache[0] == device:
            return cache[1]
        src, dst = graph.edges()
        gi = GraphIndex(src, dst, graph.num_nodes(), device)
        try:
            graph._gnb_index_cache = (device, gi)
        except AttributeError:
            pass
        return gi
