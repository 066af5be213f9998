"""gnnome_b200 -- B200-native GatedGCN message passing + edge scoring for GNNome.

``gnnome_b200.models`` / ``gnnome_b200.layers`` mirror the reference's ``models`` / ``layers``
packages; the arithmetic runs in ``csrc/libgnnome_b200.so`` (C ABI: ``include/gnnome_b200.h``)."""
from . import layers, models  # noqa: F401
from .graph import GraphIndex  # noqa: F401
from .layers.gated_gcn import get_backend, set_backend  # noqa: F401

__version__ = '0.1.0'
