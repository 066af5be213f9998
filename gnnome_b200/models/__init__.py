"""Same public names as the reference's ``models`` package (``models/__init__.py:1``).
GCNModel / GATModel / SAGEModel (ablation baselines) are outside the hot path."""
from .full_graph import GatedGCNModel, SymGatedGCNModel

__all__ = ['SymGatedGCNModel', 'GatedGCNModel']
