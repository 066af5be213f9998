"""Same public names as the reference's ``layers`` package (``layers/__init__.py:1-5``)."""
from .gated_gcn import GatedGCN, SymGatedGCN, get_backend, set_backend
from .processor import GatedGCN_processor, SymGatedGCN_processor
from .score_predictor import ScorePredictor
from .encoders import EdgeEncoder, NodeEncoder

__all__ = ['SymGatedGCN', 'GatedGCN', 'SymGatedGCN_processor', 'GatedGCN_processor', 'ScorePredictor',
           'NodeEncoder', 'EdgeEncoder']
