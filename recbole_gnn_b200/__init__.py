"""recbole_gnn_b200 — Blackwell-native bipartite graph-convolution engine behind RecBole-GNN's
layer / model / dataset API for the K-layer normalised-adjacency propagation path.

Python host (this package) over ``libb200gcn.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/b200gcn.h``).  No PyG / torch_sparse / Triton dispatch and no CPU fallback: CPU tensors raise.
"""
from .graph import GraphHandle, SparseTensor, gcn_norm
from .layers import BiGNNConv, BipartiteGCNConv, LightGCNConv, handle_from_edges
from .dataset import GeneralGraphDataset, GraphBuildMixin, InteractionDataset
from .abstract_recommender import GeneralGraphRecommender
from .models import LightGCN, NGCF, SimGCL
from . import augment, functional, ops, train

__all__ = [
    "GraphHandle", "SparseTensor", "gcn_norm", "LightGCNConv", "BipartiteGCNConv", "BiGNNConv",
    "handle_from_edges", "GeneralGraphDataset", "GraphBuildMixin", "InteractionDataset",
    "GeneralGraphRecommender", "LightGCN", "NGCF", "SimGCL", "functional", "augment", "train", "ops",
]
