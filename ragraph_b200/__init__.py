"""ragraph_b200 -- B200-native (sm_100a) implementation of RAGraph's retrieve -> gather -> propagate path.

Kernels live in ``csrc/`` behind the C ABI of ``include/ragraph_b200.h`` (libragraph_b200.so); this package
is the host-side mirror of the reference's Python interface for that path.  Importing works without a GPU;
calling any op needs the built library and CUDA tensors (no CPU fallback).
"""
from . import _lib, ops
from .csr import CSRGraph, as_csr
from .edge import (EdgeAggregator, edge_rag_forward, make_resource_graph, rating_topk, relative_edge_time_encoding, scatter_add,
                   scatter_softmax, scatter_sum)
from .layers import GCN
from .ragraph_utils import Propagation, SimilarityFunctions, TaskDecoder, ToyGraphBase
from .RAGraph import RAGraph, RAGraphFewShot
from .sampling import InverseSampling
from . import downprompt, fewshot
from .graphs import GraphedForward
from .sharded import ShardedRetriever, owner_of, shard_bounds
from .utility import normalized_adjacency_csr, process_graph_batch, process_tu_dataset

__all__ = ["_lib", "ops", "CSRGraph", "as_csr", "EdgeAggregator", "edge_rag_forward", "rating_topk", "scatter_add", "scatter_sum",
           "GCN", "Propagation", "SimilarityFunctions", "TaskDecoder", "ToyGraphBase", "RAGraph", "RAGraphFewShot", "downprompt", "fewshot",
           "relative_edge_time_encoding", "scatter_softmax", "make_resource_graph", "InverseSampling",
           "GraphedForward", "ShardedRetriever", "owner_of", "shard_bounds", "normalized_adjacency_csr", "process_graph_batch", "process_tu_dataset"]
__version__ = "0.1.0"
