"""Mirror of the fusion arithmetic of RAGraph.forward.

node:  RAGraph_node/RAGraph.py:10-63      (retrieve_weight = label_weight = 0.5, hop 3)
graph: RAGraph_graph/RAGraph.py:48-75     (0.3 / 0.3, hop 1, one query = mean of the node embeddings)
few-shot (class RAGraphFewShot): RAGraph_node_fewshot/RAGraph.py:47-83, RAGraph_graph_fewshot/RAGraph.py:46-91

Constructor differences from the reference are confined to library construction (out of the hot-path
scope): the toy-graph base is passed in (or filled with ``toy_graph_base.add_entries``) instead of being
built from a torch_geometric dataset.  forward(features, adj) keeps the reference signature.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .ragraph_utils import Propagation, TaskDecoder, ToyGraphBase


class RAGraph(nn.Module):
    def __init__(self, pretrain_model, toy_graph_base: ToyGraphBase, feture_size, num_class, emb_size,
                 finetune=True, noise_finetune=False, variant: str = "node") -> None:
        super().__init__()
        assert variant in ("node", "graph")
        self.variant = variant
        self.emb_size, self.num_class = emb_size, num_class
        self.pretrain_model = pretrain_model
        self.retrieve_weight = 0.5 if variant == "node" else 0.3
        self.label_weight = 0.5 if variant == "node" else 0.3
        self.finetune = finetune
        self.noise_finetune = noise_finetune
        if self.noise_finetune:
            assert self.finetune
        self.query_graph_hop = 3 if variant == "node" else 1
        self.toy_graph_base = toy_graph_base
        if self.finetune:
            self.decoder = TaskDecoder(emb_size, emb_size, num_class)

    def reset_parameters(self):
        self.decoder.reset_parameters()

    def forward(self, features, adj):
        pretrain_embedddings = self.pretrain_model.inference(features, adj)
        add_noise = self.training and self.noise_finetune
        query = pretrain_embedddings if self.variant == "node" else torch.mean(pretrain_embedddings, dim=0)

        if add_noise:
            # noisy branch: keep the reference's explicit [Q,k',d] tensors (extra rows / Gaussian noise)
            rag_embeddings, rag_labels = self.toy_graph_base.retrieve(query, adj, True)
            rag_label = torch.mean(rag_labels.float(), dim=1)
            rag_embedding = torch.sum(rag_embeddings, dim=1)
            fused = None
        else:
            rag_embedding = rag_label = fused = None

        if not self.finetune:
            if rag_label is None:
                _, rag_label, _ = self.toy_graph_base.retrieve_fused(query)
            return rag_label

        query_embeddings = Propagation.aggregate_k_hop_features(adj, pretrain_embedddings, self.query_graph_hop)
        if self.variant == "graph":
            query_embeddings = torch.mean(query_embeddings, dim=0, keepdim=True)
        if rag_embedding is None:
            # hidden = (1-w)*prop + w*sum_k values[idx]: gather, reduce and blend are one kernel
            hidden_embedding, rag_label, _ = self.toy_graph_base.retrieve_fused(
                query, reduce=L.REDUCE_SUM, blend_in=query_embeddings, blend_w=self.retrieve_weight)
        else:
            hidden_embedding = query_embeddings * (1 - self.retrieve_weight) + rag_embedding * self.retrieve_weight
        decode_label = torch.softmax(self.decoder(hidden_embedding), dim=1)
        return decode_label * (1 - self.label_weight) + rag_label * self.label_weight


class RAGraphFewShot(nn.Module):
    """Few-shot fusion: retrieved labels -> class id -> mean few-shot logits, decode = the backbone's second GCN layer.

    node_fewshot  (RAGraph_node_fewshot/RAGraph.py:8-83):  graph_level=False, weights 0.5/0.5 (ENZYMES) or 0.3/0.8
                  (PROTEINS), hop 3, retrieve_num 5, two-metric scores when the store carries position codes.
    graph_fewshot (RAGraph_graph_fewshot/RAGraph.py:7-91): graph_level=True -- every node of the single query graph
                  retrieves, the blended logits are averaged over the nodes ([1, C]); hop 1.
    ``pretrain_model`` provides ``encode(features, adj)`` / ``decode(hidden, adj)`` (models/gcnlayers.py:61-85)."""

    def __init__(self, pretrain_model, toy_graph_base: ToyGraphBase, emb_size, finetune=True, noise_finetune=False,
                 query_graph_hop=3, retrieve_weight=0.5, label_weight=0.5, graph_level=False) -> None:
        super().__init__()
        self.emb_size = emb_size
        self.pretrain_model = pretrain_model
        self.retrieve_weight, self.label_weight = retrieve_weight, label_weight
        self.finetune = finetune
        self.noise_finetune = noise_finetune
        if self.noise_finetune:
            assert self.finetune
        self.query_graph_hop = query_graph_hop
        self.graph_level = graph_level
        self.toy_graph_base = toy_graph_base

    def forward(self, features, adj, mean_fewshot_logits, search_positions=None):
        pretrain_embedddings = self.pretrain_model.encode(features, adj)
        add_noise = self.training and self.noise_finetune
        base = self.toy_graph_base
        mean_fewshot_logits = mean_fewshot_logits.to(torch.float32).contiguous()

        if add_noise:
            # noisy branch: the reference's explicit [Q,k',d] / [Q,k',C] tensors (extra rows or Gaussian noise)
            rag_embeddings, rag_labels = base.retrieve(pretrain_embedddings, adj, True, search_positions)
            rag_logits = torch.mean(mean_fewshot_logits[torch.argmax(rag_labels, dim=-1)], dim=1)
            idx = None
        else:
            search_positions = base.query_positions(adj, search_positions)
            _, idx = base.topk(pretrain_embedddings, base.retrieve_num, search_positions)
            # labels[idx].argmax(-1) == class_id[idx]: one bit-exact gather of the per-row class ids, then the
            # mean over k of the few-shot logits as one gather-reduce over the [C, C'] logits table
            cls = ops.gather_rows(base.class_ids().unsqueeze(1), idx).squeeze(-1)
            rag_logits = ops.gather_reduce(mean_fewshot_logits, cls, L.REDUCE_MEAN)
        if not self.finetune:
            return rag_logits

        query_embeddings = Propagation.aggregate_k_hop_features(adj, pretrain_embedddings, self.query_graph_hop)
        if idx is None:
            rag_embedding = torch.sum(rag_embeddings, dim=1)
            hidden_embedding = query_embeddings * (1 - self.retrieve_weight) + rag_embedding * self.retrieve_weight
        else:
            hidden_embedding = base.gather_reduce_blend(idx, L.REDUCE_SUM, query_embeddings, self.retrieve_weight)
        decode_logits = self.pretrain_model.decode(hidden_embedding, adj)
        label_logits = decode_logits * (1 - self.label_weight) + rag_logits * self.label_weight
        if self.graph_level:
            label_logits = label_logits.mean(dim=0).unsqueeze(0)
        return label_logits
