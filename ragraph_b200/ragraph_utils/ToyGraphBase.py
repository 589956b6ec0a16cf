"""Mirror of the reference's toy-graph vector store and its retrieve().

Reference: RAGraph_node/ragraph_utils/ToyGraphBase.py:15-119 (node), RAGraph_graph/.../ToyGraphBase.py:56-87
(graph: one 1-D query), RAGraph_node_fewshot/.../ToyGraphBase.py:47-79 (two-metric scores).

What changes on B200: the four ever-growing ``torch.cat`` tensors (:35-38, :116-119) become a pre-sized
device-resident store (capacity doubling), extended with two derived arrays the kernels want -- the fp32
inverse key norms (so keys are not re-normalised on every call, SimilarityFunctions.py:11) and the
normalised bf16 shadow read by the tcgen05 filter.  retrieve() keeps the reference signature and return
shapes; similarity + top-k is one fused launch (scores never reach HBM) and the value/label lookups are
the bit-exact gather kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L
from .. import ops


class ToyGraphBase:
    def __init__(self, pretrain_model=None, num_class: int = 3, emb_size: int = 256, query_graph_hop: int = 3,
                 device: Optional[torch.device] = None, variant: str = "node", capacity: int = 1024,
                 mode: Optional[int] = None, label_dtype: torch.dtype = torch.float32,
                 retrieve_num: Optional[int] = None) -> None:
        assert variant in ("node", "graph", "node_fewshot")
        self.variant = variant
        # inference-phase knobs, same names and defaults as the reference (:22-29)
        # construct-phase knobs (:17-18; graph variant RAGraph_graph/.../ToyGraphBase.py:20-21): 0 disables
        self.num_inverse_sample = 0 if variant == "graph" else 10
        self.num_augment_scale = 0 if variant == "graph" else 3
        # node: C + 1 (:22); graph / graph few-shot: min(3, C + 1) (RAGraph_graph/.../ToyGraphBase.py:25); node few-shot: a
        # constructor argument the caller sets to 5 (RAGraph_node_fewshot/.../ToyGraphBase.py:16,22, RAGraph.py:17,41)
        if retrieve_num is not None:
            self.retrieve_num = int(retrieve_num)
        else:
            self.retrieve_num = {"node": num_class + 1, "graph": min(3, num_class + 1), "node_fewshot": 5}[variant]
        self.noise_retrieve_num = 1
        self.num_anchors = 10
        self.dis_q = 10
        self.structure_weight = 0.001 if variant == "node_fewshot" else 0.0
        self.semantic_weight = 0.999
        # only the graph variants add Gaussian noise to the retrieved values, with std 0.01 (RAGraph_graph/.../ToyGraphBase.py:
        # 27,131-134; RAGraph_graph_fewshot the same); the node variants append random rows instead and never read it
        self.noise_std = 0.01 if variant == "graph" else 0.1
        self.toy_graph_hop = query_graph_hop - 1
        self.pretrain_model = pretrain_model
        self.mode = mode                      # None = pick per library size (see _pick_mode)

        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.emb_size, self.num_class = emb_size, num_class
        self._n = 0
        self._cap = 0
        self._label_dtype = label_dtype
        self._keys = self._values = self._labels = self._positions = None
        self._inv_norm = self._keys_tf32 = None
        self._shadow16 = {}                   # FMT_F16 / FMT_BF16 -> (shadow [n, d_pad], err_max float32 [1])
        self._derived_rows = 0                # rows [0, _derived_rows) of inv_norm / bf16 shadow are valid
        self._class_ids = None                # argmax of the label rows (few-shot fusion), valid for _class_rows rows
        self._class_rows = 0
        self.shard_lo = 0                     # global index of local row 0 (key-row sharded libraries)
        # diagnostics: with collect_stats the next topk() leaves an int32 [2] device tensor in last_stats = {rows that took
        # the second tensor-core pass, rows recomputed by the fp32 kernel} (no host sync; read it when convenient)
        self.collect_stats = False
        self.last_stats: Optional[Tensor] = None
        self._pol_free = []
        self._policy_reset()
        self._small_ws: Optional[Tensor] = None   # zero-filled scratch of the single-launch small-problem kernel
        self._small_plan = {}                     # (Q, n, k) -> workspace bytes, or 0 if the shape is not covered
        self._reserve(capacity)

    # ---- store ---------------------------------------------------------------------------
    def _reserve(self, cap: int) -> None:
        if cap <= self._cap:
            return
        cap = max(cap, 2 * self._cap)
        dev, n = self.device, self._n

        def grow(old, shape, dtype):
            new = torch.empty(shape, dtype=dtype, device=dev)
            if old is not None and n:
                new[:n].copy_(old[:n])
            return new

        self._keys = grow(self._keys, (cap, self.emb_size), torch.float32)
        self._values = grow(self._values, (cap, self.emb_size), torch.float32)
        self._labels = grow(self._labels, (cap, self.num_class), self._label_dtype)
        if self.variant == "node_fewshot":           # position codes only feed the two-metric score
            self._positions = grow(self._positions, (cap, self.num_anchors), torch.float32)
        self._inv_norm = grow(self._inv_norm, (cap,), torch.float32)
        self._keys_tf32 = None                      # rebuilt lazily at the size in use
        self._shadow16 = {}
        self._cap = cap

    def clear(self, release: bool = False) -> None:
        """Forget every library row (capacity is kept unless ``release``): the store can be refilled in place."""
        self._n = 0
        self._derived_rows = 0
        self._class_ids, self._class_rows = None, 0
        self._keys_tf32 = None
        self._shadow16 = {}
        self._policy_reset()
        if release:
            self._keys = self._values = self._labels = self._positions = self._inv_norm = None
            self._cap = 0

    def add_entries(self, keys: Tensor, values: Tensor, labels: Tensor, positions: Optional[Tensor] = None) -> None:
        """Append library rows (the torch.cat of ToyGraphBase.py:116-119)."""
        m = keys.shape[0]
        self._reserve(self._n + m)
        s = slice(self._n, self._n + m)
        self._keys[s].copy_(keys)
        self._values[s].copy_(values)
        self._labels[s].copy_(labels.to(self._label_dtype))
        if self._positions is not None:
            if positions is not None:
                self._positions[s].copy_(positions)
            else:
                self._positions[s].zero_()
        self._n += m

    def add_graph(self, embeddings: Tensor, adj, node_labels: Optional[Tensor] = None, graph_label: Optional[Tensor] = None,
                  positions: Optional[Tensor] = None) -> None:
        """Insert one resource graph: the deterministic core of ``_build_toy_graph_base`` (ToyGraphBase.py:91-119; graph
        variant RAGraph_graph/.../ToyGraphBase.py:97-129) for a graph whose backbone embeddings are given -- keys =
        F.normalize(embeddings), values = k-hop propagation of the keys over ``adj`` (toy_graph_hop = query hop - 1),
        labels = the node labels (node variants) or one_hot(graph_label) with keys / values averaged over the nodes
        (graph variant: one library row per graph).  The RNG-driven augmentation / inverse sampling in front of it are
        library-build policy and stay with the caller: call this once per augmented or sampled copy."""
        from .Propagation import Propagation
        keys = ops.rows_normalize(embeddings)
        values = Propagation.aggregate_k_hop_features(adj, keys, self.toy_graph_hop)
        if self.variant == "graph":
            if graph_label is None:
                raise RuntimeError("add_graph: the graph variant stores one_hot(graph_label) per graph")
            keys = torch.mean(keys, dim=0).unsqueeze(0)
            values = torch.mean(values, dim=0).unsqueeze(0)
            labels = torch.nn.functional.one_hot(graph_label.reshape(-1)[:1].to(torch.int64), num_classes=self.num_class)
        else:
            if node_labels is None:
                raise RuntimeError("add_graph: node_labels [n, C] needed")
            labels = node_labels
        self.add_entries(keys, values, labels, positions)

    def _build_toy_graph_base(self, features: Tensor, adj: Tensor, labels: Tensor) -> None:
        """One resource graph -> library rows, the reference's build step end to end (ToyGraphBase.py:91-119; graph variant
        RAGraph_graph/.../ToyGraphBase.py:97-129): the original graph plus ``num_augment_scale`` random augmentations, each
        embedded by ``pretrain_model.inference``, optionally reduced to ``num_inverse_sample`` nodes drawn by inverse
        importance (sub-adjacency taken from the ORIGINAL adjacency, as the reference does), then inserted with
        ``add_graph``.  ``labels`` = node labels [n, C] (node variants) or the graph label (graph variant).  Random draws
        are issued in the reference's order, so a seeded build gives the reference's library."""
        from ..sampling import InverseSampling
        from .Augmentation import Augmentation
        from .PositionAwareEncoder import PositionAwareEncoder
        from ..csr import CSRGraph, as_dense
        # the build's random edge rewrite, sub-adjacency extraction and position codes are dense by nature and run on toy
        # graphs of a few dozen nodes: a CSR adjacency (what utility.process_tu_dataset returns) is densified once here
        if isinstance(adj, CSRGraph) or (isinstance(adj, Tensor) and adj.layout != torch.strided):
            adj = as_dense(adj)
        for aug_features, aug_adj in Augmentation.augment_graph(self.num_augment_scale, features, adj):
            # few-shot backbones expose the first GCN layer as ``encode`` (RAGraph_node_fewshot/.../ToyGraphBase.py:92)
            embed = getattr(self.pretrain_model, "encode", None) if self.variant == "node_fewshot" else None
            embeddings = (embed or self.pretrain_model.inference)(aug_features, aug_adj)
            if embeddings.dim() == 3 and embeddings.shape[0] == 1:
                embeddings = embeddings[0]
            node_labels = None if self.variant == "graph" else labels
            if self.num_inverse_sample > 0:
                sample_prob = InverseSampling.compute_sample_prob(aug_adj)
                sample_mask = torch.multinomial(sample_prob, num_samples=self.num_inverse_sample, replacement=True)
                sample_adj = adj[sample_mask, :][:, sample_mask]
                embeddings = embeddings[sample_mask]
                if node_labels is not None:
                    node_labels = node_labels[sample_mask]
            else:
                sample_adj = aug_adj
            positions = None
            if self.variant != "graph":
                # the reference encodes positions for every node variant (:114) -- one CPU randint per inserted graph; the
                # codes are stored only where the two-metric score reads them
                if self._positions is not None:
                    positions = PositionAwareEncoder.encode_position_aware_code(sample_adj, self.num_anchors, self.dis_q)
                else:
                    torch.randint(low=0, high=sample_adj.shape[0], size=(int(self.num_anchors),))   # keeps the RNG stream aligned
            self.add_graph(embeddings, sample_adj, node_labels=node_labels,
                           graph_label=labels if self.variant == "graph" else None, positions=positions)

    def build_toy_graph(self, resource_graphs) -> None:
        """``resource_graphs``: iterable of (features, adj, labels) per resource graph, i.e. what the reference's loop over
        ``DataLoader(resource_dataset, batch_size=1)`` + ``process_tu_dataset`` produces (:40-45)."""
        for features, adj, labels in resource_graphs:
            self._build_toy_graph_base(features, adj, labels)

    def _refresh_derived(self, fmt16: Optional[int] = None, want_tf32: bool = False) -> None:
        n = self._n
        if self._derived_rows < n:
            lo = self._derived_rows
            self._inv_norm[lo:n].copy_(ops.row_inv_norm(self._keys[lo:n]))
            self._derived_rows = n
            self._keys_tf32 = None
            self._shadow16 = {}
        if fmt16 is not None and (fmt16 not in self._shadow16 or self._shadow16[fmt16][0].shape[0] != n):
            # normalised 16-bit image of the keys + the largest rounding-error norm of a row (certificate bound)
            err = torch.zeros(1, dtype=torch.float32, device=self.device)
            shadow, _ = ops.rows_to_shadow16(self._keys[:n], fmt16, True, err_max=err)
            self._shadow16 = {fmt16: (shadow, err)}            # one 16-bit shadow at a time (25.6 GB at 100 M x 128)
        if want_tf32 and (self._keys_tf32 is None or self._keys_tf32.shape[0] != n):
            self._keys_tf32 = ops.rows_to_tf32(self._keys[:n], True)

    _FMT_OF_MODE = {L.SIM_BF16: L.FMT_BF16, L.SIM_BF16_REFINE: L.FMT_BF16, L.SIM_F16: L.FMT_F16, L.SIM_F16_REFINE: L.FMT_F16}

    def _shadow(self, mode: int) -> Tuple[Optional[Tensor], Optional[Tensor]]:
        """(key shadow, its err_max) the similarity mode reads -- (None, None) for the fp32 path -- refreshed for the rows
        in use"""
        fmt = self._FMT_OF_MODE.get(mode)
        self._refresh_derived(fmt, mode == L.SIM_TF32)
        if mode == L.SIM_TF32:
            return self._keys_tf32, None
        return self._shadow16[fmt] if fmt is not None else (None, None)

    # reference attribute names (views of the rows in use)
    @property
    def resource_keys(self) -> Tensor: return self._keys[:self._n]
    @property
    def resource_values(self) -> Tensor: return self._values[:self._n]
    @property
    def resource_labels(self) -> Tensor: return self._labels[:self._n]
    @property
    def resource_positions(self) -> Tensor:
        return None if self._positions is None else self._positions[:self._n]
    @property
    def key_inv_norm(self) -> Tensor:
        self._refresh_derived()
        return self._inv_norm[:self._n]

    def class_ids(self) -> Tensor:
        """int64 [N]: argmax over the label columns of every library row -- what the few-shot fusion looks up per
        retrieved row (``torch.argmax(rag_labels, dim=-1)``, RAGraph_node_fewshot/RAGraph.py:54); computed once per
        library state instead of once per retrieved copy."""
        if self._class_ids is None or self._class_rows != self._n:
            self._class_ids = torch.argmax(self.resource_labels, dim=-1).contiguous()
            self._class_rows = self._n
        return self._class_ids

    def __len__(self) -> int:
        return self._n

    # ---- persistence (the reference rebuilds its library from the dataset at every construction,
    #      ToyGraphBase.py:35-38 / RAGraph.py:27-28, and never saves it) ------------------------------------------
    _FILES = (("keys", "_keys"), ("values", "_values"), ("labels", "_labels"), ("positions", "_positions"))

    def save(self, path: str, chunk_rows: int = 1 << 20) -> None:
        """Write the rows in use as raw little-endian arrays (<path>/keys.bin, values.bin, labels.bin[, positions.bin])
        plus meta.json.  Streams device -> pinned host -> file in chunks, so a 100 M-row library needs no host copy of
        itself.  Derived arrays (inverse norms, bf16 shadow) are not stored: they are one kernel pass on load."""
        import json
        import os
        os.makedirs(path, exist_ok=True)
        meta = {"format": "ragraph_b200.toygraphbase/1", "rows": self._n, "emb_size": self.emb_size,
                "num_class": self.num_class, "variant": self.variant, "label_dtype": str(self._label_dtype).split(".")[-1],
                "retrieve_num": self.retrieve_num, "noise_retrieve_num": self.noise_retrieve_num,
                "structure_weight": self.structure_weight, "semantic_weight": self.semantic_weight,
                "noise_std": self.noise_std, "toy_graph_hop": self.toy_graph_hop, "num_anchors": self.num_anchors,
                "arrays": {}}
        for name, attr in self._FILES:
            t = getattr(self, attr)
            if t is None:
                continue
            meta["arrays"][name] = {"dtype": str(t.dtype).split(".")[-1], "cols": int(t.shape[1])}
            with open(os.path.join(path, name + ".bin"), "wb") as f:
                for a in range(0, self._n, chunk_rows):
                    b = min(self._n, a + chunk_rows)
                    f.write(t[a:b].contiguous().cpu().numpy().tobytes())
        with open(os.path.join(path, "meta.json"), "w") as f:
            json.dump(meta, f, indent=1)

    @classmethod
    def load(cls, path: str, device=None, rows: Optional[Tuple[int, int]] = None, pretrain_model=None,
             mode: Optional[int] = None, chunk_rows: int = 1 << 20) -> "ToyGraphBase":
        """Rebuild a store from ``save()`` output.  ``rows=(lo, hi)`` loads only that row range -- one shard of a
        key-row sharded library (``shard_bounds``); ``shard_lo`` is set so returned indices stay global."""
        import json
        import os
        import numpy as np
        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") != "ragraph_b200.toygraphbase/1":
            raise RuntimeError(f"{path}: not a ragraph_b200 library (format {meta.get('format')!r})")
        n_all = int(meta["rows"])
        lo, hi = (0, n_all) if rows is None else (int(rows[0]), int(rows[1]))
        if not (0 <= lo <= hi <= n_all):
            raise RuntimeError(f"rows {rows} outside the stored library of {n_all} rows")
        st = cls(pretrain_model, int(meta["num_class"]), int(meta["emb_size"]), int(meta["toy_graph_hop"]) + 1, device=device,
                 variant=meta["variant"], capacity=max(hi - lo, 1), mode=mode, label_dtype=getattr(torch, meta["label_dtype"]))
        for key in ("retrieve_num", "noise_retrieve_num", "structure_weight", "semantic_weight", "noise_std", "num_anchors"):
            setattr(st, key, meta[key])
        for name, attr in cls._FILES:
            info = meta["arrays"].get(name)
            dst = getattr(st, attr)
            if info is None or dst is None:
                continue
            dt = getattr(torch, info["dtype"])
            cols = int(info["cols"])
            if dst.dtype != dt or dst.shape[1] != cols:
                raise RuntimeError(f"{path}/{name}.bin: stored {info} does not match the store layout")
            src = np.memmap(os.path.join(path, name + ".bin"), mode="r", dtype=np.dtype(info["dtype"]), shape=(n_all, cols))
            for a in range(lo, hi, chunk_rows):
                b = min(hi, a + chunk_rows)
                dst[a - lo:b - lo].copy_(torch.from_numpy(np.ascontiguousarray(src[a:b])), non_blocking=False)
        st._n = hi - lo
        st.shard_lo = lo
        return st

    def show(self):
        print('resource_keys', self.resource_keys.shape)
        print('resource_values', self.resource_values.shape)
        print('resource_labels', self.resource_labels.shape)
        print("label count distribution", torch.sum(self.resource_labels, dim=0))

    # ---- retrieval ------------------------------------------------------------------------
    # Measured crossover (B200, profiles/r1_midsize_ab.jsonl, r2_small_shapes.jsonl): the tensor-core filter + fp32 refine
    # beats the fp32 CUDA-core kernel from a few thousand keys once there is a query tile's worth of rows (cfg1: 2 708 x
    # 10.8 k x 256: 0.19 ms vs 1.07 ms); below that the single-launch small-problem kernel serves retrieve() directly.
    TC_MIN_KEYS, TC_MIN_QUERIES = 4096, 16

    def _pick_mode(self, Q: int, k: Optional[int] = None) -> int:
        if self.mode is not None:
            return self.mode
        # tensor-core filter (fp16 operands) + fp32 refine: exact-match like the fp32 kernel
        k = self.retrieve_num if k is None else k
        big = self._n >= self.TC_MIN_KEYS and Q >= self.TC_MIN_QUERIES
        ok = big and L.load().rag_sim_mode_supported(L.SIM_F16_REFINE, self.emb_size, k)
        return L.SIM_F16_REFINE if ok else L.SIM_FP32

    def topk(self, search_keys: Tensor, k: int, search_positions: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """(scores[Q,k], indices[Q,k] int64): torch.topk(cosine(search_keys, resource_keys), k) fused."""
        if self.variant == "node_fewshot" and self.structure_weight != 0.0:
            if search_positions is None:
                raise RuntimeError("node_fewshot retrieval needs search_positions (position-aware codes); retrieve() "
                                   "derives them from search_adj like the reference")
            return ops.direct(ops.cosine2_topk)(search_positions, self.resource_positions, self.structure_weight,
                                                search_keys, self.resource_keys, self.semantic_weight, k)
        if k > L.RAG_MAX_K:
            return self._topk_large(search_keys, k)
        return self.topk_local(search_keys, k, 0)

    # ---- filter-format policy of the automatic mode --------------------------------------------------------------
    # Every tensor-core configuration returns the same exact answer; they differ in speed per data distribution (B200,
    # 100 M x 128, profiles/r2_ab_fmt_kp_100m.jsonl): on spread-out (Gaussian) keys a bf16 filter is ~5 % faster than fp16
    # (narrower multipliers, higher clock under the 1 kW cap: 72.3 vs 76.3 ms), on a clustered library bf16 cannot certify a
    # single row (5.7 s per batch through the fp32 kernel) while fp16 settles it on the tensor cores (126 ms; 94 ms with
    # 32-entry candidate lists).  A store therefore STARTS safe (fp16, 16-entry lists) and adapts from the counters each call
    # leaves behind (read back asynchronously, never a host sync): many rows in the second pass -> wide lists; two calls in
    # a row whose certificates would also hold under a bf16-sized error bound -> bf16; a bf16 call with second-pass rows ->
    # back to fp16 for good.  Only large scans adapt (the format is noise below ~4 G scores per call).
    ADAPT_MIN_SCORES = 1 << 32
    WIDE_LISTS_MIN_TILE_FRACTION = 0.3        # switch to 32-entry lists when the second pass touches more query tiles than this

    def _policy_reset(self) -> None:
        self._pol = {"fmt": L.FMT_F16, "wide": False, "locked": False, "calm": 0, "gen": 0}
        self._pol_pending = []

    def _policy_poll(self) -> None:
        while self._pol_pending and self._pol_pending[0][0].query():
            _, host, Q, gen, wide_ok = self._pol_pending.pop(0)
            self._pol_free.append(host)
            p = self._pol
            if gen != p["gen"]:
                continue                                    # measured under a configuration that is already history
            n2, _, loose = (int(x) for x in host.tolist())
            if p["fmt"] == L.FMT_BF16:
                if n2 > 0.02 * Q:
                    p.update(fmt=L.FMT_F16, locked=True, calm=0, gen=p["gen"] + 1)
            elif not p["wide"] and wide_ok and -(-n2 // 256) > self.WIDE_LISTS_MIN_TILE_FRACTION * -(-Q // 256):
                # The second pass rescans the shard for the uncertified rows, packed into 256-row query tiles: its cost is the
                # FRACTION OF QUERY TILES it touches, against ~20-28 % for 32-entry lists on every tile.  Measured
                # (profiles/r2_ab_fmt_kp_*.jsonl): 100 M keys, 2 400 of 4 096 rows uncertified (10 of 16 tiles) -> wide lists
                # win 93.9 vs 125.6 ms; 12.5 M-key shards, 697 rows (3 of 16 tiles) -> 16-entry lists win 12.8 vs 14.5 ms.
                p.update(wide=True, calm=0, gen=p["gen"] + 1)
            elif n2 == 0 and loose == 0 and not p["locked"] and not p["wide"]:
                p["calm"] += 1
                if p["calm"] >= 2:
                    p.update(fmt=L.FMT_BF16, calm=0, gen=p["gen"] + 1)
            else:
                p["calm"] = 0

    def auto_state(self) -> dict:
        """what the automatic mode currently runs (reporting)"""
        p = self._pol
        return {"filter": "bf16" if p["fmt"] == L.FMT_BF16 else "fp16", "wide_lists": bool(p["wide"])}

    def topk_local(self, search_keys: Tensor, k: int, idx_offset: int = 0) -> Tuple[Tensor, Tensor]:
        """fused similarity + top-k over the rows of THIS store; returned indices = idx_offset + local row"""
        mode = self._pick_mode(search_keys.shape[0], k)
        Q = search_keys.shape[0]
        flags = 0
        adapt = self.mode is None and mode == L.SIM_F16_REFINE and Q * self._n >= self.ADAPT_MIN_SCORES
        if adapt and torch.cuda.is_current_stream_capturing():
            adapt = False                       # (graphs.GraphedForward: the policy's events / pinned copies stay out of a capture)
        if adapt:
            self._policy_poll()
            if self._pol["fmt"] == L.FMT_BF16:
                mode = L.SIM_BF16_REFINE
            if self._pol["wide"]:
                flags = L.SIM_WIDE_LISTS
        shadow, err = self._shadow(mode)
        if not (adapt or self.collect_stats):
            return ops.direct(ops.cosine_topk)(search_keys, self.resource_keys, k, self._inv_norm[:self._n], shadow, mode,
                                               flags, idx_offset, err)
        s, i, stats = ops.cosine_topk_with_stats(search_keys, self.resource_keys, k, self._inv_norm[:self._n], shadow, mode,
                                                 flags, idx_offset, err)
        if self.collect_stats:
            self.last_stats = stats
        if adapt and len(self._pol_pending) < 4:
            host = self._pol_free.pop() if self._pol_free else torch.empty(3, dtype=torch.int32).pin_memory()
            host.copy_(stats[:3], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pol_pending.append((ev, host, Q, self._pol["gen"], self.emb_size <= 128 and k <= 10))
        return s, i

    def _topk_large(self, search_keys: Tensor, k: int, budget_bytes: int = 1 << 30) -> Tuple[Tensor, Tensor]:
        """k > RAG_MAX_K (the edge variant's vanilla configs ask for retrieve_num = 100000, i.e. most of the library,
        RAGraph_edge/modules/RAGraph.py:57,73): outside the fused kernels.  The scores of a query chunk are materialised
        by the CUDA similarity kernel ([chunk, N] fp32 bounded by ``budget_bytes``) and selected with torch.topk -- a
        library call on a path that degenerates to "average almost everything"; k is clamped to the library size like
        a user of the reference would have to."""
        k = min(k, self._n)
        Q = search_keys.shape[0]
        chunk = max(1, min(Q, budget_bytes // max(4 * self._n, 1)))
        scores = torch.empty((Q, k), dtype=torch.float32, device=search_keys.device)
        idx = torch.empty((Q, k), dtype=torch.int64, device=search_keys.device)
        for a in range(0, Q, chunk):
            b = min(Q, a + chunk)
            s = ops.cosine_similarity(search_keys[a:b], self.resource_keys)
            scores[a:b], idx[a:b] = torch.topk(s, k, dim=1, largest=True, sorted=True)
        return scores, idx

    def query_positions(self, search_adj, search_positions: Optional[Tensor] = None) -> Optional[Tensor]:
        """Structure codes of the query graph for the two-metric score (RAGraph_node_fewshot/.../ToyGraphBase.py:49):
        the caller's ``search_positions`` if given, else derived from the dense ``search_adj`` like the reference does
        (PositionAwareEncoder, random anchors from the CPU generator); None for the single-metric variants."""
        if search_positions is not None or self.variant != "node_fewshot" or self.structure_weight == 0.0:
            return search_positions
        from ..csr import CSRGraph, as_dense
        if isinstance(search_adj, CSRGraph) or (isinstance(search_adj, Tensor) and search_adj.layout != torch.strided):
            search_adj = as_dense(search_adj)
        if not isinstance(search_adj, Tensor):
            return None
        from .PositionAwareEncoder import PositionAwareEncoder
        return PositionAwareEncoder.encode_position_aware_code(search_adj, self.num_anchors, self.dis_q)

    def _small_retrieve(self, search_keys: Tensor, k: int, gather: bool = True):
        """The reference's real shapes (one pooled graph query against a few hundred rows, RAGraph_graph/.../ToyGraphBase.py:
        56-87) are launch-latency problems: normalise + scores + top-k + both gathers run as ONE kernel when the shape fits
        (Q <= 64, N <= 65 536, k <= 16) and no similarity mode was forced.  Returns (idx, values[idx], labels[idx]) or None."""
        if self.mode is not None or self.collect_stats or (self.variant == "node_fewshot" and self.structure_weight != 0.0):
            return None
        n, Q = self._n, search_keys.shape[0]
        key = (Q, n, k)
        need = self._small_plan.get(key)
        if need is None:
            ok = ops.retrieve_small_supported(Q, n, self.emb_size, k) and search_keys.dtype == torch.float32
            need = int(L.load().rag_retrieve_small_workspace(Q, n, self.emb_size, k)) if ok else 0
            if len(self._small_plan) > 256:
                self._small_plan.clear()
            self._small_plan[key] = need
        if need == 0 or not search_keys.is_cuda:
            return None
        if self._small_ws is None or self._small_ws.numel() < need:
            self._small_ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
        _, idx, emb, lab = ops.retrieve_small(search_keys.contiguous(), self._keys[:n], k, self._values[:n] if gather else None,
                                              self._labels[:n] if gather else None, self._small_ws)
        return idx, emb, lab

    def retrieve(self, search_keys: Tensor, search_adj, add_noise: bool, search_positions: Optional[Tensor] = None):
        """Same contract as the reference: returns (rag_embeddings[Q,k',d], rag_labels[Q,k',C]).
        node (:47-81): add_noise doubles k and appends noise_retrieve_num random rows (CPU torch.randint,
        same RNG call as the reference); graph (:56-87): a 1-D query gives Q=1, add_noise adds N(0, noise_std)
        to the gathered values (:84-85)."""
        if search_keys.dim() == 1:
            search_keys = search_keys.unsqueeze(0)
        search_positions = self.query_positions(search_adj, search_positions)
        retrieve_num = 2 * self.retrieve_num if add_noise else self.retrieve_num
        gather_rows = ops.direct(ops.gather_rows)
        small = self._small_retrieve(search_keys, retrieve_num) if search_positions is None else None
        if small is not None:
            _, rag_embeddings, rag_labels = small
        else:
            _, topk_indices = self.topk(search_keys, retrieve_num, search_positions)
            rag_embeddings = gather_rows(self.resource_values, topk_indices)
            rag_labels = gather_rows(self.resource_labels, topk_indices)
        if add_noise:
            if self.variant == "graph":
                noise = torch.normal(mean=0, std=self.noise_std, size=rag_embeddings.shape).to(rag_embeddings.device)
                rag_embeddings = rag_embeddings + noise
            else:
                noise_indices = torch.randint(0, self._n, (search_keys.shape[0], self.noise_retrieve_num))
                noise_indices = noise_indices.to(self.device)
                rag_embeddings = torch.cat([rag_embeddings, gather_rows(self.resource_values, noise_indices)], dim=1)
                rag_labels = torch.cat([rag_labels, gather_rows(self.resource_labels, noise_indices)], dim=1)
        return rag_embeddings, rag_labels

    def gather_reduce_blend(self, idx: Tensor, reduce: int = L.REDUCE_SUM, blend_in: Optional[Tensor] = None,
                            blend_w: float = 0.0) -> Tensor:
        """(1 - w) * blend_in + w * reduce_k values[idx] -- one fused launch when nothing needs a gradient.  The library rows
        carry no grad (the reference detaches them, preprompt.py:62), but ``blend_in`` does when the backbone is fine-tuned
        (few-shot: encode() is trainable, RAGraph_node_fewshot/RAGraph.py:47-69): then only the gather + reduce runs in the
        kernel and the convex blend is ordinary differentiable torch, so d loss / d blend_in = (1 - w) * dY flows."""
        gather_reduce = ops.direct(ops.gather_reduce)
        if blend_in is not None and blend_in.requires_grad and torch.is_grad_enabled():
            rag = gather_reduce(self.resource_values, idx, reduce)
            return blend_in * (1.0 - blend_w) + rag * blend_w
        return gather_reduce(self.resource_values, idx, reduce, blend_in, blend_w)

    def retrieve_fused(self, search_keys: Tensor, k: Optional[int] = None, reduce: int = L.REDUCE_SUM,
                       blend_in: Optional[Tensor] = None, blend_w: float = 0.0):
        """The reduction every caller applies next, without the [Q,k,d] round trip: returns
        (sum|mean over k of values[idx] (optionally blended with blend_in), mean over k of labels[idx], idx)."""
        if search_keys.dim() == 1:
            search_keys = search_keys.unsqueeze(0)
        k = self.retrieve_num if k is None else k
        # (the graph variant's single pooled query: one launch instead of a 142 us pass of the tiled fp32 kernel over ONE row)
        small = self._small_retrieve(search_keys, k, gather=False)
        idx = small[0] if small is not None else self.topk(search_keys, k)[1]
        gather_reduce = ops.direct(ops.gather_reduce)
        emb = self.gather_reduce_blend(idx, reduce, blend_in, blend_w)
        labels = self.resource_labels if self.resource_labels.dtype == torch.float32 else self.resource_labels.float()
        lab = gather_reduce(labels, idx, L.REDUCE_MEAN)
        return emb, lab, idx
