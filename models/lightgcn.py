"""B200-native ``LightGCN`` behind the reference's model interface.

Same constructor, ``forward(adj_mat, use_brand=True)`` 5-tuple and ``state_dict`` keys as
reference ``models/lightgcn.py:4-81``, so ``main.py`` loads it unchanged through
``get_model`` (reference ``main.py:42-50``).  The propagation is one hand-written sm_100a CSR
SpMM per layer with the layer mean fused into the last launch; the backward pass reuses the
same kernel (the normalised adjacency is symmetric).  No torch.sparse, no CPU fallback.

Besides the drop-in ``forward`` the module exposes the native fast path
(``engine()`` -> fused BPR step + Adam + full-rank top-k, see
``gcn_recommendation_b200/engine.py``).
"""
import torch
import torch.nn as nn

from gcn_recommendation_b200 import graph, ops
from gcn_recommendation_b200.engine import LightGCNEngine

from ._packing import pack_


class LightGCN(nn.Module):
    def __init__(self, num_users, num_items, num_brands, config, pretrained_item_emb=None):
        super().__init__()
        self.num_users, self.num_items, self.num_brands = num_users, num_items, num_brands
        self.embedding_dim = config.embedding_dim
        self.n_layers = config.n_layers
        self.debug = getattr(config, "debug", False)
        # RNG consumption order of reference models/lightgcn.py:15-31 (user, brand, item)
        self.user_embedding = nn.Embedding(num_users, self.embedding_dim)
        self.brand_embedding = nn.Embedding(num_brands, self.embedding_dim)
        if pretrained_item_emb is not None:
            if pretrained_item_emb.shape[1] != self.embedding_dim:
                raise ValueError(
                    f"Pretrained embedding dim ({pretrained_item_emb.shape[1]}) does not match "
                    f"model embedding dim ({self.embedding_dim}).")
            self.item_embedding = nn.Embedding.from_pretrained(
                torch.as_tensor(pretrained_item_emb, dtype=torch.float32), freeze=False)
        else:
            self.item_embedding = nn.Embedding(num_items, self.embedding_dim)
            nn.init.xavier_uniform_(self.item_embedding.weight)
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.brand_embedding.weight)
        self.final_brand_emb = None

    # ---- drop-in path -------------------------------------------------------------------
    def _tables(self):
        return [self.user_embedding.weight, self.item_embedding.weight, self.brand_embedding.weight]

    def table_block(self):
        """[N,d] view over users | items | brands (packs on first use)."""
        return pack_(self._tables())

    def forward(self, adj_mat, use_brand=True):
        """``adj_mat``: the fp32 ``torch.sparse_coo`` tensor of reference ``main.py:334-336``
        (converted to CSR once and cached) or a :class:`graph.NormAdjCSR`."""
        g = adj_mat if isinstance(adj_mat, graph.NormAdjCSR) else graph.csr_for(adj_mat)
        self.table_block()
        u0, i0, b0 = self._tables()
        final = ops.PropagateFunction.apply(g, self.n_layers, u0, i0, b0)
        fu, fi, fb = torch.split(final, [self.num_users, self.num_items, self.num_brands])
        return fu, fi, fb, u0, i0

    # ---- native fast path ---------------------------------------------------------------
    def engine(self, adj, lr=1e-3, weight_decay=1e-4, **kw):
        """Fused trainer/evaluator over this module's parameters (shares their storage)."""
        g = adj if isinstance(adj, graph.NormAdjCSR) else graph.csr_for(adj)
        return LightGCNEngine(g, self.num_users, self.num_items, self.num_brands, self.n_layers,
                              table=self.table_block(), lr=lr, weight_decay=weight_decay, **kw)
