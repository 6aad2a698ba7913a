"""B200-native ``LightGCN_Fusion`` behind the reference's model interface.

Same constructor, ``forward`` 5-tuple and ``state_dict`` keys as reference
``models/lightgcn_fusion.py:5-65``: learnable user / item-id / brand tables, a fixed content
matrix kept as a buffer, and ``item_fusion_layer = Linear(d + c, d)`` whose leaky-relu output
replaces the item rows of layer 0.  The projection runs as a hand-written kernel that never
materialises ``cat([id_emb, content], 1)``; the propagation is the shared CSR SpMM.
"""
import torch
import torch.nn as nn

from gcn_recommendation_b200 import graph, ops
from gcn_recommendation_b200.engine import LightGCNEngine

from ._packing import pack_


class LightGCN_Fusion(nn.Module):
    def __init__(self, num_users, num_items, num_brands, config, pretrained_item_emb=None):
        super().__init__()
        self.num_users, self.num_items, self.num_brands = num_users, num_items, num_brands
        self.embedding_dim = config.embedding_dim
        self.n_layers = config.n_layers
        if pretrained_item_emb is None:
            raise ValueError("LightGCN_Fusion model requires pretrained item embeddings.")
        content_dim = pretrained_item_emb.shape[1]
        # RNG consumption order of reference models/lightgcn_fusion.py:20-35
        self.user_embedding = nn.Embedding(num_users, self.embedding_dim)
        self.item_id_embedding = nn.Embedding(num_items, self.embedding_dim)
        self.brand_embedding = nn.Embedding(num_brands, self.embedding_dim)
        self.register_buffer("item_content_embedding",
                             torch.as_tensor(pretrained_item_emb, dtype=torch.float32).clone())
        self.item_fusion_layer = nn.Linear(self.embedding_dim + content_dim, self.embedding_dim)
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_id_embedding.weight)
        nn.init.xavier_uniform_(self.brand_embedding.weight)
        nn.init.xavier_uniform_(self.item_fusion_layer.weight)

    def _tables(self):
        return [self.user_embedding.weight, self.item_id_embedding.weight,
                self.brand_embedding.weight]

    def table_block(self):
        return pack_(self._tables())

    def forward(self, adj_mat, use_brand=True):
        g = adj_mat if isinstance(adj_mat, graph.NormAdjCSR) else graph.csr_for(adj_mat)
        self.table_block()
        u0, id0, b0 = self._tables()
        fused = ops.FusionProjFunction.apply(id0, self.item_content_embedding,
                                             self.item_fusion_layer.weight,
                                             self.item_fusion_layer.bias)
        final = ops.PropagateFunction.apply(g, self.n_layers, u0, fused, b0)
        fu, fi, fb = torch.split(final, [self.num_users, self.num_items, self.num_brands])
        # the L2 term regularises the raw id table, not the fused one (reference :65)
        return fu, fi, fb, u0, id0

    def engine(self, adj, lr=1e-3, weight_decay=1e-4, **kw):
        g = adj if isinstance(adj, graph.NormAdjCSR) else graph.csr_for(adj)
        fusion = dict(content=self.item_content_embedding, weight=self.item_fusion_layer.weight.data,
                      bias=self.item_fusion_layer.bias.data)
        return LightGCNEngine(g, self.num_users, self.num_items, self.num_brands, self.n_layers,
                              table=self.table_block(), lr=lr, weight_decay=weight_decay,
                              fusion=fusion, **kw)
