"""CPU port of the reference's training / evaluation step for TIMING -- TEST INFRASTRUCTURE ONLY.

The reference cannot travel to the GPU box (``/root/reference`` does not exist there), so the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` time this restatement instead.  It
issues the SAME library calls in the same composition as the reference, so its cost profile is
the reference's:

* forward: ``torch.cat`` -> K x ``torch.sparse.mm`` on the uncoalesced COO tensor ->
  ``torch.mean(torch.stack())`` -> ``torch.split``          (reference models/lightgcn.py:37-59)
* loss: two row-dot sums, ``-mean(log(sigmoid+1e-8))``, three squared norms
                                                              (reference main.py:377-398)
* ``loss.backward()``, ``torch.optim.Adam.step()``, ``loss.item()``   (reference main.py:525-528)
* evaluate: ``matmul`` + per-user ``index_put`` mask loop + ``torch.topk`` (main.py:420-426)

Pinned against the golden vectors by tests/test_oracle_golden.py::test_torch_port_*.
Never imported by the product path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lgcn_oracle as orc


class TorchPort:
    def __init__(self, train_user, train_item, num_users, num_items, num_brands, d, n_layers,
                 lr=1e-3, lam=1e-4, seed=42, init=None):
        self.U, self.I, self.B, self.d, self.K = num_users, num_items, num_brands, d, n_layers
        self.lam = lam
        a = orc.build_norm_adj(train_user, train_item, num_users, num_items, num_brands)
        N = num_users + num_items + num_brands
        rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(a["rowptr"]))
        idx = torch.from_numpy(np.vstack([rows, a["col"].astype(np.int64)]))
        # uncoalesced COO, exactly what main.py:334-336 hands to the model
        self.adj = torch.sparse_coo_tensor(idx, torch.from_numpy(a["val"]), (N, N))
        torch.manual_seed(seed)
        emb = torch.nn.Embedding
        self.user, self.brand, self.item = emb(num_users, d), emb(num_brands, d), emb(num_items, d)
        for t in (self.item, self.user, self.brand):
            torch.nn.init.xavier_uniform_(t.weight)
        if init is not None:
            with torch.no_grad():
                for t, w in zip((self.user, self.item, self.brand), init):
                    t.weight.copy_(torch.as_tensor(w))
        self.params = [self.user.weight, self.brand.weight, self.item.weight]
        self.opt = torch.optim.Adam(self.params, lr=lr)

    def forward(self):
        ego = torch.cat([self.user.weight, self.item.weight, self.brand.weight], dim=0)
        layers = [ego]
        for _ in range(self.K):
            ego = torch.sparse.mm(self.adj, ego)
            layers.append(ego)
        final = torch.mean(torch.stack(layers, dim=0), dim=0)
        return torch.split(final, [self.U, self.I, self.B])

    def step(self, users, pos, neg):
        users, pos, neg = (torch.as_tensor(x, dtype=torch.int64) for x in (users, pos, neg))
        self.opt.zero_grad()
        fu, fi, _ = self.forward()
        eu, ep, en = fu[users], fi[pos], fi[neg]
        diff = torch.sum(eu * ep, dim=1) - torch.sum(eu * en, dim=1)
        bpr = -torch.mean(torch.log(torch.sigmoid(diff) + 1e-8))
        u0, p0, n0 = self.user.weight[users], self.item.weight[pos], self.item.weight[neg]
        reg = self.lam * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / float(len(users))
        loss = bpr + reg
        loss.backward()
        self.opt.step()
        return loss.item()

    @torch.no_grad()
    def evaluate(self, eval_users, targets, train_lists, k=20, batch=1024):
        """train_lists: dict user -> list of train items (main.py:407)."""
        fu, fi, _ = self.forward()
        hits, dcg = 0, 0.0
        eval_users = list(map(int, eval_users))
        for s in range(0, len(eval_users), batch):
            bu = eval_users[s:s + batch]
            scores = torch.matmul(fu[torch.as_tensor(bu)], fi.T)
            for j, u in enumerate(bu):
                if u in train_lists:
                    scores[j, train_lists[u]] = -1e10
            top = torch.topk(scores, k=k)[1].numpy()
            for j in range(len(bu)):
                w = np.where(top[j] == targets[s + j])[0]
                if len(w):
                    hits += 1
                    dcg += 1.0 / np.log2(w[0] + 2)
        return hits / len(eval_users), dcg / len(eval_users)
