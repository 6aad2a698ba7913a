"""Synthetic interaction graphs in the reference's on-disk format.

The reference trains on ``train.parquet`` / ``test.parquet`` / ``item_brand.parquet`` /
``stats.json`` (+ optional ``item_embeddings.npy``) read by
``load_preprocessed_data`` (reference ``main.py:181-210``) and written by its offline
scripts (reference ``dataset/amazon_books_emb/prepare_data.py:126-158``).  There is no
network here, so every workload is generated: user degrees >= ``min_degree`` and item
popularity proportional to ``rank**-alpha``.

Shapes (BASELINE.json ``configs``):

* ``gowalla``  U=29 858, I=40 981, B=1, 1 027 370 unique (user, item) pairs, d=64,  K=3
* ``amazon``   U=10.3 M, I=4.4 M,  B=1, 29.5 M train edges + 1 val + 1 test per user,
  d=128, K=4
* ``tiny``     U=120, I=200 (golden fixtures / smoke)

Only numpy is used so that the same generator runs in the CPU tests, the golden-vector
script and ``bench.py``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

SHAPES = {
    # name: (num_users, num_items, num_brands, total unique interactions, dim, layers)
    "tiny": (120, 200, 1, 1_500, 64, 3),
    "small": (3_000, 4_000, 1, 60_000, 64, 3),
    "gowalla": (29_858, 40_981, 1, 1_027_370, 64, 3),
    "amazon_16th": (643_750, 275_000, 1, 3_131_250, 128, 4),
    "amazon": (10_300_000, 4_400_000, 1, 50_100_000, 128, 4),
}


@dataclass
class Interactions:
    """All unique (user, item) pairs of a synthetic dataset, already split.

    ``train_user/train_item`` is what the reference calls ``all_train_df`` (it still
    holds the validation row of every user: reference ``main.py:201-203`` peels the
    first row per user off as validation).  ``test_*`` is one row per user.
    """

    num_users: int
    num_items: int
    num_brands: int
    train_user: np.ndarray  # int64, file order
    train_item: np.ndarray
    test_user: np.ndarray
    test_item: np.ndarray

    def split_validation(self):
        """Mirror of reference ``main.py:201-203``: the FIRST row of each user in file
        order becomes the validation row, the rest is the training graph.

        Returns (train_user, train_item, val_user, val_item)."""
        u = self.train_user
        order = np.argsort(u, kind="stable")
        su = u[order]
        first = np.ones(len(su), dtype=bool)
        first[1:] = su[1:] != su[:-1]
        is_val = np.zeros(len(u), dtype=bool)
        is_val[order[first]] = True
        return (u[~is_val], self.train_item[~is_val], u[is_val], self.train_item[is_val])


def _user_degrees(rng, num_users, total, min_degree):
    """Degrees >= min_degree summing exactly to ``total`` with a heavy right tail."""
    extra = total - num_users * min_degree
    if extra < 0:
        raise ValueError("total interactions < num_users * min_degree")
    w = rng.lognormal(mean=0.0, sigma=1.0, size=num_users)
    w /= w.sum()
    add = np.floor(w * extra).astype(np.int64)
    short = extra - int(add.sum())
    if short > 0:
        add[rng.choice(num_users, size=short, replace=False if short <= num_users else True)] += 1
    return add + min_degree


def generate(shape="tiny", seed=0, alpha=0.8, min_degree=3) -> Interactions:
    """Generate unique (user, item) pairs: user degree >= ``min_degree``; item drawn with
    probability proportional to ``(rank+1)**-alpha`` where ranks are a seeded permutation of
    the item ids; rows shuffled; one row per user held out as test."""
    if isinstance(shape, str):
        U, I, B, total, _, _ = SHAPES[shape]
    else:
        U, I, B, total = shape
    rng = np.random.default_rng(seed)
    deg = _user_degrees(rng, U, total, min_degree)
    if int(deg.max()) > I:
        raise ValueError("a user degree exceeds the number of items")
    cdf = np.cumsum((np.arange(1, I + 1, dtype=np.float64)) ** (-alpha))
    cdf /= cdf[-1]
    rank_to_item = rng.permutation(I).astype(np.int64)

    users = np.repeat(np.arange(U, dtype=np.int64), deg)
    items = rank_to_item[np.searchsorted(cdf, rng.random(len(users)), side="right").clip(0, I - 1)]
    # Re-draw duplicated (user, item) pairs until every pair is unique.
    for _ in range(200):
        key = users * I + items
        order = np.argsort(key, kind="stable")
        sk = key[order]
        dup_sorted = np.zeros(len(sk), dtype=bool)
        dup_sorted[1:] = sk[1:] == sk[:-1]
        ndup = int(dup_sorted.sum())
        if ndup == 0:
            break
        dup = order[dup_sorted]
        # heavy users exhaust the popular head quickly: mix in uniform draws
        r = rng.random(ndup)
        pick = np.searchsorted(cdf, r, side="right").clip(0, I - 1)
        uni = rng.integers(0, I, size=ndup)
        items[dup] = np.where(rng.random(ndup) < 0.5, rank_to_item[pick], uni)
    else:  # pragma: no cover
        raise RuntimeError("could not make interactions unique")

    perm = rng.permutation(len(users))
    users, items = users[perm], items[perm]
    # test = one row per user (the last one in shuffled order)
    order = np.argsort(users, kind="stable")
    su = users[order]
    last = np.ones(len(su), dtype=bool)
    last[:-1] = su[1:] != su[:-1]
    is_test = np.zeros(len(users), dtype=bool)
    is_test[order[last]] = True
    return Interactions(U, I, B, users[~is_test], items[~is_test], users[is_test], items[is_test])


def side_embeddings(num_items, dim=768, seed=1) -> np.ndarray:
    """``item_embeddings.npy`` stand-in: (num_items, dim) fp32 standard normal
    (format: reference ``dataset/amazon_books_emb/prepare_data.py:141-150``)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((num_items, dim), dtype=np.float32)


def write_reference_format(inter: Interactions, out_dir, content: np.ndarray | None = None):
    """Write the files ``load_preprocessed_data`` (reference ``main.py:181-210``) reads."""
    import pandas as pd

    os.makedirs(out_dir, exist_ok=True)
    pd.DataFrame({"user_idx": inter.train_user, "item_idx": inter.train_item}).to_parquet(
        os.path.join(out_dir, "train.parquet"), index=False)
    pd.DataFrame({"user_idx": inter.test_user, "item_idx": inter.test_item}).to_parquet(
        os.path.join(out_dir, "test.parquet"), index=False)
    pd.DataFrame({"item_idx": np.arange(inter.num_items, dtype=np.int64),
                  "brand_idx": np.zeros(inter.num_items, dtype=np.int64) % max(inter.num_brands, 1)}
                 ).to_parquet(os.path.join(out_dir, "item_brand.parquet"), index=False)
    with open(os.path.join(out_dir, "stats.json"), "w") as f:
        json.dump({"num_users": inter.num_users, "num_items": inter.num_items,
                   "num_brands": inter.num_brands}, f)
    if content is not None:
        np.save(os.path.join(out_dir, "item_embeddings.npy"), content.astype(np.float32))


SHAPES["amazon_64th"] = (160_937, 68_750, 1, 782_812, 128, 4)


def generate_device(shape, device, seed=0, alpha=0.8, min_degree=3):
    """Same generator as :func:`generate`, evaluated with torch on ``device`` (the 50 M-pair
    Amazon shape takes ~2 minutes in numpy on 8 cores; on the GPU it is a second).  The random
    streams differ from numpy's, the distributional shape is identical.  Returns an
    :class:`Interactions` whose arrays are int64 torch tensors on ``device``."""
    import torch

    if isinstance(shape, str):
        U, I, B, total, _, _ = SHAPES[shape]
    else:
        U, I, B, total = shape
    gen = torch.Generator(device=device).manual_seed(seed)
    extra = total - U * min_degree
    if extra < 0:
        raise ValueError("total interactions < num_users * min_degree")
    w = torch.exp(torch.randn(U, device=device, generator=gen, dtype=torch.float64))
    add = torch.floor(w / w.sum() * extra).to(torch.int64)
    short = extra - int(add.sum().item())
    if short > 0:
        add[torch.randperm(U, device=device, generator=gen)[:short]] += 1
    deg = add + min_degree
    cdf = torch.cumsum(torch.arange(1, I + 1, device=device, dtype=torch.float64) ** (-alpha), 0)
    cdf = cdf / cdf[-1]
    rank_to_item = torch.randperm(I, device=device, generator=gen)

    def draw(n):
        r = torch.rand(n, device=device, generator=gen, dtype=torch.float64)
        return rank_to_item[torch.searchsorted(cdf, r, right=True).clamp_(0, I - 1)]

    users = torch.repeat_interleave(torch.arange(U, device=device), deg)
    items = draw(users.numel())
    for _ in range(200):
        key = users * I + items
        sk, order = torch.sort(key, stable=True)
        dup_sorted = torch.zeros_like(sk, dtype=torch.bool)
        dup_sorted[1:] = sk[1:] == sk[:-1]
        dup = order[dup_sorted]
        nd = dup.numel()
        if nd == 0:
            break
        uni = torch.randint(0, I, (nd,), device=device, generator=gen)
        coin = torch.rand(nd, device=device, generator=gen) < 0.5
        items[dup] = torch.where(coin, draw(nd), uni)
    else:  # pragma: no cover
        raise RuntimeError("could not make interactions unique")
    perm = torch.randperm(users.numel(), device=device, generator=gen)
    users, items = users[perm], items[perm]
    su, order = torch.sort(users, stable=True)
    last = torch.ones_like(su, dtype=torch.bool)
    last[:-1] = su[1:] != su[:-1]
    is_test = torch.zeros_like(last)
    is_test[order[last]] = True
    return Interactions(U, I, B, users[~is_test], items[~is_test], users[is_test], items[is_test])


def split_validation_device(inter):
    """torch/device version of :meth:`Interactions.split_validation`."""
    import torch

    u = inter.train_user
    su, order = torch.sort(u, stable=True)
    first = torch.ones_like(su, dtype=torch.bool)
    first[1:] = su[1:] != su[:-1]
    is_val = torch.zeros_like(first)
    is_val[order[first]] = True
    return u[~is_val], inter.train_item[~is_val], u[is_val], inter.train_item[is_val]
