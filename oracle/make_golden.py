"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python oracle/make_golden.py

The reference's own functions are executed -- ``main.load_preprocessed_data``,
``models.lightgcn.LightGCN``, ``models.lightgcn_fusion.LightGCN_Fusion``,
``main.bpr_loss_reg``, ``torch.optim.Adam`` as ``main.train`` builds it, ``main.evaluate`` --
on a tiny synthetic dataset written in the reference's on-disk format; the training loop
around them follows reference ``main.py:488-531`` with a recorded, deterministic batch
stream (the reference's DataLoader stream depends on its worker count).
matplotlib is not installed here, so it is mocked exactly as the survey did.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LGCN_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from gcn_recommendation_b200 import synth  # noqa: E402


def _import_reference():
    sys.modules.setdefault("matplotlib", mock.MagicMock())
    sys.modules.setdefault("matplotlib.pyplot", mock.MagicMock())
    # the reference's `models` package must win over this repo's drop-in `models`
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    import importlib
    main = importlib.import_module("main")
    lg = importlib.import_module("models.lightgcn")
    lf = importlib.import_module("models.lightgcn_fusion")
    assert os.path.realpath(main.__file__).startswith(os.path.realpath(REF))
    assert os.path.realpath(lg.__file__).startswith(os.path.realpath(REF))
    sys.path.remove(REF)
    return main, lg, lf


def _batches(rng, tu, ti, num_items, bs, n_steps):
    """Deterministic (user, pos, neg) batches: shuffled train rows, uniform negatives by
    rejection against the user's positives (semantics of reference main.py:357-363)."""
    pos_sets = {}
    for u, i in zip(tu.tolist(), ti.tolist()):
        pos_sets.setdefault(u, set()).add(i)
    out = []
    perm = rng.permutation(len(tu))
    for s in range(n_steps):
        idx = perm[(s * bs) % len(tu):][:bs]
        if len(idx) < bs:
            idx = np.concatenate([idx, perm[: bs - len(idx)]])
        u, p = tu[idx], ti[idx]
        n = np.empty_like(p)
        for j, uu in enumerate(u.tolist()):
            while True:
                c = int(rng.integers(0, num_items))
                if c not in pos_sets[uu]:
                    break
            n[j] = c
        out.append((u.copy(), p.copy(), n))
    return out


def run_case(main, model_cls, name, shape, d, K, bs, n_steps, use_brand, fusion, seed,
             min_degree=3, n_dup=0, brand_loss_weight=None):
    """``min_degree=2`` leaves users without any training edge after the test / validation
    hold-outs (isolated nodes: degree 0 -> d = 0, reference main.py:329); ``n_dup`` appends
    repeated (user, item) rows to train.parquet (multiplicity > 1 in the adjacency, summed by the
    reference's COO -> CSR conversion, main.py:321-331)."""
    import torch

    torch.set_num_threads(1)
    inter = synth.generate(shape, seed=seed, min_degree=min_degree)
    if n_dup:
        rng_d = np.random.default_rng(seed + 99)
        pick = rng_d.choice(len(inter.train_user), n_dup, replace=False)
        pick = np.concatenate([pick, pick[: n_dup // 4]])          # some pairs three times
        inter.train_user = np.concatenate([inter.train_user, inter.train_user[pick]])
        inter.train_item = np.concatenate([inter.train_item, inter.train_item[pick]])
    content = synth.side_embeddings(inter.num_items, 768, seed + 1) if fusion else None
    golden = {}
    with tempfile.TemporaryDirectory() as tmp:
        ddir = os.path.join(tmp, "processed")
        synth.write_reference_format(inter, ddir, content)
        if use_brand:
            # a non-trivial item->brand map so that the tripartite graph has edges
            import pandas as pd
            nb = 7
            ib = pd.DataFrame({"item_idx": np.arange(inter.num_items, dtype=np.int64),
                               "brand_idx": (np.arange(inter.num_items, dtype=np.int64) * 3) % nb})
            ib.to_parquet(os.path.join(ddir, "item_brand.parquet"), index=False)
            import json
            with open(os.path.join(ddir, "stats.json"), "w") as f:
                json.dump({"num_users": inter.num_users, "num_items": inter.num_items,
                           "num_brands": nb}, f)
            golden["item_brand_item"] = ib["item_idx"].values
            golden["item_brand_brand"] = ib["brand_idx"].values
            i2b = np.zeros(inter.num_items, np.int64)
            i2b[ib["item_idx"].values] = ib["brand_idx"].values
            golden["item_to_brand"] = i2b
            if brand_loss_weight is not None:
                golden["brand_loss_weight"] = np.float64(brand_loss_weight)
        dev = torch.device("cpu")
        (train_df, val_df, test_df, U, I, B, adj, item_brand_df) = main.load_preprocessed_data(
            ddir, dev, use_brand=use_brand, debug=False)
    golden.update(dict(
        num_users=U, num_items=I, num_brands=B, d=d, K=K, bs=bs, lam=1e-4, lr=1e-3,
        all_train_user=inter.train_user, all_train_item=inter.train_item,
        test_user=inter.test_user, test_item=inter.test_item,
        train_user=train_df["user_idx"].values, train_item=train_df["item_idx"].values,
        val_user=val_df["user_idx"].values, val_item=val_df["item_idx"].values,
        adj_row=adj._indices()[0].numpy().astype(np.int32),
        adj_col=adj._indices()[1].numpy().astype(np.int32),
        adj_val=adj._values().numpy(),
    ))

    cfg = types.SimpleNamespace(embedding_dim=d, n_layers=K, debug=False, device=dev,
                                learning_rate=1e-3, weight_decay=1e-4, top_k=20, batch_size=bs)
    main.config = cfg
    torch.manual_seed(42)                                     # reference main.py:607
    model = model_cls(U, I, B, cfg, pretrained_item_emb=content).to(dev)
    for k, v in model.state_dict().items():
        golden["init/" + k] = v.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)   # main.py:469

    with torch.no_grad():
        fu, fi, fb, _, _ = model(adj, use_brand=use_brand)
        golden["fwd/user"] = fu.numpy().copy()
        golden["fwd/item"] = fi.numpy().copy()
        golden["fwd/brand"] = fb.numpy().copy()

    rng = np.random.default_rng(seed + 7)
    tu, ti = train_df["user_idx"].values, train_df["item_idx"].values
    batches = _batches(rng, tu, ti, I, bs, n_steps)
    losses = []
    model.train()
    for s, (u, p, n) in enumerate(batches):                   # main.py:488-531
        users, pos, neg = torch.from_numpy(u), torch.from_numpy(p), torch.from_numpy(n)
        opt.zero_grad()
        fu, fi, fb, u0, i0 = model(adj, use_brand=use_brand)
        if brand_loss_weight is None:
            loss = main.bpr_loss_reg(fu[users], fi[pos], fi[neg], u0[users], i0[pos], i0[neg],
                                     cfg.weight_decay)
        else:
            # the brand / author BPR term of reference main.py:382-391, fed the way main.py:499-522
            # intends (item -> brand lookup of the batch's positive and negative items); the
            # reference's train() never defines item_to_brand, so the FUNCTION is what is pinned
            i2b = torch.from_numpy(golden["item_to_brand"])
            loss = main.bpr_loss_reg(fu[users], fi[pos], fi[neg], u0[users], i0[pos], i0[neg],
                                     cfg.weight_decay, brand_loss=True, final_brand_emb=fb,
                                     pos_item_brand_idx=i2b[pos], neg_item_brand_idx=i2b[neg],
                                     brand_loss_weight=brand_loss_weight)
        loss.backward()
        if s == 0:
            for k, prm in model.named_parameters():
                golden["grad1/" + k] = prm.grad.detach().numpy().copy()
        opt.step()
        losses.append(loss.item())
        if s == 0:
            for k, v in model.state_dict().items():
                if k != "item_content_embedding":
                    golden["step1/" + k] = v.detach().numpy().copy()
    for k, v in model.state_dict().items():
        if k != "item_content_embedding":
            golden["final/" + k] = v.detach().numpy().copy()
    golden["losses"] = np.asarray(losses, np.float64)
    golden["batch_users"] = np.stack([b[0] for b in batches])
    golden["batch_pos"] = np.stack([b[1] for b in batches])
    golden["batch_neg"] = np.stack([b[2] for b in batches])

    # evaluate (main.py:404-439) with torch.topk recorded
    rec = []
    real_topk = torch.topk

    def spy(*a, **kw):
        out = real_topk(*a, **kw)
        rec.append((out[0].numpy().copy(), out[1].numpy().copy()))
        return out

    with mock.patch.object(torch, "topk", spy), mock.patch.object(main, "tqdm", lambda x, **k: x):
        recall, ndcg = main.evaluate(model, val_df, train_df, adj, 20, dev, use_brand=use_brand)
    golden["eval/recall"] = np.float64(recall)
    golden["eval/ndcg"] = np.float64(ndcg)
    golden["eval/topk_scores"] = np.concatenate([r[0] for r in rec])
    golden["eval/topk_ids"] = np.concatenate([r[1] for r in rec]).astype(np.int32)
    golden["eval/users"] = np.asarray(list(dict(zip(val_df["user_idx"], val_df["item_idx"])).keys()),
                                      np.int64)
    with torch.no_grad():
        fu, fi, _, _, _ = model(adj)
        golden["eval/F_user"] = fu.numpy().copy()
        golden["eval/F_item"] = fi.numpy().copy()

    out = os.path.join(REPO, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **golden)
    print(f"wrote {out}: recall@20={recall:.4f} ndcg@20={ndcg:.4f} "
          f"loss[0]={losses[0]:.6f} loss[-1]={losses[-1]:.6f} "
          f"({os.path.getsize(out) / 1024:.0f} KiB)")


def _digest(a):
    """fp64 (sum, sum of squares) of an array: a compact witness of a large tensor."""
    a = np.asarray(a, np.float64)
    return np.asarray([a.sum(), (a * a).sum()], np.float64)


def run_config_scale_case(main, model_cls, name, shape, d, K, bs, n_steps, seed, sample_rows=256):
    """BASELINE.json configs[1] scale (the Gowalla-shape graph): ``n_steps`` recorded steps of the
    reference's own loop (reference main.py:488-531) followed by ``main.evaluate`` (main.py:404-439)
    over ALL validation users.  Inputs regenerate from ``synth.generate(shape, seed)`` and
    ``torch.manual_seed(42)``, so only the outputs are stored: the loss curve, the recorded batch
    stream (int32), the top-20 ids + scores of every validation user, recall / NDCG, and fp64
    digests + a row sample of the large tensors (adjacency, init / final parameters, final F)."""
    import torch

    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    inter = synth.generate(shape, seed=seed)
    golden = dict(shape=np.asarray(shape), seed=seed, d=d, K=K, bs=bs, lam=1e-4, lr=1e-3,
                  num_users=inter.num_users, num_items=inter.num_items, num_brands=inter.num_brands)
    with tempfile.TemporaryDirectory() as tmp:
        ddir = os.path.join(tmp, "processed")
        synth.write_reference_format(inter, ddir, None)
        dev = torch.device("cpu")
        (train_df, val_df, test_df, U, I, B, adj, _) = main.load_preprocessed_data(
            ddir, dev, use_brand=False, debug=False)
    tu, ti = train_df["user_idx"].values, train_df["item_idx"].values
    golden.update(n_train=len(tu), n_val=len(val_df), nnz=int(adj._nnz()),
                  adj_val_digest=_digest(adj._values().numpy()),
                  adj_col_digest=_digest(adj._indices()[1].numpy()),
                  train_digest=_digest(tu * I + ti),
                  val_digest=_digest(val_df["user_idx"].values * I + val_df["item_idx"].values))
    cfg = types.SimpleNamespace(embedding_dim=d, n_layers=K, debug=False, device=dev,
                                learning_rate=1e-3, weight_decay=1e-4, top_k=20, batch_size=bs)
    main.config = cfg
    torch.manual_seed(42)                                     # reference main.py:607
    model = model_cls(U, I, B, cfg).to(dev)
    rows = np.random.default_rng(seed + 3).choice(min(U, I), sample_rows, replace=False)
    golden["sample_rows"] = rows
    for k, v in model.state_dict().items():
        golden["init_digest/" + k] = _digest(v.numpy())
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)   # main.py:469
    with torch.no_grad():
        fu, fi, fb, _, _ = model(adj, use_brand=False)
        golden["fwd_digest/user"], golden["fwd_digest/item"] = _digest(fu.numpy()), _digest(fi.numpy())
        golden["fwd_sample/user"], golden["fwd_sample/item"] = fu.numpy()[rows].copy(), fi.numpy()[rows].copy()
    rng = np.random.default_rng(seed + 7)
    batches = _batches(rng, tu, ti, I, bs, n_steps)
    losses = []
    model.train()
    for s, (u, p, n) in enumerate(batches):                   # main.py:488-531
        users, pos, neg = torch.from_numpy(u), torch.from_numpy(p), torch.from_numpy(n)
        opt.zero_grad()
        fu, fi, fb, u0, i0 = model(adj, use_brand=False)
        loss = main.bpr_loss_reg(fu[users], fi[pos], fi[neg], u0[users], i0[pos], i0[neg],
                                 cfg.weight_decay)
        loss.backward()
        opt.step()
        losses.append(loss.item())
        print(f"  step {s}: loss {losses[-1]:.6f}", flush=True)
    golden["losses"] = np.asarray(losses, np.float64)
    golden["batch_users"] = np.stack([b[0] for b in batches]).astype(np.int32)
    golden["batch_pos"] = np.stack([b[1] for b in batches]).astype(np.int32)
    golden["batch_neg"] = np.stack([b[2] for b in batches]).astype(np.int32)
    for k, v in model.state_dict().items():
        golden["final_digest/" + k] = _digest(v.numpy())
        golden["final_sample/" + k] = v.numpy()[rows % v.shape[0]].copy()

    rec = []
    real_topk = torch.topk

    def spy(*a, **kw):
        out = real_topk(*a, **kw)
        rec.append((out[0].numpy().copy(), out[1].numpy().copy()))
        return out

    with mock.patch.object(torch, "topk", spy), mock.patch.object(main, "tqdm", lambda x, **k: x):
        recall, ndcg = main.evaluate(model, val_df, train_df, adj, 20, dev, use_brand=False)
    golden["eval/recall"] = np.float64(recall)
    golden["eval/ndcg"] = np.float64(ndcg)
    golden["eval/topk_scores"] = np.concatenate([r[0] for r in rec])
    golden["eval/topk_ids"] = np.concatenate([r[1] for r in rec]).astype(np.int32)
    golden["eval/users"] = np.asarray(list(dict(zip(val_df["user_idx"], val_df["item_idx"])).keys()),
                                      np.int32)
    with torch.no_grad():
        fu, fi, _, _, _ = model(adj)
        golden["evalF_digest/user"], golden["evalF_digest/item"] = _digest(fu.numpy()), _digest(fi.numpy())
    out = os.path.join(REPO, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **golden)
    print(f"wrote {out}: recall@20={recall:.6f} ndcg@20={ndcg:.6f} loss[0]={losses[0]:.6f} "
          f"loss[-1]={losses[-1]:.6f} ({os.path.getsize(out) / 1024:.0f} KiB)")


def main_():
    main, lg, lf = _import_reference()
    os.makedirs(os.path.join(REPO, "tests", "golden"), exist_ok=True)
    if "--brand-loss" in sys.argv:         # the brand BPR term on the tripartite graph
        run_case(main, lg.LightGCN, "tiny_brandloss_d64_k3", "tiny", 64, 3, 256, 3, True, False, 2,
                 brand_loss_weight=0.1)
        return
    if "--config-scale" in sys.argv:       # ~2 min of CPU: the Gowalla-shape run (configs[1])
        run_config_scale_case(main, lg.LightGCN, "gowalla_lightgcn_d64_k3", "gowalla", 64, 3, 2048, 20, 0)
        return
    run_case(main, lg.LightGCN, "tiny_lightgcn_d64_k3", "tiny", 64, 3, 256, 6, False, False, 0)
    run_case(main, lg.LightGCN, "tiny_lightgcn_d128_k4", "tiny", 128, 4, 256, 3, False, False, 1)
    run_case(main, lg.LightGCN, "tiny_lightgcn_brand_d64_k3", "tiny", 64, 3, 256, 3, True, False, 2)
    run_case(main, lf.LightGCN_Fusion, "tiny_fusion_d64_k3", "tiny", 64, 3, 256, 4, False, True, 3)
    # edge cases (oracle-only): isolated users, repeated interactions, a single layer, d = 32
    run_case(main, lg.LightGCN, "tiny_edge_d32_k1", "tiny", 32, 1, 64, 3, False, False, 4,
             min_degree=2, n_dup=40)


if __name__ == "__main__":
    main_()
