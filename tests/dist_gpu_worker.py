"""torchrun worker: REAL-NCCL parity of the sharded engines (VERDICT r01 "what's weak" 1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/dist_gpu_worker.py

One process per GPU.  Every rank checks, through NCCL collectives on real devices:

1. ``FeatureShardedEngine`` on a golden recorded from the unmodified reference: loss curve
   (rtol 2e-5), final parameters re-assembled by ``state_dict()`` (Frobenius 1e-5), user-sharded
   ``evaluate`` -> recall@20 / NDCG@20 identical to ``main.evaluate``.
2. ``RowShardedEngine`` (north-star layout, per-layer all-gather) on the same kind of golden.
3. ``FeatureShardedEngine`` + fusion item block (all-to-all column shards <-> item rows).
4. A graph large enough for the large-graph kernels (1/64-scale Amazon shape, d = 128, K = 4):
   both sharded engines against the single-GPU engine run on the same device: loss, gathered
   parameters (Frobenius 1e-6) and the top-20 of 256 users.

Exits non-zero on the first mismatch; rank 0 prints ``DIST PARITY OK world=N``.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import rel_err  # noqa: E402
from gcn_recommendation_b200 import ops, synth  # noqa: E402
from gcn_recommendation_b200.dist import FeatureShardedEngine, RowShardedEngine, column_shard  # noqa: E402
from gcn_recommendation_b200.engine import LightGCNEngine, build_mask_csr, xavier_uniform_table  # noqa: E402
from gcn_recommendation_b200.graph import NormAdjCSR  # noqa: E402
from oracle import lgcn_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def t(a, dev, dt=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt).contiguous()


def golden_case(name, dev, item_key="item_embedding.weight"):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    U, I, B, K = int(g["num_users"]), int(g["num_items"]), int(g["num_brands"]), int(g["K"])
    csr = NormAdjCSR.from_interactions(g["train_user"], g["train_item"], U, I, B, dev)
    full = torch.cat([t(g["init/user_embedding.weight"], dev), t(g["init/" + item_key], dev),
                      t(g["init/brand_embedding.weight"], dev)]).contiguous()
    return g, U, I, B, K, csr, full


def run_golden(eng, g, dev, rank, world, label, item_key="item_embedding.weight"):
    losses = []
    for s in range(len(g["losses"])):
        u, p, n = (t(g[k][s], dev, torch.int64) for k in ("batch_users", "batch_pos", "batch_neg"))
        losses.append(float(eng.bpr_step(u, p, n, use_graph=False).item()))
    assert np.allclose(losses, g["losses"], rtol=2e-5, atol=0), (label, losses, g["losses"].tolist())
    sd = eng.state_dict()                                           # collective: every rank gets it all
    for k in ("user_embedding.weight", item_key, "brand_embedding.weight"):
        mx, fro = rel_err(sd[k].numpy(), g["final/" + k])
        assert mx < 1e-3 and fro < 1e-5, (label, k, mx, fro)
    # user-sharded evaluation against main.evaluate's metrics (reference main.py:404-439)
    users, targets = orc.eval_pairs(g["val_user"], g["val_item"])
    per = -(-len(users) // world)
    mu, mt = users[rank * per:(rank + 1) * per], targets[rank * per:(rank + 1) * per]
    mr, mc = build_mask_csr(mu, g["train_user"], g["train_item"], int(g["num_users"]), dev)
    rec, ndcg, _ = eng.evaluate(t(mu, dev, torch.int64), t(mt, dev, torch.int64), mr, mc, 20)
    assert abs(rec - float(g["eval/recall"])) < 1e-12, (label, rec, float(g["eval/recall"]))
    assert abs(ndcg - float(g["eval/ndcg"])) < 1e-12, (label, ndcg, float(g["eval/ndcg"]))
    return losses


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    done = []

    # 1. feature sharding on the reference's d = 128 golden (local width 128 / world >= 16)
    g, U, I, B, K, csr, full = golden_case("tiny_lightgcn_d128_k4", dev)
    if full.shape[1] % world == 0 and full.shape[1] // world >= 16:
        eng = FeatureShardedEngine(csr, U, I, B, K, column_shard(full, rank, world), lr=float(g["lr"]),
                                   weight_decay=float(g["lam"]), batch_size=int(g["bs"]))
        run_golden(eng, g, dev, rank, world, "feature")
        done.append("feature")

    # 2. row sharding (north-star layout) on the d = 64 golden
    g, U, I, B, K, csr, full = golden_case("tiny_lightgcn_d64_k3", dev)
    eng = RowShardedEngine(csr, U, I, B, K, full.clone(), lr=float(g["lr"]), weight_decay=float(g["lam"]),
                           batch_size=int(g["bs"]))
    run_golden(eng, g, dev, rank, world, "row")
    done.append("row")

    # 3. fusion item block, feature-sharded tables + item-sharded projection
    g, U, I, B, K, csr, full = golden_case("tiny_fusion_d64_k3", dev, "item_id_embedding.weight")
    if full.shape[1] % world == 0 and full.shape[1] // world >= 16:
        ipr = -(-I // world)
        i0, i1 = min(I, rank * ipr), min(I, (rank + 1) * ipr)
        C = t(g["init/item_content_embedding"], dev)
        fus = dict(content=C[i0:i1].contiguous(), weight=t(g["init/item_fusion_layer.weight"], dev),
                   bias=t(g["init/item_fusion_layer.bias"], dev))
        eng = FeatureShardedEngine(csr, U, I, B, K, column_shard(full, rank, world), fusion=fus,
                                   lr=float(g["lr"]), weight_decay=float(g["lam"]), batch_size=int(g["bs"]))
        run_golden(eng, g, dev, rank, world, "fusion", "item_id_embedding.weight")
        sd = eng.state_dict()
        mx, fro = rel_err(sd["item_fusion_layer.weight"].numpy(), g["final/item_fusion_layer.weight"])
        assert mx < 1e-3 and fro < 1e-5, ("fusion W", mx, fro)
        done.append("fusion")

    # 4. large-graph kernels: sharded engines vs the single-GPU engine on this device
    U, I, B, total, d, K = synth.SHAPES["amazon_64th"]
    inter = synth.generate("amazon_64th", seed=0)
    tu, ti, vu, vi = inter.split_validation()
    csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
    assert not ops._spmm_plan(U + I + B, d, csr.n_long)[1], "expected the large-graph path"
    gen = torch.Generator(device=dev).manual_seed(42)
    table = xavier_uniform_table([U, I, B], d, dev, gen)
    rng = np.random.default_rng(1)
    batches = []
    for _ in range(3):
        idx = rng.integers(0, len(tu), 2048)
        batches.append(tuple(t(a, dev, torch.int64) for a in (tu[idx], ti[idx], rng.integers(0, I, 2048))))
    single = LightGCNEngine(csr, U, I, B, K, table.clone(), batch_size=2048)
    ref_losses = [float(single.bpr_step(*b, use_graph=False).item()) for b in batches]
    users = t(vu[:256], dev, torch.int64)
    mr, mc = build_mask_csr(vu[:256], tu, ti, U, dev)
    ref_ids, ref_sc = single.rate_topk(users, mr, mc, 20)
    for label in ("feature", "row"):
        if label == "feature":
            if d % world or d // world < 16:
                continue
            eng = FeatureShardedEngine(csr, U, I, B, K, column_shard(table, rank, world), batch_size=2048)
        else:
            eng = RowShardedEngine(csr, U, I, B, K, table.clone(), batch_size=2048)
        losses = [float(eng.bpr_step(*b, use_graph=False).item()) for b in batches]
        assert np.allclose(losses, ref_losses, rtol=1e-6, atol=0), (label, losses, ref_losses)
        P = eng._full_table()[:U + I + B]
        mx, fro = rel_err(P.cpu().numpy(), single.P.cpu().numpy())
        assert fro < 1e-6 and mx < 1e-4, (label, mx, fro)
        F = eng.gather_final_table()
        ids, sc = ops.score_topk(F[:U], F[U:U + I], users, mr, mc, 20)
        same = (ids == ref_ids).all(1).float().mean().item()
        assert same >= 0.97, (label, same)
        assert torch.allclose(sc, ref_sc, rtol=1e-5, atol=1e-9), label
        done.append("large-" + label)
        del eng, P, F
        torch.cuda.empty_cache()

    dist.barrier()
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok)
    assert int(ok.item()) == world
    if rank == 0:
        print(f"DIST PARITY OK world={world} cases={','.join(done)}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
