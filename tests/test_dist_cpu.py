"""world_size-2 gloo tests on CPU for the multi-GPU decompositions (host-side logic + the math
the sharded engines rely on; the kernels themselves are CUDA-only and covered by -m gpu)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lgcn_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_lightgcn_d64_k3.npz")


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = dict(np.load(GOLD))
        U, I, B, K = int(g["num_users"]), int(g["num_items"]), int(g["num_brands"]), int(g["K"])
        a = orc.build_norm_adj(g["train_user"], g["train_item"], U, I, B)
        E0 = np.concatenate([g["init/user_embedding.weight"], g["init/item_embedding.weight"],
                             g["init/brand_embedding.weight"]], 0)
        d = E0.shape[1]
        dl = d // world
        # ---- feature sharding: propagation is column-separable, dots need one all-reduce ----
        E0_loc = np.ascontiguousarray(E0[:, rank * dl:(rank + 1) * dl])
        F_loc, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0_loc, K)
        F_ref = np.concatenate([g["fwd/user"], g["fwd/item"], g["fwd/brand"]], 0)
        assert np.array_equal(F_loc, F_ref[:, rank * dl:(rank + 1) * dl])      # bit-exact slices
        u, p, n = g["batch_users"][0], g["batch_pos"][0], g["batch_neg"][0]
        part = np.stack([(F_loc[u] * F_loc[U + p]).sum(1), (F_loc[u] * F_loc[U + n]).sum(1),
                         (E0_loc[u] ** 2 + E0_loc[U + p] ** 2 + E0_loc[U + n] ** 2).sum(1)]).astype(np.float64)
        t = torch.from_numpy(part)
        dist.all_reduce(t)
        ps, ns, reg = t.numpy()
        sg = 1.0 / (1.0 + np.exp(-(ps - ns)))
        loss = float(np.mean(-np.log(sg + 1e-8)) + float(g["lam"]) * reg.sum() / len(u))
        assert abs(loss - g["losses"][0]) <= 1e-5 * abs(g["losses"][0])
        # ---- row sharding: all-gather of row blocks reproduces the table in global order ----
        N = U + I + B
        rpr = -(-N // world)
        r0, r1 = rank * rpr, min(N, (rank + 1) * rpr)
        loc = torch.zeros((rpr, d))
        loc[:r1 - r0] = torch.from_numpy(E0[r0:r1])
        full = torch.empty((rpr * world, d))
        dist.all_gather_into_tensor(full, loc)
        assert torch.equal(full[:N], torch.from_numpy(E0))
        rp = a["rowptr"][r0:r1 + 1] - a["rowptr"][r0]
        e0, e1 = a["rowptr"][r0], a["rowptr"][r1]
        Y_loc = np.zeros((r1 - r0, d), np.float32)
        full_np = full.numpy()
        import ctypes
        orc.lib().lgcn_oracle_spmm(orc._p(orc._i64(rp)), orc._p(orc._i32(a["col"][e0:e1])),
                                   orc._p(orc._f32(a["val"][e0:e1])), orc._p(orc._f32(full_np)),
                                   orc._p(Y_loc), ctypes.c_int64(r1 - r0), ctypes.c_int32(d))
        Y_ref = orc.spmm(a["rowptr"], a["col"], a["val"], E0)
        assert np.array_equal(Y_loc, Y_ref[r0:r1])
        # ---- fusion item block: column shards <-> full rows of the rank's item block ----------
        from gcn_recommendation_b200.dist import cols_to_rows, rows_to_cols
        Ei = torch.from_numpy(np.ascontiguousarray(E0[U:U + I]))            # [I, d], I odd-sized vs world
        ipr = -(-I // world)
        a_in, a_out = torch.zeros((world, ipr, dl)), torch.zeros((world, ipr, dl))
        a2a = lambda o, i: dist.all_to_all_single(o, i)                    # noqa: E731
        rows = cols_to_rows(Ei[:, rank * dl:(rank + 1) * dl].contiguous(), a_in, a_out,
                            torch.zeros((ipr, d)), a2a)
        i0, i1 = rank * ipr, min(I, (rank + 1) * ipr)
        assert torch.equal(rows[:i1 - i0], Ei[i0:i1])
        back = rows_to_cols(rows, a_in, a_out, torch.zeros((I, dl)), a2a)
        assert torch.equal(back, Ei[:, rank * dl:(rank + 1) * dl])
        # ---- reference-format checkpoints from feature-sharded engines (round trip) ----------
        from gcn_recommendation_b200.dist import FeatureShardedEngine, column_shard
        from gcn_recommendation_b200.graph import NormAdjCSR
        csr = NormAdjCSR(torch.from_numpy(a["rowptr"].astype(np.int32)),
                         torch.from_numpy(a["col"].astype(np.int32)), torch.from_numpy(a["val"]), N)
        table = torch.from_numpy(E0)
        eng = FeatureShardedEngine(csr, U, I, B, K, column_shard(table, rank, world), batch_size=8)
        sd = eng.state_dict()
        assert list(sd.keys()) == ["user_embedding.weight", "brand_embedding.weight", "item_embedding.weight"]
        for k in sd:
            assert torch.equal(sd[k], torch.from_numpy(g["init/" + k])), k
        eng2 = FeatureShardedEngine(csr, U, I, B, K, torch.zeros((N, dl)), batch_size=8)
        eng2.load_state_dict(sd)
        assert torch.equal(eng2.P, eng.P)
        # fusion variant: content sharded by item block, W / b replicated
        gen = torch.Generator().manual_seed(5)
        c = 12
        Cfull = torch.randn((I, c), generator=gen)
        W = torch.randn((d, d + c), generator=gen)
        bb = torch.randn((d,), generator=gen)
        fus = dict(content=Cfull[i0:i1].contiguous(), weight=W.clone(), bias=bb.clone(),
                   alltoall=a2a, allreduce=lambda t: dist.all_reduce(t))
        engf = FeatureShardedEngine(csr, U, I, B, K, column_shard(table, rank, world), batch_size=8,
                                    fusion=fus)
        sdf = engf.state_dict()
        assert list(sdf.keys()) == ["item_content_embedding", "user_embedding.weight",
                                    "item_id_embedding.weight", "brand_embedding.weight",
                                    "item_fusion_layer.weight", "item_fusion_layer.bias"]
        assert torch.equal(sdf["item_content_embedding"], Cfull)
        assert torch.equal(sdf["item_id_embedding.weight"], table[U:U + I])
        assert torch.equal(sdf["item_fusion_layer.weight"], W)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_decompositions_world2_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_column_shard_helper():
    from gcn_recommendation_b200.dist import column_shard
    from gcn_recommendation_b200._lib import LgcnError
    t = torch.arange(4 * 128, dtype=torch.float32).reshape(4, 128)
    parts = [column_shard(t, r, 4) for r in range(4)]
    assert all(p.is_contiguous() and p.shape == (4, 32) for p in parts)
    assert torch.equal(torch.cat(parts, 1), t)
    with pytest.raises(LgcnError):
        column_shard(t, 0, 16)          # local width 8 is not supported by the kernels
