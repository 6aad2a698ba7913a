"""Pins the CPU oracle (oracle/) against outputs of the reference itself.

tests/golden/*.npz were produced by oracle/make_golden.py running the unmodified
reference (main.load_preprocessed_data, models.*, main.bpr_loss_reg, torch Adam,
main.evaluate) on CPU.  Integer / index work must be bit-exact; fp32 propagation is
bit-exact too (sequential FMA); loss / grads / Adam within 1e-5 relative.
"""
import numpy as np
import pytest

from conftest import config_scale_inputs, digest as _digest, near_tie_rows_ok, rel_err
from oracle import lgcn_oracle as orc

# tiny_edge_d32_k1: users without any training edge (degree 0), repeated interactions
# (multiplicity 2 and 3 in the adjacency), a single layer, d = 32
CASES = ["tiny_lightgcn_d64_k3", "tiny_lightgcn_d128_k4", "tiny_lightgcn_brand_d64_k3",
         "tiny_fusion_d64_k3", "tiny_edge_d32_k1"]
TOL = 1e-5


def _adj(g):
    ib = (g["item_brand_item"], g["item_brand_brand"]) if "item_brand_item" in g else None
    return orc.build_norm_adj(g["train_user"], g["train_item"], int(g["num_users"]),
                              int(g["num_items"]), int(g["num_brands"]), ib)


def _tables(g, prefix):
    item_key = "item_id_embedding.weight" if prefix + "/item_id_embedding.weight" in g \
        else "item_embedding.weight"
    return (g[prefix + "/user_embedding.weight"], g[prefix + "/" + item_key],
            g[prefix + "/brand_embedding.weight"], item_key)


def _e0(g, prefix):
    u, i, b, key = _tables(g, prefix)
    if key == "item_id_embedding.weight":
        i_in, pre = orc.fusion_forward(i, g["init/item_content_embedding"],
                                       g[prefix + "/item_fusion_layer.weight"],
                                       g[prefix + "/item_fusion_layer.bias"])
    else:
        i_in, pre = i, None
    return np.concatenate([u, i_in, b], 0), pre


@pytest.mark.parametrize("case", CASES)
def test_validation_split_first_row_per_user(golden, case):
    g = golden(case)
    from gcn_recommendation_b200.synth import Interactions
    inter = Interactions(int(g["num_users"]), int(g["num_items"]), int(g["num_brands"]),
                         g["all_train_user"], g["all_train_item"], g["test_user"], g["test_item"])
    tu, ti, vu, vi = inter.split_validation()
    assert np.array_equal(tu, g["train_user"]) and np.array_equal(ti, g["train_item"])
    assert np.array_equal(vu, g["val_user"]) and np.array_equal(vi, g["val_item"])


@pytest.mark.parametrize("case", CASES)
def test_adjacency_bit_exact(golden, case):
    g = golden(case)
    a = _adj(g)
    N = int(g["num_users"]) + int(g["num_items"]) + int(g["num_brands"])
    rows = np.repeat(np.arange(N), np.diff(a["rowptr"])).astype(np.int32)
    assert np.array_equal(rows, g["adj_row"])
    assert np.array_equal(a["col"], g["adj_col"])
    assert np.array_equal(a["val"].view(np.uint32), g["adj_val"].view(np.uint32))


@pytest.mark.parametrize("case", CASES)
def test_forward_bit_exact_vs_torch_sparse(golden, case):
    g = golden(case)
    a = _adj(g)
    E0, _ = _e0(g, "init")
    F, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0, int(g["K"]))
    U, I = int(g["num_users"]), int(g["num_items"])
    ref = np.concatenate([g["fwd/user"], g["fwd/item"], g["fwd/brand"]], 0)
    if "fusion" in case:     # sgemm order is not pinned -> 1e-5
        mx, fro = rel_err(F, ref)
        assert mx < TOL and fro < TOL
    else:
        assert np.array_equal(F.view(np.uint32), ref.view(np.uint32))
    assert F[:U].shape == g["fwd/user"].shape and F[U:U + I].shape == g["fwd/item"].shape


@pytest.mark.parametrize("case", CASES)
def test_loss_grads_adam(golden, case):
    g = golden(case)
    a = _adj(g)
    U, I, K = int(g["num_users"]), int(g["num_items"]), int(g["K"])
    E0, pre = _e0(g, "init")
    F, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0, K)
    u0, i0, b0, item_key = _tables(g, "init")
    loss, gF, gU, gI = orc.bpr_loss(F, u0, i0, g["batch_users"][0], g["batch_pos"][0],
                                    g["batch_neg"][0], U, float(g["lam"]))
    assert abs(loss - g["losses"][0]) <= TOL * abs(g["losses"][0])
    dE0 = orc.propagate_backward(a["rowptr"], a["col"], a["val"], gF, K)
    grads = {"user_embedding.weight": dE0[:U] + gU, "brand_embedding.weight": dE0[U + I:]}
    if pre is None:
        grads[item_key] = dE0[U:U + I] + gI
    else:
        gE, gW, gb = orc.fusion_backward(i0, g["init/item_content_embedding"],
                                         g["init/item_fusion_layer.weight"], pre, dE0[U:U + I])
        grads[item_key] = gE + gI
        grads["item_fusion_layer.weight"] = gW
        grads["item_fusion_layer.bias"] = gb
    for k, v in grads.items():
        mx, fro = rel_err(v, g["grad1/" + k])
        assert mx < TOL and fro < TOL, (k, mx, fro)
    for k, gr in grads.items():
        p = g["init/" + k].copy()
        m = np.zeros_like(p)
        v = np.zeros_like(p)
        orc.adam_step(p, g["grad1/" + k], m, v, 1, lr=float(g["lr"]))
        mx, fro = rel_err(p, g["step1/" + k])
        assert mx < TOL and fro < TOL, (k, mx, fro)


@pytest.mark.parametrize("case", CASES)
def test_multi_step_training_tracks_reference(golden, case):
    """Oracle train loop (propagate -> bpr -> backward -> adam) vs the reference's losses."""
    g = golden(case)
    a = _adj(g)
    U, I, K = int(g["num_users"]), int(g["num_items"]), int(g["K"])
    names = [k[5:] for k in g if k.startswith("init/") and k != "init/item_content_embedding"]
    P = {k: g["init/" + k].copy() for k in names}
    M = {k: np.zeros_like(P[k]) for k in names}
    V = {k: np.zeros_like(P[k]) for k in names}
    fusion = "item_fusion_layer.weight" in P
    ik = "item_id_embedding.weight" if fusion else "item_embedding.weight"
    for s in range(len(g["losses"])):
        if fusion:
            it, pre = orc.fusion_forward(P[ik], g["init/item_content_embedding"],
                                         P["item_fusion_layer.weight"], P["item_fusion_layer.bias"])
        else:
            it, pre = P[ik], None
        E0 = np.concatenate([P["user_embedding.weight"], it, P["brand_embedding.weight"]], 0)
        F, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0, K)
        loss, gF, gU, gI = orc.bpr_loss(F, P["user_embedding.weight"], P[ik], g["batch_users"][s],
                                        g["batch_pos"][s], g["batch_neg"][s], U, float(g["lam"]))
        assert abs(loss - g["losses"][s]) <= 2e-5 * abs(g["losses"][s]), (s, loss, g["losses"][s])
        dE0 = orc.propagate_backward(a["rowptr"], a["col"], a["val"], gF, K)
        G = {"user_embedding.weight": dE0[:U] + gU, "brand_embedding.weight": dE0[U + I:]}
        if fusion:
            gE, gW, gb = orc.fusion_backward(P[ik], g["init/item_content_embedding"],
                                             P["item_fusion_layer.weight"], pre, dE0[U:U + I])
            G[ik] = gE + gI
            G["item_fusion_layer.weight"], G["item_fusion_layer.bias"] = gW, gb
        else:
            G[ik] = dE0[U:U + I] + gI
        for k in names:
            orc.adam_step(P[k], G[k], M[k], V[k], s + 1, lr=float(g["lr"]))
    for k in names:
        # Adam divides by sqrt(v): elements whose gradient is ~0 amplify fp32 rounding
        # differences (m/sqrt(v) ~ sign(g)), so the element-wise max is looser than the norm.
        mx, fro = rel_err(P[k], g["final/" + k])
        assert mx < 1e-3 and fro < 1e-5, (k, mx, fro)


@pytest.mark.parametrize("case", CASES)
def test_eval_topk_and_metrics(golden, case):
    g = golden(case)
    users, targets = orc.eval_pairs(g["val_user"], g["val_item"])
    assert np.array_equal(users, g["eval/users"])
    mr, mc = orc.mask_csr(users, g["train_user"], g["train_item"], int(g["num_users"]))
    ids, sc = orc.score_topk(g["eval/F_user"], g["eval/F_item"], users, mr, mc, 20)
    ref_ids, ref_sc = g["eval/topk_ids"], g["eval/topk_scores"]
    # ids must agree wherever the reference's neighbouring scores are not fp32 near-ties
    same = ids == ref_ids
    if not same.all():
        bad_rows = np.where(~same.all(1))[0]
        for r in bad_rows:
            assert sorted(ids[r].tolist()) == sorted(ref_ids[r].tolist()) or \
                np.abs(np.sort(sc[r]) - np.sort(ref_sc[r])).max() <= 1e-6 * np.abs(ref_sc[r]).max()
            gap = np.abs(np.diff(ref_sc[r]))
            assert gap.min() <= 4e-7 * np.abs(ref_sc[r]).max(), "id mismatch without a near-tie"
    mx, _ = rel_err(sc, ref_sc)
    assert mx < TOL
    rec, ndcg = orc.recall_ndcg(ids, targets)
    assert rec == pytest.approx(float(g["eval/recall"]), abs=1e-12)
    assert ndcg == pytest.approx(float(g["eval/ndcg"]), abs=1e-12)


@pytest.mark.parametrize("case", CASES[:2])
def test_torch_port_tracks_reference(golden, case):
    """oracle/torch_port.py (the timed CPU baseline) reproduces the reference's loss curve,
    final parameters and validation metrics on the golden run."""
    import torch
    from oracle.torch_port import TorchPort
    g = golden(case)
    torch.set_num_threads(1)
    U, I, B = int(g["num_users"]), int(g["num_items"]), int(g["num_brands"])
    port = TorchPort(g["train_user"], g["train_item"], U, I, B, int(g["d"]), int(g["K"]),
                     lr=float(g["lr"]), lam=float(g["lam"]), seed=42)
    assert np.array_equal(port.user.weight.detach().numpy(), g["init/user_embedding.weight"])
    assert np.array_equal(port.item.weight.detach().numpy(), g["init/item_embedding.weight"])
    losses = [port.step(g["batch_users"][s], g["batch_pos"][s], g["batch_neg"][s])
              for s in range(len(g["losses"]))]
    assert np.allclose(losses, g["losses"], rtol=1e-6, atol=0)
    assert np.allclose(port.item.weight.detach().numpy(), g["final/item_embedding.weight"],
                       rtol=0, atol=1e-7)
    users, targets = orc.eval_pairs(g["val_user"], g["val_item"])
    lists = {}
    for u, i in zip(g["train_user"].tolist(), g["train_item"].tolist()):
        lists.setdefault(u, []).append(i)
    rec, ndcg = port.evaluate(users, targets, lists, 20)
    assert rec == pytest.approx(float(g["eval/recall"]), abs=1e-12)
    assert ndcg == pytest.approx(float(g["eval/ndcg"]), abs=1e-12)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[1] scale: the Gowalla-shape run of the unmodified reference
# (oracle/make_golden.py --config-scale: 20 recorded steps of main.py:488-531 + main.evaluate)
# ---------------------------------------------------------------------------------------------
CONFIG_GOLDEN = "gowalla_lightgcn_d64_k3"


def test_config_scale_golden_pins_the_oracle(golden):
    """The oracle at the full Gowalla shape against the reference's own run: adjacency digests,
    seed-identical init, bit-exact forward rows, the 20-step loss curve (2e-5), final parameters
    (1e-5 Frobenius on the sampled rows / digests) and the top-20 of a user sample."""
    import types

    import torch

    from models.lightgcn import LightGCN
    g = golden(CONFIG_GOLDEN)
    inter, tu, ti, vu, vi = config_scale_inputs(g)
    U, I, B, d, K = inter.num_users, inter.num_items, inter.num_brands, int(g["d"]), int(g["K"])
    a = orc.build_norm_adj(tu, ti, U, I, B)
    assert len(a["col"]) == int(g["nnz"])
    assert np.array_equal(_digest(a["val"]), g["adj_val_digest"])
    assert np.array_equal(_digest(a["col"]), g["adj_col_digest"])
    torch.manual_seed(42)
    m = LightGCN(U, I, B, types.SimpleNamespace(embedding_dim=d, n_layers=K, debug=False))
    sd = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        assert np.array_equal(_digest(v), g["init_digest/" + k]), k
    P = {k: sd[k] for k in ("user_embedding.weight", "item_embedding.weight", "brand_embedding.weight")}
    rows = g["sample_rows"]
    E0 = np.concatenate([P["user_embedding.weight"], P["item_embedding.weight"], P["brand_embedding.weight"]], 0)
    F, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0, K)
    assert np.array_equal(F[:U][rows].view(np.uint32), g["fwd_sample/user"].view(np.uint32))
    assert np.array_equal(F[U:U + I][rows].view(np.uint32), g["fwd_sample/item"].view(np.uint32))
    M = {k: np.zeros_like(v) for k, v in P.items()}
    V = {k: np.zeros_like(v) for k, v in P.items()}
    for s in range(len(g["losses"])):
        E0 = np.concatenate([P["user_embedding.weight"], P["item_embedding.weight"], P["brand_embedding.weight"]], 0)
        F, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0, K)
        loss, gF, gU, gI = orc.bpr_loss(F, P["user_embedding.weight"], P["item_embedding.weight"],
                                        g["batch_users"][s], g["batch_pos"][s], g["batch_neg"][s], U,
                                        float(g["lam"]))
        assert abs(loss - g["losses"][s]) <= 2e-5 * abs(g["losses"][s]), (s, loss, g["losses"][s])
        dE0 = orc.propagate_backward(a["rowptr"], a["col"], a["val"], gF, K)
        G = {"user_embedding.weight": dE0[:U] + gU, "item_embedding.weight": dE0[U:U + I] + gI,
             "brand_embedding.weight": dE0[U + I:]}
        for k in P:
            orc.adam_step(P[k], G[k], M[k], V[k], s + 1, lr=float(g["lr"]))
    for k in P:
        ref = g["final_sample/" + k]
        mx, fro = rel_err(P[k][rows % P[k].shape[0]], ref)
        assert fro < 1e-5 and mx < 1e-3, (k, mx, fro)
        dg = _digest(P[k])
        assert np.allclose(dg, g["final_digest/" + k], rtol=1e-5, atol=1e-9), k
    # top-20 of the first 384 validation users from the oracle's own final table
    E0 = np.concatenate([P["user_embedding.weight"], P["item_embedding.weight"], P["brand_embedding.weight"]], 0)
    F, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0, K)
    users, targets = orc.eval_pairs(vu, vi)
    assert np.array_equal(users, g["eval/users"])
    n = 384
    mr, mc = orc.mask_csr(users[:n], tu, ti, U)
    ids, sc = orc.score_topk(F[:U], F[U:U + I], users[:n], mr, mc, 20)
    assert np.allclose(sc, g["eval/topk_scores"][:n], rtol=2e-5, atol=1e-9)
    assert near_tie_rows_ok(ids, g["eval/topk_ids"][:n], g["eval/topk_scores"][:n]) <= n // 50


def test_brand_bpr_term_vs_reference(golden):
    """the brand / author BPR term of reference main.py:382-391 (recorded by calling
    the reference's own bpr_loss_reg with brand_loss=True on the tripartite graph): oracle loss,
    first-step gradients and the 3-step loss curve."""
    g = golden("tiny_brandloss_d64_k3")
    a = _adj(g)
    U, I, K = int(g["num_users"]), int(g["num_items"]), int(g["K"])
    w, lam = float(g["brand_loss_weight"]), float(g["lam"])
    names = ["user_embedding.weight", "item_embedding.weight", "brand_embedding.weight"]
    P = {k: g["init/" + k].copy() for k in names}
    M = {k: np.zeros_like(P[k]) for k in names}
    V = {k: np.zeros_like(P[k]) for k in names}
    for s in range(len(g["losses"])):
        E0 = np.concatenate([P[k] for k in names], 0)
        F, _ = orc.propagate(a["rowptr"], a["col"], a["val"], E0, K)
        u, p, n = g["batch_users"][s], g["batch_pos"][s], g["batch_neg"][s]
        loss, gF, gU, gI = orc.bpr_loss(F, P[names[0]], P[names[1]], u, p, n, U, lam)
        term, gFb = orc.bpr_brand_term(F, u, p, n, g["item_to_brand"], U, I)
        loss += w * term
        assert abs(loss - g["losses"][s]) <= 2e-5 * abs(g["losses"][s]), (s, loss, g["losses"][s])
        dE0 = orc.propagate_backward(a["rowptr"], a["col"], a["val"], gF + np.float32(w) * gFb, K)
        G = {names[0]: dE0[:U] + gU, names[1]: dE0[U:U + I] + gI, names[2]: dE0[U + I:]}
        if s == 0:
            for k in names:
                mx, fro = rel_err(G[k], g["grad1/" + k])
                assert mx < TOL and fro < TOL, (k, mx, fro)
            assert np.abs(g["grad1/" + names[2]]).max() > 0          # the brand rows do get gradients
        for k in names:
            orc.adam_step(P[k], G[k], M[k], V[k], s + 1, lr=float(g["lr"]))
    for k in names:
        mx, fro = rel_err(P[k], g["final/" + k])
        assert mx < 1e-3 and fro < 1e-5, (k, mx, fro)
