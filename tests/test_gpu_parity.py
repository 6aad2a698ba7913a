"""GPU parity tests: the sm_100a kernels (through the C ABI) against the CPU oracle and the
golden vectors produced by the reference itself.

Bars (BASELINE.json north_star): bit-exact CSR construction / normalisation / top-k ids,
<= 1e-5 relative (fp32) for propagated embeddings, loss and gradients.  Propagation is in fact
bit-exact for rows below the long-row threshold (sequential FMA order).
"""
import types

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5
CASES = ["tiny_lightgcn_d64_k3", "tiny_lightgcn_d128_k4", "tiny_lightgcn_brand_d64_k3",
         "tiny_fusion_d64_k3", "tiny_edge_d32_k1"]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _orc():
    from oracle import lgcn_oracle
    return lgcn_oracle


def _graph(g, dev, **kw):
    from gcn_recommendation_b200.graph import NormAdjCSR
    ib = (g["item_brand_item"], g["item_brand_brand"]) if "item_brand_item" in g else None
    return NormAdjCSR.from_interactions(g["train_user"], g["train_item"], int(g["num_users"]),
                                        int(g["num_items"]), int(g["num_brands"]), dev,
                                        item_brand=ib, **kw)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _t(a, dev, dt=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt).contiguous()


# ---------------------------------------------------------------------------- a1 graph
@pytest.mark.parametrize("case", CASES)
def test_csr_build_bit_exact(golden, dev, case):
    g = golden(case)
    csr = _graph(g, dev)
    rows = np.repeat(np.arange(csr.n_rows), np.diff(csr.rowptr.cpu().numpy())).astype(np.int32)
    assert np.array_equal(rows, g["adj_row"])
    assert np.array_equal(csr.col.cpu().numpy(), g["adj_col"])
    assert np.array_equal(_bits(csr.val.cpu().numpy()), _bits(g["adj_val"]))


@pytest.mark.parametrize("case", CASES[:1] + CASES[2:3] + CASES[4:])
def test_csr_from_reference_coo_tensor(golden, dev, case):
    from gcn_recommendation_b200.graph import NormAdjCSR, csr_for
    g = golden(case)
    N = int(g["num_users"]) + int(g["num_items"]) + int(g["num_brands"])
    idx = torch.from_numpy(np.vstack([g["adj_row"], g["adj_col"]]).astype(np.int64))
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(g["adj_val"]), (N, N)).to(dev)  # main.py:336
    csr = NormAdjCSR.from_sparse_coo(adj)
    ref = _graph(g, dev)
    assert torch.equal(csr.rowptr, ref.rowptr) and torch.equal(csr.col, ref.col)
    assert torch.equal(csr.val, ref.val)
    assert csr_for(adj) is csr_for(adj)
    # an unsorted COO is coalesced first
    perm = torch.randperm(idx.shape[1])
    adj2 = torch.sparse_coo_tensor(idx[:, perm], torch.from_numpy(g["adj_val"])[perm], (N, N)).to(dev)
    csr2 = NormAdjCSR.from_sparse_coo(adj2)
    assert torch.equal(csr2.rowptr, ref.rowptr) and torch.equal(csr2.col, ref.col)
    assert torch.equal(csr2.val, ref.val)


# ---------------------------------------------------------------------------- a2 spmm
@pytest.mark.parametrize("d", [16, 32, 64, 128, 256])
def test_spmm_bit_exact_all_dims(dev, d):
    from gcn_recommendation_b200 import ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    orc = _orc()
    inter = synth.generate((500, 700, 2, 9000), seed=d)
    tu, ti, _, _ = inter.split_validation()
    csr = NormAdjCSR.from_interactions(tu, ti, 500, 700, 2, dev, long_row_threshold=0)
    a = orc.build_norm_adj(tu, ti, 500, 700, 2)
    rng = np.random.default_rng(d)
    X = rng.standard_normal((1202, d), dtype=np.float32)
    Y = ops.spmm(csr, _t(X, dev)).cpu().numpy()
    ref = orc.spmm(a["rowptr"], a["col"], a["val"], X)
    assert np.array_equal(_bits(Y), _bits(ref))
    # empty rows (isolated brand nodes, cold items) produce exact zeros
    deg = np.diff(a["rowptr"])
    assert (deg == 0).any() and not Y[deg == 0].any()


@pytest.mark.parametrize("d", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("kernel", ["ring", "chunk", "ring_hot", "chunk_hot"])
def test_spmm_large_graph_kernels_bit_exact(dev, d, kernel):
    """The large-graph kernels (cp.async ring kernel / 16-row register-batch chunks) forced onto
    a small graph: every epilogue bit-exact against the oracle, with short-row/long-row mixes,
    empty rows and a ragged last worker.  The ``_hot`` variants also force the streaming L2 hints
    and the column residency classes (300 hot columns, degree-1 columns evict_first): the class
    bits of the packed column index must never reach the address arithmetic."""
    from gcn_recommendation_b200 import _lib, ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    orc = _orc()
    U, I, B = 3001, 4003, 2
    inter = synth.generate((U, I, B, 70_000), seed=100 + d)
    tu, ti, _, _ = inter.split_validation()
    a = orc.build_norm_adj(tu, ti, U, I, B)
    N = U + I + B
    rng = np.random.default_rng(d)
    X = rng.standard_normal((N, d), dtype=np.float32)
    A1 = rng.standard_normal((N, d), dtype=np.float32)
    A2 = rng.standard_normal((N, d), dtype=np.float32)
    ref = orc.spmm(a["rowptr"], a["col"], a["val"], X)
    old, old_hot = ops.SPMM_FLAGS_EXTRA, ops.HOT_BYTES
    ops.SPMM_FLAGS_EXTRA = _lib.SPMM_F_BIG_PATH | (_lib.SPMM_F_NO_RING if kernel.startswith("chunk")
                                                      else _lib.SPMM_F_FORCE_RING)
    hot = kernel.endswith("_hot")
    if hot:
        ops.SPMM_FLAGS_EXTRA |= _lib.SPMM_F_STREAM_HINTS | (_lib.SPMM_F_COLD_FIRST if d == 64 else 0)
        ops.HOT_BYTES = 300 * 4 * d
    try:
        for thr in (0, 64):          # 0: every row on the sequential path; 64: hot items segmented
            csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev, long_row_threshold=thr, seg_len=32)
            short = np.diff(a["rowptr"]) <= (thr if thr else 1 << 30)
            xt = _t(X, dev)
            if hot:
                ops.spmm(csr, xt)
                c0 = csr.colval[:, 0]
                assert int((c0 < 0).sum()) > 0 and int(((c0 >> 30) & 1).sum()) > 0
                assert csr.n_hot == 300
            Y = ops.spmm(csr, xt).cpu().numpy()
            assert np.array_equal(_bits(Y[short]), _bits(ref[short]))
            assert rel_err(Y, ref)[0] < TOL
            Ya = ops.spmm(csr, xt, addend=_t(A1, dev)).cpu().numpy()
            assert np.array_equal(_bits(Ya[short]), _bits((A1 + ref)[short]))
            Ym = ops.spmm(csr, xt, mean_layers=[_t(A1, dev), _t(A2, dev), xt]).cpu().numpy()
            want = (((A1 + A2) + X) + ref) / np.float32(4)
            assert np.array_equal(_bits(Ym[short]), _bits(want[short]))
            assert rel_err(Ym, want)[0] < TOL
        # sparse-input hop (first Horner hop): x has few non-zero rows, flagged; zero rows are
        # never read (poisoned with NaN here to prove it)
        for thr in (0, 64):
            csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev, long_row_threshold=thr, seg_len=32)
            short = np.diff(a["rowptr"]) <= (thr if thr else 1 << 30)
            nzr = rng.choice(N, 600, replace=False)
            Xs = np.zeros((N, d), np.float32)
            Xs[nzr] = rng.standard_normal((600, d), dtype=np.float32)
            flag = np.zeros(N + 32, np.uint8)
            flag[nzr] = 1
            Xp = Xs.copy()
            Xp[flag[:N] == 0] = np.nan
            zr = torch.zeros(256, device=dev)
            Yx = ops.spmm(csr, _t(Xp, dev), addend=_t(A1, dev), x_rowflag=_t(flag, dev, torch.uint8),
                          zero_row=zr).cpu().numpy()
            want = A1 + orc.spmm(a["rowptr"], a["col"], a["val"], Xs)
            assert np.array_equal(_bits(Yx[short]), _bits(want[short]))
            assert rel_err(Yx, want)[0] < TOL
            # sparse OUTPUT: addend g' flagged like x (the first Horner hop); all-zero output rows
            # are reported in y_rowflag and left unwritten (the buffer is NaN-poisoned), and the
            # next hop gathers under those flags (zero rows of its input still poisoned)
            nz2 = rng.choice(N, 40, replace=False)
            G = np.zeros((N, d), np.float32)
            G[nz2] = rng.standard_normal((40, d), dtype=np.float32)
            gflag = np.zeros(N + 32, np.uint8)
            gflag[nz2] = 1
            Xp = G.copy()
            Xp[gflag[:N] == 0] = np.nan
            ft = _t(gflag, dev, torch.uint8)
            yflag = torch.full((N + 32,), 7, dtype=torch.uint8, device=dev)
            out1 = torch.full((N, d), float("nan"), device=dev)
            ops.spmm(csr, _t(Xp, dev), out=out1, addend=_t(G, dev), x_rowflag=ft, addend_rowflag=ft,
                     zero_row=zr, y_rowflag=yflag)
            hop1 = G + orc.spmm(a["rowptr"], a["col"], a["val"], G)
            yf = yflag[:N].cpu().numpy()
            assert set(np.unique(yf)) <= {0, 1}
            nz = np.abs(hop1).max(axis=1) > 0
            assert not (nz & (yf == 0)).any(), "a non-zero row was reported as zero"
            assert 40 <= yf.sum() < 0.5 * N
            o1 = out1.cpu().numpy()
            assert np.isnan(o1[yf == 0]).all(), "an all-zero row was written"
            live = (yf == 1) & short
            assert np.array_equal(_bits(o1[live]), _bits(hop1[live]))
            assert rel_err(o1[yf == 1], hop1[yf == 1])[0] < TOL
            out2 = ops.spmm(csr, out1, addend=_t(G, dev), x_rowflag=yflag, addend_rowflag=ft,
                            zero_row=zr).cpu().numpy()
            hop2 = G + orc.spmm(a["rowptr"], a["col"], a["val"], hop1)
            assert np.array_equal(_bits(out2[short]), _bits(hop2[short])) or thr
            assert rel_err(out2, hop2)[0] < TOL
        # Adam epilogue against the standalone Adam kernel fed with the same gradient
        csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev, long_row_threshold=0)
        xt = _t(X, dev)
        p1, m1, v1 = (_t(rng.standard_normal((N, d), dtype=np.float32) * s, dev) for s in (1.0, 0.01, 0.0))
        v1 = v1.abs() + 1e-4
        p2, m2, v2 = p1.clone(), m1.clone(), v1.clone()
        sc = torch.tensor([1e-3 / (1 - 0.9 ** 3), (1 - 0.999 ** 3) ** 0.5], device=dev)
        gout = torch.empty((N, d), device=dev)
        ops.spmm_adam(csr, xt, p1, m1, v1, sc, addend=_t(A1, dev), addend2=_t(A2, dev), g_out=gout)
        gwant = (ref + A1) + A2
        assert np.array_equal(_bits(gout.cpu().numpy()), _bits(gwant))
        ops.SPMM_FLAGS_EXTRA = 0
        ops.spmm_adam(csr, xt, p2, m2, v2, sc, addend=_t(A1, dev), addend2=_t(A2, dev))
        for x1, x2 in ((p1, p2), (m1, m2), (v1, v2)):
            assert torch.equal(x1, x2)
    finally:
        ops.SPMM_FLAGS_EXTRA, ops.HOT_BYTES = old, old_hot


@pytest.mark.parametrize("d", [16, 64, 128])
@pytest.mark.parametrize("path", ["small", "ring", "chunk"])
def test_spmm_layer0_override_and_adam_row_skip(dev, d, path):
    """LightGCN_Fusion plumbing of lgcn_spmm: ``x_alt`` (the item rows of layer 0 come from the
    projected block, reference models/lightgcn_fusion.py:52 without the concat) must equal the
    same call on the materialised table BIT FOR BIT in every kernel and mode; the ADAM hop's row
    skip leaves the skipped rows' p / m / v alone, hands their gradient on, and updates the others
    exactly like the unskipped call."""
    from gcn_recommendation_b200 import _lib, ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    U, I, B = 1013, 1511, 3
    inter = synth.generate((U, I, B, 26_000), seed=11 + d)
    tu, ti, _, _ = inter.split_validation()
    N = U + I + B
    gen = torch.Generator(device=dev).manual_seed(d)
    X = torch.randn((N, d), device=dev, generator=gen)
    H = torch.randn((I, d), device=dev, generator=gen)
    A1 = torch.randn((N, d), device=dev, generator=gen)
    A2 = torch.randn((N, d), device=dev, generator=gen)
    Xcat = X.clone()
    Xcat[U:U + I] = H
    alt = (H, U)
    old = ops.SPMM_FLAGS_EXTRA
    ops.SPMM_FLAGS_EXTRA = {"small": 0, "ring": _lib.SPMM_F_BIG_PATH, "chunk": _lib.SPMM_F_BIG_PATH | _lib.SPMM_F_NO_RING}[path]
    try:
        csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev, long_row_threshold=48, seg_len=32)
        assert csr.n_long > 0
        E1 = ops.spmm(csr, Xcat)
        assert torch.equal(ops.spmm(csr, X, x_alt=alt), E1)
        assert torch.equal(ops.spmm(csr, X, addend=A1, x_alt=alt), ops.spmm(csr, Xcat, addend=A1))
        assert torch.equal(ops.spmm(csr, X, mean_layers=[X], x_alt=alt, layer0_alt=alt),
                           ops.spmm(csr, Xcat, mean_layers=[Xcat]))                      # K = 1
        assert torch.equal(ops.spmm(csr, E1, mean_layers=[X, E1], layer0_alt=alt),
                           ops.spmm(csr, E1, mean_layers=[Xcat, E1]))                    # K = 2
        assert torch.equal(ops.propagate(csr, X, 3, alt=alt), ops.propagate(csr, Xcat, 3))
        with pytest.raises(_lib.LgcnError):
            ops.spmm(csr, X, addend=A1, x_rowflag=torch.ones(N + 32, dtype=torch.uint8, device=dev),
                     zero_row=torch.zeros(256, device=dev), x_alt=alt)
        # ADAM row skip
        sc = torch.tensor([1e-3 / (1 - 0.9 ** 2), (1 - 0.999 ** 2) ** 0.5], device=dev)
        p0 = torch.randn((N, d), device=dev, generator=gen)
        m0 = torch.randn((N, d), device=dev, generator=gen) * 0.01
        v0 = torch.rand((N, d), device=dev, generator=gen) * 1e-3 + 1e-5
        pf, mf, vf = p0.clone(), m0.clone(), v0.clone()
        ops.spmm_adam(csr, X, pf, mf, vf, sc, addend=A1, addend2=A2)                     # every row
        g1 = torch.empty((N, d), device=dev)
        ops.spmm_adam(csr, X, p0.clone(), m0.clone(), v0.clone(), sc, addend=A1, g_out=g1)  # A x + A1
        ps, ms, vs = p0.clone(), m0.clone(), v0.clone()
        gs = torch.full((I, d), float("nan"), device=dev)
        ops.spmm_adam(csr, X, ps, ms, vs, sc, addend=A1, addend2=A2, skip=(gs, U))
        inside = torch.zeros(N, dtype=torch.bool, device=dev)
        inside[U:U + I] = True
        for got, full, init in ((ps, pf, p0), (ms, mf, m0), (vs, vf, v0)):
            assert torch.equal(got[~inside], full[~inside]) and torch.equal(got[inside], init[inside])
        assert torch.equal(gs, g1[U:U + I])
    finally:
        ops.SPMM_FLAGS_EXTRA = old


@pytest.mark.parametrize("d", [64, 128])
def test_spmm_long_row_plan_within_tolerance(dev, d):
    """Rows longer than the threshold are summed segment-wise: deterministic, <= 1e-5."""
    from gcn_recommendation_b200 import ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    orc = _orc()
    inter = synth.generate("small", seed=3)
    tu, ti, _, _ = inter.split_validation()
    U, I, B = inter.num_users, inter.num_items, inter.num_brands
    csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev, long_row_threshold=64, seg_len=32)
    assert csr.n_long > 10 and csr.n_seg > csr.n_long
    a = orc.build_norm_adj(tu, ti, U, I, B)
    X = np.random.default_rng(0).standard_normal((U + I + B, d), dtype=np.float32)
    xt = _t(X, dev)
    Y1 = ops.spmm(csr, xt).cpu().numpy()
    Y2 = ops.spmm(csr, xt).cpu().numpy()
    assert np.array_equal(_bits(Y1), _bits(Y2)), "segment path must be deterministic"
    ref = orc.spmm(a["rowptr"], a["col"], a["val"], X)
    mx, fro = rel_err(Y1, ref)
    assert mx < TOL and fro < TOL
    short = np.diff(a["rowptr"]) <= 64
    assert np.array_equal(_bits(Y1[short]), _bits(ref[short]))
    # epilogues on the long-row path: add and mean
    add = np.random.default_rng(1).standard_normal(X.shape, dtype=np.float32)
    Ya = ops.spmm(csr, xt, addend=_t(add, dev)).cpu().numpy()
    mx, fro = rel_err(Ya, add + ref)
    assert mx < TOL and fro < TOL
    Ym = ops.spmm(csr, xt, mean_layers=[_t(add, dev), xt]).cpu().numpy()
    mx, fro = rel_err(Ym, ((add + X) + ref) / np.float32(3))
    assert mx < TOL and fro < TOL


@pytest.mark.parametrize("case", CASES[:3] + CASES[4:])
def test_propagate_bit_exact_vs_reference(golden, dev, case):
    from gcn_recommendation_b200 import ops
    g = golden(case)
    csr = _graph(g, dev)
    E0 = np.concatenate([g["init/user_embedding.weight"], g["init/item_embedding.weight"],
                         g["init/brand_embedding.weight"]], 0)
    F = ops.propagate(csr, _t(E0, dev), int(g["K"])).cpu().numpy()
    ref = np.concatenate([g["fwd/user"], g["fwd/item"], g["fwd/brand"]], 0)
    assert np.array_equal(_bits(F), _bits(ref))


# ---------------------------------------------------------------------------- drop-in models
def _model(g, case, dev):
    cfg = types.SimpleNamespace(embedding_dim=int(g["d"]), n_layers=int(g["K"]), debug=False)
    U, I, B = int(g["num_users"]), int(g["num_items"]), int(g["num_brands"])
    if "fusion" in case:
        from models.lightgcn_fusion import LightGCN_Fusion
        m = LightGCN_Fusion(U, I, B, cfg, pretrained_item_emb=g["init/item_content_embedding"])
    else:
        from models.lightgcn import LightGCN
        m = LightGCN(U, I, B, cfg)
    sd = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init/")}
    m.load_state_dict(sd)
    return m.to(dev)


def _ref_coo(g, dev):
    N = int(g["num_users"]) + int(g["num_items"]) + int(g["num_brands"])
    idx = torch.from_numpy(np.vstack([g["adj_row"], g["adj_col"]]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(g["adj_val"]), (N, N)).to(dev)


def _bpr_loss_reg(fu, fp, fn, u0, p0, n0, lam):
    """The loss the unmodified main.py applies to the drop-in's outputs (main.py:366-402)."""
    ps = torch.sum(fu * fp, dim=1)
    ns = torch.sum(fu * fn, dim=1)
    bpr = -torch.mean(torch.log(torch.sigmoid(ps - ns) + 1e-8))
    reg = lam * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / float(len(fu))
    return bpr + reg


@pytest.mark.parametrize("case", CASES)
def test_dropin_forward_backward_under_reference_loop(golden, dev, case):
    """Drive the drop-in module the way main.train does (main.py:488-531): forward on the COO
    tensor, main.py's own loss, autograd backward, torch Adam."""
    g = golden(case)
    model = _model(g, case, dev)
    adj = _ref_coo(g, dev)
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]))
    losses = []
    for s in range(len(g["losses"])):
        users = _t(g["batch_users"][s], dev, torch.int64)
        pos = _t(g["batch_pos"][s], dev, torch.int64)
        neg = _t(g["batch_neg"][s], dev, torch.int64)
        opt.zero_grad()
        fu, fi, fb, u0, i0 = model(adj, use_brand=True)
        if s == 0:
            ref = np.concatenate([g["fwd/user"], g["fwd/item"], g["fwd/brand"]], 0)
            got = torch.cat([fu, fi, fb]).detach().cpu().numpy()
            if "fusion" in case:
                mx, fro = rel_err(got, ref)
                assert mx < TOL and fro < TOL
            else:
                assert np.array_equal(_bits(got), _bits(ref))
        loss = _bpr_loss_reg(fu[users], fi[pos], fi[neg], u0[users], i0[pos], i0[neg], float(g["lam"]))
        loss.backward()
        if s == 0:
            for k, p in model.named_parameters():
                mx, fro = rel_err(p.grad.cpu().numpy(), g["grad1/" + k])
                assert mx < TOL and fro < TOL, (k, mx, fro)
        opt.step()
        losses.append(loss.item())
    assert np.allclose(losses, g["losses"], rtol=2e-5, atol=0)
    sd = model.state_dict()
    for k in sd:
        if k == "item_content_embedding":
            continue
        mx, fro = rel_err(sd[k].cpu().numpy(), g["final/" + k])
        assert mx < 1e-3 and fro < 1e-5, (k, mx, fro)
    assert list(sd.keys()) == [k[5:] for k in g if k.startswith("init/")]


# ---------------------------------------------------------------------------- native engine
@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_engine_training_matches_reference(golden, dev, case, use_graph):
    g = golden(case)
    model = _model(g, case, dev)
    csr = _graph(g, dev)
    eng = model.engine(csr, lr=float(g["lr"]), weight_decay=float(g["lam"]), batch_size=int(g["bs"]))
    losses = []
    for s in range(len(g["losses"])):
        # host (pinned) batches, as the reference's DataLoader delivers them (main.py:462-464,492)
        u = torch.from_numpy(g["batch_users"][s]).pin_memory()
        p = torch.from_numpy(g["batch_pos"][s]).pin_memory()
        n = torch.from_numpy(g["batch_neg"][s]).pin_memory()
        losses.append(float(eng.bpr_step(u, p, n, use_graph=use_graph).item()))
        if s == 0:
            sd = model.state_dict()
            for k in sd:
                if k == "item_content_embedding":
                    continue
                # the first Adam step is lr*g/(|g|+eps): elements with |g| ~ eps amplify fp32
                # rounding differences, so the element-wise bound is looser than the norm
                mx, fro = rel_err(sd[k].cpu().numpy(), g["step1/" + k])
                assert mx < 1e-3 and fro < TOL, (k, mx, fro)
    assert np.allclose(losses, g["losses"], rtol=2e-5, atol=0), (losses, g["losses"])
    sd = model.state_dict()
    for k in sd:
        if k == "item_content_embedding":
            continue
        mx, fro = rel_err(sd[k].cpu().numpy(), g["final/" + k])
        assert mx < 1e-3 and fro < 1e-5, (k, mx, fro)
    # the scatter buffers are clean again after every step
    assert not eng.G1.any() and not eng.G2.any()


@pytest.mark.parametrize("use_graph", [False, True])
def test_engine_brand_bpr_term_matches_reference(golden, dev, use_graph):
    """engine with the brand / author BPR term (reference main.py:382-391,401) on the
    tripartite graph against the run recorded from the reference's own bpr_loss_reg."""
    case = "tiny_brandloss_d64_k3"
    g = golden(case)
    model = _model(g, case, dev)
    csr = _graph(g, dev)
    eng = model.engine(csr, lr=float(g["lr"]), weight_decay=float(g["lam"]), batch_size=int(g["bs"]),
                       item_to_brand=g["item_to_brand"], brand_loss_weight=float(g["brand_loss_weight"]))
    losses = []
    for s in range(len(g["losses"])):
        u, p, n = (torch.from_numpy(g[k][s]) for k in ("batch_users", "batch_pos", "batch_neg"))
        losses.append(float(eng.bpr_step(u, p, n, use_graph=use_graph).item()))
        if s == 0:
            sd = model.state_dict()
            for k in sd:
                mx, fro = rel_err(sd[k].cpu().numpy(), g["step1/" + k])
                assert mx < 1e-3 and fro < TOL, (k, mx, fro)
    assert np.allclose(losses, g["losses"], rtol=2e-5, atol=0), (losses, g["losses"])
    sd = model.state_dict()
    for k in sd:
        mx, fro = rel_err(sd[k].cpu().numpy(), g["final/" + k])
        assert mx < 1e-3 and fro < 1e-5, (k, mx, fro)
    assert not eng.G1.any() and not eng.G2.any() and not eng.rowflag.any()


def test_engine_tail_batch_keeps_the_captured_graph_and_bad_indices_raise(golden, dev):
    """A shorter last batch runs eagerly on views of the staging buffers (no re-capture, the
    reference's DataLoader delivers one per epoch, main.py:462-464); out-of-range batch indices
    raise IndexError like the reference's gathers (main.py:496-497)."""
    case = "tiny_lightgcn_d64_k3"
    g = golden(case)
    bs = int(g["bs"])
    out = []
    for graph in (False, True):
        model = _model(g, case, dev)
        eng = model.engine(_graph(g, dev), lr=float(g["lr"]), weight_decay=float(g["lam"]), batch_size=bs)
        u, p, n = (torch.from_numpy(g[k][0]) for k in ("batch_users", "batch_pos", "batch_neg"))
        eng.bpr_step(u, p, n, use_graph=graph)
        captured = eng._graph
        l2 = eng.bpr_step(u[:100], p[:100], n[:100], use_graph=graph).item()
        assert eng._graph is captured                    # the captured step survives the tail batch
        assert eng.bs == (bs if graph else 100)          # eager mode simply re-sizes the staging
        l3 = eng.bpr_step(u, p, n, use_graph=graph).item()
        out.append((l2, l3, eng.P.clone()))
        assert not eng.G1.any() and not eng.G2.any()
    assert out[0][0] == pytest.approx(out[1][0], rel=1e-6) and out[0][1] == pytest.approx(out[1][1], rel=1e-6)
    assert rel_err(out[0][2].cpu().numpy(), out[1][2].cpu().numpy())[1] < 1e-6
    bad = torch.from_numpy(g["batch_users"][0]).clone()
    bad[5] = int(g["num_users"])
    model = _model(g, case, dev)
    eng = model.engine(_graph(g, dev), batch_size=bs)
    with pytest.raises(IndexError):
        eng.bpr_step(bad, torch.from_numpy(g["batch_pos"][0]), torch.from_numpy(g["batch_neg"][0]), use_graph=False)


@pytest.mark.parametrize("case", ["tiny_lightgcn_d64_k3", "tiny_fusion_d64_k3"])
def test_checkpoint_round_trip_in_the_reference_format(golden, dev, case, tmp_path):
    """SURVEY 8f-4 on the GPU: train with the engine, ``torch.save(engine.state_dict())`` (what
    reference main.py:550 writes), reload it the way ``main.test`` does (main.py:571:
    ``model.load_state_dict(torch.load(path))``) into a fresh drop-in module AND into a fresh
    engine: same keys / order as the reference, parameters equal to the reference's final ones,
    and evaluation from the reloaded module reproduces ``main.evaluate``'s metrics."""
    from gcn_recommendation_b200 import ops
    from gcn_recommendation_b200.engine import build_mask_csr
    orc = _orc()
    g = golden(case)
    model = _model(g, case, dev)
    csr = _graph(g, dev)
    eng = model.engine(csr, lr=float(g["lr"]), weight_decay=float(g["lam"]), batch_size=int(g["bs"]))
    for s in range(len(g["losses"])):
        eng.bpr_step(*(torch.from_numpy(g[k][s]) for k in ("batch_users", "batch_pos", "batch_neg")), use_graph=False)
    path = str(tmp_path / "best_model.pth")
    torch.save(eng.state_dict(), path)
    sd = torch.load(path)
    assert list(sd.keys()) == [k[5:] for k in g if k.startswith("init/")]
    fresh = _model(g, case, dev)                       # golden init again
    fresh.load_state_dict(sd)                          # reference main.py:571
    for k, v in fresh.state_dict().items():
        if k == "item_content_embedding":
            assert torch.equal(v.cpu(), torch.from_numpy(g["init/" + k]))
            continue
        assert torch.equal(v.cpu(), sd[k])
        mx, fro = rel_err(v.cpu().numpy(), g["final/" + k])
        assert mx < 1e-3 and fro < 1e-5, (k, mx, fro)
    eng2 = _model(g, case, dev).engine(csr, batch_size=int(g["bs"]))
    eng2.load_state_dict(sd)
    assert torch.equal(eng2.P, eng.P)
    if "fusion" in case:
        assert torch.equal(eng2.fusion["W"], eng.fusion["W"]) and torch.equal(eng2.fusion["b"], eng.fusion["b"])
    users, targets = orc.eval_pairs(g["val_user"], g["val_item"])
    mr, mc = build_mask_csr(users, g["train_user"], g["train_item"], int(g["num_users"]), dev)
    rec, ndcg, _ = eng2.evaluate(_t(users, dev, torch.int64), _t(targets, dev, torch.int64), mr, mc, 20)
    rec1, ndcg1, _ = eng.evaluate(_t(users, dev, torch.int64), _t(targets, dev, torch.int64), mr, mc, 20)
    assert (rec, ndcg) == (rec1, ndcg1)                # the reloaded engine IS the trained one
    # parameters trained here differ from the reference's in the last bits: a fp32 near-tie at
    # rank 20 may move one user's hit, nothing more
    assert abs(rec - float(g["eval/recall"])) <= 1.0 / len(users) + 1e-12
    assert abs(ndcg - float(g["eval/ndcg"])) <= 1.0 / len(users) + 1e-12


@pytest.mark.parametrize("fusion", [False, True])
def test_engine_sparse_backward_hops_are_bit_identical(golden, dev, fusion):
    """The sparse-gradient shortcuts (first Horner hop writes only its non-zero rows, second hop
    gathers under those flags) skip exact zeros only: parameters and Adam moments after several
    steps are bit-identical to the dense hops (LightGCN; the fusion variant within 1e-6)."""
    case = "tiny_fusion_d64_k3" if fusion else "tiny_lightgcn_d128_k4"
    g = golden(case)
    csr = _graph(g, dev)
    out = []
    for sparse in (False, True):
        model = _model(g, case, dev)
        eng = model.engine(csr, lr=float(g["lr"]), weight_decay=float(g["lam"]), batch_size=int(g["bs"]))
        eng.sparse_hops = sparse
        for s in range(4):
            # distinct rows per batch: the scatter's atomics then have one addend per element, so
            # the two runs can be compared bit for bit
            u = (torch.arange(48, device=dev) + 7 * s) % int(g["num_users"])
            p = (torch.arange(48, device=dev) + 11 * s) % 90
            n = 100 + (torch.arange(48, device=dev) + 13 * s) % 90
            eng.bpr_step(u, p, n, use_graph=False)
        assert bool(eng.rowflag2.any()) == sparse
        out.append((eng.P.clone(), eng.m.clone(), eng.v.clone(), eng.loss.clone()))
    for a, b in zip(*out):
        if fusion:      # dW/db are reduced with atomics: run-to-run rounding noise, not bit-stable
            assert rel_err(a.cpu().numpy(), b.cpu().numpy())[1] < 1e-6
        else:
            assert torch.equal(a, b)


def test_bpr_fused_vs_oracle_with_duplicates(dev):
    from gcn_recommendation_b200 import ops
    orc = _orc()
    rng = np.random.default_rng(5)
    U, I, d, bs = 50, 80, 64, 512               # tiny id space -> many duplicate rows per batch
    F = rng.standard_normal((U + I + 1, d), dtype=np.float32) * 0.3
    P = rng.standard_normal((U + I + 1, d), dtype=np.float32) * 0.3
    u = rng.integers(0, U, bs)
    p = rng.integers(0, I, bs)
    n = rng.integers(0, I, bs)
    loss, gF, gU, gI = orc.bpr_loss(F, P[:U], P[U:U + I], u, p, n, U, 1e-4)
    gFd = torch.zeros((U + I + 1, d), device=dev)
    gPd = torch.zeros((U + I + 1, d), device=dev)
    out = ops.bpr_fused(_t(F, dev), _t(P, dev), _t(u, dev, torch.int64), _t(p, dev, torch.int64),
                        _t(n, dev, torch.int64), U, 1e-4, grad_scale=0.25, gF=gFd, gP=gPd)
    assert abs(out.item() - loss) <= TOL * abs(loss)
    mx, fro = rel_err(gFd.cpu().numpy(), 0.25 * gF)
    assert mx < TOL and fro < TOL
    ref_gp = np.zeros_like(P)
    ref_gp[:U] = gU
    ref_gp[U:U + I] = gI
    mx, fro = rel_err(gPd.cpu().numpy(), ref_gp)
    assert mx < TOL and fro < TOL
    ops.zero_rows(gFd, gPd, _t(u, dev, torch.int64), _t(p, dev, torch.int64), _t(n, dev, torch.int64), U)
    assert not gFd.any() and not gPd.any()


def test_adam_vs_oracle(dev):
    from gcn_recommendation_b200 import ops
    orc = _orc()
    rng = np.random.default_rng(9)
    n = 4099                                     # exercises the non-multiple-of-4 tail
    p = rng.standard_normal(n, dtype=np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    pd, md, vd = _t(p, dev), _t(m, dev), _t(v, dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    sc = torch.zeros(2, device=dev)
    for t in range(1, 6):
        gr = rng.standard_normal(n, dtype=np.float32) * 1e-2
        orc.adam_step(p, gr, m, v, t)
        ops.adam_tick(step, sc, 1e-3)
        ops.adam(pd, _t(gr, dev), md, vd, sc)
    assert int(step.item()) == 5
    for a, b in ((pd, p), (md, m), (vd, v)):
        mx, fro = rel_err(a.cpu().numpy(), b)
        assert mx < TOL and fro < 1e-6


@pytest.mark.parametrize("d", [64, 128])
def test_fusion_projection_vs_oracle(dev, d):
    from gcn_recommendation_b200 import ops
    orc = _orc()
    rng = np.random.default_rng(d)
    n, c = 333, 768
    E = rng.standard_normal((n, d), dtype=np.float32) * 0.1
    C = rng.standard_normal((n, c), dtype=np.float32)
    W = rng.standard_normal((d, d + c), dtype=np.float32) * 0.05
    b = rng.standard_normal(d, dtype=np.float32) * 0.1
    gH = rng.standard_normal((n, d), dtype=np.float32)
    H, pre = orc.fusion_forward(E, C, W, b)
    gE, gW, gb = orc.fusion_backward(E, C, W, pre, gH)
    Hd = ops.fusion_proj_fwd(_t(E, dev), _t(C, dev), _t(W, dev), _t(b, dev))
    mx, fro = rel_err(Hd.cpu().numpy(), H)
    assert mx < TOL and fro < TOL
    gEd, gWd, gbd = ops.fusion_proj_bwd(_t(E, dev), _t(C, dev), _t(W, dev), Hd, _t(gH, dev))
    for a, r in ((gEd, gE), (gWd, gW), (gbd, gb)):
        mx, fro = rel_err(a.cpu().numpy(), r)
        assert mx < TOL and fro < TOL


# ---------------------------------------------------------------------------- a7 eval
@pytest.mark.parametrize("case", CASES)
def test_topk_ids_and_metrics_vs_reference(golden, dev, case):
    from gcn_recommendation_b200 import ops
    from gcn_recommendation_b200.engine import build_mask_csr
    orc = _orc()
    g = golden(case)
    users = g["eval/users"]
    targets = dict(zip(g["val_user"].tolist(), g["val_item"].tolist()))
    tg = np.asarray([targets[u] for u in users.tolist()], np.int64)
    mr, mc = build_mask_csr(users, g["train_user"], g["train_item"], int(g["num_users"]), dev)
    omr, omc = orc.mask_csr(users, g["train_user"], g["train_item"], int(g["num_users"]))
    assert np.array_equal(mr.cpu().numpy(), omr) and np.array_equal(mc.cpu().numpy(), omc)
    ids, sc = ops.score_topk(_t(g["eval/F_user"], dev), _t(g["eval/F_item"], dev),
                             _t(users, dev, torch.int64), mr, mc, 20)
    oids, osc = orc.score_topk(g["eval/F_user"], g["eval/F_item"], users, omr, omc, 20)
    assert np.array_equal(ids.cpu().numpy(), oids), "top-k ids must be bit-exact vs the oracle"
    assert np.array_equal(_bits(sc.cpu().numpy()), _bits(osc))
    # against the reference's own torch.topk: identical except inside fp32 near-ties
    ref_ids = g["eval/topk_ids"]
    diff_rows = np.where((ids.cpu().numpy() != ref_ids).any(1))[0]
    for r in diff_rows:
        gap = np.abs(np.diff(g["eval/topk_scores"][r]))
        assert gap.min() <= 4e-7 * np.abs(g["eval/topk_scores"][r]).max()
    sums = ops.eval_metrics(ids, _t(tg, dev, torch.int64)).cpu().numpy()
    assert sums[0] / len(users) == pytest.approx(float(g["eval/recall"]), abs=1e-12)
    assert sums[1] / len(users) == pytest.approx(float(g["eval/ndcg"]), abs=1e-12)


def test_topk_without_mask_and_ragged_user_tile(dev):
    from gcn_recommendation_b200 import ops
    orc = _orc()
    rng = np.random.default_rng(11)
    Fu = rng.standard_normal((97, 128), dtype=np.float32)
    Fi = rng.standard_normal((1111, 128), dtype=np.float32)
    Fi[500] = Fi[20]                             # exact tie: lower id must come first
    users = rng.permutation(97)[:71]
    ids, sc = ops.score_topk(_t(Fu, dev), _t(Fi, dev), _t(users, dev, torch.int64), None, None, 20)
    oids, osc = orc.score_topk(Fu, Fi, users, None, None, 20)
    assert np.array_equal(ids.cpu().numpy(), oids)
    assert np.array_equal(_bits(sc.cpu().numpy()), _bits(osc))


# ---------------------------------------------------------------------------- properties
def test_gowalla_shape_properties(dev):
    """Size-independent checks at the full Gowalla shape (configs[1]): the normalised adjacency is
    symmetric (<x, A y> == <A x, y>), propagation is linear, and a sampled set of rows matches
    the oracle bit-exactly."""
    from gcn_recommendation_b200 import ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    orc = _orc()
    inter = synth.generate("gowalla", seed=0)
    tu, ti, _, _ = inter.split_validation()
    U, I, B = inter.num_users, inter.num_items, inter.num_brands
    csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
    N, d = U + I + B, 64
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((N, d), device=dev, generator=gen)
    y = torch.randn((N, d), device=dev, generator=gen)
    ax, ay = ops.spmm(csr, x), ops.spmm(csr, y)
    lhs = (x.double() * ay.double()).sum().item()
    rhs = (ax.double() * y.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0)
    f1 = ops.propagate(csr, x, 3)
    f2 = ops.propagate(csr, y, 3)
    f12 = ops.propagate(csr, (x + 2 * y).contiguous(), 3)
    mx, fro = rel_err((f1 + 2 * f2).cpu().numpy(), f12.cpu().numpy())
    assert fro < TOL
    a = orc.build_norm_adj(tu, ti, U, I, B)
    assert np.array_equal(csr.col.cpu().numpy(), a["col"])
    assert np.array_equal(_bits(csr.val.cpu().numpy()), _bits(a["val"]))
    ref = orc.spmm(a["rowptr"], a["col"], a["val"], x.cpu().numpy())
    got = ax.cpu().numpy()
    short = np.diff(a["rowptr"]) <= csr.long_row_threshold
    assert np.array_equal(_bits(got[short]), _bits(ref[short]))
    mx, fro = rel_err(got, ref)
    assert mx < TOL and fro < TOL


def test_no_cpu_fallback(dev):
    from gcn_recommendation_b200 import _lib, ops
    with pytest.raises(_lib.LgcnError):
        ops.adam(torch.zeros(8), torch.zeros(8), torch.zeros(8), torch.zeros(8), torch.zeros(2))


# ---------------------------------------------------------------------------- multi-GPU
def test_feature_sharded_step_emulated_on_one_gpu(golden, dev):
    """Feature (column) sharding: P engines over d/P columns each, the dot-product all-reduce
    emulated by summing the ranks' partials in-process (no spin-waiting kernels, one GPU).  The
    concatenated result must track the unsharded engine and the reference's loss curve."""
    from gcn_recommendation_b200.dist import FeatureShardedEngine, column_shard
    g = golden("tiny_lightgcn_d64_k3")
    U, I, B, K = int(g["num_users"]), int(g["num_items"]), int(g["num_brands"]), int(g["K"])
    csr = _graph(g, dev)
    full = torch.cat([_t(g["init/user_embedding.weight"], dev), _t(g["init/item_embedding.weight"], dev),
                      _t(g["init/brand_embedding.weight"], dev)]).contiguous()
    world = 2
    pending = []

    def make_allreduce(r):
        def ar(t):                      # rank 0 stashes, rank 1 sums and shares (sequential emulation)
            pending.append(t)
            if len(pending) == world:
                tot = sum(pending)
                for x in pending:
                    x.copy_(tot)
                pending.clear()
        return ar

    engs = [FeatureShardedEngine(csr, U, I, B, K, column_shard(full, r, world), world=world, rank=r,
                                 allreduce=make_allreduce(r), lr=float(g["lr"]),
                                 weight_decay=float(g["lam"]), batch_size=int(g["bs"])) for r in range(world)]

    # the two "ranks" must interleave around the all-reduce: run phase by phase
    from gcn_recommendation_b200 import ops
    losses = []
    for s in range(len(g["losses"])):
        for e in engs:
            e.b_users.copy_(_t(g["batch_users"][s], dev, torch.int64))
            e.b_pos.copy_(_t(g["batch_pos"][s], dev, torch.int64))
            e.b_neg.copy_(_t(g["batch_neg"][s], dev, torch.int64))
        Fs = [e.propagate() for e in engs]
        for e, F in zip(engs, Fs):
            ops.bpr_partial(F, e.P, e.b_users, e.b_pos, e.b_neg, U, e.dots)
        tot = engs[0].dots + engs[1].dots
        for e, F in zip(engs, Fs):
            e.dots.copy_(tot)
            ops.bpr_apply(F, e.P, e.b_users, e.b_pos, e.b_neg, U, e.lam, e.dots, grad_scale=1.0 / (K + 1),
                          gF=e.G1, gP=e.G2, gp_includes_gf=True, sample_ws=e.sample_ws, loss_out=e.loss)
            ops.adam_tick(e.step_dev, e.adam_scalars, e.lr, e.betas)
            acc = e.G1
            for k in range(K - 1):
                acc = ops.spmm(e.g, acc, out=e.work[k % 2], addend=e.G1)
            ops.spmm_adam(e.g, acc, e.P, e.m, e.v, e.adam_scalars, addend=e.G2, betas=e.betas, eps=e.eps)
            ops.zero_rows(e.G1, e.G2, e.b_users, e.b_pos, e.b_neg, U)
        assert engs[0].loss.item() == engs[1].loss.item()
        losses.append(engs[0].loss.item())
    assert np.allclose(losses, g["losses"], rtol=2e-5, atol=0)
    P = torch.cat([e.P for e in engs], dim=1).cpu().numpy()
    ref = np.concatenate([g["final/user_embedding.weight"], g["final/item_embedding.weight"],
                          g["final/brand_embedding.weight"]], 0)
    mx, fro = rel_err(P, ref)
    assert mx < 1e-3 and fro < 1e-5


def test_feature_sharded_fusion_step_emulated_on_one_gpu(golden, dev):
    """LightGCN_Fusion on the feature-sharded engine: two "ranks" run as threads on one GPU, the
    all-to-all (column shards <-> item-block rows) and the all-reduces emulated in-process.  The
    item block is projected item-sharded; the result must track the reference's loss curve and
    final parameters, and W / b must stay bit-identical on both ranks."""
    import threading
    from gcn_recommendation_b200.dist import FeatureShardedEngine, column_shard
    g = golden("tiny_fusion_d64_k3")
    U, I, B, K = int(g["num_users"]), int(g["num_items"]), int(g["num_brands"]), int(g["K"])
    csr = _graph(g, dev)
    full = torch.cat([_t(g["init/user_embedding.weight"], dev), _t(g["init/item_id_embedding.weight"], dev),
                      _t(g["init/brand_embedding.weight"], dev)]).contiguous()
    C = _t(g["init/item_content_embedding"], dev)
    world = 2
    bar = threading.Barrier(world)
    slots = [None] * world

    def make_allreduce(r):
        def ar(t):
            slots[r] = t
            bar.wait()
            tot = slots[0] + slots[1]            # same order on both ranks -> bit-identical
            bar.wait()
            t.copy_(tot)
            bar.wait()
        return ar

    def make_alltoall(r):
        def a2a(out, inp):
            slots[r] = inp
            bar.wait()
            for s in range(world):
                out[s].copy_(slots[s][r])
            bar.wait()
        return a2a

    ipr = -(-I // world)
    engs = []
    for r in range(world):
        fus = dict(content=C[r * ipr:min(I, (r + 1) * ipr)].contiguous(),
                   weight=_t(g["init/item_fusion_layer.weight"], dev),
                   bias=_t(g["init/item_fusion_layer.bias"], dev), alltoall=make_alltoall(r))
        # one graph object per rank, as in one process per GPU: a NormAdjCSR owns the long-row
        # segment workspace of its lgcn_spmm calls, which two host threads must not share
        engs.append(FeatureShardedEngine(csr if r == 0 else _graph(g, dev), U, I, B, K,
                                         column_shard(full, r, world), world=world, rank=r,
                                         allreduce=make_allreduce(r), fusion=fus, lr=float(g["lr"]),
                                         weight_decay=float(g["lam"]), batch_size=int(g["bs"])))
    losses = [[], []]
    errs = []

    def run(r):
        try:
            torch.cuda.set_device(dev)
            for s in range(len(g["losses"])):
                loss = engs[r].bpr_step(_t(g["batch_users"][s], dev, torch.int64),
                                        _t(g["batch_pos"][s], dev, torch.int64),
                                        _t(g["batch_neg"][s], dev, torch.int64), use_graph=False)
                losses[r].append(loss.item())
        except Exception as e:  # pragma: no cover
            errs.append(repr(e))
            bar.abort()

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=120)
    assert not errs, errs
    assert losses[0] == losses[1]
    assert np.allclose(losses[0], g["losses"], rtol=2e-5, atol=0)
    P = torch.cat([e.P for e in engs], dim=1).cpu().numpy()
    ref = np.concatenate([g["final/user_embedding.weight"], g["final/item_id_embedding.weight"],
                          g["final/brand_embedding.weight"]], 0)
    mx, fro = rel_err(P, ref)
    assert mx < 1e-3 and fro < 1e-5
    assert torch.equal(engs[0].fusion["W"], engs[1].fusion["W"])
    assert torch.equal(engs[0].fusion["b"], engs[1].fusion["b"])
    mx, fro = rel_err(engs[0].fusion["W"].cpu().numpy(), g["final/item_fusion_layer.weight"])
    assert mx < 1e-3 and fro < 1e-5


# ---------------------------------------------------------------------------- f1 sampler
def test_device_sampler_epoch_is_a_permutation_with_valid_negatives(dev):
    from gcn_recommendation_b200 import ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    inter = synth.generate("small", seed=4)
    tu, ti, _, _ = inter.split_validation()
    U, I, B = inter.num_users, inter.num_items, inter.num_brands
    csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
    E = len(tu)
    assert int(csr.rowptr[U].item()) == E
    bs = 1000
    state = torch.zeros(2, dtype=torch.int64, device=dev)
    u = torch.empty(bs, dtype=torch.int64, device=dev)
    p = torch.empty_like(u)
    n = torch.empty_like(u)
    pairs, negs = [], []
    steps = -(-E // bs)
    for _ in range(steps):
        ops.sample_bpr(csr, U, I, E, 7, state, u, p, n)
        pairs.append((u * I + p).cpu().numpy())
        negs.append((u.cpu().numpy(), n.cpu().numpy()))
    keys = np.concatenate(pairs)
    truth = np.sort(tu * I + ti)
    # the first E samples are one epoch: every training interaction exactly once
    assert np.array_equal(np.sort(keys[:E]), truth)
    assert state.cpu().tolist() == [1, steps * bs - E]
    posset = set(truth.tolist())
    allu = np.concatenate([a for a, _ in negs])
    alln = np.concatenate([b for _, b in negs])
    assert alln.min() >= 0 and alln.max() < I
    assert not any((int(a) * I + int(b)) in posset for a, b in zip(allu, alln))
    # negatives are spread over the catalogue; the order inside an epoch is not the CSR order
    assert len(np.unique(alln)) > 0.9 * I
    assert not np.array_equal(keys[:E], truth)


def test_engine_trains_from_device_sampler(golden, dev):
    g = golden("tiny_lightgcn_d64_k3")
    model = _model(g, "tiny_lightgcn_d64_k3", dev)
    eng = model.engine(_graph(g, dev), batch_size=256)
    first = eng.train_steps(1).item()
    last = eng.train_steps(60).item()
    assert abs(first - 0.6931) < 5e-3 and last < first - 1e-3       # ln 2 at init, then it learns
