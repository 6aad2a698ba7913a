"""GPU tests of the tensor-core (tcgen05) rating path: ids must equal the exact fp32 kernel /
CPU oracle bit for bit (filter-and-refine with a per-user exactness certificate)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _case(dev, nu, n_items, d, seed, scale=1.0, with_mask=True):
    rng = np.random.default_rng(seed)
    Fu = (rng.standard_normal((nu + 7, d), dtype=np.float32) * scale)
    Fi = (rng.standard_normal((n_items, d), dtype=np.float32) * scale)
    # item popularity-like norm spread so that scores are not all on one scale
    Fi *= (0.2 + rng.random((n_items, 1), dtype=np.float32) * 2.0)
    users = rng.permutation(nu + 7)[:nu].astype(np.int64)
    mr = mc = None
    if with_mask:
        cnt = rng.integers(0, 40, nu)
        rp = np.zeros(nu + 1, np.int64)
        np.cumsum(cnt, out=rp[1:])
        cols = np.concatenate([np.sort(rng.choice(n_items, c, replace=False)) for c in cnt] or [np.zeros(0)])
        mr, mc = torch.from_numpy(rp).to(dev), torch.from_numpy(cols.astype(np.int32)).to(dev)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    return t(Fu), t(Fi), t(users), mr, mc, (Fu, Fi, users)


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("n_items", [20000, 8192 + 77])
def test_tc_topk_matches_exact_kernel(dev, d, n_items):
    from gcn_recommendation_b200 import ops
    Fu, Fi, users, mr, mc, _ = _case(dev, 333, n_items, d, seed=d + n_items)
    ops.STATS["tc_users"] = ops.STATS["tc_fallback_users"] = 0
    ids, sc = ops.score_topk(Fu, Fi, users, mr, mc, 20, tensor_cores=True)
    eids, esc = ops.score_topk_exact(Fu, Fi, users, mr, mc, 20)
    assert torch.equal(ids, eids)
    assert torch.equal(sc.view(torch.int32), esc.view(torch.int32))
    assert ops.STATS["tc_users"] == 333
    # the certificate should hold for nearly every user on well-separated scores
    assert ops.STATS["tc_fallback_users"] <= 33, ops.STATS


def test_tc_topk_vs_oracle_and_masks_respected(dev):
    from gcn_recommendation_b200 import ops
    from oracle import lgcn_oracle as orc
    Fu, Fi, users, mr, mc, (hFu, hFi, hu) = _case(dev, 130, 9000, 128, seed=5)
    ids, sc = ops.score_topk(Fu, Fi, users, mr, mc, 20, tensor_cores=True)
    oids, osc = orc.score_topk(hFu, hFi, hu, mr.cpu().numpy(), mc.cpu().numpy(), 20)
    assert np.array_equal(ids.cpu().numpy(), oids)
    assert np.array_equal(sc.cpu().numpy().view(np.uint32), osc.view(np.uint32))
    rp, cols = mr.cpu().numpy(), mc.cpu().numpy()
    got = ids.cpu().numpy()
    for q in range(130):
        assert not set(got[q].tolist()) & set(cols[rp[q]:rp[q + 1]].tolist())


def test_tc_near_tie_scores_fall_back_to_exact(dev):
    """Tiny, nearly tied scores (random-init scale): whatever the certificate decides, the ids
    must still be the exact ones."""
    from gcn_recommendation_b200 import ops
    Fu, Fi, users, mr, mc, _ = _case(dev, 256, 16384, 64, seed=9, scale=1e-3)
    ids, sc = ops.score_topk(Fu, Fi, users, mr, mc, 20, tensor_cores=True)
    eids, esc = ops.score_topk_exact(Fu, Fi, users, mr, mc, 20)
    assert torch.equal(ids, eids)


def test_amazon_shape_properties(dev):
    """Size-independent checks at the FULL Amazon-Books-2023 shape (BASELINE.json configs[2]:
    14.7 M nodes, 59 M entries, d=128): the normalised adjacency is symmetric, the SpMM is
    deterministic and linear, the sparse-input hop with row flags equals the dense hop, and the
    tensor-core rating returns descending scores, no masked item and the exact kernel's ids."""
    from gcn_recommendation_b200 import ops, synth
    from gcn_recommendation_b200.engine import build_mask_csr
    from gcn_recommendation_b200.graph import NormAdjCSR
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~60 GB of free HBM")
    U, I, B, total, d, K = synth.SHAPES["amazon"]
    inter = synth.generate_device("amazon", dev, seed=0)
    tu, ti, vu, vi = synth.split_validation_device(inter)
    del inter
    csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
    N = U + I + B
    assert csr.nnz == 2 * tu.numel() and csr.n_long > 0
    gen = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn((N, d), device=dev, generator=gen)
    y = torch.randn((N, d), device=dev, generator=gen)
    ax = ops.spmm(csr, x)
    ay = ops.spmm(csr, y)
    lhs = (x.double() * ay.double()).sum().item()
    rhs = (ax.double() * y.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0)              # <x, A y> == <A x, y>
    assert torch.equal(ax, ops.spmm(csr, x))                                     # deterministic
    axy = ops.spmm(csr, (x + y).contiguous())
    err = (axy - (ax + ay)).norm().item() / axy.norm().item()
    assert err < 1e-5                                                            # linear
    del ay, axy
    # sparse-input hop: flagged gathers == dense gathers
    rows = torch.randint(0, N, (6144,), device=dev, generator=gen)
    xs = torch.zeros_like(x)
    xs[rows] = x[rows]
    flag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
    flag[rows] = 1
    zero_row = torch.zeros(256, device=dev)
    dense = ops.spmm(csr, xs, addend=xs)
    sparse = ops.spmm(csr, xs, addend=xs, x_rowflag=flag, zero_row=zero_row)
    assert torch.equal(dense, sparse)
    # ... and the sparse-OUTPUT chain of the backward pass (live-list kernel): hop 1 writes only
    # the rows it can make non-zero and reports them, hop 2 gathers under those flags; the
    # result equals two dense hops bit for bit, the unwritten rows are never read (NaN-poisoned)
    yflag = torch.zeros(N + 32, dtype=torch.uint8, device=dev)
    h1 = torch.full_like(x, float("nan"))
    ops.spmm(csr, xs, out=h1, addend=xs, x_rowflag=flag, addend_rowflag=flag, zero_row=zero_row,
             y_rowflag=yflag)
    live = yflag[:N].bool()
    assert 6000 < int(live.sum()) < N // 2
    assert torch.equal(h1[live], dense[live]) and not bool(dense[~live].any())
    assert bool(torch.isnan(h1[~live]).all())
    h2 = ops.spmm(csr, h1, addend=xs, x_rowflag=yflag, addend_rowflag=flag, zero_row=zero_row)
    assert torch.equal(h2, ops.spmm(csr, dense, addend=xs))
    del dense, sparse, xs, y, h1, h2
    # rating over the full 4.4 M-item catalogue
    Fu, Fi = ax[:U], ax[U:U + I]
    eu = vu[:256].contiguous()
    mr, mc = build_mask_csr(eu.cpu().numpy(), tu.cpu().numpy(), ti.cpu().numpy(), U, dev)
    ids, sc = ops.score_topk(Fu, Fi, eu, mr, mc, 20, tensor_cores=True)
    eids, esc = ops.score_topk_exact(Fu, Fi, eu, mr, mc, 20)
    assert torch.equal(ids, eids) and torch.equal(sc.view(torch.int32), esc.view(torch.int32))
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    rp, cols, got = mr.cpu().numpy(), mc.cpu().numpy(), ids.cpu().numpy()
    for q in range(256):
        assert not set(got[q].tolist()) & set(cols[rp[q]:rp[q + 1]].tolist())


def test_tc_batched_sweep_and_item_splits_match_exact_kernel(dev):
    """A rating sweep in user batches (one prepared table + one workspace for all batches, ragged
    last batch, per-batch views of the mask CSR) and the item-split runs of small batches (few
    user tiles: the catalogue is cut into splits, per-split lists merged and certified) return
    the exact kernel's ids and score bits."""
    from gcn_recommendation_b200 import _lib, ops
    Fu, Fi, users, mr, mc, _ = _case(dev, 1000, 40000, 128, seed=21)
    eids, esc = ops.score_topk_exact(Fu, Fi, users, mr, mc, 20)
    lib = _lib.load()
    assert lib.lgcn_score_tc_launches(256, 40000) == 3 and lib.lgcn_score_tc_launches(148 * 128, 40000) == 2
    rater = ops.TcRater(40000, 128, dev, 256)
    for bu in (256, 1000, 128):
        ids, sc = ops.score_topk(Fu, Fi, users, mr, mc, 20, tensor_cores=True,
                                 rater=rater if bu <= 256 else None, batch_users=bu)
        assert torch.equal(ids, eids), bu
        assert torch.equal(sc.view(torch.int32), esc.view(torch.int32)), bu
    assert rater.prepared_for is not None
    # no mask, k = 5
    ids, sc = ops.score_topk(Fu, Fi, users, None, None, 5, tensor_cores=True, batch_users=300)
    eids, esc = ops.score_topk_exact(Fu, Fi, users, None, None, 5)
    assert torch.equal(ids, eids) and torch.equal(sc.view(torch.int32), esc.view(torch.int32))
