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
