"""Out-of-bounds audit of every kernel family through guard bands.

``compute-sanitizer`` is CLOSED on this GPU pool (gpurun answers: "compute-sanitizer is closed on
this pool and stays closed: runs under it have left GPUs needing a reset. Find a bad access with
bounds checks and asserts of your own, small cases, and a comparison with the CPU reference"), so
the memcheck VERDICT r01 asked for is replaced by this: every output AND input tensor of a call
lives inside a larger allocation with guard bands on both sides.

* output guards hold a bit pattern and must be intact after the call  -> no write outside;
* input guards hold NaN (floats) / a huge index (ints), so a read outside poisons the result or
  faults, and the results are compared with the CPU oracle / the exact kernel as usual;
* shapes are ragged on purpose (rows not a multiple of a CTA's rows, nnz not a multiple of a tile,
  users not a multiple of 128, items not a multiple of 128, n not a multiple of 4).
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
PAD = 8192          # guard elements on each side


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


class Guarded:
    """Tensors carved out of guarded allocations; ``check()`` verifies every guard band."""

    def __init__(self, dev):
        self.dev = dev
        self.bufs = []
        self.keep = []

    def _alloc(self, n, dtype, guard_value):
        buf = torch.empty(n + 2 * PAD, dtype=dtype, device=self.dev)
        buf[:PAD] = guard_value
        buf[PAD + n:] = guard_value
        return buf

    def out(self, shape, dtype=torch.float32, fill=None):
        """Output tensor: guards = a recognisable pattern, body optionally pre-filled."""
        n = int(np.prod(shape))
        pat = {torch.float32: -1.2345678e30, torch.float64: -1.2345678e300}.get(dtype, 0x5A)
        buf = self._alloc(n, dtype, pat)
        if fill is not None:
            buf[PAD:PAD + n] = fill
        self.bufs.append((buf, n, buf[:PAD].clone(), buf[PAD + n:].clone()))
        return buf[PAD:PAD + n].view(shape)

    def inp(self, array, dtype=None):
        """Input tensor copied into a buffer whose guards poison any read outside."""
        t = torch.as_tensor(np.ascontiguousarray(array))
        if dtype is not None:
            t = t.to(dtype)
        n = t.numel()
        poison = float("nan") if t.dtype.is_floating_point else (1 << 30)
        if t.dtype == torch.uint8:
            poison = 1
        buf = self._alloc(n, t.dtype, poison)
        buf[PAD:PAD + n] = t.flatten().to(self.dev)
        self.keep.append(buf)                      # raw data_ptr() users must not outlive the buffer
        return buf[PAD:PAD + n].view(t.shape)

    def check(self):
        torch.cuda.synchronize()
        for buf, n, lo, hi in self.bufs:
            assert torch.equal(buf[:PAD].view(torch.uint8), lo.view(torch.uint8)), "write below an output"
            assert torch.equal(buf[PAD + n:].view(torch.uint8), hi.view(torch.uint8)), "write past an output"


def _orc():
    from oracle import lgcn_oracle
    return lgcn_oracle


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _guarded_graph(G, csr):
    """Re-home the kernel-facing arrays of a NormAdjCSR into guarded allocations."""
    csr.rowptr_flagged = G.inp(csr.rowptr_flagged.cpu().numpy())
    cv = csr.colval.cpu().numpy()
    csr.colval = G.inp(cv)
    if csr.n_long > 0:
        csr.long_row_ids = G.inp(csr.long_row_ids.cpu().numpy())
        csr.long_rowptr = G.inp(csr.long_rowptr.cpu().numpy())
        csr.long_colval = G.inp(csr.long_colval.cpu().numpy())
        csr.long_seg_ptr = G.inp(csr.long_seg_ptr.cpu().numpy())
        if csr.long_done is not None:                   # ABI v7 work plan: counters (written), orders (read)
            csr.long_done = G.out((csr.n_long,), torch.int32, fill=0)
    for rows in (4, 8):
        for windowed in (False, True):
            order = csr.chunk_order_for(rows, windowed)
            if order is not None:
                csr._chunk_orders[(rows, windowed)] = G.inp(order.cpu().numpy())
    return csr


@pytest.mark.parametrize("d", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("path", ["small", "ring", "chunk"])
def test_spmm_stays_inside_its_buffers(dev, d, path):
    from gcn_recommendation_b200 import _lib, ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    orc = _orc()
    U, I, B = 1013, 1511, 3                          # N = 2527: ragged against every chunk size
    inter = synth.generate((U, I, B, 26_000), seed=7 + d)
    tu, ti, _, _ = inter.split_validation()
    a = orc.build_norm_adj(tu, ti, U, I, B)
    N = U + I + B
    rng = np.random.default_rng(d)
    X = rng.standard_normal((N, d), dtype=np.float32)
    A1 = rng.standard_normal((N, d), dtype=np.float32)
    ref = orc.spmm(a["rowptr"], a["col"], a["val"], X)
    old, old_large = ops.SPMM_FLAGS_EXTRA, ops.CHUNK_ORDER_LARGE
    ops.SPMM_FLAGS_EXTRA = {"small": 0, "ring": _lib.SPMM_F_BIG_PATH | _lib.SPMM_F_FORCE_RING,
                            "chunk": _lib.SPMM_F_BIG_PATH | _lib.SPMM_F_NO_RING}[path]
    ops.CHUNK_ORDER_LARGE = "all"                    # the ring / live kernels follow a (guarded) order too
    try:
        G = Guarded(dev)
        csr = _guarded_graph(G, NormAdjCSR.from_interactions(tu, ti, U, I, B, dev, long_row_threshold=48, seg_len=32))
        assert csr.n_long > 0
        csr._seg_ws[d] = G.out((csr.n_seg, d))
        short = np.diff(a["rowptr"]) <= 48
        x = G.inp(X)
        y = ops.spmm(csr, x, out=G.out((N, d)))
        G.check()
        assert np.array_equal(_bits(y.cpu().numpy()[short]), _bits(ref[short]))
        assert rel_err(y.cpu().numpy(), ref)[0] < 1e-5
        ya = ops.spmm(csr, x, out=G.out((N, d)), addend=G.inp(A1))
        ym = ops.spmm(csr, x, out=G.out((N, d)), mean_layers=[G.inp(A1), x])
        G.check()
        assert rel_err(ya.cpu().numpy(), A1 + ref)[0] < 1e-5
        assert rel_err(ym.cpu().numpy(), ((A1 + X) + ref) / np.float32(3))[0] < 1e-5
        # sparse hops: flagged input / addend, sparse output
        nz = rng.choice(N, 97, replace=False)
        Xs = np.zeros((N, d), np.float32)
        Xs[nz] = X[nz]
        flag = np.zeros(N + 32, np.uint8)
        flag[nz] = 1
        Xp = Xs.copy()
        Xp[flag[:N] == 0] = np.nan
        ft = G.inp(flag)
        zr = G.inp(np.zeros(256, np.float32))
        yflag = G.out((N,), torch.uint8, fill=7)
        h1 = ops.spmm(csr, G.inp(Xp), out=G.out((N, d), fill=float("nan")), addend=G.inp(Xs), x_rowflag=ft,
                      addend_rowflag=ft, zero_row=zr, y_rowflag=yflag)
        G.check()
        want = Xs + orc.spmm(a["rowptr"], a["col"], a["val"], Xs)
        live = yflag.cpu().numpy() == 1
        assert rel_err(h1.cpu().numpy()[live], want[live])[0] < 1e-5 and not want[~live].any()
        # Adam epilogue
        p, m, v = (G.out((N, d), fill=0.0) for _ in range(3))
        p.copy_(torch.from_numpy(A1))
        v.fill_(1e-4)
        sc = G.inp(np.asarray([1e-3, 1.0], np.float32))
        gout = G.out((N, d))
        ops.spmm_adam(csr, x, p, m, v, sc, addend=G.inp(A1), g_out=gout)
        G.check()
        assert rel_err(gout.cpu().numpy(), ref + A1)[0] < 1e-5
        assert torch.isfinite(p).all() and torch.isfinite(m).all() and torch.isfinite(v).all()
        if csr.long_done is not None:
            assert int(csr.long_done.abs().sum()) == 0, "long-row counters must be re-armed"
    finally:
        ops.SPMM_FLAGS_EXTRA, ops.CHUNK_ORDER_LARGE = old, old_large


@pytest.mark.parametrize("d", [16, 64, 128])
def test_bpr_adam_sampler_graph_kernels_stay_inside(dev, d):
    from gcn_recommendation_b200 import _lib, ops, synth
    from gcn_recommendation_b200.graph import NormAdjCSR
    orc = _orc()
    G = Guarded(dev)
    rng = np.random.default_rng(3 + d)
    U, I, B, bs = 211, 307, 2, 333
    N = U + I + B
    F = rng.standard_normal((N, d), dtype=np.float32) * 0.3
    P = rng.standard_normal((N, d), dtype=np.float32) * 0.3
    u, p, n = rng.integers(0, U, bs), rng.integers(0, I, bs), rng.integers(0, I, bs)
    gF, gP = G.out((N, d), fill=0.0), G.out((N, d), fill=0.0)
    rowflag = G.out((N,), torch.uint8, fill=0)
    ws, loss = G.out((2 * bs,)), G.out((1,))
    ut, pt, nt = G.inp(u), G.inp(p), G.inp(n)
    ops.bpr_fused(G.inp(F), G.inp(P), ut, pt, nt, U, 1e-4, grad_scale=0.25, gF=gF, gP=gP, sample_ws=ws,
                  loss_out=loss, rowflag=rowflag)
    G.check()
    lref, gFr, gU, gI = orc.bpr_loss(F, P[:U], P[U:U + I], u, p, n, U, 1e-4)
    assert abs(loss.item() - lref) <= 1e-5 * abs(lref)
    assert rel_err(gF.cpu().numpy(), 0.25 * gFr)[0] < 1e-5
    # feature-sharded halves of the same step
    dots = G.out((3 * bs,))
    ops.bpr_partial(G.inp(F), G.inp(P), ut, pt, nt, U, dots)
    ops.bpr_apply(G.inp(F), G.inp(P), ut, pt, nt, U, 1e-4, dots, gF=G.out((N, d), fill=0.0),
                  gP=G.out((N, d), fill=0.0), sample_ws=G.out((2 * bs,)), loss_out=G.out((1,)))
    ops.zero_rows(gF, gP, ut, pt, nt, U, rowflag=rowflag)
    G.check()
    assert not gF.any() and not gP.any() and not rowflag.any()
    status = G.out((1,), torch.int32, fill=0)
    ops.check_indices(status, (ut, 0, U), (pt, 0, I), (nt, 0, 5))
    G.check()
    assert int(status.item()) == int((n >= 5).sum())
    # standalone Adam, n not a multiple of 4
    nn = 4099
    pp, mm, vv = G.out((nn,), fill=1.0), G.out((nn,), fill=0.0), G.out((nn,), fill=0.0)
    step, sc = G.out((1,), torch.int64, fill=0), G.out((2,), fill=0.0)
    ops.adam_tick(step, sc, 1e-3)
    ops.adam(pp, G.inp(rng.standard_normal(nn, dtype=np.float32)), mm, vv, sc)
    G.check()
    assert torch.isfinite(pp).all() and int(step.item()) == 1
    # graph kernels: COO -> CSR, edge weights, the device sampler
    inter = synth.generate((U, I, B, 4000), seed=d)
    tu, ti, _, _ = inter.split_validation()
    a = orc.build_norm_adj(tu, ti, U, I, B)
    rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(a["rowptr"]))
    lib = _lib.load()
    nnz = len(a["col"])
    rowptr, col = G.out((N + 1,), torch.int32), G.out((nnz,), torch.int32)
    st = G.out((1,), torch.int32, fill=0)
    _lib.check(lib.lgcn_csr_from_sorted_coo(G.inp(rows).data_ptr(), G.inp(a["col"].astype(np.int64)).data_ptr(),
                                            nnz, N, N, rowptr.data_ptr(), col.data_ptr(), st.data_ptr(),
                                            _lib.stream_ptr(dev)))
    val = G.out((nnz,))
    _lib.check(lib.lgcn_edge_weights(rowptr.data_ptr(), col.data_ptr(), G.inp(a["dinv"]).data_ptr(), None,
                                     val.data_ptr(), N, _lib.stream_ptr(dev)))
    G.check()
    assert int(st.item()) == 0 and np.array_equal(rowptr.cpu().numpy(), a["rowptr"])
    assert np.array_equal(_bits(val.cpu().numpy()), _bits(a["val"]))
    csr = NormAdjCSR.from_interactions(tu, ti, U, I, B, dev)
    csr.rowptr, csr.col = G.inp(csr.rowptr.cpu().numpy()), G.inp(csr.col.cpu().numpy())
    su, sp, sn = (G.out((257,), torch.int64) for _ in range(3))
    state = G.out((2,), torch.int64, fill=0)
    for _ in range(3):
        ops.sample_bpr(csr, U, I, len(tu), 11, state, su, sp, sn)
    G.check()
    assert int(su.max()) < U and int(sp.max()) < I and int(sn.max()) < I and int(sn.min()) >= 0


@pytest.mark.parametrize("d,n_items,nu", [(128, 20011, 333), (64, 9001, 130)])
def test_scoring_kernels_stay_inside(dev, d, n_items, nu):
    """Exact and tensor-core rating (item splits, ragged last tile / user tile): guarded tables,
    outputs, candidate workspace; ids equal to the CPU oracle."""
    from gcn_recommendation_b200 import ops
    orc = _orc()
    G = Guarded(dev)
    rng = np.random.default_rng(d)
    Fu = rng.standard_normal((nu + 5, d), dtype=np.float32)
    Fi = rng.standard_normal((n_items, d), dtype=np.float32) * (0.3 + rng.random((n_items, 1), dtype=np.float32))
    users = rng.permutation(nu + 5)[:nu].astype(np.int64)
    cnt = rng.integers(0, 30, nu)
    rp = np.zeros(nu + 1, np.int64)
    np.cumsum(cnt, out=rp[1:])
    cols = np.concatenate([np.sort(rng.choice(n_items, c, replace=False)) for c in cnt]).astype(np.int32)
    oids, osc = orc.score_topk(Fu, Fi, users, rp, cols, 20)
    fu, fi, ut = G.inp(Fu), G.inp(Fi), G.inp(users)
    mr, mc = G.inp(rp), G.inp(cols)
    rater = ops.TcRater(n_items, d, dev, nu)
    rater.ws = G.out((rater.wsb,), torch.uint8)
    ids, sc, fail = G.out((nu, 20), torch.int32), G.out((nu, 20)), G.out((nu,), torch.int32)
    rater.prepare(fi)
    rater.topk(fu, fi, ut, mr, mc, 20, ids, sc, fail)
    G.check()
    ok = fail.cpu().numpy() == 0
    assert ok.mean() > 0.9
    assert np.array_equal(ids.cpu().numpy()[ok], oids[ok])
    assert np.array_equal(_bits(sc.cpu().numpy()[ok]), _bits(osc[ok]))
    eids, esc = ops.score_topk_exact(fu, fi, ut, mr, mc, 20)
    G.check()
    assert np.array_equal(eids.cpu().numpy(), oids)
    sums = G.out((2,), torch.float64, fill=0.0)
    ops.eval_metrics(ids, G.inp(rng.integers(0, n_items, nu)), sums)
    G.check()


@pytest.mark.parametrize("d", [64, 128])
def test_fusion_kernels_stay_inside(dev, d):
    from gcn_recommendation_b200 import ops
    orc = _orc()
    G = Guarded(dev)
    rng = np.random.default_rng(d)
    n, c = 333, 768                                   # 333 items: ragged 128-item tiles
    E = rng.standard_normal((n, d), dtype=np.float32) * 0.1
    C = rng.standard_normal((n, c), dtype=np.float32)
    W = rng.standard_normal((d, d + c), dtype=np.float32) * 0.05
    b = rng.standard_normal(d, dtype=np.float32) * 0.1
    gH = rng.standard_normal((n, d), dtype=np.float32)
    H, pre = orc.fusion_forward(E, C, W, b)
    gE, gW, gb = orc.fusion_backward(E, C, W, pre, gH)
    e, cc, w, bb = G.inp(E), G.inp(C), G.inp(W), G.inp(b)
    Hd = ops.fusion_proj_fwd(e, cc, w, bb, out=G.out((n, d)))
    G.check()
    assert rel_err(Hd.cpu().numpy(), H)[0] < 1e-5
    gEd, gWd, gbd = ops.fusion_proj_bwd(e, cc, w, Hd, G.inp(gH), g_eid=G.out((n, d)),
                                        gW=G.out((d, d + c), fill=0.0), gb=G.out((d,), fill=0.0))
    G.check()
    for got, ref in ((gEd, gE), (gWd, gW), (gbd, gb)):
        assert rel_err(got.cpu().numpy(), ref)[0] < 1e-5
