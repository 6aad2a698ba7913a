"""CPU-side tests (-m "not gpu"): host logic, the synthetic generator and the C-ABI surface.
No kernel is launched here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    """liblgcn_b200.so loads and exports exactly what include/lgcn.h declares."""
    from gcn_recommendation_b200 import _lib, build
    lib_path = build.build()
    hdr = open(os.path.join(ROOT, "include", "lgcn.h")).read()
    declared = sorted(set(re.findall(r"LGCN_API[^;(]*?\b(lgcn_\w+)\s*\(", hdr)))
    assert len(declared) >= 15
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "lgcn_" in l)
    assert exported == declared
    assert _lib.exported_symbols() == declared
    lib = _lib.load()
    assert lib.lgcn_abi_version() == _lib.ABI_VERSION
    assert lib.lgcn_sizeof_spmm_args() == ctypes.sizeof(_lib.SpmmArgs)
    assert b"embedding dim" in lib.lgcn_error_string(-1)


def test_spmm_plan_query_matches_documented_selection():
    """``lgcn_spmm_launches`` (host-only): Gowalla shape -> small path, 2 launches with long rows;
    Amazon shape -> large-graph kernels, 3 launches, at every sharded width."""
    from gcn_recommendation_b200 import _lib, ops
    assert ops._spmm_plan(70_840, 64, 0) == (1, True)
    assert ops._spmm_plan(70_840, 64, 1600) == (2, True)
    assert ops._spmm_plan(70_840, 64, 1600, long_done=True) == (1, True)      # combined in place
    assert ops._spmm_plan(14_700_001, 64, 2749, long_done=True) == (3, False)
    for d in (16, 32, 64, 128):
        assert ops._spmm_plan(14_700_001, d, 2749) == (3, False)
    old = ops.SPMM_FLAGS_EXTRA
    ops.SPMM_FLAGS_EXTRA = _lib.SPMM_F_BIG_PATH
    try:
        assert ops._spmm_plan(7006, 64, 10) == (3, False)
    finally:
        ops.SPMM_FLAGS_EXTRA = old
    assert _lib.load().lgcn_spmm_launches(10, 48, 0, 0, None) == -1
    # rows per chunk that lgcn_spmm_args.chunk_order permutes (0 = the kernel ignores the order)
    rows = _lib.load().lgcn_spmm_chunk_rows
    assert rows(70_840, 64, 0) == 4 and rows(70_840, 16, 0) == 4
    assert [rows(14_700_001, d, 0) for d in (16, 32, 64, 128, 256)] == [4, 8, 8, 0, 0]
    assert rows(14_700_001, 16, _lib.SPMM_F_NO_RING) == 0 and rows(10, 48, 0) == -1


def test_argument_errors_are_reported_not_thrown():
    from gcn_recommendation_b200 import _lib
    lib = _lib.load()
    a = _lib.SpmmArgs()
    a.d = 48
    assert lib.lgcn_spmm(ctypes.byref(a), None) == -1            # LGCN_E_BAD_DIM
    a.d = 64
    assert lib.lgcn_spmm(ctypes.byref(a), None) == -2            # null pointers
    assert lib.lgcn_adam(None, None, None, None, None, 4, None, 0.9, 0.999, 1e-8, None) == -2
    assert lib.lgcn_score_topk(None, None, None, 1, 1, 64, None, None, 64, None, None, None, 0, None) == -2


def test_no_cpu_fallback_on_cpu_tensors():
    from gcn_recommendation_b200 import _lib, ops
    with pytest.raises(_lib.LgcnError):
        ops.fusion_proj_fwd(torch.zeros(4, 64), torch.zeros(4, 768), torch.zeros(64, 832), torch.zeros(64))
    with pytest.raises(_lib.LgcnError):
        ops.score_topk(torch.zeros(4, 64), torch.zeros(9, 64), torch.arange(4))


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (or the reference)."""
    for base in ("gcn_recommendation_b200", "models"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                    assert "liblgcn_oracle" not in src and "/root/reference" not in src, f


@pytest.mark.parametrize("shape", ["tiny", "small"])
def test_synthetic_generator_contract(shape):
    from gcn_recommendation_b200 import synth
    U, I, B, total, _, _ = synth.SHAPES[shape]
    a = synth.generate(shape, seed=3)
    b = synth.generate(shape, seed=3)
    assert np.array_equal(a.train_user, b.train_user) and np.array_equal(a.train_item, b.train_item)
    assert len(a.train_user) + len(a.test_user) == total
    key = np.concatenate([a.train_user, a.test_user]) * I + np.concatenate([a.train_item, a.test_item])
    assert len(np.unique(key)) == total, "pairs must be unique"
    assert len(a.test_user) == U and len(np.unique(a.test_user)) == U
    tu, ti, vu, vi = a.split_validation()
    assert len(vu) == U and len(np.unique(vu)) == U
    assert np.bincount(tu, minlength=U).min() >= 1     # >= 3 interactions per user
    assert ti.max() < I and tu.max() < U


def test_reference_on_disk_format(tmp_path):
    import json
    import pandas as pd
    from gcn_recommendation_b200 import synth
    inter = synth.generate("tiny", seed=0)
    synth.write_reference_format(inter, str(tmp_path), synth.side_embeddings(inter.num_items, 8))
    tr = pd.read_parquet(tmp_path / "train.parquet")
    assert list(tr.columns) == ["user_idx", "item_idx"] and tr["user_idx"].dtype == np.int64
    assert list(pd.read_parquet(tmp_path / "item_brand.parquet").columns) == ["item_idx", "brand_idx"]
    st = json.load(open(tmp_path / "stats.json"))
    assert st == {"num_users": 120, "num_items": 200, "num_brands": 1}
    assert np.load(tmp_path / "item_embeddings.npy").shape == (200, 8)


def test_long_row_plan_host_logic():
    """Kernel layout of include/lgcn.h: packed {col,val} of the short rows behind a flagged
    rowptr, long rows moved to their own CSR and cut into segments."""
    from gcn_recommendation_b200.graph import NormAdjCSR
    deg = np.array([0, 3, 70, 1, 64, 65, 200, 0], np.int64)
    rowptr = np.zeros(len(deg) + 1, np.int32)
    rowptr[1:] = np.cumsum(deg)
    nnz = int(rowptr[-1])
    col = torch.arange(nnz, dtype=torch.int32)
    val = torch.arange(nnz, dtype=torch.float32) * 0.5
    g = NormAdjCSR(torch.from_numpy(rowptr), col, val, n_cols=nnz, long_row_threshold=64, seg_len=32)
    assert g.n_long == 3 and g.long_row_ids.tolist() == [2, 5, 6]
    assert g.long_seg_ptr.tolist() == [0, 3, 6, 13] and g.n_seg == 13
    assert g.long_rowptr.tolist() == [0, 70, 135, 335]
    flagged = g.rowptr_flagged.numpy().view(np.uint32)
    assert (flagged >> 31).tolist() == [0, 0, 1, 0, 0, 1, 1, 0, 0]
    assert (flagged & 0x7fffffff).tolist() == [0, 0, 3, 3, 4, 68, 68, 68, 68]
    assert g.colval.shape == (68, 2) and g.long_colval.shape == (335, 2)
    # short entries keep CSR order: rows 1, 3, 4
    assert g.colval[:, 0].tolist() == list(range(0, 3)) + [73] + list(range(74, 138))
    assert torch.equal(g.colval[:, 1].view(torch.float32), g.colval[:, 0].float() * 0.5)
    assert g.long_colval[:70, 0].tolist() == list(range(3, 73))
    g0 = NormAdjCSR(torch.from_numpy(rowptr), col, val, n_cols=nnz, long_row_threshold=0)
    assert g0.n_long == 0 and g0.colval.shape == (nnz, 2)
    assert torch.equal(g0.rowptr_flagged, g0.rowptr)


def test_column_residency_classes_host_logic():
    """include/lgcn.h LGCN_COL_*: the n_hot highest-degree nodes carry bit 31 in the packed column
    index, degree-1 nodes bit 30, the low 30 bits stay the column id; clearing restores it."""
    from gcn_recommendation_b200 import synth
    from oracle import lgcn_oracle as orc
    from gcn_recommendation_b200.graph import NormAdjCSR
    inter = synth.generate("tiny", seed=3)
    tu, ti, _, _ = inter.split_validation()
    U, I, B = inter.num_users, inter.num_items, inter.num_brands
    a = orc.build_norm_adj(tu, ti, U, I, B)
    g = NormAdjCSR(torch.from_numpy(a["rowptr"].astype(np.int32)), torch.from_numpy(a["col"].astype(np.int32)),
                   torch.from_numpy(a["val"]), U + I + B, long_row_threshold=0)
    plain = g.colval[:, 0].clone()
    g.mark_hot_columns(25)
    raw = g.colval[:, 0]
    assert torch.equal(raw & 0x3fffffff, plain)
    deg = np.diff(a["rowptr"])
    cols = plain.numpy()
    hot = (raw < 0).numpy()
    once = (((raw >> 30) & 1) == 1).numpy() & ~hot
    assert np.array_equal(once, deg[cols] == 1)
    kth = np.sort(deg)[-25]
    assert hot[deg[cols] > kth].all() and not hot[deg[cols] < kth].any()
    assert len(np.unique(cols[hot])) <= 25 and hot.any()
    g.mark_hot_columns(0)
    assert torch.equal(g.colval[:, 0], plain)
    assert torch.equal(g.colval[:, 1], torch.from_numpy(a["val"]).view(torch.int32))


def test_engine_checkpoint_is_reference_format():
    """``engine.state_dict()`` has the reference's keys / order / shapes and loads into the
    reference-compatible drop-in module (reference ``main.py:550,571``)."""
    import types
    from gcn_recommendation_b200.engine import LightGCNEngine
    from gcn_recommendation_b200.graph import NormAdjCSR
    from models.lightgcn import LightGCN
    U, I, B, d = 5, 7, 1, 64
    N = U + I + B
    rowptr = torch.zeros(N + 1, dtype=torch.int32)
    g = NormAdjCSR(rowptr, torch.zeros(0, dtype=torch.int32), torch.zeros(0), N)
    table = torch.randn((N, d))
    eng = LightGCNEngine(g, U, I, B, 3, table.clone())
    sd = eng.state_dict()
    m = LightGCN(U, I, B, types.SimpleNamespace(embedding_dim=d, n_layers=3, debug=False))
    assert list(sd.keys()) == list(m.state_dict().keys())
    m.load_state_dict(sd)                                   # strict: every key and shape matches
    assert torch.equal(m.item_embedding.weight, table[U:U + I])
    eng2 = LightGCNEngine(g, U, I, B, 3, torch.zeros((N, d)))
    eng2.load_state_dict(m.state_dict())
    assert torch.equal(eng2.P, table)
    from gcn_recommendation_b200._lib import LgcnError
    bad = dict(sd)
    bad["user_embedding.weight"] = torch.zeros((U + 1, d))
    with pytest.raises(LgcnError):
        eng2.load_state_dict(bad)


def test_bench_reference_arm_prints_one_contract_line():
    """``bench.py --impl reference`` (the reference's CPU path on the host cores) prints exactly
    one JSON line with the contract's keys; it needs no GPU."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--workload", "tiny", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lightgcn_epoch_s" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["value"] > 0 and d["steps"] == 2
    assert d["config"]["workload"] == "tiny"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_mask_csr_matches_groupby_lists():
    from gcn_recommendation_b200.engine import build_mask_csr
    tu = np.array([3, 1, 3, 0, 1, 3], np.int64)
    ti = np.array([9, 4, 2, 7, 1, 5], np.int64)
    rp, col = build_mask_csr(np.array([3, 2, 1], np.int64), tu, ti, 4)
    assert rp.tolist() == [0, 3, 3, 5] and col.tolist() == [2, 5, 9, 1, 4]


def test_parameter_packing_keeps_identity_and_state_dict_keys():
    import types
    from models._packing import is_packed
    from models.lightgcn import LightGCN
    cfg = types.SimpleNamespace(embedding_dim=64, n_layers=3, debug=False)
    m = LightGCN(10, 20, 1, cfg)
    params = list(m.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    block = m.table_block()
    assert is_packed(m._tables()) and block.shape == (31, 64)
    assert all(a is b for a, b in zip(params, m.parameters()))
    assert list(m.state_dict().keys()) == ["user_embedding.weight", "brand_embedding.weight",
                                           "item_embedding.weight"]
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])
    assert torch.equal(block[:10], m.user_embedding.weight) and torch.equal(block[30:], m.brand_embedding.weight)
    with pytest.raises(ValueError):
        LightGCN(10, 20, 1, cfg, pretrained_item_emb=np.zeros((20, 32), np.float32))
    from models.lightgcn_fusion import LightGCN_Fusion
    with pytest.raises(ValueError):
        LightGCN_Fusion(10, 20, 1, cfg)
    del opt


def test_epoch_history_csv_is_the_reference_format(tmp_path):
    """`<model>_epoch_history.csv` byte-identical to the reference Logger's
    ``DataFrame.to_csv(index=False)`` (reference main.py:106-126), and a committed result file of
    the reference parses back into the same rows."""
    import pandas as pd

    from gcn_recommendation_b200.history import EpochHistory, logger_name
    rows = [(5, 0.4604123456789, 0.1151, 0.0615), (10, 1 / 3, 0.6622, 0.34220000000000006),
            (150, 2.5e-05, 0.0, 1.0)]
    name = logger_name("LightGCN", False, True)
    assert name == "LightGCN_no_brand_pretrained"
    h = EpochHistory(str(tmp_path), name)
    assert h.save() is None and not os.path.exists(h.path)          # main.py:114-116
    for r in rows:
        h.log_epoch_metrics(*r)
    path = h.save()
    assert os.path.basename(path) == "LightGCN_no_brand_pretrained_epoch_history.csv"
    ref = tmp_path / "ref.csv"
    pd.DataFrame({"epoch": [r[0] for r in rows], "avg_loss": [r[1] for r in rows],
                  "recall": [r[2] for r in rows], "ndcg": [r[3] for r in rows]}).to_csv(ref, index=False)
    assert open(path, "rb").read() == open(ref, "rb").read()


def test_host_only_plan_queries():
    """Host-only entry points of the C ABI (no GPU needed): kernel selection names, launch counts,
    the scoring workspace that must serve every smaller batch, item-split launch counts."""
    import ctypes

    from gcn_recommendation_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(128)

    def name(n_rows, d, mode, flags=0, sparse_x=0):
        assert lib.lgcn_spmm_kernel_name(n_rows, d, mode, flags, sparse_x, buf, 128) == 0
        return buf.value.decode()

    assert name(70_840, 64, _lib.SPMM_PLAIN) == "spmm_chunk_kernel<64,plain,4-row chunks,nohints>"
    assert name(14_700_001, 128, _lib.SPMM_PLAIN, _lib.SPMM_F_STREAM_HINTS) == "spmm_ring_kernel<128,plain,hints>"
    assert name(14_700_001, 16, _lib.SPMM_MEAN, _lib.SPMM_F_STREAM_HINTS) == "spmm_ring_kernel<16,mean,hints>"
    assert name(14_700_001, 128, _lib.SPMM_ADAM, _lib.SPMM_F_STREAM_HINTS) == "spmm_ring_kernel<128,adam,hints>"
    assert name(14_700_001, 128, _lib.SPMM_ADAM, _lib.SPMM_F_STREAM_HINTS | _lib.SPMM_F_NO_RING).startswith(
        "spmm_chunk_kernel<128,adam,16-row")
    assert name(14_700_001, 128, _lib.SPMM_ADD, _lib.SPMM_F_STREAM_HINTS, 1) == "spmm_live_kernel<128,hints>"
    assert lib.lgcn_spmm_kernel_name(10, 48, 0, 0, 0, buf, 128) == -1                   # LGCN_E_BAD_DIM
    assert lib.lgcn_spmm_kernel_name(10, 64, 9, 0, 0, buf, 128) == -2                   # LGCN_E_BAD_ARG
    # scoring: whole waves run unsplit (2 launches), small batches are item-split (+ merge)
    assert lib.lgcn_score_tc_launches(148 * 128, 4_400_000) == 2
    assert lib.lgcn_score_tc_launches(1024, 4_400_000) == 3
    assert lib.lgcn_score_tc_launches(1024, 2000) == 2                                 # too few tiles to split
    big = lib.lgcn_score_tc_workspace(75_776, 4_400_000, 128)
    for nu in (1, 128, 1024, 18_944, 75_776):
        assert lib.lgcn_score_tc_workspace(nu, 4_400_000, 128) <= big
    assert lib.lgcn_score_tc_workspace(100, 1000, 48) == 0                              # unsupported width


def test_mask_csr_builders_agree():
    """``mask_csr_from_graph`` (validation-time masks read off the graph's user rows) equals the
    sort-based ``build_mask_csr`` and the oracle's ``mask_csr`` (reference main.py:407)."""
    import numpy as np
    import torch

    from gcn_recommendation_b200 import synth
    from gcn_recommendation_b200.engine import build_mask_csr, mask_csr_from_graph
    from oracle import lgcn_oracle as orc
    inter = synth.generate("tiny", seed=3)
    tu, ti, vu, vi = inter.split_validation()
    U, I, B = inter.num_users, inter.num_items, inter.num_brands
    a = orc.build_norm_adj(tu, ti, U, I, B)
    g = type("G", (), {})()
    g.device = torch.device("cpu")
    g.rowptr = torch.from_numpy(a["rowptr"].astype(np.int32))
    g.col = torch.from_numpy(a["col"])
    users = np.asarray([5, 0, 119, 7, 7], np.int64)
    rp1, c1 = mask_csr_from_graph(g, users, U)
    rp2, c2 = build_mask_csr(users, tu, ti, U)
    rp3, c3 = orc.mask_csr(users, tu, ti, U)
    assert np.array_equal(rp1.numpy(), rp3) and np.array_equal(c1.numpy(), c3)
    assert np.array_equal(rp2.numpy(), rp3) and np.array_equal(c2.numpy(), c3)
    assert c1.dtype == torch.int32 and rp1.dtype == torch.int64


def test_chunk_order_policy_is_host_logic(monkeypatch):
    """``ops._set_chunk_order`` (no GPU): small graphs always get the global order; HBM-streaming
    graphs get the windowed order only for MEAN / ADAM launches and the flagged walk with dense
    output, never at d = 128 (one worker per warp) -- profiles/r02_chunk_order_ab.txt."""
    from gcn_recommendation_b200 import _lib, ops

    class FakeGraph:
        def __init__(self, n_rows):
            self.n_rows = n_rows
            self.asked = []

        def chunk_order_for(self, rows, windowed):
            self.asked.append((rows, windowed))
            return ("order", rows, windowed)

    monkeypatch.setattr(ops, "ptr", lambda t, *a, **k: 1234 if t is not None else None)

    def run(n_rows, d, mode, x_rowflag=None, y_rowflag=None, policy="auto"):
        monkeypatch.setattr(ops, "CHUNK_ORDER_LARGE", policy)
        g = FakeGraph(n_rows)
        a = _lib.SpmmArgs()
        a.n_rows, a.d, a.mode, a.flags = n_rows, d, mode, 0
        a.x_rowflag, a.y_rowflag = x_rowflag, y_rowflag
        ops._set_chunk_order(a, g)
        return g.asked

    assert run(70_840, 64, ops.SPMM_PLAIN) == [(4, False)]                 # small graph: every mode
    assert run(70_840, 64, ops.SPMM_ADAM) == [(4, False)]
    big = 14_700_001
    assert run(big, 16, ops.SPMM_PLAIN) == [] and run(big, 16, ops.SPMM_ADD) == []
    assert run(big, 16, ops.SPMM_MEAN) == [(4, True)] and run(big, 32, ops.SPMM_ADAM) == [(8, True)]
    assert run(big, 64, ops.SPMM_ADD, x_rowflag=1) == [(8, True)]          # hop 2: flagged in, dense out
    assert run(big, 64, ops.SPMM_ADD, x_rowflag=1, y_rowflag=1) == []      # hop 1: sparse out
    assert run(big, 128, ops.SPMM_MEAN) == []                              # one worker per warp
    assert run(big, 16, ops.SPMM_PLAIN, policy="all") == [(4, True)]
    assert run(big, 16, ops.SPMM_MEAN, policy="off") == []
