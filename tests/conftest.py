import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
        return cache[name]

    return load


def rel_err(a, b):
    """max|a-b| / max|b| and Frobenius relative error (the 1e-5 bar of BASELINE.json)."""
    import numpy as np

    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(float(np.abs(b).max()), 1e-30)
    fro = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    return float(np.abs(a - b).max() / den), fro


def digest(a):
    """fp64 (sum, sum of squares): the compact witness oracle/make_golden.py stores for big tensors."""
    import numpy as np

    a = np.asarray(a, np.float64)
    return np.asarray([a.sum(), (a * a).sum()], np.float64)


def config_scale_inputs(g):
    """Regenerate the inputs of the config-scale golden (synth + seed) and check them against the
    stored digests.  Returns (inter, tu, ti, vu, vi)."""
    import numpy as np

    from gcn_recommendation_b200 import synth
    inter = synth.generate(str(g["shape"]), seed=int(g["seed"]))
    tu, ti, vu, vi = inter.split_validation()
    I = inter.num_items
    assert len(tu) == int(g["n_train"]) and len(vu) == int(g["n_val"])
    assert np.array_equal(digest(tu * I + ti), g["train_digest"])
    assert np.array_equal(digest(vu * I + vi), g["val_digest"])
    return inter, tu, ti, vu, vi


def near_tie_rows_ok(ids, ref_ids, ref_sc, tol=2e-6):
    """Rows whose ids differ from the reference's torch.topk must differ only inside fp32
    near-ties: at every differing position the reference's score has a neighbour (or the cut-off
    at rank k) within ``tol`` relative.  Returns the number of differing rows."""
    import numpy as np

    bad_rows = np.where((ids != ref_ids).any(1))[0]
    for r in bad_rows:
        s = ref_sc[r].astype(np.float64)
        scale = np.abs(s).max()
        for j in np.where(ids[r] != ref_ids[r])[0]:
            near = j == len(s) - 1                       # swap across the rank-k cut-off
            if j > 0:
                near |= abs(s[j] - s[j - 1]) <= tol * scale
            if j + 1 < len(s):
                near |= abs(s[j] - s[j + 1]) <= tol * scale
            assert near, f"user row {r}: id mismatch at rank {j} without a near-tie"
    return len(bad_rows)
