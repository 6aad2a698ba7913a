"""Epoch history in the reference's artefact format.

The reference's ``Logger`` (reference ``main.py:89-126``) collects one row per evaluated epoch --
``log_epoch_metrics(epoch, avg_loss, recall, ndcg)`` (``main.py:106-111``, called every
``val_interval`` epochs from ``main.py:536-545``) -- and ``save`` writes them with
``DataFrame.to_csv(index=False)`` to ``<results_dir>/<model_name>_epoch_history.csv`` with the
columns ``epoch,avg_loss,recall,ndcg`` (``main.py:119-126``).  This writer produces the same file
byte for byte (pandas formats floats with ``repr``) without pandas or matplotlib; the PNG plot of
``main.py:130-169`` is observability only and is not reproduced.
"""
from __future__ import annotations

import os


class EpochHistory:
    COLUMNS = ("epoch", "avg_loss", "recall", "ndcg")

    def __init__(self, results_dir, model_name):
        self.results_dir, self.model_name = results_dir, model_name
        self.rows = []
        self.batch_losses = []

    def log_batch_loss(self, loss):
        """reference ``main.py:101-104``."""
        self.batch_losses.append(float(loss))

    def log_epoch_metrics(self, epoch, avg_loss, recall, ndcg):
        """reference ``main.py:106-111``."""
        self.rows.append((int(epoch), float(avg_loss), float(recall), float(ndcg)))

    @property
    def path(self):
        return os.path.join(self.results_dir, f"{self.model_name}_epoch_history.csv")

    def save(self):
        """reference ``main.py:113-126``: nothing is written when no epoch was logged."""
        if not self.rows:
            return None
        os.makedirs(self.results_dir, exist_ok=True)
        with open(self.path, "w", newline="") as f:
            f.write(",".join(self.COLUMNS) + "\n")
            for e, l, r, n in self.rows:
                f.write(f"{e},{l!r},{r!r},{n!r}\n")
        return self.path


def logger_name(model_name, use_brand, use_pretrained_emb):
    """File-name stem of reference ``main.py:444-446`` (``LightGCN_no_brand_pretrained`` ...)."""
    return f"{model_name}_{'brand' if use_brand else 'no_brand'}" + ("_pretrained" if use_pretrained_emb else "")
