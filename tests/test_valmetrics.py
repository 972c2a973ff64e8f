"""CPU: the oracle restatement of the reference's validation scoring / metrics path against the `results`
dict captured from the reference's own `on_validation_epoch_end` (tests/golden/valmetrics.npz), and the AUC
restatement against sklearn (the library the reference calls)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_path as R
from tests._golden import GOLD


def load_case(tag):
    z = np.load(os.path.join(GOLD, "valmetrics.npz"))
    n = int(z[f"{tag}__n_batches"])
    batches = [(torch.from_numpy(z[f"{tag}__uid{i}"]), torch.from_numpy(z[f"{tag}__score{i}"]), torch.from_numpy(z[f"{tag}__label{i}"]))
               for i in range(n)]
    warm = None
    if f"{tag}__warm_int" in z.files:
        warm = set(int(x) for x in z[f"{tag}__warm_int"]) | set(str(x) for x in z[f"{tag}__warm_str"])
    res = {}
    for key in z.files:
        if key.startswith(f"{tag}__res__"):
            _, _, grp, name = key.split("__")
            res.setdefault(grp, {})[name] = float(z[key])
    return batches, warm, res


@pytest.mark.parametrize("tag", ["plain", "split"])
def test_oracle_metrics_match_reference(tag):
    batches, warm, ref = load_case(tag)
    us, ss, ls = [], [], []
    for u, s, l in batches:
        a, b, c = R.validation_pairs(u, s, l)
        us.extend(a.tolist()); ss.extend(b); ls.extend(c)
    got = R.validation_metrics(us, ss, ls, k=10, user_in_train_set=warm)
    assert set(got) == set(ref)
    for grp in ref:
        assert set(got[grp]) == set(ref[grp]), grp
        for k, v in ref[grp].items():
            assert got[grp][k] == pytest.approx(v, rel=1e-9, abs=1e-12), (grp, k)


def test_auc_restatement_equals_sklearn():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    for n in (2, 3, 10, 257):
        for _ in range(20):
            y = rng.integers(0, 2, size=n).astype(np.float32)
            if len(set(y.tolist())) < 2:
                y[0], y[1] = 0, 1
            p = np.round(rng.random(n).astype(np.float32), 1 if n > 3 else 3)   # many ties
            assert R.roc_auc(y, p) == pytest.approx(roc_auc_score(y, p), rel=1e-12, abs=1e-15)
