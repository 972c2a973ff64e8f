"""-m gpu: device-side validation metrics vs the `results` dict captured from the reference's own
on_validation_epoch_end (tests/golden/valmetrics.npz) and vs the oracle on larger random inputs."""
import math

import numpy as np
import pytest
import torch

from oracle import ref_path as R
from tests.test_valmetrics import load_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, ref, tag=""):
    assert set(got) == set(ref)
    for g in ref:
        assert set(got[g]) == set(ref[g]), g
        for k, v in ref[g].items():
            a = got[g][k]
            if isinstance(v, float) and math.isnan(v):
                assert math.isnan(a), (tag, g, k)
                continue
            tol = 1e-5 if k == "LogLoss" else 1e-11   # log-loss: float32 terms, summation order differs from numpy's
            assert abs(a - v) <= tol * max(1.0, abs(v)), (tag, g, k, a, v)


@pytest.mark.parametrize("tag", ["plain", "split"])
def test_device_metrics_match_reference(tag):
    from news_recsys_b200.metrics import ValidationMetrics
    batches, warm, ref = load_case(tag)
    vm = ValidationMetrics(k=10, user_in_train_set=warm)
    for u, s, l in batches:
        vm.update(u.to(DEV), s.to(DEV), l.to(DEV))
    _close(vm.compute(), ref, tag)


@pytest.mark.parametrize("k", [10, 3])
def test_device_metrics_match_oracle_on_random_logs(k):
    """40k samples, 3k users (some with one impression, some with hundreds), heavy score ties, warm/cold split."""
    from news_recsys_b200.metrics import ValidationMetrics
    rng = np.random.default_rng(k)
    warm = set(range(0, 1500)) | {str(x) for x in range(1500, 1600)}
    vm = ValidationMetrics(k=k, user_in_train_set=warm)
    us, ss, ls = [], [], []
    for b in range(10):
        u = torch.from_numpy(np.minimum(rng.zipf(1.3, size=4096), 3000).astype(np.int64))
        s = torch.from_numpy(np.clip(np.round(rng.random(4096), 2), 0.01, 0.99).astype(np.float32)).view(-1, 1)
        l = torch.from_numpy((rng.random((4096, 2)) < 0.15).astype(np.float32))
        vm.update(u.to(DEV), s.to(DEV), l.to(DEV))
        a, bb, c = R.validation_pairs(u, s, l)
        us += a.tolist(); ss += list(bb); ls += list(c)
    _close(vm.compute(), R.validation_metrics(us, ss, ls, k, warm), f"k={k}")


def test_train_set_without_numeric_ids_marks_every_user_cold():
    """a non-empty user_in_train_set none of whose entries is a user id — the reference tests
    `uid not in set and str(uid) not in set` (base_model.py:364-366), so EVERY user is cold; the device path used to treat
    the unusable set as absent (every user warm)."""
    from news_recsys_b200.metrics import ValidationMetrics
    rng = np.random.default_rng(7)
    train = {"U12", "abc", "007"}                     # "007" != str(7): not a match either
    vm = ValidationMetrics(k=10, user_in_train_set=train)
    u = torch.from_numpy(rng.integers(1, 40, size=2048).astype(np.int64))
    s = torch.from_numpy(rng.random(2048).astype(np.float32)).view(-1, 1)
    l = torch.from_numpy((rng.random((2048, 2)) < 0.2).astype(np.float32))
    vm.update(u.to(DEV), s.to(DEV), l.to(DEV))
    a, b, c = R.validation_pairs(u, s, l)
    _close(vm.compute(), R.validation_metrics(a.tolist(), list(b), list(c), 10, train), "non-numeric train set")


def test_model_validation_hooks_match_oracle():
    """BaseModel.validation_step / on_validation_epoch_end (same names as the reference) on a Deep model."""
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.synthetic import mind_config, synth_batch
    rows = {"user_id": 60, "item_id": 200, "category": 18, "subcategory": 70, "user_click_category": 18}
    cfg = mind_config("deep", rows)
    torch.manual_seed(0)
    model = Deep(cfg).to(DEV)
    us, ss, ls = [], [], []
    for i in range(4):
        b = {k: v.to(DEV) for k, v in synth_batch(cfg, 256, seed=i, label_p=0.3).items()}
        model.validation_step(b, i)
        a, bb, c = R.validation_pairs(b["user_id"], model.inference(b), b["label"])
        us += a.tolist(); ss += list(bb); ls += list(c)
    _close(model.on_validation_epoch_end(), R.validation_metrics(us, ss, ls, 10, None), "hooks")
    with pytest.raises(Exception):
        from news_recsys_b200.metrics import ValidationMetrics
        ValidationMetrics().update(torch.zeros(4, dtype=torch.int64), torch.zeros(4), torch.zeros(4, 2))   # CPU tensors: loud
