"""ORACLE / TEST INFRASTRUCTURE — golden vectors for the validation scoring / metrics path.

Runs the reference's OWN `BaseModel.validation_step` (base_model.py:320-330) and `on_validation_epoch_end` (:333-528)
on a reference Deep model over a few synthetic batches.  The method only prints / logs its `results` dict, so the
dict is captured from the method's frame when it returns (`sys.setprofile`) — no arithmetic is supplied from outside.

    python oracle/make_golden_valmetrics.py      # build container only; writes tests/golden/valmetrics.npz
"""
import os as _os
import sys as _sys

# set iteration order inside the reference depends on the string hash seed: pin it so that regenerating reproduces the fixture
if _os.environ.get("PYTHONHASHSEED") != "0":
    _os.environ["PYTHONHASHSEED"] = "0"
    _os.execv(_sys.executable, [_sys.executable] + _sys.argv)

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import synth_batch, GOLD, CFGS  # noqa: E402  (sets up refshim + reference on sys.path)

import yaml  # noqa: E402


def run(tag, with_warm):
    from src.model.sort.deep.model import Deep
    path = os.path.join(CFGS, "train_cf_deep.yaml")
    cfg = yaml.safe_load(open(path))
    torch.manual_seed(42)
    model = Deep(path)
    model.user_scores_dict = {}
    gen = torch.Generator().manual_seed(77)
    uids, scores, labels = [], [], []
    for b in range(6):
        batch = synth_batch(cfg, 64, gen, label_p=0.3)
        batch["user_id"] = torch.randint(1, 40, (64,), generator=gen)       # ~10 impressions per user, some with 1
        if b == 0:
            batch["user_id"][:12] = 39                                           # a user with more than k items
        with torch.no_grad():
            sc = model.inference(batch)
        if b == 1:                                                               # exact score ties inside users
            sc = torch.round(sc * 8) / 8
            model.inference = (lambda fixed: (lambda _b: fixed))(sc)
        model.validation_step(batch, b)
        if b == 1:
            del model.inference
        uids.append(batch["user_id"].clone()); scores.append(sc.clone()); labels.append(batch["label"].clone())
    if with_warm:
        model.user_in_train_set = set(list(range(1, 20)) + [str(u) for u in range(20, 25)])   # ints and strings (:366)
    captured = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "on_validation_epoch_end":
            captured.update(frame.f_locals.get("results") or {})

    model.current_epoch = 0
    sys.setprofile(prof)
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            model.on_validation_epoch_end()
    finally:
        sys.setprofile(None)
    assert captured, "results not captured"
    out = {}
    for i, (u, s, l) in enumerate(zip(uids, scores, labels)):
        out[f"{tag}__uid{i}"], out[f"{tag}__score{i}"], out[f"{tag}__label{i}"] = u.numpy(), s.numpy(), l.numpy()
    out[f"{tag}__n_batches"] = np.array(len(uids))
    if with_warm:
        out[f"{tag}__warm_int"] = np.array(sorted(x for x in model.user_in_train_set if isinstance(x, int)))
        out[f"{tag}__warm_str"] = np.array(sorted(x for x in model.user_in_train_set if isinstance(x, str)))
    for grp, d in captured.items():
        for k, v in d.items():
            out[f"{tag}__res__{grp}__{k}"] = np.array(float(v), dtype=np.float64)
    print(tag, {g: {k: round(float(v), 4) for k, v in d.items()} for g, d in captured.items()})
    return out


def main():
    out = {}
    out.update(run("plain", False))
    out.update(run("split", True))
    np.savez_compressed(os.path.join(GOLD, "valmetrics.npz"), **out)
    print("wrote valmetrics.npz")


if __name__ == "__main__":
    main()
