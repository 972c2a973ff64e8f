"""-m gpu: the device-resident feature file — batches assembled on the GPU equal the host path, which
tests/test_ingest.py pins to the reference's own DataReader + default collate."""
import os

import numpy as np
import pytest
import torch

from tests._golden import GOLD

pytestmark = pytest.mark.gpu
DEV = "cuda"
CFG = os.path.join(GOLD, "configs", "train_cf_deep_hist.yaml")
TXT = os.path.join(GOLD, "ingest_features.txt")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    from news_recsys_b200.ingest import DeviceFeatureFile, FeatureFile, compile_feature_file
    out = str(tmp_path_factory.mktemp("ingest") / "features.nrxf")
    compile_feature_file(CFG, TXT, out)
    ff = FeatureFile(out)
    return ff, DeviceFeatureFile(ff, DEV)


@pytest.mark.parametrize("id_dtype", [torch.int64, torch.int32])
def test_device_assembly_equals_host_pack(files, id_dtype):
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.trainer import BatchLayout
    ff, dff = files
    layout = BatchLayout(Deep(CFG), 48, id_dtype)
    rows = np.random.default_rng(3).permutation(200)[:48]
    for kw in (dict(rows=rows), dict(start=152), dict(rows=np.array([3] * 47 + [199]))):
        want = ff.pack(layout, torch.zeros(layout.nbytes, dtype=torch.uint8), **kw)
        got = torch.full((layout.nbytes,), 0, dtype=torch.uint8, device=DEV)
        dkw = dict(rows=torch.from_numpy(kw["rows"]).to(DEV)) if "rows" in kw else kw
        dff.assemble(layout, got, **dkw)
        assert torch.equal(got.cpu(), want)
    # a row outside the file becomes an all-padding sample instead of a wild read
    bad = torch.tensor([0] * 47 + [10**9], device=DEV)
    got = torch.zeros(layout.nbytes, dtype=torch.uint8, device=DEV)
    dff.assemble(layout, got, rows=bad)
    v = layout.views(got)
    assert int(v["user_id"][-1]) == 0 and float(v["user_history_mask"][-1].sum()) == 0 and float(v["label"][-1].abs().sum()) == 0
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    dff.assemble(layout, got, rows=bad, status=st)
    assert int(st.item()) == 4                       # ... and is reported through the status word when one is given
    dff.assemble(layout, got, rows=torch.arange(48, device=DEV), status=st.zero_())
    assert int(st.item()) == 0
    from news_recsys_b200._lib import NrxError
    with pytest.raises(NrxError):
        dff.assemble(layout, got, start=180)


def test_trainer_raises_on_rows_outside_the_device_file(files):
    """a shuffled-row index past the end of the device-resident file used to become a silent fake negative."""
    from news_recsys_b200._lib import NrxError
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.trainer import FusedTrainer
    ff, dff = files
    torch.manual_seed(3)
    tr = FusedTrainer(Deep(CFG).to(DEV), 64, kind="deep", id_dtype=torch.int32)
    rows = torch.arange(64, device=DEV)
    tr.load_rows(dff, rows=rows)
    tr.step()
    tr.check_status()
    rows[5] = 200                                    # the file has 200 rows
    tr.load_rows(dff, rows=rows)
    tr.step()
    with pytest.raises(NrxError, match="outside the device-resident feature file"):
        tr.check_status()


def test_training_from_the_device_file_equals_host_fed_training(files):
    """FusedTrainer.load_rows(device file, device permutation) + step == feed(host-packed blob), bit for bit."""
    from news_recsys_b200.model.sort.deep.model import Deep
    from news_recsys_b200.trainer import FusedTrainer
    ff, dff = files
    perm = np.random.default_rng(5).permutation(200)
    out = {}
    for mode in ("host", "device"):
        torch.manual_seed(3)
        tr = FusedTrainer(Deep(CFG).to(DEV), 64, kind="deep", id_dtype=torch.int32)
        dperm = torch.from_numpy(perm).to(DEV)
        losses = []
        for i in range(3):
            if mode == "host":
                blob = ff.pack(tr.layout, torch.zeros(tr.layout.nbytes, dtype=torch.uint8), rows=perm[i * 64:(i + 1) * 64])
                tr.load_blob(blob.to(DEV))
            else:
                tr.load_rows(dff, rows=dperm[i * 64:(i + 1) * 64])
            losses.append(float(tr.step().item()))
        out[mode] = (losses, {k: v.detach().clone() for k, v in tr.model.state_dict().items()})
    assert out["host"][0] == out["device"][0]
    for k, v in out["host"][1].items():
        assert torch.equal(out["device"][1][k], v), k
