"""Import shim (test infrastructure only): stands in for the `lightning` package,
which is absent from this image, so that /root/reference's modules import
UNMODIFIED when oracle/make_golden.py generates the golden fixtures.
Contains no arithmetic: LightningModule == nn.Module + no-op logging hooks."""
import torch
import torch.nn as nn


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.logger = None
        self.trainer = None
        self.current_epoch = 0

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")


class Trainer:  # never used by the golden generator
    def __init__(self, *a, **k):
        raise RuntimeError("refshim: Trainer is not available")


class LightningDataModule:
    pass


def seed_everything(seed, workers=False):
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed
