"""Host-side LR schedule with the reference's semantics (src/model/model_utils/lr_schedule.py:6-28):
constant lrs[0] before milestones[0], cosine to lrs[1] until milestones[1], then constant."""
import math

from torch.optim.lr_scheduler import LRScheduler


class CosinDecayLR(LRScheduler):
    def __init__(self, optimizer, lrs=(1e-3, 1e-5), milestones=(2000, 5000)):
        if len(lrs) != 2 or len(milestones) != 2:
            raise AssertionError("CosinDecayLR takes exactly 2 lrs and 2 milestones")
        self.lrs = list(lrs)
        self.milestones = list(milestones)
        super().__init__(optimizer)

    def lr_at(self, step: int) -> float:
        lo, hi = self.milestones
        if step < lo:
            return self.lrs[0]
        if step >= hi:
            return self.lrs[1]
        t = (step - lo) / max(1, hi - lo)
        return self.lrs[1] + (self.lrs[0] - self.lrs[1]) * 0.5 * (1.0 + math.cos(math.pi * t))

    def get_lr(self):
        lr = self.lr_at(self.last_epoch)
        return [lr for _ in self.optimizer.param_groups]
