"""LR — mirror of reference src/model/sort/lr/model.py (forward :24-27): sigmoid(sum of all
D=1 embeddings), no bias.  K1 + SUM field logit + sigmoid, all in libnrx."""
import torch

from .... import ops
from ...BaseModel.base_model import BaseModel


class LR(BaseModel):
    def __init__(self, config_path):
        super().__init__(config_path)
        self.score_fc = torch.sum  # kept for attribute parity with the reference (:17)

    def get_inp_embedding(self, batch):
        features, _, _ = self.get_embeddings_from_batch(batch, self.user_feature_names | self.item_feature_names)
        return features

    def forward(self, x):
        feats, dims, _ = self.get_embeddings_from_batch(x, self.user_feature_names | self.item_feature_names)
        cols, c = [], 0
        for d in dims:
            cols.append(c)
            c += d
        logit = ops.FieldLogitFn.apply(feats, cols, dims, ops.L.FIELD_SUM)
        return ops.SigmoidFn.apply(None, logit)  # shape [B], like the reference
