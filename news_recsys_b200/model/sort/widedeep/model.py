"""WideDeep — mirror of reference src/model/sort/widedeep/model.py: `WideDeepModel` (:14-27,
sigmoid(sum(wide_x) + bias + MLP(deep_x))) and `WideDeep.get_inp_embedding` (:53-69: column 0 of each
wide field -> wide_x, the remaining columns and every other field -> deep_x).

Execution: K1 writes the concatenated features once; the wide sum reads column 0 of the wide fields in
place (nrx_field_logit WIDE).  The reference's slice+cat copies that build deep_x are avoided by moving
the column selection to the weight side: the tower runs on the full feature row with a first-layer
weight whose wide columns are zero (a [128, ΣD] scatter of the [128, ΣD - n_wide] parameter, per step),
so no [B, ΣD] activation copy is ever made and autograd returns the gradient in the parameter's shape."""
import torch
import torch.nn as nn

from .... import ops
from ...BaseModel.base_model import BaseModel
from ...model_utils.utils import MLP


class WideDeepModel(nn.Module):
    def __init__(self, input_dim, hidden_dims=(32, 32, 1)):
        super().__init__()
        self.wide_network = torch.sum  # attribute parity with the reference (:19)
        self.deep_network = MLP(dims=[input_dim] + list(hidden_dims))
        self.bias = nn.Parameter(torch.zeros(1))

    def forward(self, wide_x, deep_x):
        """Reference signature (:24-27) on already split tensors."""
        nw = wide_x.shape[1]
        wide = ops.FieldLogitFn.apply(wide_x.contiguous(), list(range(nw)), [1] * nw, ops.L.FIELD_WIDE)
        deep = self.deep_network(deep_x.contiguous())
        return ops.SigmoidFn.apply(self.bias, wide, deep.view(-1)).view(-1, 1)


class WideDeep(BaseModel):
    def __init__(self, config_path):
        super().__init__(config_path)
        self.wide_feature_names = set(self.config.wide_and_deep_cfg.wide_feature_names)
        self.score_fc = WideDeepModel(
            input_dim=self.user_input_dim + self.item_input_dim - len(self.wide_feature_names),
            hidden_dims=[128, 128, 128, 64, 1])

    def _split_cols(self, dims, fnames):
        wide_cols, deep_cols, s = [], [], 0
        for d, n in zip(dims, fnames):
            if n in self.wide_feature_names:
                wide_cols.append(s)
                deep_cols += list(range(s + 1, s + d))
            else:
                deep_cols += list(range(s, s + d))
            s += d
        return wide_cols, deep_cols

    def get_inp_embedding(self, batch):
        features, dims, fnames = self.get_embeddings_from_batch(batch, self.user_feature_names | self.item_feature_names)
        wide_cols, deep_cols = self._split_cols(dims, fnames)
        return features[:, wide_cols], features[:, deep_cols]

    def forward(self, x):
        features, dims, fnames = self.get_embeddings_from_batch(x, self.user_feature_names | self.item_feature_names)
        wide_cols, deep_cols = self._split_cols(dims, fnames)
        wide = ops.FieldLogitFn.apply(features, wide_cols, [1] * len(wide_cols), ops.L.FIELD_WIDE)
        lin = self.score_fc.deep_network.linears()
        w0 = lin[0].weight
        idx = torch.as_tensor(deep_cols, device=features.device)
        w0_full = w0.new_zeros(w0.shape[0], features.shape[1]).index_copy(1, idx, w0)  # zero columns for the wide weights
        ws = [w0_full] + [m.weight for m in lin[1:]]
        bs = [m.bias for m in lin]
        deep = ops.TowerFn.apply(features, None, len(lin), *ws, *bs)
        return ops.SigmoidFn.apply(self.score_fc.bias, wide, deep.view(-1)).view(-1, 1)
