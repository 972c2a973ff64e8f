"""DeepFM — the reference ships no DeepFM class; its docs sketch the config
(documents/config_file_introduction.md:157-176: `deepfm_cfg{fm_feature_names, fm_dim}`).  It is composed
from the reference's own parts, as BASELINE.json config 2 asks: the FM logit of `FMModel.forward`
(src/model/sort/fm/model.py:18-25, column 0 = first-order weight, columns 1.. = latent vector, as in
fm/model.py:48-59) over `fm_feature_names`, plus the `MLP` logit (model_utils/utils.py:6-17,
[ΣD,128,128,128,64,1] like sort/deep/model.py:29) over ALL embeddings, one sigmoid.
state_dict: `score_fc.bias`, `score_fc.deep_network.network.{0..8}.*` (the WideDeep naming)."""
import torch
import torch.nn as nn

from .... import ops
from ...BaseModel.base_model import BaseModel
from ...model_utils.utils import MLP


class DeepFMModel(nn.Module):
    def __init__(self, input_dim, hidden_dims=(32, 32, 1)):
        super().__init__()
        self.deep_network = MLP(dims=[input_dim] + list(hidden_dims))
        self.bias = nn.Parameter(torch.zeros(1))


class DeepFM(BaseModel):
    def __init__(self, config_path):
        super().__init__(config_path)
        cfg = self.config.get("deepfm_cfg", None)
        names = self.user_feature_names | self.item_feature_names
        self.fm_feature_names = set(cfg.fm_feature_names) if cfg and cfg.get("fm_feature_names") else set(names)
        self.score_fc = DeepFMModel(input_dim=self.user_input_dim + self.item_input_dim, hidden_dims=[128, 128, 128, 64, 1])

    def get_inp_embedding(self, batch):
        features, _, _ = self.get_embeddings_from_batch(batch, self.user_feature_names | self.item_feature_names)
        return features

    def fm_fields(self, dims, fnames):
        cols, fdims, s = [], [], 0
        for d, n in zip(dims, fnames):
            if n in self.fm_feature_names:
                cols.append(s)
                fdims.append(d)
            s += d
        return cols, fdims

    def forward(self, x):
        features, dims, fnames = self.get_embeddings_from_batch(x, self.user_feature_names | self.item_feature_names)
        cols, fdims = self.fm_fields(dims, fnames)
        fm = ops.FieldLogitFn.apply(features, cols, fdims, ops.L.FIELD_FM)
        deep = self.score_fc.deep_network(features)
        return ops.SigmoidFn.apply(self.score_fc.bias, fm, deep.view(-1)).view(-1, 1)
