"""DCN — mirror of reference src/model/sort/dcn/model.py: `DCNModel` (:15-29: 3 DCN-v1 cross layers, then
MLP([2d,128,128,128,64,1]) on cat[x, cross]) and `DCN` (:31-48).  K1 -> K5 (cross + concat) -> K4 (tower)."""
import torch.nn as nn

from .... import ops
from ...BaseModel.base_model import BaseModel
from ...model_utils.utils import MLP
from .dcn_arch import DCNNet


class DCNModel(nn.Module):
    def __init__(self, input_dim, cross_num_layers=3, deep_hidden_dims=(32, 32, 1)):
        super().__init__()
        self.cross_net = DCNNet(input_dim=input_dim, num_layers=cross_num_layers)
        self.score_fc = MLP(dims=[input_dim * 2] + list(deep_hidden_dims))

    def forward(self, x):
        return ops.SigmoidFn.apply(None, self.score_fc(self.cross_net.cat_forward(x))).view(-1, 1)


class DCN(BaseModel):
    def __init__(self, config_path):
        super().__init__(config_path)
        self.score_fc = DCNModel(input_dim=self.user_input_dim + self.item_input_dim, cross_num_layers=3,
                                 deep_hidden_dims=[128, 128, 128, 64, 1])

    def get_inp_embedding(self, batch):
        features, _, _ = self.get_embeddings_from_batch(batch, self.user_feature_names | self.item_feature_names)
        return features

    def forward(self, x):
        return self.score_fc(self.get_inp_embedding(x))
