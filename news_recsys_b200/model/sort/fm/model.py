"""FM — mirror of reference src/model/sort/fm/model.py: `FMModel` (:12-26, a single bias
parameter), `FM.get_inp_embedding` (:48-59, w = column 0, v = columns 1.. of every field) and
`FM.forward` (:43-45).  When every field is a plain sparse feature of equal width (the shipped
train_cf_fm.yaml) the gather, the FM sum and the sigmoid run as ONE kernel (nrx_fm_fused_fwd);
otherwise K1 -> nrx_field_logit (FM) -> sigmoid."""
import torch
import torch.nn as nn

from .... import ops
from ...BaseModel.base_model import BaseModel


class FMModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(1))

    def forward(self, w, v):
        """Reference signature (:18-26) for callers that hold w [B,F] and v [B,F,D-1] already."""
        B, F = w.shape
        x = torch.cat([w.unsqueeze(-1), v], dim=2).reshape(B, -1).contiguous()
        D = v.shape[2] + 1
        logit = ops.FieldLogitFn.apply(x, [f * D for f in range(F)], [D] * F, ops.L.FIELD_FM)
        return ops.SigmoidFn.apply(self.bias, logit).view(-1, 1)


class FM(BaseModel):
    def __init__(self, config_path):
        super().__init__(config_path)
        self.score_fc = FMModel()

    def get_inp_embedding(self, batch):
        features, dims, _ = self.get_embeddings_from_batch(batch, self.user_feature_names | self.item_feature_names)
        w, v, s = [], [], 0
        for d in dims:
            w.append(features[:, s:s + 1])
            v.append(features[:, s + 1:s + d])
            s += d
        return torch.cat(w, dim=1), torch.stack(v, dim=1)

    def forward(self, x):
        names = self.user_feature_names | self.item_feature_names
        fb, dims, _, out_dim = self.bind_features(x, names)
        if fb is None:
            raise ValueError("FM.forward: none of the model's features is present in the batch")
        fb.status = self._id_status(fb.device)   # out-of-table ids raise at the next check_ids() (nn.Embedding would raise here)
        tnames = list(self.embedding_tables.keys())
        ws = [self.embedding_tables[t].weight for t in tnames]
        if ops.fm_fused_eligible(fb.specs, self._weights()):
            if torch.is_grad_enabled():
                return ops.FmFusedFn.apply(fb, out_dim, tnames, self.score_fc.bias, *ws)
            prob, _, _, _ = ops.fm_fused_fwd(fb, self.score_fc.bias, status=fb.status)
            return prob.view(-1, 1)
        if len(set(dims)) != 1:
            raise RuntimeError("stack expects each tensor to be equal size (FM needs equal field widths, fm/model.py:58)")
        feats = ops.EmbedPoolFn.apply(fb, out_dim, tnames, *ws) if torch.is_grad_enabled() else ops.embed_pool_fwd(fb, out_dim, status=fb.status)
        cols, c = [], 0
        for d in dims:
            cols.append(c)
            c += d
        logit = ops.FieldLogitFn.apply(feats, cols, dims, ops.L.FIELD_FM)
        return ops.SigmoidFn.apply(self.score_fc.bias, logit).view(-1, 1)
