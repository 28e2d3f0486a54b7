"""Predictors — mirror of modules/predictor.py (BasePredictor :54-93) on the C-ABI kernels."""
import torch
import torch.nn as nn

from .. import compute_dtype, ops
from .modules import FCNet, as_compute, _no_training


def set_predictor(predictor_type: str, v_dim: int, embed_dim: int, hidden_dim: int, ans_dim: int, device: str,
                  cls_layer: int, dropout: float, c_len: int, neg_slope: float):
    if predictor_type == 'base':
        return BasePredictor(v_dim=v_dim, hidden_dim=hidden_dim, ans_dim=ans_dim, device=device,
                             cls_layer=cls_layer, dropout=dropout).to(device)
    if predictor_type in ('base-cap', 'q-cap'):
        raise NotImplementedError(f"predictor_type='{predictor_type}' is outside the accelerated VQA forward path "
                                  "(q-cap is broken in the reference, SURVEY.md F8)")
    return None            # like the reference: unknown types (e.g. 'none') give no predictor


class BasePredictor(nn.Module):
    """Σ_K v → v_net → ⊙ q → classifier FCNet(H→2H→A)   (predictor.py:54-93)"""

    def __init__(self, v_dim: int, hidden_dim: int, ans_dim: int, device: str, cls_layer: int = 2,
                 dropout: float = 0.5):
        super().__init__()
        self.device = device
        self.v_net = FCNet(v_dim, hidden_dim)
        self.classifier = FCNet(in_dim=hidden_dim, mid_dim=2 * hidden_dim, out_dim=ans_dim, layer=cls_layer,
                                dropout=dropout)

    def forward(self, batch):
        _no_training(self)
        q = batch['q'].to(self.device)
        if 'v_sum' in batch:                       # encoder already produced Σ_K (fused pooling kernels)
            v = batch['v_sum'].to(self.device)
        else:
            v = batch['v'].to(self.device)
            dtype = compute_dtype()
            v = as_compute(v, dtype)
            B, K, V = v.shape
            ones = torch.zeros((B * K, 1), dtype=torch.float32, device=v.device)   # softmax(0)·K = 1
            _, vs, _ = ops.attention_pool(ones, 0.0, v, False, True, False)         # (1/K) Σ_K v
            v = vs.float() * K
        joint = self.v_net(v, mul=q.float().contiguous())        # ReLU(W v) ⊙ q   (predictor.py:88-91)
        return self.classifier(joint, out_dtype=torch.float32)
