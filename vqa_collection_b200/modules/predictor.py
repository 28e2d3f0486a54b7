"""Predictors — mirror of modules/predictor.py (BasePredictor :54-93, BaseCaptionPredictor :95-140,
PredictorwithCaption :144-213) on the C-ABI kernels."""
import torch
import torch.nn as nn

from .. import compute_dtype, ops
from .modules import FCNet, LReLUNet, CaptionEmbedding, SentenceEmbedding, as_compute, _no_training


def set_predictor(predictor_type: str, v_dim: int, embed_dim: int, hidden_dim: int, ans_dim: int, device: str,
                  cls_layer: int, dropout: float, c_len: int, neg_slope: float):
    if predictor_type == 'base':
        return BasePredictor(v_dim=v_dim, hidden_dim=hidden_dim, ans_dim=ans_dim, device=device,
                             cls_layer=cls_layer, dropout=dropout).to(device)
    if predictor_type == 'q-cap':
        return PredictorwithCaption(embed_dim=embed_dim, c_len=c_len, v_dim=v_dim, hidden_dim=hidden_dim,
                                    ans_dim=ans_dim, device=device, cls_layer=cls_layer, dropout=dropout,
                                    neg_slope=neg_slope).to(device)
    if predictor_type == 'base-cap':
        return BaseCaptionPredictor(v_dim=v_dim, embed_dim=embed_dim, hidden_dim=hidden_dim, ans_dim=ans_dim,
                                    device=device, cls_layer=cls_layer, dropout=dropout).to(device)
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

    def _pooled(self, batch):
        """Σ_K v (predictor.py:85)"""
        if 'v_sum' in batch:                       # encoder already produced Σ_K (fused pooling kernels)
            return batch['v_sum'].to(self.device)
        v = batch['v'].to(self.device)
        dtype = compute_dtype()
        v = as_compute(v, dtype)
        B, K, V = v.shape
        ones = torch.zeros((B * K, 1), dtype=torch.float32, device=v.device)   # softmax(0)·K = 1
        _, vs, _ = ops.attention_pool(ones, 0.0, v, False, True, False)         # (1/K) Σ_K v
        return vs.float() * K

    def forward(self, batch):
        _no_training(self)
        q = batch['q'].to(self.device)
        joint = self.v_net(self._pooled(batch), mul=q.float().contiguous())        # ReLU(W v) ⊙ q   (predictor.py:88-91)
        return self.classifier(joint, out_dtype=torch.float32)


class BaseCaptionPredictor(BasePredictor):
    """BasePredictor + a caption embedding added to the visual one: joint = q ⊙ (c_net(GRU_last(c)) + v_net(Σ_K v))
    (predictor.py:95-140).  The add and the ⊙ q are v_net's epilogue (add-after-activation, then mul)."""

    def __init__(self, v_dim: int, embed_dim: int, hidden_dim: int, ans_dim: int, device: str, cls_layer: int = 2,
                 dropout: float = 0.5):
        super().__init__(v_dim, hidden_dim, ans_dim, device, cls_layer, dropout)
        self.c_rnn = SentenceEmbedding(in_dim=embed_dim, hidden_dim=hidden_dim, rnn_layer=1, device=device, rnn_type='GRU')
        self.c_net = FCNet(hidden_dim, hidden_dim, dropout=dropout)

    def forward(self, batch):
        _no_training(self)
        q = batch['q'].to(self.device)
        c = self.c_net(self.c_rnn(batch['c'].to(self.device)), out_dtype=torch.float32)       # [B,H] f32
        v = self._pooled(batch)
        joint = self.v_net(v, mul=q.float().contiguous(), add=c.contiguous(), add_after_act=True)   # q ⊙ (ReLU(W v) + c)
        return self.classifier(joint, out_dtype=torch.float32)


class PredictorwithCaption(nn.Module):
    """'Generating Question Relevant Captions to Aid VQA' predictor (predictor.py:144-213, BASELINE config 5).

    Every stage is a bias-free Linear + LeakyReLU (LReLUNet) on the tcgen05 GEMM with the element-wise neighbours
    folded into its epilogue: ``c ⊙ vq`` is the ``mul`` operand of vq_net, ``q ⊙ (v + c)`` the ``add``-after-
    activation + ``mul`` operands of vqc_net, the final Sigmoid the classifier's epilogue.  Line 203's
    ``(joint.repeat ⊙ V).sum(1)`` equals ``joint ⊙ Σ_K V`` and is computed on the pooled vector."""

    def __init__(self, embed_dim: int, c_len: int, v_dim: int, hidden_dim: int, ans_dim: int, device: str,
                 cls_layer: int = 2, dropout: float = 0.5, neg_slope: float = 0.01):
        super().__init__()
        self.device = device
        self.v_net = LReLUNet(v_dim, hidden_dim, neg_slope)
        self.caption_embedding = CaptionEmbedding(v_dim=hidden_dim, q_dim=hidden_dim, c_dim=embed_dim,
                                                  hidden_dim=hidden_dim, max_len=c_len, device=device, dropout=dropout)
        self.c_net = LReLUNet(hidden_dim, hidden_dim, neg_slope)
        self.vq_net = LReLUNet(hidden_dim, hidden_dim, neg_slope)
        self.joint_net = LReLUNet(hidden_dim, hidden_dim, neg_slope)
        self.vqc_net = LReLUNet(hidden_dim, hidden_dim, neg_slope)
        self.classifier = nn.Sequential(LReLUNet(hidden_dim, ans_dim, neg_slope), nn.Sigmoid())

    def forward(self, batch):
        _no_training(self)
        for i in batch:
            if torch.is_tensor(batch[i]):
                batch[i] = batch[i].to(self.device)
        dtype = compute_dtype()
        q = batch['q'].float().contiguous()
        batch['v'] = self.v_net(batch['v'])                                   # [B,K,H]  (predictor.py:188)
        V = batch['v']
        B, K, H = V.shape
        zeros = torch.zeros((B * K, 1), dtype=torch.float32, device=V.device)       # softmax(0) = 1/K
        _, vmean, _ = ops.attention_pool(zeros, 0.0, V.contiguous(), False, True, False)
        v = as_compute((vmean.float() * K).contiguous(), dtype)               # Σ_K V  (:191)
        c = self.caption_embedding(v, q, batch['c'])                          # [B,H]  (:192)
        c = self.c_net(c, out_dtype=torch.float32)                            # (:197)
        self.c_grad = c
        cv = self.vq_net(v, mul=c)                                            # LReLU(W_vq v) ⊙ c  (:196,201)
        z = self.joint_net(cv, out_dtype=torch.float32)                       # (:201)
        v2 = ops.softmax_mul(z.contiguous(), v)                               # softmax_H ⊙ Σ_K V  (:202-203)
        joint = self.vqc_net(v2, add=c, add_after_act=True, mul=q)            # q ⊙ (LReLU(W v) + c)  (:208-209)
        self.logit_grad = joint
        return self.classifier[0](joint, out_dtype=torch.float32, sigmoid=True)      # (:213)
