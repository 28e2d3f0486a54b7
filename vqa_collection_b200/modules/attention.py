"""Attention blocks — mirror of modules/attention.py on the C-ABI kernels.

MultiplyAttention ('new', attention.py:54-86): the W_v projection, the ⊙ with the W_q
projection and the 1-wide logit layer run as ONE tcgen05 GEMM with a row-reduction
epilogue (the [B,K,H] projection never reaches HBM); softmax over the K regions is the
pooling kernel.
"""
import torch
import torch.nn as nn

from .. import compute_dtype, ops
from .modules import FCNet, PreparedCache, as_compute, wn_linear, _no_training
from ..engine import weight_norm_scale


def set_att(att_type):
    return {
        'base': ConcatAttention,
        'new': MultiplyAttention
    }[att_type]


class MultiplyAttention(nn.Module):
    """softmax_K( w · (ReLU(W_v v) ⊙ ReLU(W_q q)) + b )   (attention.py:54-86)"""

    def __init__(self, v_dim, q_dim, hidden_dim, dropout=0.2):
        super().__init__()
        self.W_v = FCNet(v_dim, hidden_dim)
        self.W_q = FCNet(q_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout)              # identity in eval (attention.py:74)
        self.linear = wn_linear(q_dim, 1)               # uses q_dim like the reference (:66)
        self._cache = PreparedCache()

    def _logit_vector(self):
        lin = self.linear
        return self._cache.get("wlin", (lin.weight_v, lin.weight_g), lambda: (
            lin.weight_v.detach().float().reshape(-1) * weight_norm_scale(lin.weight_v, lin.weight_g)).contiguous())

    def logit_parts(self, v, q):
        """→ (parts f32 [B*K, n_parts], x in compute dtype [B,K,V])"""
        _no_training(self)
        dtype = compute_dtype()
        B, K, V = v.shape
        x = as_compute(v, dtype)
        qp = self.W_q(q, out_dtype=torch.float32)                            # [B,H] f32
        (W, s, b), = self.W_v.prepared(dtype)
        parts = ops.linear(x.view(B * K, V), W, s, b, relu=True, mul=qp.contiguous(), mul_row_div=K,
                           logit_w=self._logit_vector())
        return parts, x

    def logits(self, v, q):
        parts, _ = self.logit_parts(v, q)
        B, K = v.shape[0], v.shape[1]
        return (parts.sum(1) + self.linear.bias.detach().float()).view(B, K, 1)

    # -- caption decoder: many queries against the SAME regions (generator.py:172 inside the time loop) --------
    def project(self, v):
        """region half of the logits, once per caption batch: ReLU(W_v v + b) → ([B*K,H] compute dtype, x)"""
        _no_training(self)
        dtype = compute_dtype()
        B, K, V = v.shape
        x = as_compute(v, dtype)
        (W, s, b), = self.W_v.prepared(dtype)
        return ops.linear(x.view(B * K, V), W, s, b, relu=True, out_dtype=dtype), x

    def step_parts(self, proj, q, K):
        """logit parts [b*K,1] of one step for the first b = q.shape[0] samples of ``proj``"""
        qp = self.W_q(q, out_dtype=torch.float32)
        return ops.attention_logits(proj, qp.contiguous(), self._logit_vector(), K, mode=0)

    def step_weights(self, dtype):
        """(mode, W_q, scale, bias, logit vector) of the hidden-state half, for vqa_caption_decode_steps"""
        (W, s, b), = self.W_q.prepared(dtype)
        return 0, W, s, b, self._logit_vector()

    def forward(self, v, q):
        """v [batch, num_objs, v_dim], q [batch, q_dim] → [batch, num_objs, 1] (f32)"""
        parts, x = self.logit_parts(v, q)
        att, _, _ = ops.attention_pool(parts, float(self.linear.bias.detach()), x, True, False, False)
        return att.unsqueeze(2)


class ConcatAttention(nn.Module):
    """softmax_K( w2 · ReLU(W1 [v;q] + b1) + b2 )   (attention.py:18-51, att_type='base').

    W1 [v;q] = W1[:, :V] v + W1[:, V:] q: the q-half is ONE [B,H] GEMM (not K copies of q);
    it enters the v-half GEMM as an additive row-broadcast epilogue operand, and the 1-wide
    second layer is that GEMM's row-reduction epilogue — like MultiplyAttention, the
    [B,K,H] hidden layer never reaches HBM.  Parameters keep the reference's names."""

    def __init__(self, v_dim, q_dim, hidden_dim):
        super().__init__()
        self.sequence = nn.Sequential(wn_linear(v_dim + q_dim, hidden_dim), nn.ReLU(), wn_linear(hidden_dim, 1))
        self.v_dim = v_dim
        self._cache = PreparedCache()

    @property
    def linear(self):                                   # the logit layer, same role as MultiplyAttention.linear
        return self.sequence[2]

    def _prepared(self, dtype):
        l0, l2 = self.sequence[0], self.sequence[2]

        def build():
            V = self.v_dim
            v = l0.weight_v.detach()
            s1 = weight_norm_scale(l0.weight_v, l0.weight_g)
            H = v.shape[0]
            sv = torch.full((H,), s1, dtype=torch.float32, device=v.device)
            wlin = (l2.weight_v.detach().float().reshape(-1) * weight_norm_scale(l2.weight_v, l2.weight_g)).contiguous()
            return (v[:, :V].to(dtype).contiguous(), v[:, V:].to(dtype).contiguous(), sv,
                    l0.bias.detach().float().contiguous(), wlin)
        return self._cache.get(("concat", dtype), (l0.weight_v, l0.weight_g, l0.bias, l2.weight_v, l2.weight_g), build)

    def logit_parts(self, v, q):
        """→ (parts f32 [B*K, n_parts], x in compute dtype [B,K,V])"""
        _no_training(self)
        dtype = compute_dtype()
        B, K, V = v.shape
        x = as_compute(v, dtype)
        W1v, W1q, sv, b1, wlin = self._prepared(dtype)
        qadd = ops.linear(as_compute(q, dtype), W1q, sv, b1, relu=False, out_dtype=torch.float32)     # [B,H]
        parts = ops.linear(x.view(B * K, V), W1v, sv, None, relu=True, add=qadd, add_row_div=K, logit_w=wlin)
        return parts, x

    def logits(self, v, q):
        parts, _ = self.logit_parts(v, q)
        B, K = v.shape[0], v.shape[1]
        return (parts.sum(1) + self.linear.bias.detach().float()).view(B, K, 1)

    # -- caption decoder: many queries against the SAME regions (generator.py:172 inside the time loop) --------
    def project(self, v):
        """region half of the hidden layer, once per caption batch: W1[:, :V] v (pre-activation) → ([B*K,H], x)"""
        _no_training(self)
        dtype = compute_dtype()
        B, K, V = v.shape
        x = as_compute(v, dtype)
        W1v, _, sv, _, _ = self._prepared(dtype)
        return ops.linear(x.view(B * K, V), W1v, sv, None, relu=False, out_dtype=dtype), x

    def step_parts(self, proj, q, K):
        """logit parts [b*K,1] of one step for the first b = q.shape[0] samples of ``proj``"""
        dtype = compute_dtype()
        _, W1q, sv, b1, wlin = self._prepared(dtype)
        qadd = ops.linear(as_compute(q, dtype), W1q, sv, b1, relu=False, out_dtype=torch.float32)
        return ops.attention_logits(proj, qadd, wlin, K, mode=1)

    def step_weights(self, dtype):
        """(mode, W1[:, V:], scale, bias, logit vector) of the hidden-state half, for vqa_caption_decode_steps"""
        _, W1q, sv, b1, wlin = self._prepared(dtype)
        return 1, W1q, sv, b1, wlin

    def forward(self, v, q):
        """v [batch, num_objs, v_dim], q [batch, q_dim] → [batch, num_objs, 1] (f32)"""
        parts, x = self.logit_parts(v, q)
        att, _, _ = ops.attention_pool(parts, float(self.linear.bias.detach()), x, True, False, False)
        return att.unsqueeze(2)
