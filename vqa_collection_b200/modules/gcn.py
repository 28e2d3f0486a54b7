"""Graph convolution — mirror of modules/gcn.py on the C-ABI kernels.

Only conv_type='corr' is runnable in the reference (GCN.forward passes 3 arguments,
gcn.py:205-210), so only CorrelatedGraphConv is built.  Like the reference,
``GCN.gcn`` is a plain Python list: the layer tensors are NOT registered parameters
(gcn.py:188-190, F3) so state_dict keys stay identical to old checkpoints.
"""
import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from .. import compute_dtype, ops
from .modules import DotProduct, PreparedCache, as_compute, _no_training
from ..engine import prepare_gcn_layer


def get_graph_conv(conv_type):
    return {
        'base': BaseGraphConv,
        'direct': DirectedGraphConv,
        'corr': CorrelatedGraphConv
    }[conv_type]


class BaseGraphConv(nn.Module):
    def __init__(self, in_dim, out_dim, num_labels, bias=True):
        raise NotImplementedError("conv_type='base' cannot run in the reference either (gcn.py:38 vs :205); "
                                  "only 'corr' is built")


class DirectedGraphConv(nn.Module):
    """parameter container of gcn.py:54-76 (weight.{0,1,2}: Linear(bias=False); bias [L,out])"""

    def __init__(self, in_dim, out_dim, num_labels, dir_num=3):
        super().__init__()
        self.out_dim = out_dim
        self.dir_num = dir_num
        self.weight = nn.ModuleList([nn.Linear(in_dim, out_dim, bias=False) for _ in range(dir_num)])
        self.bias = Parameter(torch.FloatTensor(num_labels, out_dim))
        stdv = 1. / math.sqrt(self.out_dim)
        self.bias.data.uniform_(-stdv, stdv)

    def forward(self, feature, graph):
        raise NotImplementedError("conv_type='direct' cannot run in the reference either (gcn.py:109 vs :205)")


class CorrelatedGraphConv(DirectedGraphConv):
    """gcn.py:112-168.  forward = one wide tcgen05 GEMM x·[W0+W1;W2;Wa;Wb]ᵀ + the
    relation-masked graph-attention kernel."""

    def __init__(self, in_dim, out_dim, num_labels, dir_num=3):
        super().__init__(in_dim, out_dim, num_labels, dir_num)
        assert dir_num == 3 and in_dim == out_dim
        self.dot_product = DotProduct(in_dim, in_dim, out_dim)
        self.softmax = nn.Softmax(dim=1)
        self._cache = PreparedCache()
        self._K = 36

    def prepared(self, dtype):
        params = list(self.parameters())
        return self._cache.get(("gcn", dtype, self._K), params, lambda: prepare_gcn_layer(
            {k: v for k, v in self.state_dict().items()}, dtype, self.bias.device, self._K))

    def run(self, x, graph, att=None, want_out=True, want_vsum=False, want_alpha=False):
        """x: RAW features [B,K,V]; att f32 [B,K] or None (feature = att ⊙ x).
        Returns (ReLU(layer output) or None, Σ_K or None, α or None) — the ReLU of
        GCN.forward (gcn.py:212) is fused."""
        dtype = compute_dtype()
        B, K, V = x.shape
        if K != self._K:
            self._K = K
        P = self.prepared(dtype)
        labels = graph if graph.dtype == torch.uint8 else graph.to(torch.uint8)
        xc = as_compute(x, dtype).view(B * K, V)
        if "Wg3" in P:                                   # bf16: merged algebra + tcgen05 graph attention
            Y = ops.linear(xc, P["Wg3"])
            return ops.graph_attention_merged(Y, xc, att, labels.contiguous(), P["wvec"], P["gat_c0"],
                                              P["label_bias_lp"], P["num_labels"], K, want_out, want_vsum, want_alpha)
        Y = ops.linear(xc, P["Wg"])
        return ops.graph_attention(Y, att, labels.contiguous(), P["label_bias"], P["ba"], P["bb"], K,
                                   want_out, want_vsum, want_alpha)

    def forward(self, feature, graph, get_alpha):
        raise NotImplementedError("call GCN.forward: the layer kernel fuses the GCN's ReLU (gcn.py:211-212)")


class GCN(nn.Module):
    """Relation encoder (gcn.py:171-215): conv layers + dropout (identity in eval) + ReLU."""

    def __init__(self, in_dim: int, out_dim: int, num_labels: int, device: str, conv_layer: int = 1,
                 conv_type: str = 'corr', dropout: float = 0.5):
        super().__init__()
        GraphConv = get_graph_conv(conv_type)
        self.gcn = [GraphConv(in_dim, out_dim, num_labels).to(device)]
        for _ in range(conv_layer - 1):
            self.gcn.append(GraphConv(out_dim, out_dim, num_labels).to(device))
        self.dropout = nn.Dropout(dropout)

    def __repr__(self):
        return '\n'.join(layer.__repr__() for layer in self.gcn)

    def forward(self, feature, graph, get_alpha, att=None, want_vsum=False):
        """feature [batch, num_objs, in_dim], graph [batch, num_objs, num_objs] (labels 0..11).
        ``att`` (extension): f32 [B,K] — `feature` is then the RAW x and the layer input is
        att ⊙ x, applied inside the kernel."""
        _no_training(self)
        alphas, vsum = [], None
        for i, layer in enumerate(self.gcn):
            last = i == len(self.gcn) - 1
            feature, vsum, alpha = layer.run(feature, graph, att if i == 0 else None, True,
                                             want_vsum and last, get_alpha)
            alphas.append(alpha)
        if want_vsum:
            return feature, vsum, alphas
        if get_alpha:
            return feature, alphas
        return feature
