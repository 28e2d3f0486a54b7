"""Encoders — mirror of modules/encoder.py (BaseEncoder :96-183, RelationEncoder :186-272)."""
import torch
import torch.nn as nn

from .. import ops
from .attention import set_att
from .gcn import GCN
from .modules import FCNet, SentenceEmbedding, PretrainedWordEmbedding, PreparedCache, _no_training


def set_encoder(encoder_type: str, ntoken: int, v_dim: int, embed_dim: int, hidden_dim: int, device: str,
                dropout: float, rnn_type: str, rnn_layer: int, att_type: str, conv_type: str, conv_layer: int,
                vocab_path: str = ''):
    if encoder_type == 'base':
        model = BaseEncoder(ntoken=ntoken, v_dim=v_dim, embed_dim=embed_dim, hidden_dim=hidden_dim, device=device,
                            dropout=dropout, rnn_type=rnn_type, rnn_layer=rnn_layer, att_type=att_type)
    elif encoder_type == 'relation':
        model = RelationEncoder(ntoken=ntoken, v_dim=v_dim, embed_dim=embed_dim, hidden_dim=hidden_dim,
                                device=device, dropout=dropout, rnn_type=rnn_type, rnn_layer=rnn_layer,
                                att_type=att_type, conv_type=conv_type, conv_layer=conv_layer)
    else:
        raise NotImplementedError(f"encoder_type='{encoder_type}' is outside the accelerated VQA forward path")
    if vocab_path != '':
        model.embedding = PretrainedWordEmbedding(vocab_path=vocab_path, device=device)      # encoder.py:56-57
    return model.to(device)


class BaseEncoder(nn.Module):
    """embedding → GRU → top-down attention → v_att·v, q_net   (encoder.py:96-183)"""

    def __init__(self, ntoken: int, embed_dim: int, hidden_dim: int, rnn_layer: int, v_dim: int, device: str,
                 dropout: float = 0.5, rnn_type: str = 'GRU', att_type: str = 'base'):
        super().__init__()
        self.device = device
        self.embedding = nn.Embedding(ntoken + 1, embed_dim, padding_idx=ntoken)
        self.q_rnn = SentenceEmbedding(in_dim=embed_dim, hidden_dim=hidden_dim, rnn_layer=rnn_layer, dropout=0.0,
                                       device=device, rnn_type=rnn_type)
        self.attention = set_att(att_type)(v_dim=v_dim, q_dim=hidden_dim, hidden_dim=hidden_dim)
        self.q_net = FCNet(hidden_dim, hidden_dim)
        self._cache = PreparedCache()

    def _attend(self, batch, want_v, want_vsum):
        """shared front: returns (x compute-dtype [B,K,V], q_emb f32, att f32 [B,K], vsum, vatt)"""
        _no_training(self)
        v = batch['img'].to(self.device)
        q_tok = batch['q'].to(self.device)
        q_emb = self.q_rnn.forward_tokens(q_tok, self.embedding.weight)           # [B,H] f32
        parts, x = self.attention.logit_parts(v, q_emb)
        att, vsum, vatt = ops.attention_pool(parts, float(self.attention.linear.bias.detach()), x,
                                             True, want_vsum, want_v)
        return x, q_emb, att, vsum, vatt

    def base_forward(self, batch, _want_v=True):
        x, q_emb, att, vsum, vatt = self._attend(batch, _want_v, True)
        out = {
            'v': vatt,                                   # [batch, num_objs, v_dim]  (v_att * v)
            'q': self.q_net(q_emb, out_dtype=torch.float32),      # [batch, hidden_dim]
            'v_att': att.unsqueeze(2),                   # [batch, num_objs, 1]
            'v_sum': vsum,                               # extension: Σ_K v for the predictor
        }
        # caption keys are passed through when present (encoder.py:155-156,172); the reference
        # raises KeyError without them, this mirror only needs them for the (out-of-scope) decoder
        if 'c' in batch:
            c_target = batch['c'].to(self.device)
            out['c'] = torch.nn.functional.embedding(c_target, self.embedding.weight.detach())
            out['c_target'] = c_target
        if 'cap_len' in batch:
            out['cap_len'] = batch['cap_len'].to(self.device)
        return out

    def forward(self, batch):
        return self.base_forward(batch)


class RelationEncoder(BaseEncoder):
    """BaseEncoder + relation GCNs over the attended regions (encoder.py:186-272): the spatial branch over
    batch['graph'] (default) and/or the implicit branch over the fully connected graph (``use_imp``; labels 1 off the
    diagonal, 0 on it — encoder.py:231-234,255).  Their outputs are summed (encoder.py:257,264).
    Extension: when batch has 'bbox' (+'wh') and no 'graph', the labels are computed on device."""

    def __init__(self, ntoken: int, embed_dim: int, hidden_dim: int, rnn_layer: int, v_dim: int, device: str,
                 dropout: float = 0.5, rnn_type: str = 'GRU', att_type: str = 'base', conv_layer: int = 1,
                 conv_type: str = 'corr', use_imp: bool = False, use_spa: bool = True, use_sem: bool = False,
                 num_objs: int = 36):
        super().__init__(ntoken, embed_dim, hidden_dim, rnn_layer, v_dim, device, dropout, rnn_type, att_type)
        assert use_imp or use_spa or use_sem, 'Should use at least one relation'

        def make():
            return GCN(in_dim=v_dim, out_dim=v_dim, num_labels=12, device=device, conv_layer=conv_layer,
                       conv_type=conv_type, dropout=dropout)
        self.implicit_encoder = make() if use_imp else None
        self.spatial_encoder = make() if use_spa else None
        self.num_objs = num_objs
        self._imp_labels = None

    def graph_labels(self, batch):
        if 'graph' in batch:
            g = batch['graph'].to(self.device)
            return g if g.dtype == torch.uint8 else g.to(torch.uint8)        # loader gives float64 (dataset.py:102)
        bbox = batch['bbox'].to(self.device).float().contiguous()
        w, h = batch['wh']
        return ops.relation_labels(bbox, float(w), float(h))

    def implicit_labels(self, B, K, device):
        """the fully connected graph of encoder.py:231-234 as labels [B,K,K] u8 (1 off the diagonal)"""
        if self._imp_labels is None or self._imp_labels.shape[:2] != (B, K) or self._imp_labels.device != device:
            g = (torch.ones(K, K) - torch.eye(K)).to(torch.uint8)
            self._imp_labels = g.unsqueeze(0).repeat(B, 1, 1).contiguous().to(device)
        return self._imp_labels

    def forward(self, batch, graph_alpha=False):
        x, q_emb, att, _, _ = self._attend(batch, False, False)
        B, K, V = x.shape
        new_v = vsum = alphas = None
        branches = []
        if self.implicit_encoder is not None:
            branches.append((self.implicit_encoder, self.implicit_labels(B, K, x.device)))
        if self.spatial_encoder is not None:
            branches.append((self.spatial_encoder, self.graph_labels(batch)))
        for gcn, labels in branches:
            v_b, vsum_b, alphas = gcn(x, labels, graph_alpha, att=att, want_vsum=True)   # like the reference, the LAST branch's α
            if new_v is None:
                new_v, vsum = v_b, vsum_b
            else:
                ops.add_(new_v, v_b)
                ops.add_(vsum, vsum_b)
        if graph_alpha:
            return alphas if alphas is not None else []
        if new_v is None:                              # use_sem only: no encoder exists for it (encoder.py:247 zeros)
            new_v = torch.zeros((B, K, V), dtype=x.dtype, device=x.device)
            vsum = torch.zeros((B, V), dtype=x.dtype, device=x.device)
        out = {'v': new_v, 'q': self.q_net(q_emb, out_dtype=torch.float32), 'v_att': att.unsqueeze(2), 'v_sum': vsum}
        if 'c' in batch:
            c_target = batch['c'].to(self.device)
            out['c'] = torch.nn.functional.embedding(c_target, self.embedding.weight.detach())
            out['c_target'] = c_target
        if 'cap_len' in batch:
            out['cap_len'] = batch['cap_len'].to(self.device)
        return out
