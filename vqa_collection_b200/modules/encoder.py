"""Encoders — mirror of modules/encoder.py (BaseEncoder :96-183, RelationEncoder :186-272)."""
import torch
import torch.nn as nn

from .. import compute_dtype, ops
from .attention import set_att
from .gcn import GCN
from .modules import FCNet, SentenceEmbedding, PretrainedWordEmbedding, PreparedCache, as_compute, _no_training


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
    """BaseEncoder + spatial-relation GCN over batch['graph']   (encoder.py:186-272).
    Extension: when batch has 'bbox' (+'wh') and no 'graph', the labels are computed on device."""

    def __init__(self, ntoken: int, embed_dim: int, hidden_dim: int, rnn_layer: int, v_dim: int, device: str,
                 dropout: float = 0.5, rnn_type: str = 'GRU', att_type: str = 'base', conv_layer: int = 1,
                 conv_type: str = 'corr', use_imp: bool = False, use_spa: bool = True, use_sem: bool = False,
                 num_objs: int = 36):
        super().__init__(ntoken, embed_dim, hidden_dim, rnn_layer, v_dim, device, dropout, rnn_type, att_type)
        assert use_imp or use_spa or use_sem, 'Should use at least one relation'
        if use_imp or use_sem:
            raise NotImplementedError("only the spatial relation branch is built (reference default, encoder.py:202-204)")
        self.implicit_encoder = None
        self.spatial_encoder = GCN(in_dim=v_dim, out_dim=v_dim, num_labels=12, device=device, conv_layer=conv_layer,
                                   conv_type=conv_type, dropout=dropout)

    def graph_labels(self, batch):
        if 'graph' in batch:
            g = batch['graph'].to(self.device)
            return g if g.dtype == torch.uint8 else g.to(torch.uint8)        # loader gives float64 (dataset.py:102)
        bbox = batch['bbox'].to(self.device).float().contiguous()
        w, h = batch['wh']
        return ops.relation_labels(bbox, float(w), float(h))

    def forward(self, batch, graph_alpha=False):
        x, q_emb, att, _, _ = self._attend(batch, False, False)
        labels = self.graph_labels(batch)
        new_v, vsum, alphas = self.spatial_encoder(x, labels, graph_alpha, att=att, want_vsum=True)
        if graph_alpha:
            return alphas
        out = {'v': new_v, 'q': self.q_net(q_emb, out_dtype=torch.float32), 'v_att': att.unsqueeze(2), 'v_sum': vsum}
        if 'c' in batch:
            c_target = batch['c'].to(self.device)
            out['c'] = torch.nn.functional.embedding(c_target, self.embedding.weight.detach())
            out['c_target'] = c_target
        if 'cap_len' in batch:
            out['cap_len'] = batch['cap_len'].to(self.device)
        return out
