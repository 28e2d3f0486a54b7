"""Caption head — mirror of modules/generator.py (set_decoder :12-37, DecoderModule :40-120,
BaseDecoder :123-181) on the C-ABI kernels.

Only what the reference can run is built: decoder_type='base' (nn.GRUCell or nn.LSTMCell).
``BUTDDecoder.decode`` has no return statement in the reference (generator.py:240-266), so
``DecoderModule.forward`` raises a TypeError on it; decoder_type='butd' raises here as well.

The time loop of the teacher-forced forward is restructured around what does NOT depend on the
hidden state:
  * the region half of the decoder's attention (ReLU(W_v v), or W1[:, :V] v for att_type='base')
    is one GEMM per caption batch instead of one per step (``attention.project``);
  * the previous-word half of the GRUCell input GEMM, W_ih[:, :E]·prev + b_ih, is one GEMM over
    all (sample, step) pairs (teacher forcing: every ``prev`` is known up front);
  * the word logits Linear(h_t) of ALL steps are one [Σ_t batch_t, Hd]×[ntoken, Hd]ᵀ GEMM written
    straight in pack_padded_sequence order — the padded [B, max_len, ntoken] tensor of
    generator.py:88 is never built.
Per step that leaves: W_q GEMM → vqa_attention_logits → vqa_attention_pool → W_ih[:, E:] GEMM
(+ the hoisted half as its additive epilogue operand) → W_hh GEMM → vqa_gru_cell.
"""
import torch
import torch.nn as nn

from .. import compute_dtype, ops
from .attention import set_att
from .modules import PreparedCache, as_compute, _no_training


def set_decoder(decoder_type: str, ntoken: int, embed_dim: int, hidden_dim: int, v_dim: int, max_len: int,
                device: str, dropout: float, rnn_type: str, att_type: str):
    if decoder_type == 'none':
        return
    return {
        'base': BaseDecoder,
        'butd': BUTDDecoder
    }[decoder_type](ntoken=ntoken, embed_dim=embed_dim, hidden_dim=hidden_dim, v_dim=v_dim, max_len=max_len,
                    device=device, dropout=dropout, rnn_type=rnn_type, att_type=att_type).to(device)


class DecoderModule(nn.Module):
    h_num = 1

    def init_hidden(self, batch_size):
        """Initialize hidden states (generator.py:45-49)."""
        init = torch.zeros((batch_size, self.hidden_dim), device=self.device)
        if self.rnn_type == 'LSTM':
            return [(init, init)] * self.h_num
        return [init] * self.h_num

    def select_hidden(self, h, batch_size):
        for i in range(len(h)):
            if self.rnn_type == 'LSTM':
                h[i] = (h[i][0][:batch_size], h[i][1][:batch_size])
            else:
                h[i] = h[i][:batch_size]
        return h

    def decode(self, v, v_mean, prev, h):
        pass


class BaseDecoder(DecoderModule):
    """Base generator based on "Show, Attend and Tell" (generator.py:123-181); rnn_type 'GRU' (nn.GRUCell) or 'LSTM'
    (nn.LSTMCell: the hidden state is the pair (h, c), attention and word logits read h)."""

    step_loop_in_python = False          # True: compose the per-step ops from Python (same kernels; tests run both)

    def __init__(self, ntoken: int, embed_dim: int, hidden_dim: int, v_dim: int, max_len: int, device: str,
                 dropout: float = 0.5, rnn_type: str = 'GRU', att_type: str = 'base'):
        super().__init__()
        assert rnn_type == 'LSTM' or rnn_type == 'GRU'
        self.rnn_type, self.hidden_dim, self.max_len, self.ntoken, self.device = rnn_type, hidden_dim, max_len, ntoken, device
        self.embed_dim, self.v_dim = embed_dim, v_dim
        self.h_num = 1
        rnn_cls = nn.LSTMCell if rnn_type == 'LSTM' else nn.GRUCell
        self.rnn = rnn_cls(input_size=embed_dim + v_dim, hidden_size=hidden_dim)         # parameters only
        self.attention = set_att(att_type)(v_dim=v_dim, q_dim=hidden_dim, hidden_dim=hidden_dim)
        self.fcnet = nn.Linear(hidden_dim, ntoken)
        self.dropout = nn.Dropout(dropout)                                              # identity in eval
        self._cache = PreparedCache()
        self._init_weights()

    def _init_weights(self):
        self.fcnet.bias.data.fill_(0)
        self.fcnet.weight.data.uniform_(-0.1, 0.1)

    def prepared(self, dtype):
        r, f = self.rnn, self.fcnet

        def build():
            E = self.embed_dim
            E_pad = (E + 63) // 64 * 64
            w_ih = r.weight_ih.detach()
            w_prev = torch.zeros((w_ih.shape[0], E_pad), dtype=dtype, device=w_ih.device)
            w_prev[:, :E] = w_ih[:, :E].to(dtype)
            return dict(E_pad=E_pad, w_prev=w_prev, w_att=w_ih[:, E:].to(dtype).contiguous(),
                        w_hh=r.weight_hh.detach().to(dtype).contiguous(),
                        b_ih=r.bias_ih.detach().float().contiguous(), b_hh=r.bias_hh.detach().float().contiguous(),
                        w_fc=f.weight.detach().to(dtype).contiguous(), b_fc=f.bias.detach().float().contiguous())
        return self._cache.get(("dec", dtype), (r.weight_ih, r.weight_hh, r.bias_ih, r.bias_hh, f.weight, f.bias), build)

    def _pad_prev(self, prev2d, E_pad, dtype):
        """[rows,E] word embeddings → [rows,E_pad] in the compute dtype (zero-padded GEMM operand)"""
        if prev2d.shape[1] == E_pad and prev2d.dtype == dtype and prev2d.is_contiguous():
            return prev2d
        out = torch.zeros((prev2d.shape[0], E_pad), dtype=dtype, device=prev2d.device)
        out[:, : prev2d.shape[1]] = prev2d.to(dtype)
        return out

    def decode(self, v, v_mean, prev, h):
        """One decoding step (generator.py:168-181; the call tools/caption.py:93 makes).
        v [b,K,V], prev [b,E] embedded previous word, h = [h0 f32 [b,Hd]] → ([h'], word logits f32 [b,ntoken], att [b,K,1])"""
        _no_training(self)
        dtype = compute_dtype()
        P = self.prepared(dtype)
        lstm = self.rnn_type == 'LSTM'
        state = h[0]
        h0 = (state[0] if lstm else state).to(self.device).float().contiguous()
        v = v.to(self.device)
        parts, x = self.attention.logit_parts(v, h0)
        att, att_v, _ = ops.attention_pool(parts, float(self.attention.linear.bias.detach()), x, True, True, False)
        gi = ops.linear(self._pad_prev(prev.to(self.device), P["E_pad"], dtype), P["w_prev"], bias=P["b_ih"],
                        out_dtype=torch.float32)
        gi = ops.linear(att_v, P["w_att"], add=gi, out_dtype=torch.float32)
        h_new = h0.clone()
        h_lp = torch.empty(h0.shape, dtype=dtype, device=h0.device)
        if lstm:
            gates = ops.linear(as_compute(h0, dtype), P["w_hh"], bias=P["b_hh"], add=gi, out_dtype=torch.float32)
            c_new = state[1].to(self.device).float().clone().contiguous()
            ops.lstm_cell(gates, c_new, h_new, h_lp)
        else:
            gh = ops.linear(as_compute(h0, dtype), P["w_hh"], bias=P["b_hh"], out_dtype=torch.float32)
            ops.gru_cell(gi, gh, h_new, h_lp)
        output = ops.linear(h_lp, P["w_fc"], bias=P["b_fc"], out_dtype=torch.float32)
        return [(h_new, c_new) if lstm else h_new], output, att.unsqueeze(2)

    def forward(self, batch):
        """Teacher-forced pass (generator.py:66-120) → {'predict': [Σ_t batch_t, ntoken] f32, 'target': [Σ_t batch_t]}
        in pack_padded_sequence order.  Captions are sorted by decreasing length with a STABLE sort on the host
        (the reference's ``cap_len.sort`` leaves the order of ties to the backend; its ``.tolist()`` of the lengths
        is a host round trip anyway)."""
        _no_training(self)
        dtype = compute_dtype()
        P = self.prepared(dtype)
        dev = self.device
        v, caption, target = batch['v'].to(dev), batch['c'].to(dev), batch['c_target'].to(dev)
        cap_len, sort_id = torch.sort(batch['cap_len'].cpu(), dim=0, descending=True, stable=True)
        decode_len = (cap_len - 1).tolist()                      # no step at the <end> position (generator.py:93)
        B, K, V = v.shape
        T, Hd, E = max(decode_len), self.hidden_dim, caption.shape[2]
        batches = [sum(l > t for l in decode_len) for t in range(T)]
        offs = [0]
        for bt in batches:
            offs.append(offs[-1] + bt)
        sort_dev = sort_id.to(dev)
        v = as_compute(v.index_select(0, sort_dev), dtype)
        caption = caption.index_select(0, sort_dev)
        target = target.index_select(0, sort_dev)

        # hoisted out of the time loop (see the module docstring)
        proj, x = self.attention.project(v)                                             # [B*K,Hd]
        prev_all = self._pad_prev(caption[:, :T].reshape(B * T, E), P["E_pad"], dtype)
        lstm = self.rnn_type == 'LSTM'
        ng = 4 if lstm else 3
        gi_prev = ops.linear(prev_all, P["w_prev"], bias=P["b_ih"], out_dtype=torch.float32).view(B, T * ng * Hd)
        att_bias = float(self.attention.linear.bias.detach())

        h = torch.zeros((B, Hd), dtype=torch.float32, device=dev)
        c = torch.zeros((B, Hd), dtype=torch.float32, device=dev) if lstm else None
        h_in = torch.zeros((B, Hd), dtype=dtype, device=dev)
        if self.step_loop_in_python:
            h_all = torch.empty((offs[-1], Hd), dtype=dtype, device=dev)                 # every h_t, packed order
            for t, bt in enumerate(batches):
                h_in = h_in[:bt]
                parts = self.attention.step_parts(proj, h_in, K)
                _, att_v, _ = ops.attention_pool(parts, att_bias, x[:bt], False, True, False)
                gi = ops.linear(att_v, P["w_att"], add=gi_prev[:bt, t * ng * Hd:(t + 1) * ng * Hd], out_dtype=torch.float32)
                gh = ops.linear(h_in, P["w_hh"], bias=P["b_hh"], add=gi if lstm else None, out_dtype=torch.float32)
                h_in = h_all[offs[t]:offs[t + 1]]
                if lstm:
                    ops.lstm_cell(gh, c[:bt], h[:bt], h_in)
                else:
                    ops.gru_cell(gi, gh, h[:bt], h_in)
        else:                                                                            # the same steps in one C call
            mode, w_q, q_scale, q_bias, logit_w = self.attention.step_weights(dtype)
            h_all = ops.caption_decode_steps(x, proj, batches, mode, w_q, q_scale, q_bias, logit_w, att_bias, gi_prev,
                                             P["w_att"], P["w_hh"], P["b_hh"], h, h_in, c=c)
        predict = ops.linear(h_all, P["w_fc"], bias=P["b_fc"], out_dtype=torch.float32)
        # the targets are the words after <start> (generator.py:115), packed like the predictions
        tgt = torch.cat([target[:bt, t + 1] for t, bt in enumerate(batches)])
        return {'predict': predict, 'target': tgt}


class BUTDDecoder(DecoderModule):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("decoder_type='butd' cannot run in the reference either: BUTDDecoder.decode has no "
                                  "return statement (generator.py:240-266), DecoderModule.forward unpacks None; only "
                                  "'base' is built")
