"""GPU parity of the caption head (BaseDecoder teacher-forced forward + decode step,
generator.py:66-181) against goldens of the real reference and, for cases the fixtures do not hold
(ties in the caption lengths, att_type='base'), against the CPU oracle that those goldens pin."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import vqa_oracle as O

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def build_model(cfg, W, device="cuda"):
    from vqa_collection_b200.modules.wrapper import set_model
    m = set_model(encoder_type="base", predictor_type="base", decoder_type=cfg.decoder, ntoken=cfg.ntoken, v_dim=cfg.v_dim,
                  embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, decoder_hidden_dim=cfg.decoder_hidden_dim,
                  rnn_layer=cfg.rnn_layer, ans_dim=cfg.ans_dim, cls_layer=2, c_len=cfg.c_len, device=device, dropout=0.2,
                  rnn_type=cfg.rnn_type, att_type=cfg.att_type)
    m.load_state_dict(W, strict=True)
    return m.eval()


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["decoder_small", "decoder_full", "decoder_lstm_small"])
def test_caption_head_matches_reference(golden_dir, name, precision):
    import vqa_collection_b200 as pkg
    pkg.set_precision(precision)
    try:
        z = np.load(os.path.join(golden_dir, name + ".npz"))
        meta = ast.literal_eval(str(z["meta"]))
        cfg = O.Config(**meta["cfg"])
        W = O.make_weights(cfg, meta["wseed"])
        batch = O.make_decoder_batch(cfg, meta["B"], meta["bseed"])
        m = build_model(cfg, W)
        tol = 1e-5 if precision == "fp32" else 1e-2
        cs = meta["col_stride"]
        with torch.no_grad():
            predict, cap = m(batch)                                  # Wrapper.forward: both heads (wrapper.py:64-74)
            cap2 = m.forward_cap(batch)
            enc = m.encoder(batch)
            h0 = torch.rand((meta["B"], cfg.decoder_hidden_dim), generator=torch.Generator().manual_seed(meta["bseed"])) - 0.5
            if cfg.rnn_type == "LSTM":                                # nn.LSTMCell: the state is the pair (h, c)
                c0 = torch.rand((meta["B"], cfg.decoder_hidden_dim), generator=torch.Generator().manual_seed(meta["bseed"] + 1)) - 0.5
                h1, word, att = m.generator.decode(v=enc["v"], v_mean=None, prev=enc["c"][:, 3], h=[(h0.cuda(), c0.cuda())])
                assert relerr(h1[0][1], z["step_c"]) < tol
                h1 = [h1[0][0]]
            else:
                h1, word, att = m.generator.decode(v=enc["v"], v_mean=None, prev=enc["c"][:, 3], h=[h0.cuda()])
        assert relerr(predict, z["logits"]) < tol
        assert cap["predict"].dtype == torch.float32 and cap["predict"].shape == (int((z["cap_len"] - 1).sum()), cfg.ntoken)
        assert torch.equal(cap["predict"], cap2["predict"])                              # deterministic
        assert np.array_equal(cap["target"].cpu().numpy(), z["cap_target"])
        assert relerr(cap["predict"][:, ::cs], z["cap_predict_sub"]) < tol
        assert relerr(torch.logsumexp(cap["predict"], 1), z["cap_lse"]) < tol
        loss = torch.nn.functional.cross_entropy(cap["predict"], cap["target"])          # wrapper.py:32-36
        assert abs(loss.item() - float(z["cap_loss"])) < tol * abs(float(z["cap_loss"]))
        assert relerr(h1[0], z["step_h"]) < tol and h1[0].dtype == torch.float32
        assert relerr(word[:, ::cs], z["step_word_sub"]) < tol
        assert relerr(att[:, :, 0], z["step_att"]) < tol and att.shape == (meta["B"], 36, 1)
        if precision == "fp32":
            assert np.array_equal(cap["predict"].argmax(1).cpu().numpy(), z["cap_argmax"])
            assert np.array_equal(word.argmax(1).cpu().numpy(), z["step_word_argmax"])
    finally:
        pkg.set_precision("bf16")


@pytest.mark.parametrize("rnn_type", ["GRU", "LSTM"])
@pytest.mark.parametrize("att_type", ["new", "base"])
def test_caption_head_ties_and_concat_attention_vs_oracle(att_type, rnn_type):
    """B = 40 > c_len - 1: equal caption lengths occur (stable order), batch_t shrinks step by step;
    att_type='base' runs the decoder's ConcatAttention (mode 1 of vqa_attention_logits)"""
    import vqa_collection_b200 as pkg
    from dataclasses import replace
    pkg.set_precision("fp32")
    try:
        cfg = replace(O.SMALL_DECODER, att_type=att_type, rnn_type=rnn_type)
        W = O.make_weights(cfg, 2222)
        batch = O.make_decoder_batch(cfg, 40, 7003)
        assert len(set(batch["cap_len"].tolist())) < 40
        m = build_model(cfg, W)
        with torch.no_grad():
            want_logits, enc = O.forward(batch, W, cfg)
            want = O.base_decoder_forward(enc, W, cfg)
            predict, cap = m(batch)
        assert relerr(predict, want_logits) < 1e-5
        assert torch.equal(cap["target"].cpu(), want["target"])
        assert relerr(cap["predict"], want["predict"]) < 1e-5
        # the per-step ops composed from Python launch the same kernels in the same order as vqa_caption_decode_steps
        m.generator.step_loop_in_python = True
        with torch.no_grad():
            cap_py = m.forward_cap(batch)
        m.generator.step_loop_in_python = False
        assert torch.equal(cap_py["predict"], cap["predict"]) and torch.equal(cap_py["target"], cap["target"])
        # one caption only (the tools/caption.py beam-search shape): batch of 1, a single step
        one = {k: (v[:1] if torch.is_tensor(v) else v) for k, v in batch.items()}
        one["cap_len"] = torch.tensor([2])
        with torch.no_grad():
            _, enc1 = O.forward(one, W, cfg)
            want1 = O.base_decoder_forward(enc1, W, cfg)
            got1 = m.forward_cap(one)
        assert got1["predict"].shape == (1, cfg.ntoken) and relerr(got1["predict"], want1["predict"]) < 1e-5
    finally:
        pkg.set_precision("bf16")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", [0, 1])
def test_attention_logits_and_gru_cell_ops(dtype, mode):
    from vqa_collection_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, K, Hd = 7, 36, 72                                             # Hd not a multiple of 256: lane tail
    proj = (torch.rand((B * K + 36, Hd), generator=g) - (0.5 if mode else 0.0)).to(dtype)      # extra rows are ignored
    q = torch.rand((B, Hd), generator=g) - (0.5 if mode else 0.0)
    w = torch.rand((Hd,), generator=g) - 0.5
    got = ops.attention_logits(proj.cuda(), q.cuda(), w.cuda(), K, mode)
    p = proj[:B * K].float().view(B, K, Hd)
    want = ((p * q[:, None]) if mode == 0 else torch.relu(p + q[:, None])) @ w
    assert got.shape == (B * K, 1) and relerr(got.view(B, K), want) < 1e-5
    assert ops.attention_logits(proj.cuda(), q[:0].cuda(), w.cuda(), K, mode).shape == (0, 1)   # empty step
    # GRUCell gate update against torch.nn.GRUCell fed the same pre-activations
    H = 40
    cell = torch.nn.GRUCell(16, H)
    x, h = torch.rand((B, 16), generator=g), torch.rand((B, H), generator=g) - 0.5
    with torch.no_grad():
        want_h = cell(x, h)
        gi = torch.nn.functional.linear(x, cell.weight_ih, cell.bias_ih)
        gh = torch.nn.functional.linear(h, cell.weight_hh, cell.bias_hh)
    h_dev = h.cuda().clone()
    h_lp = torch.zeros((B, 2 * H), dtype=dtype, device="cuda")[:, :H]                        # row-strided view
    ops.gru_cell(gi.cuda(), gh.cuda(), h_dev, h_lp)
    assert relerr(h_dev, want_h) < 1e-5
    assert relerr(h_lp.float(), want_h) < (1e-5 if dtype == torch.float32 else 4e-3)
