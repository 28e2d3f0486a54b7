"""Pin the CPU oracle against outputs of the REAL reference (tests/golden/*.npz,
made by tests/golden/make_golden.py from /root/reference)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import vqa_oracle as O


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    return z, meta


def _relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("name", ["updown_small", "regat_small", "updown_full", "regat_full", "concat_small",
                                  "concat_full", "basecap_small", "basecap_full", "gru2_small", "lstm2_small", "lstm_full",
                                  "regat_imp_small", "imp_only_small", "regat2_small"])
def test_forward_matches_reference(golden_dir, name):
    z, meta = _load(golden_dir, name)
    cfg = O.Config(**meta["cfg"])
    W = O.make_weights(cfg, meta["wseed"])
    batch = O.make_batch(cfg, meta["B"], meta["bseed"])
    with torch.no_grad():
        logits, enc = O.forward(batch, W, cfg)
        score, label, _ = O.forward_vqa(batch, W, cfg)
    # fp32 restatement vs the reference's fp32 library kernels: 1e-5 of max-norm
    assert _relerr(enc["q_emb"].numpy(), z["q_emb"]) < 1e-5
    assert _relerr(enc["v_att"].numpy()[:, :, 0], z["v_att"]) < 1e-5
    assert _relerr(enc["q"].numpy(), z["q"]) < 1e-5
    assert _relerr(enc["v"].numpy()[:, :, ::16], z["v_sub"]) < 1e-5
    assert _relerr(enc["v"].sum(1).numpy(), z["v_sum"]) < 1e-5
    assert _relerr(logits.numpy(), z["logits"]) < 1e-5
    assert np.array_equal(label.numpy(), z["label"])
    assert np.allclose(score.sum(1).numpy(), z["score_sum"])
    if cfg.relation:
        assert _relerr(enc["alpha"][0].numpy(), z["alpha"]) < 1e-5
        assert np.array_equal(batch["graph"].numpy().astype(np.uint8), z["graph"])


@pytest.mark.parametrize("name", ["qcap_small", "qcap_full"])
def test_qcap_forward_matches_repaired_reference(golden_dir, name):
    """config 5: the real reference with ONE method repaired (CaptionEmbedding.forward_all, see make_golden.py)"""
    z, meta = _load(golden_dir, name)
    assert meta["repaired"] == "CaptionEmbedding.forward_all"
    cfg = O.Config(**meta["cfg"])
    W = O.make_weights(cfg, meta["wseed"])
    batch = O.make_batch(cfg, meta["B"], meta["bseed"])
    with torch.no_grad():
        enc = O.base_encoder(batch, W)
        V = O.lrelu_net(enc["v"], W, "predictor.v_net", cfg.neg_slope)
        vsum = V.sum(1)
        c_emb = O.caption_embedding(vsum, enc["q"], enc["c"], W)
        predict, _ = O.forward(batch, W, cfg)
        score, label, _ = O.forward_vqa(batch, W, cfg)
    assert _relerr(enc["v_att"].numpy()[:, :, 0], z["v_att"]) < 1e-5
    assert _relerr(enc["q"].numpy(), z["q"]) < 1e-5
    assert _relerr(vsum.numpy(), z["vsum"]) < 1e-5
    assert _relerr(c_emb.numpy(), z["c_emb"]) < 1e-5
    assert _relerr(O.lrelu_net(c_emb, W, "predictor.c_net", cfg.neg_slope).numpy(), z["c_grad"]) < 1e-5
    assert _relerr(predict.numpy(), z["predict"]) < 1e-5
    assert np.array_equal(label.numpy(), z["label"])
    assert np.allclose(score.sum(1).numpy(), z["score_sum"])


@pytest.mark.parametrize("name", ["updown_full", "concat_full"])
def test_attention_logits_match_reference(golden_dir, name):
    z, meta = _load(golden_dir, name)
    cfg = O.Config(**meta["cfg"])
    W = O.make_weights(cfg, meta["wseed"])
    batch = O.make_batch(cfg, meta["B"], meta["bseed"])
    with torch.no_grad():
        lg = O.attention_logits(batch["img"], torch.from_numpy(z["q_emb"]), W)
    assert _relerr(lg.numpy()[:, :, 0], z["att_logits"]) < 1e-5


@pytest.mark.parametrize("name", ["train_small", "train_full"])
def test_training_gradients_match_reference(golden_dir, name):
    """loss and every parameter gradient of the REAL reference's get_loss + backward (dropout p=0)"""
    z, meta = _load(golden_dir, name)
    cfg = O.Config(**meta["cfg"])
    W = O.make_weights(cfg, meta["wseed"])
    batch = O.make_batch(cfg, meta["B"], meta["bseed"])
    loss, logits, grads = O.loss_and_grads(batch, W, cfg)
    assert abs(float(loss) - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    keys = [k[5:] for k in z.files if k.startswith("grad:")]
    assert sorted(keys) == sorted(grads.keys()) and len(keys) == 26
    gmax = max(float(np.abs(z["grad:" + k]).max()) for k in keys)
    for k in keys:
        g = grads[k].numpy().reshape(-1)[:: meta["stride"]]
        ref = z["grad:" + k].reshape(-1)
        if k == "encoder.attention.linear.bias":
            # the softmax is shift invariant: this gradient is identically 0, what autograd leaves is rounding noise
            assert np.abs(ref).max() < 1e-5 * gmax and np.abs(g).max() < 1e-5 * gmax
            continue
        # 5e-5 of the gradient's own max-norm: fp32 backward sums run in a different order than the reference's
        # graph (worst: the weight_g scalars, Σ dW⊙v with cancellation — 1.2e-5 observed)
        assert np.abs(g - ref).max() <= 5e-5 * np.abs(ref).max(), k
        assert abs(np.linalg.norm(grads[k].numpy().astype(np.float64)) - float(z["norm:" + k])) <= 5e-5 * float(z["norm:" + k]), k


def test_float64_truth_is_close(golden_dir):
    """The float64 run of the oracle brackets the reference's own fp32 error."""
    z, meta = _load(golden_dir, "updown_small")
    cfg = O.Config(**meta["cfg"])
    W = O.to_dtype(O.make_weights(cfg, meta["wseed"]), torch.float64)
    batch = O.make_batch(cfg, meta["B"], meta["bseed"])
    batch["img"] = batch["img"].double()
    with torch.no_grad():
        logits, _ = O.forward(batch, W, cfg)
    assert _relerr(logits.numpy(), z["logits"]) < 1e-4


def test_relation_grid_exact(golden_dir):
    z, meta = _load(golden_dir, "relation")
    boxes = O.make_boxes(48, 36, meta["grid_seed"], meta["W"], meta["H"], grid=True)
    got = O.relation_graph_batch(boxes, meta["W"], meta["H"])
    assert np.array_equal(got, z["grid_graph"])
    one = O.relation_graph(boxes[0], meta["W"], meta["H"])
    assert one.dtype == np.float64 and np.array_equal(one.astype(np.uint8), z["grid_graph"][0])
    assert np.all(np.diagonal(got, axis1=1, axis2=2) == 0)


def test_relation_continuous(golden_dir):
    z, meta = _load(golden_dir, "relation")
    boxes = O.make_boxes(16, 36, meta["cont_seed"], meta["W"], meta["H"], grid=False)
    got = O.relation_graph_batch(boxes, meta["W"], meta["H"])
    # same numpy float32 arctan2 on this host: expected exact; gate is statistical (H2)
    assert (got != z["cont_graph"]).mean() < 1e-4


def test_relation_known_answers(golden_dir):
    z, meta = _load(golden_dir, "relation")
    ab, ba = O.spatial_relation_pairs(z["ka_a"], z["ka_b"], meta["W"], meta["H"])
    assert np.array_equal(np.stack([ab, ba], 1), z["ka_labels"])
    # the table of SURVEY.md §8a-R, hand-pinned (first 16 = the compass, a→b / b→a)
    compass = [(3, 7), (11, 7), (10, 6), (10, 6), (9, 5), (9, 5), (8, 4), (8, 4),
               (7, 3), (7, 11), (6, 10), (6, 10), (5, 9), (5, 9), (4, 8), (4, 8)]
    assert [tuple(x) for x in z["ka_labels"][:16].tolist()] == compass
    rest = [(1, 2), (2, 1), (1, 2), (1, 2), (3, 3), (1, 2), (0, 0), (3, 7),
            (3, 3), (3, 7), (9, 5), (10, 6)]
    assert [tuple(x) for x in z["ka_labels"][16:].tolist()] == rest


@pytest.mark.parametrize("name", ["decoder_small", "decoder_full", "decoder_lstm_small"])
def test_caption_decoder_matches_reference(golden_dir, name):
    """SURVEY §8f f3 (second half): the teacher-forced BaseDecoder forward and one decode() step of the real
    reference (decoder_type='base', the main.py default) — ragged caption lengths, packed output order."""
    z, meta = _load(golden_dir, name)
    cfg = O.Config(**meta["cfg"])
    W = O.make_weights(cfg, meta["wseed"])
    batch = O.make_decoder_batch(cfg, meta["B"], meta["bseed"])
    assert np.array_equal(batch["cap_len"].numpy(), z["cap_len"])
    cs = meta["col_stride"]
    with torch.no_grad():
        logits, enc = O.forward(batch, W, cfg)
        cap = O.base_decoder_forward(enc, W, cfg)
        h0 = torch.rand((meta["B"], cfg.decoder_hidden_dim), generator=torch.Generator().manual_seed(meta["bseed"])) - 0.5
        if cfg.rnn_type == "LSTM":
            c0 = torch.rand((meta["B"], cfg.decoder_hidden_dim), generator=torch.Generator().manual_seed(meta["bseed"] + 1)) - 0.5
            (h1, c1), word, att = O.base_decoder_step(enc["v"], enc["c"][:, 3], h0, W, c=c0)
            assert _relerr(c1.numpy(), z["step_c"]) < 1e-5
        else:
            h1, word, att = O.base_decoder_step(enc["v"], enc["c"][:, 3], h0, W)
    assert _relerr(logits.numpy(), z["logits"]) < 1e-5                       # the VQA head is unchanged by the caption head
    assert cap["predict"].shape[0] == int((z["cap_len"] - 1).sum())
    assert np.array_equal(cap["target"].numpy(), z["cap_target"])
    assert _relerr(cap["predict"].numpy()[:, ::cs], z["cap_predict_sub"]) < 1e-5
    assert _relerr(torch.logsumexp(cap["predict"], 1).numpy(), z["cap_lse"]) < 1e-5
    assert np.array_equal(cap["predict"].argmax(1).numpy(), z["cap_argmax"])
    loss = torch.nn.functional.cross_entropy(cap["predict"], cap["target"])   # wrapper.py:32-36
    assert abs(loss.item() - float(z["cap_loss"])) < 1e-5 * abs(float(z["cap_loss"]))
    assert _relerr(h1.numpy(), z["step_h"]) < 1e-5
    assert _relerr(word.numpy()[:, ::cs], z["step_word_sub"]) < 1e-5
    assert np.array_equal(word.argmax(1).numpy(), z["step_word_argmax"])
    assert _relerr(att.numpy()[:, :, 0], z["step_att"]) < 1e-5
