"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference (read-only; never copied
into this repo), loads the seeded synthetic weights of
``oracle.vqa_oracle.make_weights`` into the reference's own ``Wrapper`` (GCN
layer tensors are assigned object-by-object because they are not registered
parameters), runs the reference's own forward on the seeded
batches of ``make_batch`` and stores ONLY outputs (inputs and weights are
re-derivable from the seeds).  The reference cannot travel to the GPU box, the
fixtures do.
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np
import torch

from oracle import vqa_oracle as O

from modules.wrapper import set_model          # noqa: E402  (the reference)
from util.relation import relation_graph, spatial_relation   # noqa: E402


def build_reference(cfg: O.Config, W: dict):
    m = set_model(encoder_type="relation" if cfg.relation else "base",
                  predictor_type=cfg.predictor, decoder_type=cfg.decoder, ntoken=cfg.ntoken,
                  v_dim=cfg.v_dim, embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim,
                  decoder_hidden_dim=cfg.decoder_hidden_dim, rnn_layer=cfg.rnn_layer, ans_dim=cfg.ans_dim, cls_layer=2,
                  c_len=cfg.c_len, device="cpu", dropout=0.2, neg_slope=cfg.neg_slope, rnn_type=cfg.rnn_type,
                  att_type=cfg.att_type, conv_layer=cfg.conv_layer, conv_type="corr")
    if cfg.relation and (cfg.use_imp or not cfg.use_spa):
        # set_model cannot reach use_imp / use_spa (encoder.py:36-48 does not pass them): build the reference's own
        # RelationEncoder directly and put it in the Wrapper
        from modules.encoder import RelationEncoder as RefRelationEncoder
        m.encoder = RefRelationEncoder(ntoken=cfg.ntoken, embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim,
                                       rnn_layer=cfg.rnn_layer, v_dim=cfg.v_dim, device="cpu", dropout=0.2,
                                       rnn_type=cfg.rnn_type, att_type=cfg.att_type, conv_layer=cfg.conv_layer,
                                       conv_type="corr", use_imp=cfg.use_imp, use_spa=cfg.use_spa)
    sd = {k: v for k, v in W.items() if not k.startswith("gcn")}
    m.load_state_dict(sd, strict=True)
    if cfg.relation:
        for name, enc in (("gcn", m.encoder.spatial_encoder), ("gcn_imp", m.encoder.implicit_encoder)):
            if enc is None:
                continue
            for i, layer in enumerate(enc.gcn):
                lsd = {k[len(f"{name}.{i}."):]: v for k, v in W.items() if k.startswith(f"{name}.{i}.")}
                layer.load_state_dict(lsd, strict=True)
    return m.eval()


def run_model(name, cfg, B, wseed, bseed):
    W = O.make_weights(cfg, wseed)
    batch = O.make_batch(cfg, B, bseed)
    m = build_reference(cfg, W)
    ref_batch = {k: v for k, v in batch.items() if k not in ("bbox", "wh")}
    with torch.no_grad():
        enc = m.encoder(ref_batch)
        logits = m.predictor(enc)
        score, label, _ = m.forward_vqa(ref_batch)
        out = {
            "logits": logits.numpy(), "label": label.numpy(), "score_sum": score.sum(1).numpy(),
            "v_att": enc["v_att"].numpy()[:, :, 0], "q": enc["q"].numpy(),
            # encoder 'v' is [B,K,V]: keep every 16th channel + the K-sum
            "v_sub": enc["v"].numpy()[:, :, ::16].copy(), "v_sum": enc["v"].sum(1).numpy(),
        }
        # the GRU state is not returned by the reference encoder: recompute it with the
        # reference's own sub-modules
        out["q_emb"] = m.encoder.q_rnn(m.encoder.embedding(batch["q"])).numpy()
        out["att_logits"] = m.encoder.attention.logits(batch["img"], torch.from_numpy(out["q_emb"])).numpy()[:, :, 0]
        if cfg.relation:
            out["alpha"] = m.encoder(ref_batch, True)[0].numpy()
            out["graph"] = batch["graph"].numpy().astype(np.uint8)
    meta = dict(cfg=cfg.as_dict(), B=B, wseed=wseed, bseed=bseed)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array(repr(meta)), **out)
    print(name, {k: v.shape for k, v in out.items()})


def repaired_forward_all(self, v, q, c):
    """Minimal repair of CaptionEmbedding.forward_all (modules.py:291-297), which raises UnboundLocalError as
    written: `word_hidden` / `cap_hidden` are read before assignment and
    SentenceEmbedding.forward_all takes one argument and returns only `output`.  The repair keeps every live
    line and supplies the one missing value — the word GRU's final hidden state, taken from the same nn.GRU
    call forward_all makes (modules.py:149-151)."""
    self.word_rnn.rnn.flatten_parameters()
    output, word_hidden = self.word_rnn.rnn(c, self.word_rnn.init_hidden(c.size(0)))   # [B,T,H], [1,B,H]
    word_hidden = self.attention(word_hidden, v, q)                                    # modules.py:293
    word_hidden = word_hidden.repeat(20, 1, 1).transpose(0, 1)                         # modules.py:294
    output = self.caption_rnn.forward_all(word_hidden * output)                        # modules.py:295
    output = self.fcnet(output)                                                        # modules.py:296
    return output


def run_qcap(name, cfg, B, wseed, bseed):
    """Config 5: the REAL reference (encoder, LReLUNet, CaptionAttention, both GRUs, PredictorwithCaption.forward)
    with the one broken method replaced by `repaired_forward_all`.  Grad mode stays on: predictor.py:199,211 call
    retain_grad(), which raises under no_grad."""
    from modules.modules import CaptionEmbedding
    assert cfg.c_len == 20                           # modules.py:294 hard-codes repeat(20, ...)
    W = O.make_weights(cfg, wseed)
    batch = O.make_batch(cfg, B, bseed)
    m = build_reference(cfg, W)
    CaptionEmbedding.forward_all = repaired_forward_all
    ref_batch = {k: v for k, v in batch.items() if k not in ("bbox", "wh")}
    enc = m.encoder(ref_batch)
    enc_q, v_att = enc["q"].detach().clone(), enc["v_att"].detach().clone()
    # sub-module outputs through the reference's own modules, for stage-wise pinning
    pr = m.predictor
    V = pr.v_net(enc["v"])
    vsum = V.sum(1)
    c_emb = pr.caption_embedding(vsum, enc["q"], enc["c"])
    predict = pr(enc)                                # mutates enc['v'] in place (predictor.py:188)
    score, label, _ = m.forward_vqa(ref_batch)
    out = {"predict": predict.detach().numpy(), "label": label.numpy(), "score_sum": score.sum(1).detach().numpy(),
           "v_att": v_att.numpy()[:, :, 0], "q": enc_q.numpy(), "vsum": vsum.detach().numpy(),
           "c_emb": c_emb.detach().numpy(), "c_grad": pr.c_grad.detach().numpy(),
           "joint": pr.logit_grad.detach().numpy()}
    meta = dict(cfg=cfg.as_dict(), B=B, wseed=wseed, bseed=bseed, repaired="CaptionEmbedding.forward_all")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array(repr(meta)), **out)
    print(name, {k: v.shape for k, v in out.items()})


def run_train(name, cfg, B, wseed, bseed, full_grads):
    """Wrapper.get_loss + loss.backward() of the REAL reference in train mode with every nn.Dropout
    set to p=0 / not in-place (the in-place dropout after ReLU breaks autograd in torch 2.x)."""
    W = O.make_weights(cfg, wseed)
    batch = O.make_batch(cfg, B, bseed)
    m = build_reference(cfg, W).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p, mod.inplace = 0.0, False
    ref_batch = {k: v for k, v in batch.items() if k not in ("bbox", "wh")}
    loss, writes = m.get_loss(ref_batch)
    loss.backward()
    out = {"loss": np.array(loss.item(), dtype=np.float64), "score": np.array(writes["train/score"], dtype=np.float64)}
    for k, p in m.named_parameters():
        g = p.grad.detach().numpy().astype(np.float32)
        out["norm:" + k] = np.array(np.linalg.norm(g.astype(np.float64)))
        out["grad:" + k] = g if full_grads else g.reshape(-1)[::97].copy()
    meta = dict(cfg=cfg.as_dict(), B=B, wseed=wseed, bseed=bseed, full_grads=full_grads, stride=1 if full_grads else 97)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array(repr(meta)), **out)
    print(name, "loss", loss.item(), "params", len(out) // 2 - 1)


def run_decoder(name, cfg, B, wseed, bseed, col_stride):
    """Wrapper.forward with the default caption head (decoder_type='base', main.py:87): VQA logits + the
    teacher-forced word logits of DecoderModule.forward in pack_padded_sequence order, and one decode() step
    (the call tools/caption.py:93 makes).  Word logits: every ``col_stride``-th column + row argmax + row
    logsumexp + the cross-entropy of wrapper.py:32-36 (keeps the fixture small at ntoken=20000)."""
    W = O.make_weights(cfg, wseed)
    batch = O.make_decoder_batch(cfg, B, bseed)
    m = build_reference(cfg, W)
    ref_batch = {k: v for k, v in batch.items() if k not in ("bbox", "wh")}
    with torch.no_grad():
        predict, cap = m(ref_batch)
        cap2 = m.forward_cap(ref_batch)
        assert torch.equal(cap["predict"], cap2["predict"])
        loss_cap = torch.nn.functional.cross_entropy(cap["predict"], cap["target"])
        enc = m.encoder(ref_batch)
        h0 = [torch.rand((B, cfg.decoder_hidden_dim), generator=torch.Generator().manual_seed(bseed)) - 0.5]
        if cfg.rnn_type == "LSTM":                          # nn.LSTMCell state is the pair (h, c)
            h0 = [(h0[0], torch.rand((B, cfg.decoder_hidden_dim), generator=torch.Generator().manual_seed(bseed + 1)) - 0.5)]
        h1, word, att = m.generator.decode(v=enc["v"], v_mean=enc["v"].mean(1), prev=enc["c"][:, 3], h=h0)
    out = {"logits": predict.numpy(), "cap_predict_sub": cap["predict"].numpy()[:, ::col_stride].copy(),
           "cap_argmax": cap["predict"].argmax(1).numpy(), "cap_lse": torch.logsumexp(cap["predict"], 1).numpy(),
           "cap_target": cap["target"].numpy(), "cap_loss": np.array(loss_cap.item(), dtype=np.float64),
           "cap_len": batch["cap_len"].numpy(),
           "step_h": (h1[0][0] if cfg.rnn_type == "LSTM" else h1[0]).numpy(), "step_word_sub": word.numpy()[:, ::col_stride].copy(),
           "step_word_argmax": word.argmax(1).numpy(), "step_att": att.numpy()[:, :, 0]}
    if cfg.rnn_type == "LSTM":
        out["step_c"] = h1[0][1].numpy()
    meta = dict(cfg=cfg.as_dict(), B=B, wseed=wseed, bseed=bseed, col_stride=col_stride)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array(repr(meta)), **out)
    print(name, {k: v.shape for k, v in out.items()}, float(loss_cap))


def run_relation():
    W_, H_ = 640, 480
    boxes = O.make_boxes(48, 36, 4242, W_, H_, grid=True)
    graphs = np.stack([relation_graph(boxes[i], W_, H_) for i in range(boxes.shape[0])]).astype(np.uint8)
    cont = O.make_boxes(16, 36, 4243, W_, H_, grid=False)
    cgraphs = np.stack([relation_graph(cont[i], W_, H_) for i in range(cont.shape[0])]).astype(np.uint8)
    # known-answer table of SURVEY.md §8a-R evaluated through the reference itself
    a = np.array([290, 190, 310, 210], dtype=np.float32)
    shifts = [(50, 0), (50, 20), (50, 50), (20, 50), (0, 50), (-20, 50), (-50, 50), (-50, 20),
              (-50, 0), (-50, -20), (-50, -50), (-20, -50), (0, -50), (20, -50), (50, -50), (50, -20)]
    ka_a, ka_b, ka_l = [], [], []
    for dx, dy in shifts:
        b = a + np.array([dx, dy, dx, dy], dtype=np.float32)
        ka_a.append(a); ka_b.append(b); ka_l.append(spatial_relation(a, b, W_, H_))
    a2 = np.array([10, 10, 100, 100], dtype=np.float32)
    for b in ([20, 20, 50, 50], [0, 0, 200, 200], [10, 10, 100, 100], [10, 10, 50, 50],
              [20, 10, 110, 100], [50, 50, 50, 50], [600, 440, 630, 470], [120, 10, 200, 100]):
        b = np.array(b, dtype=np.float32)
        ka_a.append(a2); ka_b.append(b); ka_l.append(spatial_relation(a2, b, W_, H_))
    a3 = np.array([0, 0, 10, 10], dtype=np.float32)
    for b in ([20, 20, 30, 30], [20, 0, 30, 10], [0, 20, 10, 30], [200, 200, 203, 203]):
        b = np.array(b, dtype=np.float32)
        ka_a.append(a3); ka_b.append(b); ka_l.append(spatial_relation(a3, b, W_, H_))
    np.savez_compressed(os.path.join(HERE, "relation.npz"),
                        meta=np.array(repr(dict(W=W_, H=H_, grid_seed=4242, cont_seed=4243))),
                        grid_graph=graphs, cont_graph=cgraphs,
                        ka_a=np.stack(ka_a), ka_b=np.stack(ka_b), ka_labels=np.array(ka_l, dtype=np.uint8))
    print("relation", graphs.shape, cgraphs.shape, np.array(ka_l).T)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "--regat2-only" in sys.argv:
        run_model("regat2_small", O.SMALL_REGAT2, 8, 1111, 3005)
        sys.exit(0)
    if "--imp-only" in sys.argv:
        run_model("regat_imp_small", O.SMALL_REGAT_IMP, 8, 1111, 3003)
        run_model("imp_only_small", O.SMALL_IMP_ONLY, 8, 1111, 3004)
        sys.exit(0)
    if "--rnn-only" in sys.argv:
        run_model("gru2_small", O.SMALL_GRU2, 8, 1111, 9001)
        run_model("lstm2_small", O.SMALL_LSTM2, 8, 1111, 9002)
        run_model("lstm_full", O.FULL_LSTM, 4, 1111, 9003)
        sys.exit(0)
    if "--basecap-only" in sys.argv:
        run_model("basecap_small", O.SMALL_BASECAP, 8, 1111, 8001)
        run_model("basecap_full", O.FULL_BASECAP, 4, 1111, 8002)
        sys.exit(0)
    if "--decoder-lstm-only" in sys.argv:
        run_decoder("decoder_lstm_small", O.SMALL_DECODER_LSTM, 8, 1111, 7003, 1)
        sys.exit(0)
    if "--decoder-only" in sys.argv:
        run_decoder("decoder_small", O.SMALL_DECODER, 8, 1111, 7001, 1)
        run_decoder("decoder_full", O.FULL_DECODER, 5, 1111, 7002, 16)
        sys.exit(0)
    if "--qcap-only" in sys.argv:
        run_qcap("qcap_small", O.SMALL_QCAP, 8, 1111, 6001)
        run_qcap("qcap_full", O.FULL_QCAP, 4, 1111, 6002)
        sys.exit(0)
    run_relation()
    run_model("updown_small", O.SMALL, 8, 1111, 2001)
    run_model("regat_small", O.SMALL_REGAT, 8, 1111, 3001)
    run_model("updown_full", O.FULL, 4, 1111, 2002)
    run_model("regat_full", O.FULL_REGAT, 4, 1111, 3002)
    run_model("concat_small", O.SMALL_CONCAT, 8, 1111, 4001)
    run_model("concat_full", O.FULL_CONCAT, 4, 1111, 4002)
    run_train("train_small", O.SMALL, 16, 1111, 5001, True)
    run_train("train_full", O.FULL, 8, 1111, 5002, False)
    run_qcap("qcap_small", O.SMALL_QCAP, 8, 1111, 6001)
    run_qcap("qcap_full", O.FULL_QCAP, 4, 1111, 6002)
    run_decoder("decoder_small", O.SMALL_DECODER, 8, 1111, 7001, 1)
    run_decoder("decoder_full", O.FULL_DECODER, 5, 1111, 7002, 16)
    run_decoder("decoder_lstm_small", O.SMALL_DECODER_LSTM, 8, 1111, 7003, 1)
    run_model("basecap_small", O.SMALL_BASECAP, 8, 1111, 8001)
    run_model("basecap_full", O.FULL_BASECAP, 4, 1111, 8002)
    run_model("gru2_small", O.SMALL_GRU2, 8, 1111, 9001)
    run_model("lstm2_small", O.SMALL_LSTM2, 8, 1111, 9002)
    run_model("lstm_full", O.FULL_LSTM, 4, 1111, 9003)
    run_model("regat_imp_small", O.SMALL_REGAT_IMP, 8, 1111, 3003)
    run_model("imp_only_small", O.SMALL_IMP_ONLY, 8, 1111, 3004)
    run_model("regat2_small", O.SMALL_REGAT2, 8, 1111, 3005)
