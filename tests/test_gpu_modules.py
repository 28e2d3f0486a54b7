"""GPU parity of the drop-in module API (attention / gcn / predictor / encoder / wrapper)
against the golden outputs of the real reference — the tests read like the reference's
own usage: set_model(**hparams) → load_state_dict → forward_vqa / get_att / encoder(batch, True)."""
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
    m = set_model(encoder_type="relation" if cfg.relation else "base", predictor_type=cfg.predictor, decoder_type="none",
                  ntoken=cfg.ntoken, v_dim=cfg.v_dim, embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim,
                  decoder_hidden_dim=0, rnn_layer=cfg.rnn_layer, ans_dim=cfg.ans_dim, cls_layer=2, c_len=cfg.c_len, device=device,
                  dropout=0.2, rnn_type=cfg.rnn_type, att_type=cfg.att_type, conv_layer=cfg.conv_layer, conv_type="corr")
    if cfg.relation and (cfg.use_imp or not cfg.use_spa):
        # like the reference, set_model does not reach use_imp / use_spa: the encoder is built directly
        from vqa_collection_b200.modules.encoder import RelationEncoder
        m.encoder = RelationEncoder(ntoken=cfg.ntoken, embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim,
                                    rnn_layer=cfg.rnn_layer, v_dim=cfg.v_dim, device=device, dropout=0.2, rnn_type=cfg.rnn_type,
                                    att_type=cfg.att_type, conv_layer=cfg.conv_layer, conv_type="corr", use_imp=cfg.use_imp,
                                    use_spa=cfg.use_spa).to(device)
    m.load_state_dict({k: v for k, v in W.items() if not k.startswith("gcn")}, strict=True)
    if cfg.relation:
        for name, enc in (("gcn", m.encoder.spatial_encoder), ("gcn_imp", m.encoder.implicit_encoder)):
            if enc is None:
                continue
            for i, layer in enumerate(enc.gcn):
                layer.load_state_dict({k[len(f"{name}.{i}."):]: v for k, v in W.items() if k.startswith(f"{name}.{i}.")})
    return m.eval()


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["updown_small", "regat_small", "updown_full", "regat_full", "concat_small",
                                  "concat_full", "basecap_small", "basecap_full", "gru2_small", "lstm2_small", "lstm_full",
                                  "regat_imp_small", "imp_only_small", "regat2_small"])
def test_wrapper_api_matches_reference(golden_dir, name, precision):
    import vqa_collection_b200 as pkg
    pkg.set_precision(precision)
    try:
        z = np.load(os.path.join(golden_dir, name + ".npz"))
        meta = ast.literal_eval(str(z["meta"]))
        cfg = O.Config(**meta["cfg"])
        W = O.make_weights(cfg, meta["wseed"])
        batch = O.make_batch(cfg, meta["B"], meta["bseed"])
        ref_batch = {k: v for k, v in batch.items() if k not in ("bbox", "wh")}     # CPU tensors, like the loader
        m = build_model(cfg, W)
        tol = 1e-5 if precision == "fp32" else 1e-2
        with torch.no_grad():
            predict, att = m.get_att(ref_batch)                      # module-level kernels
            score, label, target = m.forward_vqa(ref_batch)         # fused engine (module-level path for 'base-cap')
            p2, cap = m(ref_batch)
        assert cap is None and att.shape == (meta["B"], 36, 1)
        assert relerr(att[:, :, 0], z["v_att"]) < tol
        assert relerr(predict, z["logits"]) < tol and relerr(p2, z["logits"]) < tol
        assert score.shape == (meta["B"], cfg.ans_dim) and target.dtype == torch.float32
        if precision == "fp32":
            assert np.array_equal(label.cpu().numpy(), z["label"])
            assert np.allclose(score.sum(1).cpu().numpy(), z["score_sum"])
        if cfg.relation:
            with torch.no_grad():
                alphas = m.encoder(ref_batch, True)                  # tools/visualize.py:81 usage
            assert relerr(alphas[0], z["alpha"]) < tol
            # on-device labels from boxes give the same answers as the loader's precomputed graph
            bb = {k: v for k, v in batch.items() if k != "graph"}
            with torch.no_grad():
                _, label2, _ = m.forward_vqa(bb)
            assert torch.equal(label2, label)
    finally:
        pkg.set_precision("bf16")


def test_module_level_blocks_fp32(golden_dir):
    """set_att / FCNet / GCN used on their own, like the reference's building blocks"""
    import vqa_collection_b200 as pkg
    from vqa_collection_b200.modules.attention import set_att
    from vqa_collection_b200.modules.modules import FCNet
    pkg.set_precision("fp32")
    try:
        cfg = O.SMALL_REGAT
        W = O.make_weights(cfg, 1111)
        batch = O.make_batch(cfg, 5, 77)
        m = build_model(cfg, W)
        with torch.no_grad():
            q_emb = O.question_embedding(batch["q"], W)
            want_logits = O.multiply_attention_logits(batch["img"], q_emb, W)
            got_logits = m.encoder.attention.logits(batch["img"].cuda(), q_emb.cuda())
            assert relerr(got_logits, want_logits) < 1e-5
            got_att = m.encoder.attention(batch["img"].cuda(), q_emb.cuda())
            assert relerr(got_att, torch.softmax(want_logits, 1)) < 1e-5
            # FCNet on a 3-D input
            f = m.encoder.attention.W_v
            want = O.fcnet1(batch["img"], W, "encoder.attention.W_v")
            assert relerr(f(batch["img"].cuda()), want) < 1e-5
            # stand-alone GCN on arbitrary features (att=None)
            feat = torch.rand((5, 36, cfg.v_dim), generator=torch.Generator().manual_seed(1)) * 0.05
            want_f, want_a = O.gcn(feat, batch["graph"].float(), W, 1)
            got_f, got_a = m.encoder.spatial_encoder(feat.cuda(), batch["graph"].cuda(), True)
            assert relerr(got_f, want_f) < 1e-5 and relerr(got_a[0], want_a[0]) < 1e-5
            # predictor on an encoder dict without the fused 'v_sum' key (reference dict format)
            enc = O.relation_encoder(batch, W, 1)
            want_p = O.base_predictor(enc, W)
            got_p = m.predictor({"v": enc["v"].cuda(), "q": enc["q"].cuda()})
            assert relerr(got_p, want_p) < 1e-5
        assert isinstance(f, FCNet) and set_att("new") is type(m.encoder.attention)
    finally:
        pkg.set_precision("bf16")


def test_util_relation_drop_in():
    from vqa_collection_b200.util.relation import relation_graph, spatial_relation
    boxes = O.make_boxes(1, 36, 8, 640, 480)[0]
    g = relation_graph(boxes, 640, 480)
    assert g.dtype == np.float64 and g.shape == (36, 36)
    assert np.array_equal(g, O.relation_graph(boxes, 640, 480))
    assert spatial_relation(np.float32([290, 190, 310, 210]), np.float32([340, 190, 360, 210]), 640, 480) == (3, 7)


@pytest.mark.parametrize("cls_layer", [1, 3])
def test_other_classifier_depths_take_the_module_path(cls_layer):
    """the fused engine is built for the 2-layer classifier (main.py default); other depths must not reach it"""
    import vqa_collection_b200 as pkg
    from vqa_collection_b200.modules.wrapper import set_model
    pkg.set_precision("fp32")
    try:
        cfg = O.SMALL
        torch.manual_seed(3)
        m = set_model(encoder_type="base", predictor_type="base", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
                      embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, decoder_hidden_dim=0, rnn_layer=1,
                      ans_dim=cfg.ans_dim, cls_layer=cls_layer, c_len=cfg.c_len, device="cuda", dropout=0.2, rnn_type="GRU",
                      att_type="new").eval()
        assert m.engine() is None
        batch = O.make_batch(cfg, 5, 12)
        with torch.no_grad():
            score, label, target = m.forward_vqa({k: v for k, v in batch.items() if torch.is_tensor(v)})
            predict, _ = m({k: v for k, v in batch.items() if torch.is_tensor(v)})
        # the same stack in plain torch on the module's own parameters
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        enc = O.base_encoder(batch, sd)
        x = O.fcnet1(enc["v"].sum(1), sd, "predictor.v_net") * enc["q"]
        lin = [i for i, mod in enumerate(m.predictor.classifier.main) if isinstance(mod, torch.nn.Linear)]
        for i in lin:
            p = f"predictor.classifier.main.{i}"
            x = torch.relu(torch.nn.functional.linear(x, O.wn_weight(sd[p + ".weight_v"], sd[p + ".weight_g"]), sd[p + ".bias"]))
        assert relerr(predict, x) < 1e-5
        assert torch.equal(label.cpu(), x.argmax(1))
    finally:
        pkg.set_precision("bf16")


def test_pretrained_word_embedding_matches_embedding_layer(tmp_path):
    """set_model(pretrained_embed_path=...) (main.py:83 default): same answers as an nn.Embedding holding the same table,
    through the fused engine and through the module-level path"""
    import vqa_collection_b200 as pkg
    from vqa_collection_b200.modules.wrapper import set_model
    pkg.set_precision("fp32")
    try:
        cfg = O.SMALL
        n_words = cfg.ntoken - 3                                     # + <oov>,<start>,<end>,<pad> = ntoken + 1 rows
        table = torch.randn((n_words, cfg.embed_dim), generator=torch.Generator().manual_seed(9))
        path = str(tmp_path / "glove.txt")
        with open(path, "w") as f:
            for i in range(n_words):
                f.write(f"w{i} " + " ".join(f"{x:.6f}" for x in table[i].tolist()) + "\n")
        kw = dict(encoder_type="base", predictor_type="base", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
                  embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, decoder_hidden_dim=0, rnn_layer=1, ans_dim=cfg.ans_dim,
                  cls_layer=2, c_len=cfg.c_len, device="cuda", dropout=0.2, rnn_type="GRU", att_type="new")
        W = O.make_weights(cfg, 1111)
        m_glove = set_model(pretrained_embed_path=path, **kw)
        m_glove.load_state_dict({k: v for k, v in W.items() if k != "encoder.embedding.weight"}, strict=True)
        W2 = dict(W)
        W2["encoder.embedding.weight"] = m_glove.encoder.embedding.vocab.clone()
        m_plain = set_model(**kw)
        m_plain.load_state_dict(W2, strict=True)
        batch = {k: v for k, v in O.make_batch(cfg, 6, 31).items() if torch.is_tensor(v)}
        with torch.no_grad():
            s1, l1, _ = m_glove.eval().forward_vqa(batch)
            s2, l2, _ = m_plain.eval().forward_vqa(batch)
            p1, _ = m_glove(batch)
            p2, _ = m_plain(batch)
            want, _ = O.forward(batch, W2, cfg)
        assert torch.equal(l1, l2) and torch.equal(s1, s2) and torch.equal(p1, p2)
        assert relerr(p1, want) < 1e-5
    finally:
        pkg.set_precision("bf16")
