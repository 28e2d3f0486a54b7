"""CPU oracle for the VQA forward hot path of Jayie/vqa-collection.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a CPU restatement
(torch-CPU tensor ops for the floating-point contractions, numpy for the box
geometry) of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU baseline.
Nothing under ``vqa_collection_b200/`` imports it.

Parity status: PINNED.  ``tests/golden/*.npz`` hold outputs of the real
reference (``/root/reference``, imported unmodified by
``tests/golden/make_golden.py``) on seeded inputs; ``tests/test_oracle_golden.py``
checks every function here against them.  The one exception is the float32
``arctan2`` rounding inside ``spatial_relation`` (numpy SIMD path, host
dependent): pinned only on integer-grid boxes and the known-answer
table of SURVEY.md §8a-R.

Config 5 (predictor 'q-cap', ``qcap_predictor`` below) is PINNED ONLY TO A REPAIRED REFERENCE: the
reference's ``CaptionEmbedding.forward_all`` (modules.py:291-297) raises as written;
``tests/golden/make_golden.py`` swaps in the minimal repair documented there for that ONE method and
runs everything else (LReLUNet, CaptionAttention, both GRUs, PredictorwithCaption) unmodified.
Parity for that method itself is therefore unpinned; every other op of config 5 is pinned.

Every function cites the reference file:line it follows (paths relative to the
reference repository root).  All functions compute in the dtype of their
inputs, so feeding float64 weights/inputs yields a float64 "truth" run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# configuration / synthetic data
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class Config:
    """Hyper-parameters of the path (reference defaults: main.py:60-80)."""
    ntoken: int = 20000
    v_dim: int = 2048
    embed_dim: int = 300
    hidden_dim: int = 1024
    ans_dim: int = 3129
    num_objs: int = 36
    q_len: int = 14
    c_len: int = 20
    num_labels: int = 12
    conv_layer: int = 1
    relation: bool = False          # encoder_type 'relation' vs 'base'
    use_imp: bool = False           # RelationEncoder(use_imp=True): + the implicit (fully connected) branch, encoder.py:202,252-257
    use_spa: bool = True            # the spatial branch (encoder.py:203,260-264)
    att_type: str = "new"           # 'new' = MultiplyAttention (CLI default main.py:67), 'base' = ConcatAttention
    predictor: str = "base"         # 'base' = BasePredictor, 'q-cap' = PredictorwithCaption (config 5)
    neg_slope: float = 0.01         # LeakyReLU slope of the q-cap predictor's own LReLUNets (predictor.py:159)
    rnn_type: str = "GRU"           # question encoder cell (main.py:66): 'GRU' or 'LSTM'
    rnn_layer: int = 1              # stacked layers of the question encoder (main.py:72)
    decoder: str = "none"           # 'base' = BaseDecoder caption head (generator.py:123-181; main.py:87 default)
    decoder_hidden_dim: int = 512   # main.py:88

    def as_dict(self):
        return asdict(self)


FULL = Config()
FULL_REGAT = Config(relation=True)
# small config: every dimension shrunk, same structure (fast CPU tests)
SMALL = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200,
               relation=False)
SMALL_REGAT = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200,
                     relation=True)
SMALL_REGAT2 = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, relation=True, conv_layer=2)
SMALL_REGAT_IMP = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, relation=True, use_imp=True)
SMALL_IMP_ONLY = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, relation=True, use_imp=True,
                        use_spa=False)
SMALL_CONCAT = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, att_type="base")
FULL_CONCAT = Config(att_type="base")
SMALL_QCAP = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, predictor="q-cap")
FULL_QCAP = Config(predictor="q-cap")
SMALL_GRU2 = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, rnn_layer=2)
SMALL_LSTM2 = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, rnn_type="LSTM", rnn_layer=2)
FULL_LSTM = Config(rnn_type="LSTM")
SMALL_BASECAP = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, predictor="base-cap")
FULL_BASECAP = Config(predictor="base-cap")
SMALL_DECODER = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, decoder="base",
                       decoder_hidden_dim=64)
FULL_DECODER = Config(decoder="base")
SMALL_DECODER_LSTM = Config(ntoken=100, v_dim=256, embed_dim=64, hidden_dim=128, ans_dim=200, decoder="base",
                            decoder_hidden_dim=64, rnn_type="LSTM")


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def make_weights(cfg: Config, seed: int = 1111, sharpen_att: float = 100.0,
                 sharpen_cls: float = 4.0, sharpen_gcn: float = 4.0) -> dict:
    """Seeded synthetic weights, keyed by the reference's parameter names
    (probed listing of ``Wrapper.state_dict()``), plus the
    unregistered GCN layer tensors (gcn.py:188-190) under ``gcn.{i}.*``.

    Distributions follow the reference's default initialisers (nn.Linear /
    nn.GRU / nn.Embedding defaults; gcn.py:69-76 for the label bias) but are
    drawn by this function from ONE torch.Generator so that the same seed gives
    the same weights on any host.  ``weight_g`` is ``‖v‖_F`` (weight_norm's
    init) times an optional "trained-like" sharpening factor so
    that attention and answer distributions are not flat.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    H, V, E, A = cfg.hidden_dim, cfg.v_dim, cfg.embed_dim, cfg.ans_dim
    w = {}

    emb = torch.randn((cfg.ntoken + 1, E), generator=g, dtype=torch.float32)
    emb[cfg.ntoken].zero_()                       # padding_idx row (encoder.py:128)
    w["encoder.embedding.weight"] = emb
    kb = 1.0 / math.sqrt(H)
    ng = 4 if cfg.rnn_type == "LSTM" else 3                 # gates per cell (nn.LSTM [i;f;g;o] / nn.GRU [r;z;n])
    for l in range(cfg.rnn_layer):
        w[f"encoder.q_rnn.rnn.weight_ih_l{l}"] = _uniform(g, (ng * H, E if l == 0 else H), kb)
        w[f"encoder.q_rnn.rnn.weight_hh_l{l}"] = _uniform(g, (ng * H, H), kb)
        w[f"encoder.q_rnn.rnn.bias_ih_l{l}"] = _uniform(g, (ng * H,), kb)
        w[f"encoder.q_rnn.rnn.bias_hh_l{l}"] = _uniform(g, (ng * H,), kb)

    def wn_linear(prefix, out_dim, in_dim, gscale=1.0):
        b = 1.0 / math.sqrt(in_dim)
        v = _uniform(g, (out_dim, in_dim), b)
        w[prefix + ".bias"] = _uniform(g, (out_dim,), b)
        w[prefix + ".weight_g"] = (torch.norm(v) * gscale).reshape(())
        w[prefix + ".weight_v"] = v

    if cfg.att_type == "new":
        wn_linear("encoder.attention.W_v.main.0", H, V)
        wn_linear("encoder.attention.W_q.main.0", H, H)
        wn_linear("encoder.attention.linear", 1, H, sharpen_att)
    else:                                          # ConcatAttention (attention.py:27-31)
        wn_linear("encoder.attention.sequence.0", H, V + H)
        # the concat logit layer sees un-gated ReLU features (much larger than the ⊙ product): a tenth
        # of the sharpening gives a comparably peaked softmax
        wn_linear("encoder.attention.sequence.2", 1, H, 0.1 * sharpen_att)
    wn_linear("encoder.q_net.main.0", H, H)
    if cfg.predictor == "q-cap":
        # PredictorwithCaption (predictor.py:148-184): bias-free nn.Linear + LeakyReLU stages and two GRUs
        def lrelu_linear(prefix, out_dim, in_dim, gain=1.0):
            w[prefix + ".main.0.weight"] = _uniform(g, (out_dim, in_dim), gain / math.sqrt(in_dim))

        def gru(prefix, in_dim):
            w[prefix + ".weight_ih_l0"] = _uniform(g, (3 * H, in_dim), kb)
            w[prefix + ".weight_hh_l0"] = _uniform(g, (3 * H, H), kb)
            w[prefix + ".bias_ih_l0"] = _uniform(g, (3 * H,), kb)
            w[prefix + ".bias_hh_l0"] = _uniform(g, (3 * H,), kb)

        # "trained-like" gains: the default init shrinks activations ~3x per bias-free layer and the softmax over
        # H divides by H again, which would leave every gate at its constant mid value (σ≈0.5, softmax≈uniform)
        lrelu_linear("predictor.v_net", H, V, 4.0)
        gru("predictor.caption_embedding.word_rnn.rnn", E)
        gru("predictor.caption_embedding.caption_rnn.rnn", H)
        lrelu_linear("predictor.caption_embedding.attention.W_v", H, H, 20.0)
        lrelu_linear("predictor.caption_embedding.attention.W_q", H, H, 100.0)
        lrelu_linear("predictor.caption_embedding.fcnet", H, H, 3.0)
        lrelu_linear("predictor.c_net", H, H, 3.0)
        lrelu_linear("predictor.vq_net", H, H, 2.0)
        lrelu_linear("predictor.joint_net", H, H, 60.0)
        lrelu_linear("predictor.vqc_net", H, H, 40.0)
        w["predictor.classifier.0.main.0.weight"] = _uniform(g, (A, H), 4.0 * sharpen_cls / math.sqrt(H))
    else:
        wn_linear("predictor.v_net.main.0", H, V)
        wn_linear("predictor.classifier.main.0", 2 * H, H)
        wn_linear("predictor.classifier.main.3", A, 2 * H, sharpen_cls)
        if cfg.predictor == "base-cap":
            # BaseCaptionPredictor (predictor.py:95-140): + a caption GRU (last state) and FCNet(H, H)
            p = "predictor.c_rnn.rnn"
            w[p + ".weight_ih_l0"] = _uniform(g, (3 * H, E), kb)
            w[p + ".weight_hh_l0"] = _uniform(g, (3 * H, H), kb)
            w[p + ".bias_ih_l0"] = _uniform(g, (3 * H,), kb)
            w[p + ".bias_hh_l0"] = _uniform(g, (3 * H,), kb)
            wn_linear("predictor.c_net.main.0", H, H)

    if cfg.relation:
        for i in range(cfg.conv_layer):
            p = f"gcn.{i}."
            bv = 1.0 / math.sqrt(V)
            w[p + "bias"] = _uniform(g, (cfg.num_labels, V), bv)
            for d in range(3):
                w[p + f"weight.{d}.weight"] = _uniform(g, (V, V), bv)
            for nm in ("wa", "wb"):
                w[p + f"dot_product.{nm}.weight"] = _uniform(g, (V, V), bv) * sharpen_gcn
                w[p + f"dot_product.{nm}.bias"] = _uniform(g, (V,), bv)

    if cfg.decoder == "base":
        # BaseDecoder (generator.py:155-166): nn.GRUCell(E+V -> Hd), its own attention over (v, h), nn.Linear(Hd, ntoken).
        # Drawn last so that every earlier tensor is the same with and without the caption head.
        Hd = cfg.decoder_hidden_dim
        kd = 1.0 / math.sqrt(Hd)
        w["generator.rnn.weight_ih"] = _uniform(g, (ng * Hd, E + V), kd)       # nn.GRUCell / nn.LSTMCell (rnn_type)
        w["generator.rnn.weight_hh"] = _uniform(g, (ng * Hd, Hd), kd)
        w["generator.rnn.bias_ih"] = _uniform(g, (ng * Hd,), kd)
        w["generator.rnn.bias_hh"] = _uniform(g, (ng * Hd,), kd)
        if cfg.att_type == "new":
            wn_linear("generator.attention.W_v.main.0", Hd, V)
            wn_linear("generator.attention.W_q.main.0", Hd, Hd)
            wn_linear("generator.attention.linear", 1, Hd, sharpen_att)
        else:
            wn_linear("generator.attention.sequence.0", Hd, V + Hd)
            wn_linear("generator.attention.sequence.2", 1, Hd, 0.1 * sharpen_att)
        w["generator.fcnet.weight"] = _uniform(g, (cfg.ntoken, Hd), 0.1 * sharpen_cls)     # generator.py:165 (±0.1)
        w["generator.fcnet.bias"] = _uniform(g, (cfg.ntoken,), 0.1)

    if cfg.relation and cfg.use_imp:
        # the implicit branch's own GCN (encoder.py:211-219), unregistered tensors like the spatial one: gcn_imp.{i}.*
        for i in range(cfg.conv_layer):
            p = f"gcn_imp.{i}."
            bv = 1.0 / math.sqrt(V)
            w[p + "bias"] = _uniform(g, (cfg.num_labels, V), bv)
            for d in range(3):
                w[p + f"weight.{d}.weight"] = _uniform(g, (V, V), bv)
            for nm in ("wa", "wb"):
                w[p + f"dot_product.{nm}.weight"] = _uniform(g, (V, V), bv) * sharpen_gcn
                w[p + f"dot_product.{nm}.bias"] = _uniform(g, (V,), bv)
    return w


def make_boxes(B: int, K: int, seed: int, W: int = 640, H: int = 480,
               grid: bool = True) -> np.ndarray:
    """Seeded synthetic boxes [B,K,4] float32 (x0,y0,x1,y1).  ``grid=True``:
    integer-grid boxes on which label parity is bit-exact;
    ``grid=False``: continuous-uniform boxes for the statistical gate."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if grid:
        x0 = torch.randint(0, W - 16, (B, K), generator=g).float()
        y0 = torch.randint(0, H - 16, (B, K), generator=g).float()
        bw = torch.randint(8, 201, (B, K), generator=g).float()
        bh = torch.randint(8, 201, (B, K), generator=g).float()
    else:
        x0 = torch.rand((B, K), generator=g) * (W - 16)
        y0 = torch.rand((B, K), generator=g) * (H - 16)
        bw = 8 + torch.rand((B, K), generator=g) * 192
        bh = 8 + torch.rand((B, K), generator=g) * 192
    x1 = torch.minimum(x0 + bw, torch.tensor(float(W - 1)))
    y1 = torch.minimum(y0 + bh, torch.tensor(float(H - 1)))
    return torch.stack([x0, y0, x1, y1], dim=2).numpy().astype(np.float32)


def make_batch(cfg: Config, B: int, seed: int, W: int = 640, H: int = 480) -> dict:
    """Seeded synthetic batch in the reference's wire format (dataset.py:96-104):
    img f32 [B,K,V] ~U[0,1); q int64 [B,T]; c int64 [B,c_len]; cap_len; a f64
    sparse soft scores; bbox f32 [B,K,4] (integer grid) and graph f64 [B,K,K]
    = relation_graph(bbox) when cfg.relation."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    batch = {
        "img": torch.rand((B, cfg.num_objs, cfg.v_dim), generator=g, dtype=torch.float32),
        "q": torch.randint(0, cfg.ntoken, (B, cfg.q_len), generator=g),
        "c": torch.randint(0, cfg.ntoken, (B, cfg.c_len), generator=g),
        "cap_len": torch.full((B,), cfg.c_len, dtype=torch.int64),
    }
    a = torch.zeros((B, cfg.ans_dim), dtype=torch.float64)
    idx = torch.randint(0, cfg.ans_dim, (B, 3), generator=g)
    val = torch.randint(1, 4, (B, 3), generator=g).double() / 3.0
    a.scatter_(1, idx, val)
    batch["a"] = a
    if cfg.relation:
        bbox = make_boxes(B, cfg.num_objs, seed + 7919, W, H, grid=True)
        batch["bbox"] = torch.from_numpy(bbox)
        batch["wh"] = (W, H)
        graph = np.stack([relation_graph(bbox[i], W, H) for i in range(B)])
        batch["graph"] = torch.from_numpy(graph)          # float64, like the loader
    return batch


def make_decoder_batch(cfg: Config, B: int, seed: int) -> dict:
    """make_batch with ragged caption lengths in [2, c_len] (the decoder stops one step before each
    caption's <end>, generator.py:93).  Distinct lengths while B < c_len, so that the descending sort
    of generator.py:76 has no ties to order; beyond that ties occur (stable order, see base_decoder_forward)."""
    batch = make_batch(cfg, B, seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 104729)
    if B < cfg.c_len:
        batch["cap_len"] = (torch.randperm(cfg.c_len - 1, generator=g)[:B] + 2).to(torch.int64)
    else:
        batch["cap_len"] = torch.randint(2, cfg.c_len + 1, (B,), generator=g)
    return batch


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------
def weight_norm_scale(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """s = g / ‖v‖_F — the scalar of weight_norm(dim=None) (modules.py:38;
    torch ``_weight_norm(v, g, -1)`` = ``v * (g / v.norm())``)."""
    return g / torch.norm(v)


def wn_weight(v, g):
    return v * weight_norm_scale(v, g)


def fcnet1(x, W, prefix):
    """1-layer FCNet: ReLU(Linear_wn(x)) (modules.py:35-38,55)."""
    wt = wn_weight(W[prefix + ".main.0.weight_v"], W[prefix + ".main.0.weight_g"])
    return torch.relu(F.linear(x, wt, W[prefix + ".main.0.bias"]))


def lstm_all(x, W, prefix, layer=0):
    """one nn.LSTM layer (batch_first, h0 = c0 = 0, modules.py:121-130,139-146), every time step.
    Gate order [i; f; g; o]; c' = σ(f)c + σ(i)tanh(g); h' = σ(o)tanh(c')."""
    w_ih, w_hh = W[prefix + f".weight_ih_l{layer}"], W[prefix + f".weight_hh_l{layer}"]
    b_ih, b_hh = W[prefix + f".bias_ih_l{layer}"], W[prefix + f".bias_hh_l{layer}"]
    B, T, _ = x.shape
    Hd = w_hh.shape[1]
    h = torch.zeros((B, Hd), dtype=x.dtype)
    c = torch.zeros((B, Hd), dtype=x.dtype)
    gi_all = F.linear(x, w_ih, b_ih)
    outs = []
    for t in range(T):
        gates = gi_all[:, t] + F.linear(h, w_hh, b_hh)
        i, f, g, o = gates[:, :Hd], gates[:, Hd:2 * Hd], gates[:, 2 * Hd:3 * Hd], gates[:, 3 * Hd:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)


def rnn_stack_last(x, W, prefix, rnn_type="GRU", n_layer=1):
    """SentenceEmbedding.forward (modules.py:155-159): stacked nn.GRU / nn.LSTM layers (each layer reads the full
    output sequence of the one below; inter-layer dropout is training-only), last time step of the top layer."""
    for l in range(n_layer):
        x = lstm_all(x, W, prefix, l) if rnn_type == "LSTM" else gru_all(x, W, prefix, l)
    return x[:, -1]


def gru_all(x, W, prefix, layer=0):
    """one nn.GRU layer (batch_first, h0=0), every time step: ``output`` [B,T,H] of
    SentenceEmbedding.forward_all (modules.py:147-152); the final hidden state is output[:, -1]."""
    w_ih, w_hh = W[prefix + f".weight_ih_l{layer}"], W[prefix + f".weight_hh_l{layer}"]
    b_ih, b_hh = W[prefix + f".bias_ih_l{layer}"], W[prefix + f".bias_hh_l{layer}"]
    B, T, _ = x.shape
    Hd = w_hh.shape[1]
    h = torch.zeros((B, Hd), dtype=x.dtype)
    gi_all = F.linear(x, w_ih, b_ih)
    outs = []
    for t in range(T):
        gi = gi_all[:, t]
        gh = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :Hd] + gh[:, :Hd])
        z = torch.sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
        n = torch.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
        h = (1.0 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs, 1)


def gru_last(x, W, prefix="encoder.q_rnn.rnn"):
    """nn.GRU(1 layer, batch_first, h0=0), last time step (modules.py:139-159).
    Gate order [r; z; n]; n uses r ⊙ (W_hn h + b_hn) (torch GRU semantics)."""
    w_ih, w_hh = W[prefix + ".weight_ih_l0"], W[prefix + ".weight_hh_l0"]
    b_ih, b_hh = W[prefix + ".bias_ih_l0"], W[prefix + ".bias_hh_l0"]
    B, T, _ = x.shape
    Hd = w_hh.shape[1]
    h = torch.zeros((B, Hd), dtype=x.dtype)
    gi_all = F.linear(x, w_ih, b_ih)                       # [B,T,3H]
    for t in range(T):
        gi = gi_all[:, t]
        gh = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :Hd] + gh[:, :Hd])
        z = torch.sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
        n = torch.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
        h = (1.0 - z) * n + z * h
    return h


def question_embedding(q_tokens, W):
    """embedding → RNN last state (encoder.py:159-160); cell type and depth are read off the parameter names/shapes."""
    emb = W["encoder.embedding.weight"][q_tokens]
    p = "encoder.q_rnn.rnn"
    n_layer = sum(1 for k in W if k.startswith(p + ".weight_hh_l"))
    lstm = W[p + ".weight_hh_l0"].shape[0] == 4 * W[p + ".weight_hh_l0"].shape[1]
    if n_layer == 1 and not lstm:
        return gru_last(emb, W)
    return rnn_stack_last(emb, W, p, "LSTM" if lstm else "GRU", n_layer)


def multiply_attention_logits(v, q, W, prefix="encoder.attention"):
    """MultiplyAttention.logits (attention.py:68-76), dropout = identity (eval)."""
    vp = fcnet1(v, W, prefix + ".W_v")                       # [B,K,H]
    qp = fcnet1(q, W, prefix + ".W_q").unsqueeze(1)          # [B,1,H]
    joint = vp * qp
    wl = wn_weight(W[prefix + ".linear.weight_v"], W[prefix + ".linear.weight_g"])
    return F.linear(joint, wl, W[prefix + ".linear.bias"])   # [B,K,1]


def multiply_attention(v, q, W, prefix="encoder.attention"):
    """MultiplyAttention.forward: softmax over the K regions (attention.py:78-86)."""
    return torch.softmax(multiply_attention_logits(v, q, W, prefix), dim=1)


def concat_attention_logits(v, q, W, prefix="encoder.attention"):
    """ConcatAttention.logits (attention.py:33-40): w2·ReLU(W1 [v;q] + b1) + b2."""
    vq = torch.cat((v, q.unsqueeze(1).repeat(1, v.size(1), 1)), 2)
    w1 = wn_weight(W[prefix + ".sequence.0.weight_v"], W[prefix + ".sequence.0.weight_g"])
    w2 = wn_weight(W[prefix + ".sequence.2.weight_v"], W[prefix + ".sequence.2.weight_g"])
    hid = torch.relu(F.linear(vq, w1, W[prefix + ".sequence.0.bias"]))
    return F.linear(hid, w2, W[prefix + ".sequence.2.bias"])          # [B,K,1]


def attention_logits(v, q, W, prefix="encoder.attention"):
    """set_att dispatch (attention.py:11-15) by the parameter names present"""
    if prefix + ".sequence.0.weight_v" in W:
        return concat_attention_logits(v, q, W, prefix)
    return multiply_attention_logits(v, q, W, prefix)


def base_encoder(batch, W):
    """BaseEncoder.base_forward (encoder.py:146-181) minus the caption keys."""
    v = batch["img"]
    q = question_embedding(batch["q"], W)
    v_att = torch.softmax(attention_logits(v, q, W), dim=1)           # attention.py:51,86
    v = v_att * v
    qn = fcnet1(q, W, "encoder.q_net")
    out = {"v": v, "q": qn, "v_att": v_att, "q_emb": q}
    if "c" in batch:
        out["c"] = W["encoder.embedding.weight"][batch["c"]]          # encoder.py:172 (embedded caption tokens)
        out["c_target"] = batch["c"]                                  # encoder.py:155,178
    if "cap_len" in batch:
        out["cap_len"] = batch["cap_len"]
    return out


def directed_conv(feature, graph, W, p):
    """DirectedGraphConv.conv (gcn.py:93-107): W2 f + adj·(W0 f) + adj·(W1 f)
    + Σ_j bias[label_ij]  (label 0, incl. the diagonal, contributes bias[0])."""
    adj = (graph != 0).to(feature.dtype)
    out = F.linear(feature, W[p + "weight.2.weight"])
    for d in range(2):
        out = out + torch.bmm(adj, F.linear(feature, W[p + f"weight.{d}.weight"]))
    labels = graph.long()
    return out + W[p + "bias"][labels].sum(2)


def relation_alpha(feature, adj, W, p):
    """CorrelatedGraphConv.relation_alpha (gcn.py:119-128; DotProduct
    modules.py:86-95): softmax over dim=1 (the ROW index i) of adj·ReLU(ABᵀ)."""
    a = F.linear(feature, W[p + "dot_product.wa.weight"], W[p + "dot_product.wa.bias"])
    b = F.linear(feature, W[p + "dot_product.wb.weight"], W[p + "dot_product.wb.bias"])
    alpha = torch.relu(torch.bmm(a, b.transpose(1, 2)))
    alpha = torch.bmm(adj, alpha)
    return torch.softmax(alpha, dim=1)


def corr_graph_conv(feature, graph, W, p):
    """CorrelatedGraphConv.forward (gcn.py:152-168) → (output, alpha)."""
    adj = (graph != 0).to(feature.dtype)
    conv = directed_conv(feature, graph, W, p)
    alpha = relation_alpha(feature, adj, W, p)
    return torch.bmm(alpha, conv), alpha


def gcn(feature, graph, W, n_layer=1, prefix="gcn"):
    """GCN.forward (gcn.py:199-215): per layer conv → dropout(identity) → ReLU."""
    alphas = []
    for i in range(n_layer):
        feature, alpha = corr_graph_conv(feature, graph, W, f"{prefix}.{i}.")
        alphas.append(alpha)
        feature = torch.relu(feature)
    return feature, alphas


def relation_encoder(batch, W, n_layer=1, use_imp=False, use_spa=True):
    """RelationEncoder.forward (encoder.py:236-272): the implicit branch (fully connected graph, ones - eye, its own GCN)
    and / or the spatial branch (batch['graph']), outputs summed; ``alpha`` = the last branch's (encoder.py:256,263)."""
    out = base_encoder(batch, W)
    v = out["v"]
    output_v = torch.zeros_like(v)
    alphas = []
    if use_imp:
        K = v.shape[1]
        g_imp = (torch.ones(K, K) - torch.eye(K)).to(v.dtype).repeat(v.shape[0], 1, 1)       # encoder.py:231-234,255
        new_v, alphas = gcn(v, g_imp, W, n_layer, prefix="gcn_imp")
        output_v = output_v + new_v
    if use_spa:
        new_v, alphas = gcn(v, batch["graph"].to(v.dtype), W, n_layer)
        output_v = output_v + new_v
    out["v"] = output_v
    out["alpha"] = alphas
    return out


def base_predictor(enc, W):
    """BasePredictor.forward (predictor.py:81-93): Σ_K v → v_net → ⊙q →
    classifier FCNet(H→2H→A, 2 layers, final ReLU modules.py:55)."""
    v = enc["v"].sum(1)
    v = fcnet1(v, W, "predictor.v_net")
    joint = enc["q"] * v
    p = "predictor.classifier.main"
    w0 = wn_weight(W[p + ".0.weight_v"], W[p + ".0.weight_g"])
    w3 = wn_weight(W[p + ".3.weight_v"], W[p + ".3.weight_g"])
    h = torch.relu(F.linear(joint, w0, W[p + ".0.bias"]))
    return torch.relu(F.linear(h, w3, W[p + ".3.bias"]))


def base_caption_predictor(enc, W):
    """BaseCaptionPredictor.forward (predictor.py:116-140): c = FCNet(GRU_last(embedded caption));
    joint = q ⊙ (c + v_net(Σ_K v)); same classifier as BasePredictor."""
    v = fcnet1(enc["v"].sum(1), W, "predictor.v_net")
    c = fcnet1(gru_last(enc["c"], W, "predictor.c_rnn.rnn"), W, "predictor.c_net")
    joint = enc["q"] * (c + v)
    p = "predictor.classifier.main"
    w0 = wn_weight(W[p + ".0.weight_v"], W[p + ".0.weight_g"])
    w3 = wn_weight(W[p + ".3.weight_v"], W[p + ".3.weight_g"])
    h = torch.relu(F.linear(joint, w0, W[p + ".0.bias"]))
    return torch.relu(F.linear(h, w3, W[p + ".3.bias"]))


def lrelu_net(x, W, prefix, slope):
    """LReLUNet (modules.py:62-78): LeakyReLU(Linear(x)), no bias."""
    return F.leaky_relu(F.linear(x, W[prefix + ".main.0.weight"]), slope)


def caption_embedding(v, q, c, W, prefix="predictor.caption_embedding"):
    """CaptionEmbedding.forward (modules.py:291-306) with the minimal repair of forward_all (SURVEY.md F8 /
    §8c): the word GRU's final state (not an undefined variable) gates its own outputs.
        out_w = GRU_word(c)                  [B,T,H]   (modules.py:292; all steps)
        a     = σ(h_w ⊙ LReLU(W_v v) + h_w ⊙ LReLU(W_q q)),  h_w = out_w[:, -1]   (CaptionAttention, modules.py:225-243;
                its LReLUNets use the default slope 0.01, dropout = identity in eval)
        out_c = GRU_cap(a[:,None,:] ⊙ out_w) [B,T,H]   (modules.py:294-295)
        c_emb = max_t LReLU(W_f out_c)       [B,H]     (modules.py:296,306)"""
    out_w = gru_all(c, W, prefix + ".word_rnn.rnn")
    h_w = out_w[:, -1]
    a = torch.sigmoid(h_w * lrelu_net(v, W, prefix + ".attention.W_v", 0.01)
                      + h_w * lrelu_net(q, W, prefix + ".attention.W_q", 0.01))
    out_c = gru_all(a.unsqueeze(1) * out_w, W, prefix + ".caption_rnn.rnn")
    return lrelu_net(out_c, W, prefix + ".fcnet", 0.01).max(dim=1)[0]


def qcap_predictor(enc, W, slope):
    """PredictorwithCaption.forward (predictor.py:186-213); returns the sigmoid outputs [B,A]."""
    V = lrelu_net(enc["v"], W, "predictor.v_net", slope)              # [B,K,H]   :188
    v = V.sum(1)                                                      # :191
    c = caption_embedding(v, enc["q"], enc["c"], W)                   # :192
    vq = lrelu_net(v, W, "predictor.vq_net", slope)                   # :196
    c = lrelu_net(c, W, "predictor.c_net", slope)                     # :197
    joint = torch.softmax(lrelu_net(c * vq, W, "predictor.joint_net", slope), 1)          # :201-202
    v = (joint.unsqueeze(1) * V).sum(1)                               # :203  (= joint ⊙ Σ_K V)
    v = lrelu_net(v, W, "predictor.vqc_net", slope)                   # :208
    joint = enc["q"] * (v + c)                                        # :209
    return torch.sigmoid(F.leaky_relu(F.linear(joint, W["predictor.classifier.0.main.0.weight"]), slope))   # :213


def gru_cell(x, h, W, prefix):
    """nn.GRUCell (generator.py:158-159): gate order [r; z; n], n = tanh(W_in x + b_in + r ⊙ (W_hn h + b_hn))."""
    Hd = h.shape[1]
    gi = F.linear(x, W[prefix + ".weight_ih"], W[prefix + ".bias_ih"])
    gh = F.linear(h, W[prefix + ".weight_hh"], W[prefix + ".bias_hh"])
    r = torch.sigmoid(gi[:, :Hd] + gh[:, :Hd])
    z = torch.sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
    n = torch.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
    return (1.0 - z) * n + z * h


def lstm_cell(x, h, c, W, prefix):
    """nn.LSTMCell: gate order [i; f; g; o]; c' = σ(f)c + σ(i)tanh(g); h' = σ(o)tanh(c')."""
    Hd = h.shape[1]
    gates = F.linear(x, W[prefix + ".weight_ih"], W[prefix + ".bias_ih"]) + F.linear(h, W[prefix + ".weight_hh"], W[prefix + ".bias_hh"])
    i, f, g, o = gates[:, :Hd], gates[:, Hd:2 * Hd], gates[:, 2 * Hd:3 * Hd], gates[:, 3 * Hd:]
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c), c


def base_decoder_step(v, prev, h, W, prefix="generator", c=None):
    """BaseDecoder.decode (generator.py:168-181), dropout = identity (eval):
    att = attention(v, h); att_v = Σ_K att·v; h' = GRUCell([prev; att_v], h); word = Linear(h').
    With ``c`` (rnn_type='LSTM'): (h', c') = LSTMCell([prev; att_v], (h, c)) and the return value is ((h', c'), word, att)."""
    att = torch.softmax(attention_logits(v, h, W, prefix + ".attention"), dim=1)
    att_v = (att * v).sum(1)
    x = torch.cat([prev, att_v], dim=1)
    if c is not None:
        h, c = lstm_cell(x, h, c, W, prefix + ".rnn")
        return (h, c), F.linear(h, W[prefix + ".fcnet.weight"], W[prefix + ".fcnet.bias"]), att
    h = gru_cell(x, h, W, prefix + ".rnn")
    return h, F.linear(h, W[prefix + ".fcnet.weight"], W[prefix + ".fcnet.bias"]), att


def base_decoder_forward(enc, W, cfg: Config, prefix="generator"):
    """DecoderModule.forward (generator.py:66-120), teacher forced: captions sorted by decreasing length
    (generator.py:76-79; ties keep their batch order — torch's CPU sort is stable), step t runs on the
    ``batch_t`` captions longer than t+1, and the per-step word logits / targets come back in
    pack_padded_sequence order (time-major, generator.py:117-118) without building the padded
    [B, max_len, ntoken] tensor.  ``enc``: encoder output with 'v', 'c' (embedded caption), 'c_target', 'cap_len'."""
    cap_len, sort_id = torch.sort(enc["cap_len"], dim=0, descending=True, stable=True)
    v, caption, target = enc["v"][sort_id], enc["c"][sort_id], enc["c_target"][sort_id]
    decode_len = (cap_len - 1).tolist()
    h = torch.zeros((v.shape[0], cfg.decoder_hidden_dim), dtype=v.dtype)
    c = torch.zeros_like(h) if cfg.rnn_type == "LSTM" else None
    predict, tgt, atts = [], [], []
    for t in range(max(decode_len)):
        bt = sum(l > t for l in decode_len)
        if c is not None:
            (h, c), word, att = base_decoder_step(v[:bt], caption[:bt, t], h[:bt], W, prefix, c=c[:bt])
        else:
            h, word, att = base_decoder_step(v[:bt], caption[:bt, t], h[:bt], W, prefix)
        predict.append(word)
        tgt.append(target[:bt, t + 1])                     # targets are the words after <start> (generator.py:115)
        atts.append(att)
    return {"predict": torch.cat(predict, 0), "target": torch.cat(tgt, 0), "att": atts}


def compute_score(predict, target):
    """compute_score (wrapper.py:8-22): lowest-index argmax → one-hot ⊙ target."""
    label = torch.max(predict, 1)[1]
    one_hot = torch.zeros_like(target)
    one_hot.scatter_(1, label.view(-1, 1), 1)
    return one_hot * target, label


def forward(batch, W, cfg: Config):
    """Wrapper.forward / get_att composition (wrapper.py:64-74,107-110)."""
    enc = (relation_encoder(batch, W, cfg.conv_layer, cfg.use_imp, cfg.use_spa) if cfg.relation
           else base_encoder(batch, W))
    if cfg.predictor == "q-cap":
        logits = qcap_predictor(enc, W, cfg.neg_slope)
    elif cfg.predictor == "base-cap":
        logits = base_caption_predictor(enc, W)
    else:
        logits = base_predictor(enc, W)
    return logits, enc


def forward_vqa(batch, W, cfg: Config):
    """Wrapper.forward_vqa (wrapper.py:113-118) → (score, label, target)."""
    logits, _ = forward(batch, W, cfg)
    target = batch["a"].to(logits.dtype)
    score, label = compute_score(logits, target)
    return score, label, target


def get_loss(batch, W, cfg: Config):
    """Wrapper.get_loss (wrapper.py:76-105) for the VQA head with dropout = identity:
    instance_bce_with_logits (wrapper.py:25-29) = mean BCE-with-logits × number of answers."""
    logits, _ = forward(batch, W, cfg)
    target = batch["a"].to(logits.dtype)
    loss = F.binary_cross_entropy_with_logits(logits, target) * target.size(1)
    return loss, logits


def loss_and_grads(batch, W, cfg: Config):
    """What train.py:103,108 leave in ``.grad`` after get_loss + backward (torch autograd over the
    restated forward; GCN tensors are not trainable in the reference)."""
    Wg = {k: (v.detach().clone().requires_grad_(True) if not k.startswith("gcn.") else v) for k, v in W.items()}
    loss, logits = get_loss(batch, Wg, cfg)
    loss.backward()
    grads = {k: v.grad for k, v in Wg.items() if not k.startswith("gcn.")}
    return loss.detach(), logits.detach(), grads


def to_dtype(W: dict, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in W.items()}


# ----------------------------------------------------------------------------
# spatial relation labels (util/relation.py)
# ----------------------------------------------------------------------------
def spatial_relation_pairs(a: np.ndarray, b: np.ndarray, w, h):
    """Vectorised spatial_relation (relation.py:3-45) over N box pairs.
    a, b: [N,4] float32.  Returns (label_ab, label_ba) uint8 [N].

    dtype flow mirrors the reference with float32 boxes: intersection, areas,
    IoU, centres, centre distance, arctan2/rad2deg/-90/%360//45 all float32;
    ‖(w,h)‖ is float64 (python ints) so the ``dist <= 0.5`` compare is float64.
    """
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    n = a.shape[0]
    I = np.stack([np.maximum(a[:, 0], b[:, 0]), np.maximum(a[:, 1], b[:, 1]),
                  np.minimum(a[:, 2], b[:, 2]), np.minimum(a[:, 3], b[:, 3])], axis=1)
    lab_ab = np.zeros(n, dtype=np.uint8)
    lab_ba = np.zeros(n, dtype=np.uint8)
    done = np.zeros(n, dtype=bool)

    inside = np.all(I == b, axis=1)                      # relation.py:24
    lab_ab[inside], lab_ba[inside] = 1, 2
    done |= inside
    covered = np.all(I == a, axis=1) & ~done             # relation.py:25
    lab_ab[covered], lab_ba[covered] = 2, 1
    done |= covered

    area = lambda x: (x[:, 3] - x[:, 1]) * (x[:, 2] - x[:, 0])   # unclamped (F6)
    with np.errstate(divide="ignore", invalid="ignore"):
        ai = area(I)
        iou = ai / (area(a) + area(b) - ai)
    overlap = (iou >= np.float32(0.5)) & ~done           # relation.py:30
    lab_ab[overlap], lab_ba[overlap] = 3, 3
    done |= overlap

    two = np.float32(2)
    ca = np.stack([a[:, 0] + (a[:, 2] - a[:, 0]) / two, a[:, 1] + (a[:, 3] - a[:, 1]) / two], 1)
    cb = np.stack([b[:, 0] + (b[:, 2] - b[:, 0]) / two, b[:, 1] + (b[:, 3] - b[:, 1]) / two], 1)
    d = ca - cb
    nrm = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])       # float32 (np.linalg.norm)
    dist = nrm.astype(np.float64) / np.linalg.norm([w, h])
    near = (dist <= 0.5) & ~done                         # relation.py:37-38
    e = cb - ca
    delta = np.rad2deg(np.arctan2(e[:, 0], e[:, 1])) - np.float32(90)   # relation.py:40
    idx = lambda x: (np.ceil((x % np.float32(360)) / np.float32(45)) + 3).astype(np.int64)
    i1 = idx(delta)
    i2 = idx(delta + np.float32(180))
    lab_ab[near] = i1[near]
    lab_ba[near] = i2[near]
    return lab_ab, lab_ba


def relation_graph(bbox: np.ndarray, w, h) -> np.ndarray:
    """relation_graph (relation.py:65-80): float64 [K,K]; one evaluation per
    unordered pair i<j writes [i,j] and [j,i]; the diagonal stays 0."""
    K = bbox.shape[0]
    iu, ju = np.triu_indices(K, 1)
    lab_ab, lab_ba = spatial_relation_pairs(bbox[iu], bbox[ju], w, h)
    out = np.zeros((K, K), dtype=np.float64)
    out[iu, ju] = lab_ab
    out[ju, iu] = lab_ba
    return out


def relation_graph_batch(bbox: np.ndarray, w, h) -> np.ndarray:
    """[B,K,4] → uint8 [B,K,K] (the on-device wire format of the label kernel)."""
    B, K, _ = bbox.shape
    iu, ju = np.triu_indices(K, 1)
    a = bbox[:, iu].reshape(-1, 4)
    b = bbox[:, ju].reshape(-1, 4)
    lab_ab, lab_ba = spatial_relation_pairs(a, b, w, h)
    out = np.zeros((B, K, K), dtype=np.uint8)
    out[:, iu, ju] = lab_ab.reshape(B, -1)
    out[:, ju, iu] = lab_ba.reshape(B, -1)
    return out
