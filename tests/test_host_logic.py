"""CPU tests of the host side: C-ABI surface, parameter names, weight preparation, the
comparison-only octant rule, sharding (incl. a world_size-2 gloo run).  No compute calls
into the library (there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import vqa_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from vqa_collection_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(libpath):
    from vqa_collection_b200 import _lib
    header = open(os.path.join(ROOT, "include", "vqa_b200.h")).read()
    declared = set(re.findall(r"\b(vqa_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(libpath)
    for name in declared:
        assert hasattr(lib, name), name
    lib.vqa_abi_version.restype = ctypes.c_int
    assert lib.vqa_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    """ctypes mirrors have one field per member of the C structs, in order"""
    from vqa_collection_b200 import _lib
    header = open(os.path.join(ROOT, "include", "vqa_b200.h")).read()
    for cname, cls in (("vqa_linear_args", _lib.LinearArgs), ("vqa_gru_args", _lib.GruArgs),
                       ("vqa_graph_attention_args", _lib.GraphAttentionArgs), ("vqa_forward_args", _lib.ForwardArgs),
                       ("vqa_train_args", _lib.TrainArgs), ("vqa_forward_host_args", _lib.ForwardHostArgs)):
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname + ";", header).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            for part in stmt.split(","):
                part = re.sub(r"\[[^\]]*\]", "", part)                # array members: p_v[VQA_TRAIN_LAYERS]
                names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
        assert names == [f[0] for f in cls._fields_], cname


def test_library_fails_loudly_without_gpu(libpath):
    from vqa_collection_b200 import _lib
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = lib.vqa_relation_labels(None, None, 1, 36, 640.0, 480.0, None, None)
    assert rc != 0 and len(lib.vqa_last_error()) > 0           # an error code, never a CPU fallback
    from vqa_collection_b200 import ops
    with pytest.raises(RuntimeError):
        ops.relation_labels(torch.zeros((1, 36, 4)), 640, 480)
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros((8, 64)), torch.zeros((8, 64)))


REF_KEYS = [
    "encoder.embedding.weight", "encoder.q_rnn.rnn.weight_ih_l0", "encoder.q_rnn.rnn.weight_hh_l0",
    "encoder.q_rnn.rnn.bias_ih_l0", "encoder.q_rnn.rnn.bias_hh_l0",
    "encoder.attention.W_v.main.0.bias", "encoder.attention.W_v.main.0.weight_g", "encoder.attention.W_v.main.0.weight_v",
    "encoder.attention.W_q.main.0.bias", "encoder.attention.W_q.main.0.weight_g", "encoder.attention.W_q.main.0.weight_v",
    "encoder.attention.linear.bias", "encoder.attention.linear.weight_g", "encoder.attention.linear.weight_v",
    "encoder.q_net.main.0.bias", "encoder.q_net.main.0.weight_g", "encoder.q_net.main.0.weight_v",
    "predictor.v_net.main.0.bias", "predictor.v_net.main.0.weight_g", "predictor.v_net.main.0.weight_v",
    "predictor.classifier.main.0.bias", "predictor.classifier.main.0.weight_g", "predictor.classifier.main.0.weight_v",
    "predictor.classifier.main.3.bias", "predictor.classifier.main.3.weight_g", "predictor.classifier.main.3.weight_v",
]


@pytest.mark.parametrize("enc", ["base", "relation"])
def test_drop_in_parameter_names_and_checkpoint_loading(enc):
    """state_dict keys are the reference's (SURVEY.md §8b probed listing) → old checkpoints load"""
    from vqa_collection_b200.modules.wrapper import set_model
    cfg = O.SMALL_REGAT if enc == "relation" else O.SMALL
    m = set_model(encoder_type=enc, predictor_type="base", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
                  embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, rnn_layer=1, ans_dim=cfg.ans_dim, cls_layer=2,
                  c_len=20, device="cpu", dropout=0.2, rnn_type="GRU", att_type="new", conv_layer=1, conv_type="corr")
    assert list(m.state_dict().keys()) == REF_KEYS               # GCN tensors stay unregistered (F3)
    W = O.make_weights(cfg, 1111)
    m.load_state_dict({k: v for k, v in W.items() if not k.startswith("gcn.")}, strict=True)
    if enc == "relation":
        layer = m.encoder.spatial_encoder.gcn[0]                 # indexable plain list like the reference
        assert sorted(layer.state_dict().keys()) == sorted(k[6:] for k in W if k.startswith("gcn.0."))
        layer.load_state_dict({k[6:]: v for k, v in W.items() if k.startswith("gcn.0.")}, strict=True)
        names = m.reference_named_weights()
        assert set(names) == set(W)
    assert m.train_step_supported() == (enc == "base")
    if enc == "relation":
        with pytest.raises(NotImplementedError):
            m.get_loss({})
    m.train()
    with pytest.raises(NotImplementedError):                     # training path must not silently run eval math
        m.encoder.q_net(torch.zeros(2, cfg.hidden_dim))


def test_unsupported_configurations_raise():
    from vqa_collection_b200.modules.wrapper import set_model
    from vqa_collection_b200.modules.attention import set_att, MultiplyAttention, ConcatAttention
    from vqa_collection_b200.modules.gcn import get_graph_conv, CorrelatedGraphConv
    assert set_att("new") is MultiplyAttention and set_att("base") is ConcatAttention
    assert get_graph_conv("corr") is CorrelatedGraphConv
    with pytest.raises(KeyError):
        set_att("nope")
    kw = dict(encoder_type="base", predictor_type="base", ntoken=10, v_dim=8, embed_dim=8, hidden_dim=8,
              decoder_hidden_dim=8, rnn_layer=1, ans_dim=4, cls_layer=2, c_len=5, device="cpu", att_type="new")
    with pytest.raises(NotImplementedError):                     # BUTDDecoder.decode returns None in the reference
        set_model(decoder_type="butd", **kw)
    m = set_model(decoder_type="base", rnn_type="LSTM", **kw)    # nn.LSTMCell caption head: 4 gate blocks, (h, c) state
    assert m.state_dict()["generator.rnn.weight_ih"].shape == (32, 16)
    (h0, c0), = m.generator.init_hidden(3)
    assert h0.shape == c0.shape == (3, 8)
    # LSTM / stacked question encoders (main.py:66,72): the reference's parameter names, never on the fused engine / train step
    m = set_model(decoder_type="none", rnn_type="LSTM", **{**kw, "rnn_layer": 2})
    names = [k for k in m.state_dict() if k.startswith("encoder.q_rnn.rnn.")]
    assert sorted(names) == sorted(f"encoder.q_rnn.rnn.{w}_l{l}" for l in (0, 1) for w in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
    assert m.state_dict()["encoder.q_rnn.rnn.weight_ih_l1"].shape == (32, 8) and m.engine() is None
    assert not m.train_step_supported()


def test_caption_head_parameter_names_and_checkpoint_loading():
    """decoder_type='base' (main.py:87 default): the generator's state_dict keys are the reference's (the oracle's
    weight dict, pinned by tests/golden/decoder_*.npz, loads strictly); use_mtl adds log_vars (wrapper.py:50)"""
    from vqa_collection_b200.modules.wrapper import set_model
    cfg = O.SMALL_DECODER
    kw = dict(encoder_type="base", predictor_type="base", decoder_type="base", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
              embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, decoder_hidden_dim=cfg.decoder_hidden_dim, rnn_layer=1,
              ans_dim=cfg.ans_dim, cls_layer=2, c_len=cfg.c_len, device="cpu", dropout=0.2, rnn_type="GRU", att_type="new")
    m = set_model(**kw)
    W = O.make_weights(cfg, 1111)
    assert sorted(m.state_dict().keys()) == sorted(W.keys())
    m.load_state_dict(W, strict=True)
    assert not m.use_mtl and not m.train_step_supported()
    assert [h.shape for h in m.generator.init_hidden(3)] == [(3, cfg.decoder_hidden_dim)]
    m2 = set_model(use_mtl=True, **kw)
    assert "log_vars" in m2.state_dict() and m2.use_mtl
    with pytest.raises(RuntimeError):                            # no CPU fallback behind the caption head either
        m.eval().generator.decode(torch.zeros(2, 36, cfg.v_dim), None, torch.zeros(2, cfg.embed_dim),
                                  m.generator.init_hidden(2))


def test_qcap_parameter_names_and_checkpoint_loading():
    """predictor_type='q-cap' (BASELINE config 5) keeps the reference's parameter names: the oracle's weight dict
    (keyed like the reference's state_dict, pinned by tests/golden/qcap_*.npz) loads strictly"""
    from vqa_collection_b200.modules.wrapper import set_model
    cfg = O.SMALL_QCAP
    W = O.make_weights(cfg, 1111)
    m = set_model(encoder_type="base", predictor_type="q-cap", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
                  embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, decoder_hidden_dim=0, rnn_layer=1,
                  ans_dim=cfg.ans_dim, cls_layer=2, c_len=cfg.c_len, device="cpu", dropout=0.2,
                  neg_slope=cfg.neg_slope, rnn_type="GRU", att_type="new")
    sd = m.state_dict()
    assert set(sd) == set(W), set(sd) ^ set(W)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(W[k].shape), k
    m.load_state_dict(W, strict=True)
    # predictor_type='base-cap' (predictor.py:95-140): the reference's names again; never on the fused Up-Down engine / train step
    cfg2 = O.SMALL_BASECAP
    m2 = set_model(encoder_type="base", predictor_type="base-cap", decoder_type="none", ntoken=cfg2.ntoken, v_dim=cfg2.v_dim,
                   embed_dim=cfg2.embed_dim, hidden_dim=cfg2.hidden_dim, decoder_hidden_dim=0, rnn_layer=1,
                   ans_dim=cfg2.ans_dim, cls_layer=2, c_len=cfg2.c_len, device="cpu", dropout=0.2, rnn_type="GRU", att_type="new")
    W2 = O.make_weights(cfg2, 1111)
    assert sorted(m2.state_dict().keys()) == sorted(W2.keys())
    m2.load_state_dict(W2, strict=True)
    assert not m2.train_step_supported()


def test_concat_attention_parameter_names():
    """att_type='base' keeps the reference's names (encoder.attention.sequence.{0,2}.*)"""
    from vqa_collection_b200.modules.wrapper import set_model
    from vqa_collection_b200.engine import prepare_weights
    cfg = O.SMALL_CONCAT
    m = set_model(encoder_type="base", predictor_type="base", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
                  embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, rnn_layer=1, ans_dim=cfg.ans_dim, cls_layer=2,
                  c_len=20, device="cpu", dropout=0.2, rnn_type="GRU", att_type="base", conv_layer=1, conv_type="corr")
    W = O.make_weights(cfg, 1111)
    assert set(m.state_dict().keys()) == set(W.keys())
    m.load_state_dict(W, strict=True)
    assert m.state_dict()["encoder.attention.sequence.0.weight_v"].shape == (cfg.hidden_dim, cfg.v_dim + cfg.hidden_dim)
    P = prepare_weights(W, torch.float32, "cpu", False)
    assert P["att_concat"] and P["Wv"].shape == (cfg.hidden_dim, cfg.v_dim) and P["W1q"].shape == (cfg.hidden_dim,) * 2
    assert torch.equal(torch.cat([P["Wv"], P["W1q"]], 1), W["encoder.attention.sequence.0.weight_v"])


def test_prepare_weights_layout():
    from vqa_collection_b200.engine import prepare_weights, weight_norm_scale
    cfg = O.SMALL_REGAT
    W = O.make_weights(cfg, 7)
    P = prepare_weights(W, torch.bfloat16, "cpu", True)
    H, V = cfg.hidden_dim, cfg.v_dim
    assert P["E_pad"] % 64 == 0 and P["emb"].shape == (cfg.ntoken + 1, P["E_pad"])
    assert torch.all(P["emb"][:, cfg.embed_dim:] == 0) and torch.all(P["w_ih"][:, cfg.embed_dim:] == 0)
    # bf16 + K=36 + V%128==0: merged ReGAT algebra [W0+W1 ; W2 ; WbᵀWa] (vqa_graph_attention layout 1)
    assert P["Wqq"].shape == (2 * H, H) and P["Wg3"].shape == (3 * V, V) and "Wg" not in P
    wa, wb = W["gcn.0.dot_product.wa.weight"].double(), W["gcn.0.dot_product.wb.weight"].double()
    ba, bb = W["gcn.0.dot_product.wa.bias"].double(), W["gcn.0.dot_product.wb.bias"].double()
    x = torch.rand((5, V), dtype=torch.float64)
    dot = (x @ wa.t() + ba) @ (x @ wb.t() + bb).t()                       # modules.py:92-95
    q = x @ (wb.t() @ wa).t()
    merged = q @ x.t() + (x @ (wa.t() @ bb))[:, None] + (x @ (wb.t() @ ba))[None, :] + ba.dot(bb)
    assert torch.allclose(dot, merged, rtol=1e-10, atol=1e-10)
    assert torch.equal(P["Wg3"][2 * V:], (wb.t() @ wa).float().to(torch.bfloat16))
    assert torch.equal(P["wvec"][0], (wa.t() @ bb).float().to(torch.bfloat16)) and torch.all(P["wvec"][2:] == 0)
    assert abs(P["gat_c0"] - float(ba.dot(bb))) < 1e-12 and P["label_bias_lp"].shape == (16, V)
    P32 = prepare_weights(W, torch.float32, "cpu", True)
    assert P32["Wg"].shape == (4 * V, V) and "Wg3" not in P32
    s = weight_norm_scale(W["encoder.attention.W_v.main.0.weight_v"], W["encoder.attention.W_v.main.0.weight_g"])
    assert torch.all(P["sv"] == s)
    # weight_norm scale comes from the same torch CPU op as the reference hook (H9)
    ref = (W["encoder.attention.W_v.main.0.weight_g"] / torch.norm(W["encoder.attention.W_v.main.0.weight_v"])).item()
    assert s == ref
    w01 = (W["gcn.0.weight.0.weight"] + W["gcn.0.weight.1.weight"]).to(torch.bfloat16)
    assert torch.equal(P["Wg3"][:V], w01)
    wl = W["encoder.attention.linear.weight_v"] * O.weight_norm_scale(W["encoder.attention.linear.weight_v"],
                                                                       W["encoder.attention.linear.weight_g"])
    assert torch.allclose(P["wlin"], wl.reshape(-1))


def _octant_rule(ex, ey):
    """numpy mirror of the comparison-only octant rule in csrc/relation.cu"""
    ab = np.zeros(ex.shape, np.uint8)
    ba = np.zeros(ex.shape, np.uint8)

    def put(m, a, b):
        ab[m], ba[m] = a, b
    pos, neg, zx = ex > 0, ex < 0, ex == 0
    put(pos & (ey == 0), 3, 7)
    put(pos & (ey < 0) & (-ey <= ex), 4, 8)
    put(pos & (ey < 0) & (-ey > ex), 5, 9)
    put(pos & (ey > 0) & (ex <= ey), 10, 6)
    put(pos & (ey > 0) & (ex > ey), 11, 7)
    put(neg & (ey == 0), 7, 3)
    put(neg & (ey < 0) & (ex >= ey), 6, 10)
    put(neg & (ey < 0) & (ex < ey), 7, 11)
    put(neg & (ey > 0) & (ey <= -ex), 8, 4)
    put(neg & (ey > 0) & (ey > -ex), 9, 5)
    put(zx & (ey < 0), 5, 9)
    put(zx & (ey >= 0), 9, 5)
    return ab, ba


def test_octant_rule_equals_arctan2_pipeline_on_grid():
    """the rule the kernel uses == the reference's float32 arctan2 pipeline on half-integer offsets
    (exact boundaries hit often, never merely approached: SURVEY.md H2)"""
    r = np.arange(-200, 201, dtype=np.float32) * 0.5
    ex, ey = [a.ravel() for a in np.meshgrid(r, r)]
    delta = np.rad2deg(np.arctan2(ex, ey)) - np.float32(90)
    idx = lambda x: (np.ceil((x % np.float32(360)) / np.float32(45)) + 3).astype(np.uint8)
    ab, ba = _octant_rule(ex, ey)
    assert np.array_equal(ab, idx(delta)) and np.array_equal(ba, idx(delta + np.float32(180)))


def _iou_rule(ai, d):
    """the division-free form of `fl32(ai/d) >= 0.5` that csrc/relation.cu evaluates (float32 ops only)"""
    ai, d = ai.astype(np.float32), d.astype(np.float32)
    x = ai - np.float32(0.5) * d
    t = d * np.float32(-2.0 ** -26)
    return np.where(d > 0, x >= t, np.where(d < 0, x <= t, (d == 0) & (ai > 0)))


def test_division_free_iou_rule_equals_float32_division():
    """relation.py:28-30 compares the float32 quotient with 0.5; the kernel never divides"""
    rng = np.random.default_rng(5)
    n = 2_000_000
    # d of either sign over many binades, ai clustered within a few ulp of d/2 plus a broad spread
    d = (rng.standard_normal(n) * np.exp2(rng.integers(-8, 24, n))).astype(np.float32)
    half = (np.float32(0.5) * d).astype(np.float32)
    near = half.copy()
    for _ in range(4):                                   # walk up to +-4 ulp away from d/2
        step = rng.integers(-1, 2, n)
        near = np.where(step > 0, np.nextafter(near, np.float32(np.inf)),
                        np.where(step < 0, np.nextafter(near, np.float32(-np.inf)), near)).astype(np.float32)
    broad = (d * rng.uniform(-2, 2, n).astype(np.float32)).astype(np.float32)
    ai = np.where(rng.random(n) < 0.7, near, broad).astype(np.float32)
    d[:1000] = 0.0                                       # +-inf / NaN quotients
    ai[:300] = 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = (ai / d) >= np.float32(0.5)
    got = _iou_rule(ai, d)
    assert np.array_equal(ref, got), int((ref != got).sum())
    assert 0.05 < ref.mean() < 0.95                      # the sample straddles the threshold


def test_near_threshold_equals_sqrt_divide_compare(libpath):
    """relation.py:37-38: float32 norm, float64 divide, `<= 0.5`; the kernel compares the squared distance
    with one precomputed float32 (vqa_relation_near_threshold)"""
    lib = ctypes.CDLL(libpath)
    lib.vqa_relation_near_threshold.restype = ctypes.c_float
    lib.vqa_relation_near_threshold.argtypes = [ctypes.c_float, ctypes.c_float]
    for w, h in [(640, 480), (500, 375), (1, 1), (7, 3), (1920, 1080), (333.5, 250.25), (3, 4), (1e-3, 2e-3)]:
        s_max = np.float32(lib.vqa_relation_near_threshold(w, h))
        diag = np.linalg.norm([float(np.float32(w)), float(np.float32(h))])
        s = np.full(65, s_max, dtype=np.float32)
        for k in range(32):                              # 32 floats either side of the threshold
            s[33 + k:] = np.nextafter(s[33 + k:], np.float32(np.inf))
            s[:32 - k] = np.nextafter(s[:32 - k], np.float32(-np.inf))
        ref = (np.sqrt(s).astype(np.float64) / diag) <= 0.5
        assert np.array_equal(ref, s <= s_max), (w, h)
        assert ref[32] and not ref[33]


def test_shard_bounds_cover_rows_exactly():
    from vqa_collection_b200.parallel import shard_bounds, shard_batch
    for n in (0, 1, 7, 1024, 1027):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    batch = O.make_batch(O.SMALL_REGAT, 10, 3)
    s = shard_batch(batch, 4, 1)
    assert s["img"].shape[0] == 3 and s["graph"].shape[0] == 3 and s["wh"] == batch["wh"]
    assert torch.equal(s["q"], batch["q"][3:6])


GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import vqa_oracle as O
from vqa_collection_b200.parallel import shard_batch, gather_rows, reduce_score
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = O.SMALL
W = O.make_weights(cfg, 1111)
batch = O.make_batch(cfg, 11, 5)                   # ragged: 6 + 5 rows
with torch.no_grad():
    full_score, full_label, _ = O.forward_vqa(batch, W, cfg)
    local = shard_batch(batch, world, rank)
    score, label, _ = O.forward_vqa(local, W, cfg)   # the CPU oracle stands in for the per-rank engine
labels = gather_rows(label, 11)
total = reduce_score(score.sum())
assert torch.equal(labels, full_label), (labels, full_label)
assert torch.allclose(total, full_score.sum())
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_sharded_forward_equals_unsharded(tmp_path):
    """world_size-2 gloo run: sharded result == unsharded result, no data-path collective"""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


GLOO_TRAIN_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import vqa_oracle as O
from vqa_collection_b200.parallel import shard_batch, FlatGradients, average_gradients_
from vqa_collection_b200.training import param_names, GRU_PARAMS
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = O.SMALL
W = O.make_weights(cfg, 1111)
batch = O.make_batch(cfg, 12, 5)                    # equal shards: 6 + 6 rows
_, _, full = O.loss_and_grads(batch, W, cfg)        # single-process gradients on the whole batch
_, _, local = O.loss_and_grads(shard_batch(batch, world, rank), W, cfg)   # the CPU oracle stands in for the per-rank step
names = param_names()
fg = FlatGradients([tuple(W[n].shape) for n in names], "cpu")
for v, n in zip(fg.views, names):
    v.copy_(local[n])
# the exchange step (SUM + divide under gloo, AVG under NCCL) in the TWO buckets of training.py: the weight-normed layers
# (final before the BPTT) and then GRU + embedding — contiguous ranges of the flat buffer that together cover it
head_off = fg.offset_of(len(GRU_PARAMS))
assert 0 < head_off < fg.flat.numel() and fg.offset_of(len(names)) == fg.flat.numel()
assert fg.views[len(GRU_PARAMS)].data_ptr() == fg.flat[head_off:].data_ptr()
average_gradients_(fg.flat[head_off:])
average_gradients_(fg.flat[:head_off])
gmax = max(float(full[n].abs().max()) for n in names)
for v, n in zip(fg.views, names):
    ref = full[n]
    if n == "encoder.attention.linear.bias":       # identically 0 up to rounding: the softmax is shift invariant
        assert float(v.abs().max()) < 1e-5 * gmax and float(ref.abs().max()) < 1e-5 * gmax
        continue
    assert torch.allclose(v, ref, rtol=1e-4, atol=1e-6 * float(ref.abs().max()) + 1e-12), (n, float((v - ref).abs().max()))
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gradient_average_equals_single_process(tmp_path):
    """config 4 exchange step on CPU: averaged shard gradients == gradients of the concatenated batch"""
    script = tmp_path / "worker_train.py"
    script.write_text(GLOO_TRAIN_WORKER)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29532", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_host_pack_is_round_to_nearest_even(libpath):
    """the e2e path's host-side f32→bf16 packing (host_pack.cpp) == torch's RNE cast, bit for bit"""
    lib = ctypes.CDLL(libpath)
    lib.vqa_packpool_create.restype = ctypes.c_void_p
    lib.vqa_packpool_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    pool = ctypes.c_void_p(lib.vqa_packpool_create(3))
    g = np.random.default_rng(0)
    x = (g.standard_normal(1 << 18) * 10.0 ** g.integers(-30, 30, 1 << 18)).astype(np.float32)
    x = np.concatenate([x, np.array([0, -0.0, np.inf, -np.inf, 1e-45, 3.4e38, 1.00390625, 1.01171875, 1.005859375], np.float32),
                        g.random(77).astype(np.float32)])
    y = np.empty(x.shape, np.uint16)
    lib.vqa_packpool_run(pool, x.ctypes.data, y.ctypes.data, x.size)
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(y, ref)
    lib.vqa_packpool_destroy.argtypes = [ctypes.c_void_p]
    lib.vqa_packpool_destroy(pool)


def test_flat_gradient_views():
    from vqa_collection_b200.parallel import FlatGradients
    shapes = [(3, 5), (), (7,), (2, 2, 2)]
    fg = FlatGradients(shapes, "cpu")
    assert [tuple(v.shape) for v in fg.views] == shapes and fg.matches(shapes, "cpu") and not fg.matches(shapes[:2], "cpu")
    fg.views[2].fill_(3.0)
    assert fg.flat.sum() == 21.0 and all(v.data_ptr() % 256 == fg.flat.data_ptr() % 256 for v in fg.views)


def test_host_path_tuning_scales_with_cores_per_rank(monkeypatch):
    """the packing pool and the packed/raw split of the host path follow this rank's share of the host cores"""
    from vqa_collection_b200 import engine
    assert [engine.host_raw_chunk_period(c) for c in (32, 16, 8, 4, 2, 1)] == [0, 4, 3, 2, 1, 1]
    n = engine.host_cores_per_rank()
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert engine.host_cores_per_rank() == max(1, n // 8)
    monkeypatch.setenv("VQA_B200_PACK_THREADS", "3")
    assert engine.host_pack_threads() == 3
    monkeypatch.delenv("VQA_B200_PACK_THREADS")
    assert engine.host_pack_threads() == max(1, n // 8)


def _write_glove(path, n_words, dim, seed=3):
    g = torch.Generator().manual_seed(seed)
    table = torch.randn((n_words, dim), generator=g)
    with open(path, "w") as f:
        for i in range(n_words):
            f.write(f"w{i} " + " ".join(f"{x:.6f}" for x in table[i].tolist()) + "\n")
    return table


def test_pretrained_word_embedding_drop_in(tmp_path):
    """main.py:83 default: a GloVe text file replaces encoder.embedding (encoder.py:56-57, modules.py:166-199) — rows in
    file order + 4 zero rows, frozen, NOT in the state_dict"""
    from vqa_collection_b200.modules.wrapper import set_model
    from vqa_collection_b200.modules.modules import PretrainedWordEmbedding
    path = str(tmp_path / "glove.txt")
    table = _write_glove(path, 12, 8)
    m = set_model(encoder_type="base", predictor_type="base", decoder_type="none", ntoken=15, v_dim=16, embed_dim=8,
                  hidden_dim=8, rnn_layer=1, ans_dim=4, cls_layer=2, c_len=5, device="cpu", att_type="new",
                  pretrained_embed_path=path)
    emb = m.encoder.embedding
    assert isinstance(emb, PretrainedWordEmbedding) and (emb.vocab_len, emb.vocab_dim) == (16, 8)
    assert "encoder.embedding.weight" not in m.state_dict() and not list(emb.parameters())
    assert torch.allclose(emb.vocab[:12], table, atol=1e-6) and float(emb.vocab[12:].abs().max()) == 0.0
    s = torch.tensor([[0, 11, 15], [3, 3, 12]])
    out = emb(s)
    assert out.shape == (2, 3, 8) and torch.equal(out, emb.vocab[s])
    assert "encoder.embedding.weight" in m.reference_named_weights()          # the fused engine still gets the table


def test_fused_optimizer_host_side():
    """constructor checks mirror torch.optim.Adamax; CPU tensors raise (no CPU fallback); nothing to do without grads"""
    from vqa_collection_b200 import optim
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(ValueError):
        optim.Adamax([p], lr=-1.0)
    with pytest.raises(ValueError):
        optim.Adamax([p], betas=(1.0, 0.999))
    o = optim.Adamax([{"params": [p]}], lr=0.002)
    assert o.defaults == dict(lr=0.002, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    o.step()                                                            # no gradients: no library call
    assert float(optim.clip_grad_norm_([p], 0.25)) == 0.0
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        o.step()
    with pytest.raises(RuntimeError):
        optim.clip_grad_norm_([p], 0.25)
    with pytest.raises(NotImplementedError):
        optim.clip_grad_norm_([p], 0.25, norm_type=1)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` needs no GPU: the reference's own CPU path (baseline/_ref, staged by build(); the oracle
    port only when that copy is absent) on the full 1024-question batch, ONE JSON line on stdout with the keys the driver
    reads (metric / value / unit / config of the b200 arm, impl, cpu_baseline, e2e with zero copy bytes) and the ReGAT block"""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(out.stdout.strip().splitlines()) == 1, out.stdout[:500]          # nothing but the JSON line on stdout
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "questions/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("VQA forward questions/sec") and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["config"]["workload"].startswith("Up-Down VQA forward") and line["config"]["batch_per_gpu"] == 1024
    cb = line["cpu_baseline"]
    staged = os.path.isdir(os.path.join(root, "baseline", "_ref", "modules"))
    assert cb["kind"] == ("reference" if staged else "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "1024 questions per step" in cb["sample"] and line["same_config"] is True
    rg = line["regat"]
    assert rg["value"] > 0 and rg["config"]["workload"].startswith("ReGAT") and rg["cpu_baseline"]["kind"] == cb["kind"]
    assert line["e2e"] == {"value": line["value"], "unit": "questions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_fp16_plane_pair_format_and_three_product_rule():
    """The number format of precision 'fp32tc' (VQA_F16X2), restated in numpy: x = hi + lo'·2^-11 with hi = fp16(x),
    lo' = fp16((x - hi)·2^11).  (1) the pair carries >= 21 significant bits wherever hi is a normal fp16 number, and never
    loses the residual to fp16's subnormals (which an unscaled residual does); (2) the three products the kernels issue —
    hi·hi + 2^-11 (hi·lo' + lo'·hi), summed in fp32 — match a float64 dot product to fp32-sum accuracy, two orders of
    magnitude closer than the bf16 operands of the default mode; (3) the token table of the GRU is an exact regrouping of
    W_ih·x_t + b_ih (gather of a GEMM == GEMM of a gather)."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(200000) * np.logspace(-4, 3, 200000)).astype(np.float32)
    hi = x.astype(np.float16)
    lo = ((x - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    back = hi.astype(np.float64) + lo.astype(np.float64) / 2048.0
    normal = np.abs(x) > 6.2e-5
    rel = np.abs(back - x.astype(np.float64))[normal] / np.abs(x.astype(np.float64))[normal]
    assert rel.max() < 2.0 ** -21
    lo_unscaled = (x - hi.astype(np.float32)).astype(np.float16)                 # what the scale factor avoids
    back_u = hi.astype(np.float64) + lo_unscaled.astype(np.float64)
    small = normal & (np.abs(x) < 1e-2)
    assert (np.abs(back_u - x)[small] / np.abs(x)[small]).max() > 16 * rel[np.abs(x[normal]) < 1e-2].max()
    # three-product rule on a GEMM of the path's depth
    K, M, N = 2048, 64, 48
    A = np.maximum(rng.standard_normal((M, K)), 0).astype(np.float32)           # post-ReLU activations
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)

    def split(t):
        h = t.astype(np.float16)
        return h.astype(np.float32), ((t - h.astype(np.float32)) * np.float32(2048.0)).astype(np.float16).astype(np.float32)
    ah, al = split(A)
    wh, wl = split(W)
    got = (ah @ wh.T) + (ah @ wl.T + al @ wh.T) * np.float32(2.0 ** -11)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    scale = np.abs(ref).max()
    e_split = np.abs(got - ref).max() / scale
    e_f32 = np.abs(A @ W.T - ref).max() / scale
    bf = lambda t: torch.from_numpy(t).to(torch.bfloat16).float().numpy()
    e_bf16 = np.abs(bf(A) @ bf(W).T - ref).max() / scale
    assert e_split < 4 * max(e_f32, 2e-7) and e_split < 1e-6 and e_bf16 > 100 * e_split
    # token table: W_ih·emb[token] + b == (emb·W_ihᵀ + b)[token]
    emb = rng.standard_normal((50, 24)).astype(np.float32)
    w_ih = rng.standard_normal((36, 24)).astype(np.float32)
    b = rng.standard_normal(36).astype(np.float32)
    tok = rng.integers(0, 50, (7, 5))
    table = emb @ w_ih.T + b
    assert np.array_equal(table[tok], (emb[tok].reshape(-1, 24) @ w_ih.T + b).reshape(7, 5, 36))
