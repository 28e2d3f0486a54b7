"""End-to-end GPU parity: VQAEngine (one C call per forward) against the CPU oracle and
the committed golden outputs of the real reference."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import vqa_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine_mod():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from vqa_collection_b200 import engine
    return engine


def relerr(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def margin_ok_mask(ref_logits, err_abs):
    """samples whose reference top-1/top-2 gap exceeds 4x the observed logit error (H1b)"""
    s = np.sort(ref_logits, 1)
    return (s[:, -1] - s[:, -2]) > 4.0 * err_abs


GOLDEN = ["updown_small", "regat_small", "updown_full", "regat_full", "concat_small", "concat_full"]


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp32tc"])
@pytest.mark.parametrize("name", GOLDEN)
def test_engine_matches_reference_golden(engine_mod, golden_dir, name, precision):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    cfg = O.Config(**meta["cfg"])
    W = O.make_weights(cfg, meta["wseed"])
    batch = O.make_batch(cfg, meta["B"], meta["bseed"])
    eng = engine_mod.VQAEngine(W, relation=cfg.relation, precision=precision)
    kw = dict(labels=batch["graph"].cuda(), want_alpha=True) if cfg.relation else {}
    exact = precision in ("fp32", "fp32tc")                    # fp32tc: fp32-class arithmetic on the tensor cores, same gates
    out = eng.forward(batch["img"].cuda(), batch["q"].cuda(), want_v=precision != "fp32tc", want_q=True, **kw)
    tol = 1e-5 if exact else 1e-2
    assert relerr(out["att"], z["v_att"]) < tol
    assert relerr(out["q"], z["q"]) < tol
    if precision != "fp32tc":                                  # (that engine does not materialise the encoder output 'v')
        assert relerr(out["v"].float()[:, :, ::16], z["v_sub"]) < tol
    assert relerr(out["logits"], z["logits"]) < tol
    if cfg.relation:
        assert relerr(out["alpha"], z["alpha"]) < tol
    label = out["label"].cpu().numpy()
    if exact:
        assert np.array_equal(label, z["label"])                 # bit-exact answers
    else:
        err = np.abs(out["logits"].cpu().numpy() - z["logits"]).max()
        ok = margin_ok_mask(z["logits"], err)
        assert np.array_equal(label[ok], z["label"][ok])
    # the kernel's own argmax rule: lowest index among its own maxima
    assert torch.equal(out["label"].cpu(), torch.max(out["logits"].cpu(), 1)[1])
    assert eng.last_launches > 0


@pytest.mark.parametrize("precision,relation,B", [("fp32", False, 64), ("bf16", False, 256),
                                                   ("fp32", True, 16), ("bf16", True, 64)])
def test_engine_matches_oracle_full_dims(engine_mod, precision, relation, B):
    cfg = O.FULL_REGAT if relation else O.FULL
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, B, 4321)
    with torch.no_grad():
        ref_logits, enc = O.forward(batch, W, cfg)
        _, ref_label, _ = O.forward_vqa(batch, W, cfg)
    eng = engine_mod.VQAEngine(W, relation=relation, precision=precision)
    kw = {}
    if relation:
        # labels computed on device from the boxes (k7) inside the same forward
        kw = dict(bbox=batch["bbox"].cuda(), wh=batch["wh"], want_alpha=True)
    out = eng.forward(batch["img"].cuda(), batch["q"].cuda(), **kw)
    tol = 1e-5 if precision == "fp32" else 1e-2
    assert relerr(out["att"], enc["v_att"][:, :, 0]) < tol
    assert relerr(out["logits"], ref_logits) < tol
    if relation:
        assert np.array_equal(out["labels"].cpu().numpy(), batch["graph"].numpy().astype(np.uint8))
        assert relerr(out["alpha"], enc["alpha"][0]) < tol
    label = out["label"].cpu().numpy()
    ref_label = ref_label.numpy()
    if precision == "fp32":
        assert np.array_equal(label, ref_label)
    else:
        err = np.abs(out["logits"].cpu().numpy() - ref_logits.numpy()).max()
        ok = margin_ok_mask(ref_logits.numpy(), err)
        assert ok.mean() > 0.5
        assert np.array_equal(label[ok], ref_label[ok])


def test_engine_sharded_equals_unsharded(engine_mod):
    """Data-parallel sharding is exact: rows [r*B/N, (r+1)*B/N) of the unsharded result."""
    cfg = O.FULL
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, 256, 99)
    eng = engine_mod.VQAEngine(W, relation=False, precision="bf16")
    img, q = batch["img"].cuda(), batch["q"].cuda()
    full = eng.forward(img, q)
    full_logits = full["logits"].clone()
    for r in range(2):
        part = eng.forward(img[r * 128:(r + 1) * 128], q[r * 128:(r + 1) * 128])
        assert torch.equal(part["logits"], full_logits[r * 128:(r + 1) * 128])


@pytest.mark.parametrize("pack_on_host", [True, False])
@pytest.mark.parametrize("relation", [False, True])
def test_engine_host_path(engine_mod, relation, pack_on_host):
    """vqa_forward_host: wire-format host buffers → answers; host bf16 packing is bit-identical to the device cast"""
    cfg = O.FULL_REGAT if relation else O.FULL
    W = O.make_weights(cfg, 1111)
    B = 300 if not relation else 70                       # ragged last chunk
    batch = O.make_batch(cfg, B, 98)
    eng = engine_mod.VQAEngine(W, relation=relation, precision="bf16")
    kw = dict(labels=batch["graph"].cuda()) if relation else {}
    dev = eng.forward(batch["img"].cuda(), batch["q"].cuda(), **kw)
    hkw = dict(labels_h=batch["graph"].to(torch.uint8)) if relation else {}
    for _ in range(2):                                    # second call reuses the context's buffers
        label_h, h2d, d2h = eng.forward_host(batch["img"].pin_memory(), batch["q"].pin_memory(), chunk=64,
                                             pack_on_host=pack_on_host, **hkw)
        assert torch.equal(label_h, dev["label"].cpu())
        assert torch.equal(eng.last_host_outputs["logits"], dev["logits"])
    if pack_on_host:                                      # hybrid staging: every 3rd chunk raw f32 + device cast
        label_x, h2d_x, _ = eng.forward_host(batch["img"].pin_memory(), batch["q"].pin_memory(), chunk=32,
                                             raw_chunk_period=3, **hkw)
        assert torch.equal(label_x, dev["label"].cpu()) and torch.equal(eng.last_host_outputs["logits"], dev["logits"])
        assert h2d_x > h2d
        a1 = eng.forward_host_async(batch["img"], batch["q"], **hkw)          # two batches in flight
        a2 = eng.forward_host_async(batch["img"], batch["q"], **hkw)
        assert torch.equal(a1.result()[0], dev["label"].cpu()) and torch.equal(a2.result()[0], dev["label"].cpu())
    per = 2 if pack_on_host else 4
    assert h2d == batch["img"].numel() * per + batch["q"].numel() * 8 + (B * 36 * 36 if relation else 0) and d2h == B * 8
    if relation:                                          # boxes instead of precomputed labels
        label_b, _, _ = eng.forward_host(batch["img"], batch["q"], bbox_h=batch["bbox"], wh=batch["wh"], pack_on_host=pack_on_host)
        assert torch.equal(label_b, dev["label"].cpu())


def test_engine_host_path_fp32(engine_mod):
    cfg = O.SMALL
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, 33, 98)
    eng = engine_mod.VQAEngine(W, relation=False, precision="fp32")
    dev = eng.forward(batch["img"].cuda(), batch["q"].cuda())
    label_h, h2d, _ = eng.forward_host(batch["img"], batch["q"], chunk=8)
    assert torch.equal(label_h, dev["label"].cpu()) and h2d == batch["img"].numel() * 4 + batch["q"].numel() * 8


def test_engine_full_batch_properties(engine_mod):
    """BASELINE size (B=1024): determinism, non-negativity (final ReLU, F7), attention rows sum to 1."""
    cfg = O.FULL
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, 1024, 7)
    eng = engine_mod.VQAEngine(W, relation=False, precision="bf16")
    img, q = eng.resident(batch["img"].cuda()), batch["q"].cuda()
    a = eng.forward(img, q)
    la, aa = a["logits"].clone(), a["att"].clone()
    b = eng.forward(img, q)
    assert torch.equal(la, b["logits"]) and torch.equal(aa, b["att"])
    assert float(la.min()) >= 0.0
    assert torch.allclose(aa.sum(1), torch.ones(1024, device="cuda"), atol=1e-5)
    assert torch.equal(b["label"], torch.max(b["logits"], 1)[1])


def test_regat_full_batch_properties(engine_mod):
    """BASELINE config 3 size (ReGAT, B=1024), size-independent properties: on-device labels == the oracle's relation graphs
    for every image, determinism, alpha columns sum to 1 (softmax over the ROW index, gcn.py:125), and the 8-way
    data-parallel split (128 questions per GPU) reproduces its rows of the unsharded result bit for bit."""
    cfg = O.FULL_REGAT
    W = O.make_weights(cfg, 1111)
    B = 1024
    g = torch.Generator().manual_seed(77)
    img = torch.rand((B, cfg.num_objs, cfg.v_dim), generator=g).to(torch.bfloat16).cuda()
    q = torch.randint(0, cfg.ntoken, (B, cfg.q_len), generator=g).cuda()
    boxes = O.make_boxes(B, cfg.num_objs, 78, 640, 480, grid=True)
    bbox = torch.from_numpy(boxes).cuda()
    eng = engine_mod.VQAEngine(W, relation=True, precision="bf16")
    a = eng.forward(img, q, bbox=bbox, wh=(640, 480), want_alpha=True)
    logits, alpha, labels, answers = a["logits"].clone(), a["alpha"].clone(), a["labels"].clone(), a["label"].clone()
    assert np.array_equal(labels.cpu().numpy(), O.relation_graph_batch(boxes, 640, 480).astype(np.uint8))
    b = eng.forward(img, q, bbox=bbox, wh=(640, 480), want_alpha=True)
    assert torch.equal(b["logits"], logits) and torch.equal(b["alpha"], alpha) and torch.equal(b["label"], answers)
    assert float(logits.min()) >= 0.0
    assert torch.allclose(alpha.sum(1), torch.ones((B, cfg.num_objs), device="cuda"), atol=1e-4)
    assert torch.equal(answers, torch.max(logits, 1)[1])
    for r in (0, 3, 7):                                                   # ranks of an 8-GPU job
        lo, hi = r * 128, (r + 1) * 128
        part = eng.forward(img[lo:hi], q[lo:hi], bbox=bbox[lo:hi], wh=(640, 480))
        assert torch.equal(part["logits"], logits[lo:hi]) and torch.equal(part["label"], answers[lo:hi])


@pytest.mark.parametrize("precision", ["fp32", "fp32tc"])
@pytest.mark.parametrize("relation", [False, True])
@pytest.mark.parametrize("B", [0, 1, 37, 129])
def test_engine_empty_single_and_ragged_batches(engine_mod, relation, B, precision):
    """edge cases of the data-parallel split: an empty shard, one question, sizes that are not tile multiples (fp32tc: an odd
    number of 128-row blocks leaves the last CTA pair of the recurrent kernel half empty)"""
    cfg = O.SMALL_REGAT if relation else O.SMALL
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, max(B, 1), 77 + B)
    batch = {k: (v[:B] if torch.is_tensor(v) else v) for k, v in batch.items()}
    eng = engine_mod.VQAEngine(W, relation=relation, precision=precision)
    kw = dict(bbox=batch["bbox"].cuda(), wh=batch["wh"]) if relation else {}
    out = eng.forward(batch["img"].cuda(), batch["q"].cuda(), **kw)
    assert out["logits"].shape == (B, cfg.ans_dim) and out["label"].shape == (B,) and out["att"].shape == (B, cfg.num_objs)
    if B == 0:
        assert eng.last_launches == 0
        return
    with torch.no_grad():
        ref, enc = O.forward(batch, W, cfg)
    assert relerr(out["logits"], ref) < 1e-5
    assert relerr(out["att"], enc["v_att"][:, :, 0]) < 1e-5
    assert torch.equal(out["label"].cpu(), ref.argmax(1))


def test_engine_host_path_bf16_feature_cache(engine_mod):
    """a host feature cache kept in the resident bf16 format goes straight to HBM: same answers as the device path"""
    cfg = O.SMALL
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, 150, 41)
    eng = engine_mod.VQAEngine(W, relation=False, precision="bf16")
    img_lp = batch["img"].to(torch.bfloat16)
    ref = eng.forward(img_lp.cuda(), batch["q"].cuda())
    label_h, h2d, d2h = eng.forward_host(img_lp.pin_memory(), batch["q"].pin_memory(), chunk=64)
    assert torch.equal(label_h, ref["label"].cpu())
    assert torch.equal(eng.last_host_outputs["logits"], ref["logits"])
    assert h2d == img_lp.numel() * 2 + batch["q"].numel() * 8 and d2h == 150 * 8
    with pytest.raises(TypeError):
        engine_mod.VQAEngine(W, relation=False, precision="fp32").forward_host(img_lp, batch["q"])
