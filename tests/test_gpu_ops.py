"""GPU parity tests, op by op, through the C ABI (ctypes) against the CPU oracle.

Bar: bit-exact for labels / indices; ‖Δ‖∞/‖ref‖∞ ≤ 1e-5 in fp32 mode and ≤ 1e-2 in
bf16 mode (BASELINE.json north_star) for floating point.
"""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import vqa_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2}


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from vqa_collection_b200 import ops as _ops
    return _ops


def relerr(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---------------------------------------------------------------------------- relation labels
def test_relation_labels_golden(ops, golden_dir):
    z = np.load(os.path.join(golden_dir, "relation.npz"))
    meta = ast.literal_eval(str(z["meta"]))
    boxes = O.make_boxes(48, 36, meta["grid_seed"], meta["W"], meta["H"], grid=True)
    got = ops.relation_labels(torch.from_numpy(boxes).cuda(), meta["W"], meta["H"]).cpu().numpy()
    assert np.array_equal(got, z["grid_graph"])            # vs the REAL reference's output
    # known-answer table: pair (a,b) as a 2-box image
    pair = np.stack([z["ka_a"], z["ka_b"]], 1).astype(np.float32)         # [N,2,4]
    lab = ops.relation_labels(torch.from_numpy(pair).cuda(), meta["W"], meta["H"]).cpu().numpy()
    assert np.array_equal(np.stack([lab[:, 0, 1], lab[:, 1, 0]], 1), z["ka_labels"])
    assert np.all(lab[:, 0, 0] == 0) and np.all(lab[:, 1, 1] == 0)


@pytest.mark.parametrize("B,K", [(1, 36), (257, 36), (64, 1), (33, 64), (1024, 36)])
def test_relation_labels_grid_exact(ops, B, K):
    boxes = O.make_boxes(B, K, 77 + B + K, 640, 480, grid=True)
    want = O.relation_graph_batch(boxes, 640, 480)
    got = ops.relation_labels(torch.from_numpy(boxes).cuda(), 640, 480).cpu().numpy()
    assert got.shape == (B, K, K) and got.dtype == np.uint8
    assert np.array_equal(got, want)


def test_relation_labels_edge_cases(ops):
    # empty batch
    got = ops.relation_labels(torch.zeros((0, 36, 4), device="cuda"), 640, 480)
    assert got.shape == (0, 36, 36)
    # identical boxes, zero-area boxes, diagonal-disjoint boxes (F6 quirks)
    b = np.array([[[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [5, 5, 5, 5], [200, 200, 203, 203]]], np.float32)
    want = O.relation_graph_batch(b, 640, 480)
    got = ops.relation_labels(torch.from_numpy(b).cuda(), 640, 480).cpu().numpy()
    assert np.array_equal(got, want)
    assert got[0, 0, 1] == 1 and got[0, 1, 0] == 2          # identical → (1,2)
    assert got[0, 0, 2] == 3 and got[0, 2, 0] == 3          # diagonal-disjoint → (3,3)
    # per-image (w,h)
    boxes = O.make_boxes(8, 36, 5, 640, 480, grid=True)
    wh = np.array([[640, 480], [320, 240], [1000, 50], [64, 48]] * 2, np.float32)
    want = np.stack([O.relation_graph_batch(boxes[i:i + 1], wh[i, 0], wh[i, 1])[0] for i in range(8)])
    got = ops.relation_labels(torch.from_numpy(boxes).cuda(), wh=torch.from_numpy(wh).cuda()).cpu().numpy()
    assert np.array_equal(got, want)
    # K out of range is an error, not a silent fallback
    with pytest.raises(RuntimeError):
        ops.relation_labels(torch.zeros((1, 65, 4), device="cuda"), 640, 480)
    with pytest.raises(RuntimeError):
        ops.relation_labels(torch.zeros((1, 36, 4)), 640, 480)          # CPU tensor


def test_relation_labels_continuous_statistical(ops):
    boxes = O.make_boxes(512, 36, 991, 640, 480, grid=False)
    want = O.relation_graph_batch(boxes, 640, 480)
    got = ops.relation_labels(torch.from_numpy(boxes).cuda(), 640, 480).cpu().numpy()
    assert (got != want).mean() < 1e-5                       # near-boundary pairs only (H2)


def test_relation_labels_host_form(ops):
    boxes = O.make_boxes(100, 36, 12, 640, 480, grid=True)
    got = ops.relation_labels_host(boxes, 640, 480)
    assert np.array_equal(got, O.relation_graph_batch(boxes, 640, 480))


def test_relation_full_size_properties(ops):
    """BASELINE size and beyond: size-independent properties (zero diagonal, label range,
    opposite-direction pairing, determinism)."""
    boxes = torch.from_numpy(O.make_boxes(16384, 36, 3, 640, 480, grid=True)).cuda()
    lab = ops.relation_labels(boxes, 640, 480)
    lab2 = ops.relation_labels(boxes, 640, 480)
    assert torch.equal(lab, lab2)
    assert int(lab.max()) <= 11
    assert int(torch.diagonal(lab, dim1=1, dim2=2).max()) == 0
    lt = lab.transpose(1, 2)
    d = (lab >= 4) & (lt >= 4)
    opp = torch.where(lab <= 7, lab + 4, lab - 4)
    assert torch.equal(lt[d], opp[d])                          # direction labels come in opposite pairs
    assert torch.equal((lab == 1), (lt == 2)) and torch.equal((lab == 0), (lt == 0))


# ---------------------------------------------------------------------------- linear
def _lin_ref(A, W, scale, bias, relu, mul, div, logit_w):
    y = A.double() @ W.double().t()
    if scale is not None:
        y = y * scale.double()
    if bias is not None:
        y = y + bias.double()
    if relu:
        y = torch.relu(y)
    if mul is not None:
        y = y * mul.double().repeat_interleave(div, 0)[: y.shape[0]]
    if logit_w is not None:
        return y * logit_w.double()
    return y


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (72, 128, 128), (300, 200, 192), (1024, 3129, 2048),
                                   (36 * 64, 1024, 2048), (8, 3072, 1024), (1, 64, 64)])
def test_linear_store(ops, dtype, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((M, K), generator=g).to(dtype)
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(dtype)
    scale = torch.rand((N,), generator=g) + 0.5
    bias = torch.randn((N,), generator=g)
    want = _lin_ref(A.float(), W.float(), scale, bias, True, None, 1, None)
    for out_dtype in (torch.float32, dtype):
        got = ops.linear(A.cuda(), W.cuda(), scale.cuda(), bias.cuda(), relu=True, out_dtype=out_dtype)
        assert got.shape == (M, N) and got.dtype == out_dtype
        # operands are exact in `dtype`, so only accumulation order / output rounding differ
        tol = (1e-5 if dtype == torch.float32 else 1e-4) if out_dtype == torch.float32 else 1e-2
        assert relerr(got, want) < tol, (dtype, out_dtype)
    # no scale / bias / relu
    got = ops.linear(A.cuda(), W.cuda(), out_dtype=torch.float32)
    assert relerr(got, _lin_ref(A.float(), W.float(), None, None, False, None, 1, None)) < (
        1e-5 if dtype == torch.float32 else 1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Bn,K36,N,K", [(4, 36, 1024, 2048), (7, 36, 128, 256), (3, 36, 300, 64)])
def test_linear_mul_and_logit_reduction(ops, dtype, Bn, K36, N, K):
    g = torch.Generator().manual_seed(Bn + N + K)
    M = Bn * K36
    A = torch.rand((M, K), generator=g).to(dtype)
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(dtype)
    scale = torch.full((N,), 0.9)
    bias = torch.randn((N,), generator=g) * 0.1
    mul_c = torch.rand((Bn, N), generator=g)
    lw = torch.randn((N,), generator=g)
    # elementwise multiplier, row-broadcast over the K36 regions
    got = ops.linear(A.cuda(), W.cuda(), scale.cuda(), bias.cuda(), relu=True, mul=mul_c.cuda(),
                     mul_row_div=K36, out_dtype=torch.float32)
    want = _lin_ref(A.float(), W.float(), scale, bias, True, mul_c, K36, None)
    tol = 1e-5 if dtype == torch.float32 else 1e-4
    assert relerr(got, want) < tol
    # row reduction with the logit vector
    parts = ops.linear(A.cuda(), W.cuda(), scale.cuda(), bias.cuda(), relu=True, mul=mul_c.cuda(),
                       mul_row_div=K36, logit_w=lw.cuda())
    pw = ops.linear_part_width(dtype)
    assert parts.shape == (M, (N + pw - 1) // pw)
    full = _lin_ref(A.float(), W.float(), scale, bias, True, mul_c, K36, lw)
    assert relerr(parts.sum(1), full.sum(1)) < tol
    for p in range(parts.shape[1]):
        assert relerr(parts[:, p], full[:, p * pw:(p + 1) * pw].sum(1)) < 2 * tol


def test_linear_rejects_bad_arguments(ops):
    A = torch.zeros((8, 60), device="cuda", dtype=torch.bfloat16)
    W = torch.zeros((8, 60), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.linear(A, W)                                      # K % 64 != 0
    with pytest.raises(RuntimeError):
        ops.linear(A.cpu(), W.cpu())                          # no CPU fallback
    with pytest.raises(TypeError):
        ops.linear(A.half(), W.half())


# ---------------------------------------------------------------------------- GRU
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfg,B", [(O.SMALL, 5), (O.FULL, 16)])
def test_gru_last_state(ops, dtype, cfg, B):
    from vqa_collection_b200.engine import prepare_weights
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, B, 5)
    P = prepare_weights(W, dtype, "cuda", False)
    h = ops.gru_last_state(batch["q"].cuda(), P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"])
    with torch.no_grad():
        want = O.question_embedding(batch["q"], W)
    assert relerr(h, want) < TOL[dtype]
    if dtype == torch.bfloat16:
        # persistent fused kernel (gate-interleaved weights): one launch for all T steps
        packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
        h2, h2_lp = ops.gru_last_state(batch["q"].cuda(), P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"],
                                       want_lp=True, packed=packed)
        assert relerr(h2, want) < TOL[dtype]
        assert relerr(h2_lp, want) < TOL[dtype]
        h3 = ops.gru_last_state(batch["q"].cuda(), P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=packed)
        assert torch.equal(h2, h3)                              # deterministic


@pytest.mark.parametrize("B,T", [(1, 1), (130, 3), (1024, 14), (1300, 14)])
def test_gru_persistent_shapes(ops, B, T):
    """ragged batch, single step, and a batch larger than one co-resident grid (chunked)"""
    from vqa_collection_b200.engine import prepare_weights
    cfg = O.FULL
    W = O.make_weights(cfg, 1111)
    P = prepare_weights(W, torch.bfloat16, "cuda", False)
    g = torch.Generator().manual_seed(B + T)
    q = torch.randint(0, cfg.ntoken + 1, (B, T), generator=g)           # includes the padding row
    packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
    h = ops.gru_last_state(q.cuda(), P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=packed)
    idx = torch.arange(0, B, max(1, B // 64))
    with torch.no_grad():
        want = O.question_embedding(q[idx], W)
    assert relerr(h[idx.cuda()], want) < 1e-2


@pytest.mark.parametrize("B,T", [(1, 1), (5, 2), (130, 3), (1024, 14), (1300, 14)])
def test_gru_token_table(ops, B, T):
    """token-table form of the fused GRU (vqa_gru_args.d_gi_table): no gather, no x-part GEMM; against the oracle, against the
    x-part form, deterministic; every tile choice (one / two row blocks per CTA pair, 32 / 64 units)"""
    from vqa_collection_b200.engine import prepare_weights
    cfg = O.FULL
    W = O.make_weights(cfg, 1111)
    P = prepare_weights(W, torch.bfloat16, "cuda", False)
    assert P["gi_table"].dtype == torch.float16 and P["gi_table"].shape == (P["emb"].shape[0], 3 * cfg.hidden_dim)
    g = torch.Generator().manual_seed(B + T)
    q = torch.randint(0, cfg.ntoken + 1, (B, T), generator=g)           # includes the padding row
    packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
    args = (q.cuda(), P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"])
    h_x = ops.gru_last_state(*args, packed=packed)
    h, h_lp = ops.gru_last_state(*args, packed=packed, gi_table=P["gi_table"], want_lp=True)
    idx = torch.arange(0, B, max(1, B // 64))
    with torch.no_grad():
        want = O.question_embedding(q[idx], W)
    e_tab, e_x = relerr(h[idx.cuda()], want), relerr(h_x[idx.cuda()], want)
    print(f"token table: err {e_tab:.2e} (x-part form {e_x:.2e})")
    assert e_tab < 1e-2
    assert relerr(h_lp[idx.cuda()], want) < 1e-2
    assert relerr(h, h_x) < 1e-2
    assert torch.equal(h, ops.gru_last_state(*args, packed=packed, gi_table=P["gi_table"]))
    for forced in ("64x1", "32x1", "64x2", "32x2"):
        import subprocess, sys, os
        code = (
            "import torch, sys; sys.path.insert(0, '.');\n"
            "from oracle import vqa_oracle as O; from vqa_collection_b200 import ops; from vqa_collection_b200.engine import prepare_weights\n"
            f"W = O.make_weights(O.FULL, 1111); P = prepare_weights(W, torch.bfloat16, 'cuda', False)\n"
            f"g = torch.Generator().manual_seed({B + T}); q = torch.randint(0, O.FULL.ntoken + 1, ({B}, {T}), generator=g)\n"
            "h = ops.gru_last_state(q.cuda(), P['emb'], P['w_ih'], P['b_ih'], P['w_hh'], P['b_hh'], packed=(P['wx_packed'], P['wh_packed'], P['bias_packed']), gi_table=P['gi_table'])\n"
            "torch.save(h.cpu(), sys.argv[1])\n")
        if B != 1024:
            continue                                                    # the forced tiles: at the full batch only
        out = f"/tmp/gru_tab_{forced}_{B}.pt"
        env = dict(os.environ, VQA_B200_GRU_CFG=forced)
        subprocess.run([sys.executable, "-c", code, out], check=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        hf = torch.load(out)
        assert relerr(hf[idx], want) < 1e-2, forced
        assert relerr(hf, h.cpu()) < 2e-3, forced                      # same arithmetic, other tile shapes


# ---------------------------------------------------------------------------- fused answer selection
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(1024, 3129, 2048), (37, 3129, 256), (300, 200, 128), (5, 64, 64)])
def test_linear_fused_argmax(ops, dtype, M, N, K):
    """lowest-index argmax of every output row selected in the GEMM's epilogue (wrapper.py:14, torch.max's tie rule):
    many N tiles per row block, row tails, ties inside a tile and across tiles, an all-zero (post-ReLU) row"""
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.rand((M, K), generator=g) - 0.3).to(dtype)
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(dtype)
    bias = torch.randn((N,), generator=g) * 0.1
    # duplicated weight rows give exactly equal logits far apart (different N tiles) and next to each other
    W[N - 1] = W[3]
    W[5] = W[3]
    bias[N - 1] = bias[5] = bias[3] = 2.0                      # ... and large enough to be the maximum of many rows
    A[0] = 0                                                   # row 0: logits = bias only
    out, label = ops.linear(A.cuda(), W.cuda(), None, bias.cuda(), relu=True, out_dtype=torch.float32, want_argmax=True)
    assert label.dtype == torch.int64 and label.shape == (M,)
    assert torch.equal(label, torch.max(out, 1)[1])            # the rule, on the kernel's own logits
    assert int(label[0]) == 3                                  # first of the three equal maxima
    ties = (out[:, 3] == out.max(1)[0]).sum().item()
    assert ties >= 1
    bias0 = torch.full((N,), -1.0)
    out0, label0 = ops.linear(torch.zeros((M, K)).to(dtype).cuda(), W.cuda(), None, bias0.cuda(), relu=True,
                              out_dtype=torch.float32, want_argmax=True)
    assert float(out0.abs().max()) == 0.0 and int(label0.abs().max()) == 0       # all-zero rows -> index 0


# ---------------------------------------------------------------------------- attention pooling
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,K,V,P", [(3, 36, 2048, 4), (1, 36, 256, 1), (5, 10, 64, 2), (2, 64, 512, 8)])
def test_attention_pool(ops, dtype, B, K, V, P):
    g = torch.Generator().manual_seed(B * K + V)
    parts = torch.randn((B * K, P), generator=g)
    x = torch.rand((B, K, V), generator=g).to(dtype)
    bias = 0.3
    att, vsum, vatt = ops.attention_pool(parts.cuda(), bias, x.cuda(), True, True, True)
    want_att = torch.softmax(parts.sum(1).view(B, K) + bias, 1)
    want_v = want_att.unsqueeze(2).double() * x.double()
    assert relerr(att, want_att) < 1e-5
    assert relerr(vatt, want_v) < TOL[dtype]
    assert relerr(vsum, want_v.sum(1)) < TOL[dtype]
    att2, vsum2, none = ops.attention_pool(parts.cuda(), bias, x.cuda(), True, True, False)
    assert none is None and relerr(vsum2, want_v.sum(1)) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,K,V,P", [(1, 36, 2048, 4), (37, 36, 2048, 4), (300, 36, 2048, 1), (9, 64, 1024, 2), (11, 7, 512, 3),
                                     (1024, 36, 2048, 4)])
def test_attention_pool_stream_kernel(ops, dtype, B, K, V, P):
    """att + vsum (the Up-Down forward's form) takes the persistent bulk-copy streaming kernel (V <= 2048): fewer images
    than SMs, many ring wrap-arounds per CTA, a short last chunk (K = 64), one chunk per image (K = 7), idle consumer warps"""
    g = torch.Generator().manual_seed(B * K + V + P)
    parts = torch.randn((B * K, P), generator=g) * 2
    x = torch.rand((B, K, V), generator=g).to(dtype)
    att, vsum, none = ops.attention_pool(parts.cuda(), -0.2, x.cuda(), True, True, False)
    want_att = torch.softmax(parts.sum(1).view(B, K) - 0.2, 1)
    want = (want_att.unsqueeze(2).double() * x.double()).sum(1)
    assert none is None and relerr(att, want_att) < 1e-5
    assert relerr(vsum, want) < TOL[dtype]
    # per-row check (a max-norm over the whole batch would hide a wrong image): every image on its own
    err = ((vsum.double().cpu() - want).abs().amax(1) / want.abs().amax(1))
    assert float(err.max()) < TOL[dtype] * 2
    _, vsum2, _ = ops.attention_pool(parts.cuda(), -0.2, x.cuda(), False, True, False)
    assert torch.equal(vsum2, vsum)                                        # deterministic, att output optional


# ---------------------------------------------------------------------------- backward GEMM forms
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(512, 2048, 3129), (1000, 304, 520), (36 * 64, 1024, 2048)])
def test_linear_trans_w_is_dX(ops, dtype, M, N, K):
    """dX[m,n] = Σ_k dY[m,k]·W[k,n]: W read as stored by the forward ([out=K, in=N] row-major), K tail zero-filled"""
    g = torch.Generator().manual_seed(M + N + K)
    ld = (K + 7) // 8 * 8
    dY = torch.zeros((M, ld))
    dY[:, :K] = torch.randn((M, K), generator=g)
    dY[:, K:] = 7.0                                            # padding must never be read
    W = torch.randn((K, N), generator=g) / K ** 0.5
    saved = torch.relu(torch.randn((M, N), generator=g))       # ReLU output of the previous layer
    acc = torch.randn((M, N), generator=g)
    dYd, Wd = dY.to(dtype).cuda(), W.to(dtype).cuda()
    want = (dYd[:, :K].float().cpu().double() @ Wd.float().cpu().double()) * 0.37 + acc.double()
    want = torch.where(saved > 0, want, torch.zeros_like(want))
    scale = torch.full((N,), 0.37).cuda()
    got = ops.linear(dYd[:, :K], Wd, scale, None, trans_w=True, add=acc.cuda(), mask=saved.to(dtype).cuda(),
                     out_dtype=torch.float32)
    assert got.shape == (M, N)
    assert relerr(got, want) < (1e-5 if dtype == torch.float32 else 2e-3)   # operands already rounded: only accumulation differs


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(3129, 2048, 512), (1024, 2048, 36 * 100), (200, 136, 77 * 8)])
def test_linear_trans_a_is_dW(ops, dtype, M, N, K):
    """dW[m,n] = Σ_k dY[k,m]·X[k,n]: both operands read row-major with the contraction as the row index"""
    g = torch.Generator().manual_seed(M * 3 + N + K)
    ldm = (M + 7) // 8 * 8
    dY = torch.zeros((K, ldm))
    dY[:, :M] = torch.randn((K, M), generator=g)
    X = torch.randn((K, N), generator=g)
    dYd, Xd = dY.to(dtype).cuda(), X.to(dtype).cuda()
    want = dYd[:, :M].float().cpu().double().t() @ Xd.float().cpu().double()
    got = ops.linear(dYd[:, :M], Xd, trans_a=True, trans_w=True, out_dtype=torch.float32)
    assert got.shape == (M, N)
    assert relerr(got, want) < (1e-5 if dtype == torch.float32 else 2e-3)


# ---------------------------------------------------------------------------- argmax
def test_argmax_lowest_index_on_ties(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn((257, 3129), generator=g))         # ~50 % exact zeros (F7)
    x[0] = 0.0                                                    # all tied → index 0
    x[1, 100] = x[1, 3000] = 9.0                                  # tie → 100
    x[2, 3128] = 10.0
    got = ops.argmax_rows(x.cuda()).cpu()
    assert torch.equal(got, torch.max(x, 1)[1])
    assert got[0] == 0 and got[1] == 100 and got[2] == 3128


# ---------------------------------------------------------------------------- graph attention
@pytest.mark.parametrize("dtype,merged", [(torch.float32, False), (torch.bfloat16, False), (torch.bfloat16, True)])
@pytest.mark.parametrize("cfg,B", [(O.SMALL_REGAT, 6), (O.FULL_REGAT, 3), (O.SMALL_REGAT, 301)])
def test_graph_attention_layer(ops, dtype, merged, cfg, B):
    """layout 0 = FFMA kernel on [P|S|A'|B'] (fp32 parity path), layout 1 = tcgen05 kernel on the merged
    [P|S|Q] projection (bf16 product path); B=301 makes every persistent CTA loop over >1 image."""
    from vqa_collection_b200.engine import prepare_gcn_layer
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, B, 9)
    g = torch.Generator().manual_seed(4)
    att = torch.softmax(torch.randn((B, cfg.num_objs), generator=g) * 2, 1)
    x = batch["img"]
    feature = att.unsqueeze(2) * x
    with torch.no_grad():
        want, alpha = O.corr_graph_conv(feature, batch["graph"].float(), W, "gcn.0.")
        want = torch.relu(want)
    Pl = prepare_gcn_layer({k[6:]: v for k, v in W.items() if k.startswith("gcn.0.")}, dtype, "cuda", merged=merged)
    xd = x.to(dtype).cuda().view(B * cfg.num_objs, cfg.v_dim)
    labels = batch["graph"].to(torch.uint8).cuda()

    def run(att_d, *want_flags):
        if merged:
            Y = ops.linear(xd, Pl["Wg3"])
            return ops.graph_attention_merged(Y, xd, att_d, labels, Pl["wvec"], Pl["gat_c0"], Pl["label_bias_lp"],
                                              Pl["num_labels"], cfg.num_objs, *want_flags)
        Y = ops.linear(xd, Pl["Wg"])
        return ops.graph_attention(Y, att_d, labels, Pl["label_bias"], Pl["ba"], Pl["bb"], cfg.num_objs, *want_flags)

    out, vsum, al = run(att.cuda(), True, True, True)
    tol = TOL[dtype]
    assert relerr(al, alpha) < tol
    assert relerr(out, want) < tol
    assert relerr(vsum, want.sum(1)) < tol
    # only the K-sum requested (the whole-path configuration)
    none, vsum2, none2 = run(att.cuda(), False, True, False)
    assert none is None and none2 is None and torch.equal(vsum2, vsum)
    # attention = None means the features are used as they are
    with torch.no_grad():
        want1, _ = O.corr_graph_conv(x, batch["graph"].float(), W, "gcn.0.")
    out1, _, _ = run(None, True, False, False)
    # (un-attended features are 36x larger: α is far more peaked, so bf16 gets 3x the slack)
    assert relerr(out1, torch.relu(want1)) < (tol if dtype == torch.float32 else 3 * tol)


def test_answer_scores(ops):
    g = torch.Generator().manual_seed(4)
    B, A = 67, 3129
    target = torch.zeros((B, A))
    idx = torch.randint(0, A, (B, 3), generator=g)
    target.scatter_(1, idx, torch.randint(1, 4, (B, 3), generator=g).float() / 3.0)
    logits = torch.randn((B, A), generator=g)
    logits[:, 7] = logits.max() + 1.0                  # a tie between two answers: the lowest index wins
    logits[:, 3000] = logits[:, 7]
    ref_score, ref_label = O.compute_score(logits, target)
    label = ops.argmax_rows(logits.cuda())
    dense, row, total = ops.answer_scores(label, target.cuda(), want_dense=True, want_sum=True)
    assert torch.equal(label.cpu(), ref_label) and torch.equal(dense.cpu(), ref_score)
    assert torch.equal(row.cpu(), ref_score.sum(1)) and abs(float(total) - float(ref_score.sum())) < 1e-4
    _, row2, _ = ops.answer_scores(label, target.cuda(), want_dense=False)
    assert torch.equal(row2, row)


# ---------------------------------------------------------------------------- fp32-class tensor-core mode (VQA_F16X2)
def test_split_f32_planes(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn((38, 100), generator=g) * torch.logspace(-6, 3, 100)          # element count % 8 == 0 (16-byte planes)
    x[0, :4] = torch.tensor([0.0, 1.0, -65504.0, 6.0e-8])
    p = ops.split_f32(x.cuda()).cpu()
    hi = x.to(torch.float16)
    lo = ((x - hi.float()) * 2048.0).to(torch.float16)
    assert torch.equal(p[0], hi) and torch.equal(p[1], lo)
    back = p[0].double() + p[1].double() / 2048.0
    ok = x.abs() > 1e-4
    assert float(((back - x.double()).abs() / x.double().abs())[ok].max()) < 2.0 ** -21


@pytest.mark.parametrize("M,N,K", [(36864, 1024, 2048), (1024, 3129, 2048), (1024, 2048, 1024), (300, 200, 128), (5, 64, 64)])
def test_linear_split(ops, M, N, K):
    """three-product fp16 split GEMM against float64: an order of magnitude inside the fp32 gate (1e-5), every tile shape
    (CTA pairs, 256 / 192 / 128 / 64 wide), row / column tails, fused epilogues, plane-pair output, fused argmax"""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.rand((M, K), generator=g) * 4.0                    # post-ReLU features
    A[A < 1.0] = 0.0
    W = torch.randn((N, K), generator=g) / K ** 0.5
    bias = torch.randn((N,), generator=g) * 0.1
    scale = torch.rand((N,), generator=g) + 0.5
    A2, W2 = ops.split_f32(A.cuda()), ops.split_f32(W.cuda())
    rows = torch.arange(0, M, max(1, M // 256))
    ref = torch.relu((A[rows].double() @ W.double().t()) * scale.double() + bias.double())
    out, label = ops.linear_split(A2, W2, scale.cuda(), bias.cuda(), relu=True, want_argmax=True)
    err = relerr(out[rows.cuda()], ref)
    f32 = relerr(torch.relu((A[rows] @ W.t()) * scale + bias), ref)
    print(f"linear_split {M}x{N}x{K}: err {err:.2e} (cpu fp32 {f32:.2e})")
    assert err < 5e-6
    assert torch.equal(label, torch.max(out, 1)[1])
    planes = ops.linear_split(A2, W2, scale.cuda(), bias.cuda(), relu=True, out_split=True)
    assert torch.equal(planes, ops.split_f32(out))
    if N % 256 == 0:
        lw = torch.randn((N,), generator=g)
        mul = torch.rand(((M + 35) // 36, N), generator=g)
        parts = ops.linear_split(A2, W2, scale.cuda(), bias.cuda(), relu=True, mul=mul.cuda(), mul_row_div=36, logit_w=lw.cuda())
        want = ((ref * mul[rows // 36].double()) * lw.double()).sum(1)
        assert relerr(parts.sum(1)[rows.cuda()], want) < 5e-6
