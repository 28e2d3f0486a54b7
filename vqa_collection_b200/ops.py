"""Tensor-level wrappers over the C ABI (include/vqa_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every
function passes raw device pointers + sizes to the library and returns torch
tensors allocated with the caching allocator.  Inputs must already be CUDA
tensors; there is no CPU path (non-CUDA input raises, like a device mismatch
does in the reference).
"""
import ctypes as C

import torch

from . import _lib as L

_TORCH_DTYPE = {L.VQA_F32: torch.float32, L.VQA_BF16: torch.bfloat16}


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return L.VQA_F32
    if t == torch.bfloat16:
        return L.VQA_BF16
    raise TypeError(f"vqa_collection_b200: unsupported dtype {t} (float32 or bfloat16)")


def torch_dtype(code: int) -> torch.dtype:
    return _TORCH_DTYPE[code]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require(t, dtype=None, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"vqa_collection_b200: {name} must be a CUDA tensor (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"vqa_collection_b200: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous() and not (t.dim() == 2 and t.stride(1) == 1):      # row-strided 2-D views are fine
        raise RuntimeError(f"vqa_collection_b200: {name} must be contiguous")
    return t


def device_info():
    lib = L.load()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    L.check(lib.vqa_device_info(C.byref(a), C.byref(b), C.byref(c)))
    return {"sm_count": a.value, "cc": (b.value, c.value)}


def relation_labels(bbox, img_w=None, img_h=None, wh=None):
    """bbox f32 [B,K,4] (CUDA) → uint8 [B,K,K].  util/relation.py:65-80 batched."""
    lib = L.load()
    _require(bbox, torch.float32, "bbox")
    if bbox.dim() != 3 or bbox.shape[2] != 4:
        raise ValueError("bbox must be [B,K,4]")
    B, K = bbox.shape[0], bbox.shape[1]
    if wh is not None:
        _require(wh, torch.float32, "wh")
        if tuple(wh.shape) != (B, 2):
            raise ValueError("wh must be [B,2]")
    elif img_w is None or img_h is None:
        raise ValueError("relation_labels needs (img_w, img_h) or a per-image wh tensor")
    out = torch.empty((B, K, K), dtype=torch.uint8, device=bbox.device)
    L.check(lib.vqa_relation_labels(_ptr(bbox), _ptr(wh), B, K, float(img_w or 0), float(img_h or 0),
                                    _ptr(out), _stream()))
    return out


def relation_labels_host(bbox_np, img_w, img_h):
    """Host-buffer form: numpy f32 [B,K,4] → numpy uint8 [B,K,K] (H2D + kernel + D2H)."""
    import numpy as np
    lib = L.load()
    bbox_np = np.ascontiguousarray(bbox_np, dtype=np.float32)
    B, K = bbox_np.shape[0], bbox_np.shape[1]
    out = np.empty((B, K, K), dtype=np.uint8)
    L.check(lib.vqa_relation_labels_host(bbox_np.ctypes.data_as(C.c_void_p), B, K, float(img_w), float(img_h),
                                         out.ctypes.data_as(C.c_void_p)))
    return out


def cast_to_bf16(x):
    lib = L.load()
    _require(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    L.check(lib.vqa_cast_f32_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()))
    return out


def cast_to_f32(x):
    lib = L.load()
    _require(x, torch.bfloat16, "x")
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    L.check(lib.vqa_cast_bf16_to_f32(_ptr(x), _ptr(out), x.numel(), _stream()))
    return out


def linear_part_width(dtype: torch.dtype) -> int:
    return L.load().vqa_linear_part_width(dtype_code(dtype))


def linear(A, W, scale=None, bias=None, relu=False, mul=None, mul_row_div=1, logit_w=None,
           out_dtype=None, N=None, add=None, add_row_div=1, trans_a=False, trans_w=False, mask=None, out=None,
           leaky_slope=0.0, add_after_act=False, sigmoid=False, want_argmax=False):
    """Fused weight-normed linear layer (modules.py:13-60), see vqa_linear in the header.

    A [M,K], W [N_rows,K] (same dtype); returns [M,N] (out_dtype) or, with ``logit_w``,
    the f32 row-reduction parts [M, n_parts].  ``N`` < W.shape[0] uses only the first N
    rows of W (row-padded weights).  Backward forms: ``trans_w`` → W is [K,N] (y = A·W),
    ``trans_a`` (with trans_w) → A is [K,M] (y = Aᵀ·W); ``mask`` [M,N] zeroes y where mask ≤ 0.
    ``leaky_slope`` turns the ReLU into LeakyReLU (LReLUNet, modules.py:62-78); ``add_after_act`` applies ``add``
    after the activation; ``sigmoid`` applies the logistic function last.  ``want_argmax`` (f32 store form): also returns the
    lowest-index argmax of every output row (int64 [M]; wrapper.py:14), selected in the GEMM's epilogue → (out, label).
    """
    lib = L.load()
    _require(A, None, "A")
    _require(W, A.dtype, "W")
    if A.dim() != 2 or W.dim() != 2:
        raise ValueError(f"linear: shape mismatch A{tuple(A.shape)} W{tuple(W.shape)}")
    if trans_a and not trans_w:
        raise ValueError("linear: trans_a needs trans_w")
    M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
    Kw, Nw = (W.shape[0], W.shape[1]) if trans_w else (W.shape[1], W.shape[0])
    if K != Kw:
        raise ValueError(f"linear: shape mismatch A{tuple(A.shape)} W{tuple(W.shape)} (trans_a={trans_a}, trans_w={trans_w})")
    N = Nw if N is None else N
    code = dtype_code(A.dtype)
    for nm, t in (("scale", scale), ("bias", bias), ("logit_w", logit_w)):
        if t is not None:
            _require(t, torch.float32, nm)
            if t.numel() < N:
                raise ValueError(f"linear: {nm} has {t.numel()} < N={N} elements")
    a = L.LinearArgs()
    a.d_A, a.lda, a.d_W, a.ldw = A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0)
    a.M, a.N, a.K, a.dtype = M, N, K, code
    a.trans_a, a.trans_w = int(bool(trans_a)), int(bool(trans_w))
    a.d_scale, a.d_bias, a.relu = _ptr(scale), _ptr(bias), int(bool(relu))
    a.leaky_slope, a.add_after_act, a.sigmoid = float(leaky_slope), int(bool(add_after_act)), int(bool(sigmoid))
    if mul is not None:
        _require(mul, torch.float32, "mul")
        a.d_mul, a.ld_mul, a.mul_row_div = mul.data_ptr(), mul.stride(0), int(mul_row_div)
    else:
        a.mul_row_div = 1
    if add is not None:
        _require(add, torch.float32, "add")
        a.d_add, a.ld_add, a.add_row_div = add.data_ptr(), add.stride(0), int(add_row_div)
    else:
        a.add_row_div = 1
    if mask is not None:
        _require(mask, None, "mask")
        a.d_mask, a.ld_mask, a.mask_dtype = mask.data_ptr(), mask.stride(0), dtype_code(mask.dtype)
    if logit_w is not None:
        pw = lib.vqa_linear_part_width(code)
        n_parts = (N + pw - 1) // pw
        out = torch.empty((M, n_parts), dtype=torch.float32, device=A.device)
        a.d_logit_w, a.ldo, a.out_dtype = logit_w.data_ptr(), n_parts, L.VQA_F32
    else:
        if out is None:
            out_dtype = A.dtype if out_dtype is None else out_dtype
            out = torch.empty((M, N), dtype=out_dtype, device=A.device)
        a.ldo, a.out_dtype = out.stride(0), dtype_code(out.dtype)
    a.d_out = out.data_ptr()
    label = None
    if want_argmax:
        if logit_w is not None or out.dtype != torch.float32:
            raise ValueError("linear: want_argmax needs the f32 store form")
        label = torch.empty((M,), dtype=torch.int64, device=A.device)
        amax_ws = torch.zeros((lib.vqa_linear_argmax_workspace_bytes(M),), dtype=torch.uint8, device=A.device)
        a.d_argmax_label, a.d_argmax_ws = label.data_ptr(), amax_ws.data_ptr()
    L.check(lib.vqa_linear(C.byref(a), _stream()))
    return (out, label) if want_argmax else out


def split_f32(x):
    """f32 tensor → the fp16 plane pair of the fp32-class tensor-core mode (VQA_F16X2): out[0] = fp16(x),
    out[1] = fp16((x - out[0])·2^11), shape [2, *x.shape]."""
    lib = L.load()
    _require(x, torch.float32, "x")
    x = x.contiguous()
    out = torch.empty((2,) + tuple(x.shape), dtype=torch.float16, device=x.device)
    L.check(lib.vqa_split_f32(_ptr(x), out[0].data_ptr(), out[1].data_ptr(), x.numel(), _stream()))
    return out


def linear_split(A2, W2, scale=None, bias=None, relu=False, mul=None, mul_row_div=1, logit_w=None, out_split=False,
                 add=None, add_row_div=1, want_argmax=False):
    """vqa_linear in the fp32-class tensor-core mode: A2 [2,M,K], W2 [2,N,K] fp16 plane pairs (split_f32); three
    tcgen05.mma per k-step, fp32 accumulation.  Returns f32 [M,N], the f32 logit parts (``logit_w``) or, with ``out_split``,
    the plane pair [2,M,N] of the result (the next layer's operand)."""
    lib = L.load()
    _require(A2, torch.float16, "A2")
    _require(W2, torch.float16, "W2")
    if A2.dim() != 3 or W2.dim() != 3 or A2.shape[0] != 2 or W2.shape[0] != 2 or A2.shape[2] != W2.shape[2]:
        raise ValueError(f"linear_split: shape mismatch A2{tuple(A2.shape)} W2{tuple(W2.shape)}")
    if not (A2.is_contiguous() and W2.is_contiguous()):
        raise RuntimeError("linear_split: plane pairs must be contiguous")
    _, M, K = A2.shape
    N = W2.shape[1]
    a = L.LinearArgs()
    a.d_A, a.lda, a.d_W, a.ldw = A2.data_ptr(), K, W2.data_ptr(), K
    a.M, a.N, a.K, a.dtype = M, N, K, L.VQA_F16X2
    for nm, t in (("scale", scale), ("bias", bias), ("logit_w", logit_w)):
        if t is not None:
            _require(t, torch.float32, nm)
    a.d_scale, a.d_bias, a.relu = _ptr(scale), _ptr(bias), int(bool(relu))
    a.mul_row_div = a.add_row_div = 1
    if mul is not None:
        _require(mul, torch.float32, "mul")
        a.d_mul, a.ld_mul, a.mul_row_div = mul.data_ptr(), mul.stride(0), int(mul_row_div)
    if add is not None:
        _require(add, torch.float32, "add")
        a.d_add, a.ld_add, a.add_row_div = add.data_ptr(), add.stride(0), int(add_row_div)
    if logit_w is not None:
        pw = lib.vqa_linear_part_width(L.VQA_F16X2)
        n_parts = (N + pw - 1) // pw
        out = torch.empty((M, n_parts), dtype=torch.float32, device=A2.device)
        a.d_logit_w, a.ldo, a.out_dtype = logit_w.data_ptr(), n_parts, L.VQA_F32
    elif out_split:
        out = torch.empty((2, M, N), dtype=torch.float16, device=A2.device)
        a.ldo, a.out_dtype = N, L.VQA_F16X2
    else:
        out = torch.empty((M, N), dtype=torch.float32, device=A2.device)
        a.ldo, a.out_dtype = N, L.VQA_F32
    a.d_out = out.data_ptr()
    label = None
    if want_argmax:
        if logit_w is not None or out_split:
            raise ValueError("linear_split: want_argmax needs the f32 store form")
        label = torch.empty((M,), dtype=torch.int64, device=A2.device)
        amax_ws = torch.zeros((lib.vqa_linear_argmax_workspace_bytes(M),), dtype=torch.uint8, device=A2.device)
        a.d_argmax_label, a.d_argmax_ws = label.data_ptr(), amax_ws.data_ptr()
    L.check(lib.vqa_linear(C.byref(a), _stream()))
    return (out, label) if want_argmax else out


def attention_pool_split(parts, bias, x2, want_att=True):
    """softmax over the K regions + attention-weighted feature sum in the fp32-class mode: x2 = fp16 plane pair [2,B,K,V];
    returns (att f32 [B,K] or None, the pooled features as a plane pair [2,B,V])"""
    lib = L.load()
    _require(parts, torch.float32, "parts")
    _require(x2, torch.float16, "x2")
    _, B, K, V = x2.shape
    att = torch.empty((B, K), dtype=torch.float32, device=x2.device) if want_att else None
    vsum = torch.empty((2, B, V), dtype=torch.float16, device=x2.device)
    L.check(lib.vqa_attention_pool(_ptr(parts), parts.shape[1], float(bias), x2.data_ptr(), B, K, V, L.VQA_F16X2,
                                   _ptr(att), vsum.data_ptr(), None, _stream()))
    return att, vsum


def gru_last_state_split(tokens, gi_table, w_hh2, b_hh, wh_packed2=None):
    """Question encoder in the fp32-class mode (VQA_F16X2): tokens int64 [B,T]; gi_table f32 [rows, 3H] = W_ih·emb[v] + b_ih;
    w_hh2 the plane pair [2,3H,H] of W_hh (wh_packed2: the same in the gate-interleaved packing -> GEMM + gate update in
    one kernel, all steps in one launch).  Returns the last state, f32 [B,H]."""
    lib = L.load()
    _require(tokens, torch.int64, "tokens")
    _require(gi_table, torch.float32, "gi_table")
    _require(w_hh2, torch.float16, "w_hh2")
    _require(b_hh, torch.float32, "b_hh")
    B, T = tokens.shape
    H = w_hh2.shape[-1]
    ws_bytes = lib.vqa_gru_workspace_bytes(B, T, H, 64, L.VQA_F16X2)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=tokens.device)
    h = torch.empty((B, H), dtype=torch.float32, device=tokens.device)
    a = L.GruArgs()
    a.d_tokens, a.B, a.T, a.H, a.E_pad, a.ntoken_rows, a.dtype = tokens.data_ptr(), B, T, H, 64, gi_table.shape[0], L.VQA_F16X2
    a.d_gi_table, a.d_w_hh, a.d_b_hh = gi_table.data_ptr(), w_hh2.data_ptr(), b_hh.data_ptr()
    if wh_packed2 is not None:
        _require(wh_packed2, torch.float16, "wh_packed2")
        a.d_wh_packed = wh_packed2.data_ptr()
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    a.d_h_last = h.data_ptr()
    L.check(lib.vqa_gru_last_state(C.byref(a), _stream()))
    return h


def gru_last_state(tokens, emb, w_ih, b_ih, w_hh, b_hh, want_lp=False, packed=None, gi_table=None):
    """Embedding gather + GRU last state (encoder.py:159-160, modules.py:139-159).

    tokens int64 [B,T]; emb [rows,E_pad]; w_ih [3H,E_pad]; w_hh [3H,H] (all one dtype);
    biases f32.  Returns f32 [B,H] (and the low-precision copy when ``want_lp``).
    """
    lib = L.load()
    _require(tokens, torch.int64, "tokens")
    _require(emb, None, "emb")
    _require(w_ih, emb.dtype, "w_ih")
    _require(w_hh, emb.dtype, "w_hh")
    _require(b_ih, torch.float32, "b_ih")
    _require(b_hh, torch.float32, "b_hh")
    B, T = tokens.shape
    H = w_hh.shape[1]
    E_pad = emb.shape[1]
    if w_ih.shape != (3 * H, E_pad) or w_hh.shape != (3 * H, H):
        raise ValueError("gru_last_state: weight shapes do not match")
    code = dtype_code(emb.dtype)
    ws_bytes = lib.vqa_gru_workspace_bytes(B, T, H, E_pad, code)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=tokens.device)
    h = torch.empty((B, H), dtype=torch.float32, device=tokens.device)
    h_lp = torch.empty((B, H), dtype=emb.dtype, device=tokens.device) if want_lp else None
    a = L.GruArgs()
    a.d_tokens, a.B, a.T, a.H, a.E_pad, a.ntoken_rows, a.dtype = tokens.data_ptr(), B, T, H, E_pad, emb.shape[0], code
    a.d_emb, a.d_w_ih, a.d_b_ih, a.d_w_hh, a.d_b_hh = (emb.data_ptr(), w_ih.data_ptr(), b_ih.data_ptr(),
                                                       w_hh.data_ptr(), b_hh.data_ptr())
    if packed is not None:                 # (wx_packed, wh_packed, bias_packed) from engine.pack_gru
        a.d_wx_packed, a.d_wh_packed, a.d_bias_packed = (packed[0].data_ptr(), packed[1].data_ptr(),
                                                         packed[2].data_ptr())
    if gi_table is not None:               # engine.gru_token_table: the token-table form of the fused kernel
        _require(gi_table, torch.float16, "gi_table")
        if gi_table.shape != (emb.shape[0], 3 * H):
            raise ValueError("gru_last_state: gi_table must be [rows of emb, 3H]")
        a.d_gi_table = gi_table.data_ptr()
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    a.d_h_last, a.d_h_last_lp = h.data_ptr(), (h_lp.data_ptr() if want_lp else None)
    L.check(lib.vqa_gru_last_state(C.byref(a), _stream()))
    return (h, h_lp) if want_lp else h


def gru_sequence(x, w_ih, b_ih, w_hh, b_hh, packed=None, want_last=False):
    """1-layer GRU over dense inputs, every hidden state (SentenceEmbedding.forward_all, modules.py:147-152).

    x [B,T,E_pad] (E_pad = w_ih.shape[1], zero padded); returns out [B,T,H] in x.dtype
    (and the f32 last state [B,H] when ``want_last``)."""
    lib = L.load()
    _require(x, None, "x")
    _require(w_ih, x.dtype, "w_ih")
    _require(w_hh, x.dtype, "w_hh")
    _require(b_ih, torch.float32, "b_ih")
    _require(b_hh, torch.float32, "b_hh")
    B, T, E_pad = x.shape
    H = w_hh.shape[1]
    if w_ih.shape != (3 * H, E_pad) or w_hh.shape != (3 * H, H):
        raise ValueError("gru_sequence: weight shapes do not match")
    code = dtype_code(x.dtype)
    ws_bytes = lib.vqa_gru_workspace_bytes(B, T, H, E_pad, code)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=x.device)
    out = torch.empty((B, T, H), dtype=x.dtype, device=x.device)
    h = torch.empty((B, H), dtype=torch.float32, device=x.device) if want_last else None
    a = L.GruArgs()
    a.B, a.T, a.H, a.E_pad, a.dtype = B, T, H, E_pad, code
    a.d_x, a.d_out_all = x.data_ptr(), out.data_ptr()
    a.d_w_ih, a.d_b_ih, a.d_w_hh, a.d_b_hh = w_ih.data_ptr(), b_ih.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr()
    if packed is not None:
        a.d_wx_packed, a.d_wh_packed, a.d_bias_packed = (packed[0].data_ptr(), packed[1].data_ptr(),
                                                         packed[2].data_ptr())
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    a.d_h_last = _ptr(h)
    L.check(lib.vqa_gru_last_state(C.byref(a), _stream()))
    return (out, h) if want_last else out


def lstm_sequence(x, w_ih, b_ih, w_hh, b_hh, want_all=True, want_last=False):
    """1-layer LSTM over dense inputs, h0 = c0 = 0 (SentenceEmbedding with rnn_type='LSTM', modules.py:121-159).

    x [B,T,E_pad] (E_pad = w_ih.shape[1], zero padded); w_ih [4H,E_pad], w_hh [4H,H] in x.dtype; biases f32 [4H].
    Returns (every hidden state [B,T,H] in x.dtype or None, f32 last state [B,H] or None)."""
    lib = L.load()
    _require(x, None, "x")
    _require(w_ih, x.dtype, "w_ih")
    _require(w_hh, x.dtype, "w_hh")
    _require(b_ih, torch.float32, "b_ih")
    _require(b_hh, torch.float32, "b_hh")
    B, T, E_pad = x.shape
    H = w_hh.shape[1]
    if w_ih.shape != (4 * H, E_pad) or w_hh.shape != (4 * H, H) or not (want_all or want_last):
        raise ValueError("lstm_sequence: weight shapes do not match")
    code = dtype_code(x.dtype)
    ws_bytes = lib.vqa_lstm_workspace_bytes(B, T, H, code)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=x.device)
    out = torch.empty((B, T, H), dtype=x.dtype, device=x.device) if want_all else None
    h = torch.empty((B, H), dtype=torch.float32, device=x.device) if want_last else None
    a = L.LstmArgs()
    a.B, a.T, a.H, a.E_pad, a.dtype = B, T, H, E_pad, code
    a.d_x, a.d_w_ih, a.d_b_ih, a.d_w_hh, a.d_b_hh = x.data_ptr(), w_ih.data_ptr(), b_ih.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr()
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    a.d_h_last, a.d_out_all = _ptr(h), _ptr(out)
    L.check(lib.vqa_lstm_sequence(C.byref(a), _stream()))
    return out, h


def caption_gate_scale(out_w, p, r, want_a=False):
    """a = σ(h_w⊙p + h_w⊙r), h_w = out_w[:, -1]; returns a[:, None, :] ⊙ out_w (modules.py:225-243,294-295)."""
    lib = L.load()
    _require(out_w, None, "out_w")
    _require(p, torch.float32, "p")
    _require(r, torch.float32, "r")
    B, T, H = out_w.shape
    in2 = torch.empty_like(out_w)
    a = torch.empty((B, H), dtype=torch.float32, device=out_w.device) if want_a else None
    L.check(lib.vqa_caption_gate_scale(out_w.data_ptr(), p.data_ptr(), r.data_ptr(), B, T, H, dtype_code(out_w.dtype),
                                       in2.data_ptr(), _ptr(a), _stream()))
    return (in2, a) if want_a else in2


def seq_max(e):
    """max over the time axis of [B,T,H] (modules.py:306)."""
    lib = L.load()
    _require(e, None, "e")
    B, T, H = e.shape
    out = torch.empty((B, H), dtype=e.dtype, device=e.device)
    L.check(lib.vqa_seq_max(e.data_ptr(), B, T, H, dtype_code(e.dtype), out.data_ptr(), _stream()))
    return out


def softmax_mul(z, v):
    """softmax over the last axis of z (f32 [B,H]) times v [B,H] (predictor.py:202-203)."""
    lib = L.load()
    _require(z, torch.float32, "z")
    _require(v, None, "v")
    B, H = z.shape
    out = torch.empty_like(v)
    L.check(lib.vqa_softmax_mul(z.data_ptr(), v.data_ptr(), B, H, dtype_code(v.dtype), out.data_ptr(), _stream()))
    return out


def attention_logits(proj, q, w, K, mode=0):
    """Attention logits of one decoding step from the region projection made once per caption batch
    (BaseDecoder.decode → attention.py:70-76 / :33-40, see vqa_attention_logits).

    proj [B*K, Hd] (compute dtype; may hold MORE rows than B*K — only the first q.shape[0]*K are read),
    q f32 [B,Hd], w f32 [Hd]; mode 0 = MultiplyAttention, 1 = ConcatAttention.  Returns f32 [B*K, 1]
    (the ``parts`` operand of attention_pool; the logit layer's bias is added there)."""
    lib = L.load()
    _require(proj, None, "proj")
    _require(q, torch.float32, "q")
    _require(w, torch.float32, "w")
    B, Hd = q.shape
    if proj.dim() != 2 or proj.shape[1] != Hd or proj.shape[0] < B * K or w.numel() != Hd:
        raise ValueError(f"attention_logits: shape mismatch proj{tuple(proj.shape)} q{tuple(q.shape)} w{tuple(w.shape)}")
    out = torch.empty((B * K, 1), dtype=torch.float32, device=q.device)
    L.check(lib.vqa_attention_logits(proj.data_ptr(), proj.stride(0), q.data_ptr(), q.stride(0), w.data_ptr(), B, K, Hd,
                                     int(mode), dtype_code(proj.dtype), out.data_ptr(), _stream()))
    return out


def gru_cell(gi, gh, h, h_lp):
    """nn.GRUCell gate update (generator.py:158-159,178): gi = W_ih x + b_ih, gh = W_hh h + b_hh (f32 [B,3H]);
    ``h`` f32 [B,H] is updated IN PLACE, ``h_lp`` [B,H] (compute dtype, row-strided view allowed) receives the copy
    the next GEMMs read."""
    lib = L.load()
    _require(gi, torch.float32, "gi")
    _require(gh, torch.float32, "gh")
    _require(h, torch.float32, "h")
    _require(h_lp, None, "h_lp")
    B, H = h.shape
    if gi.shape != (B, 3 * H) or gh.shape != (B, 3 * H) or h_lp.shape != (B, H) or not (gi.is_contiguous() and
                                                                                        gh.is_contiguous() and h.is_contiguous()):
        raise ValueError("gru_cell: shape mismatch")
    L.check(lib.vqa_gru_cell(gi.data_ptr(), gh.data_ptr(), h.data_ptr(), B, H, dtype_code(h_lp.dtype), h.data_ptr(),
                             h_lp.data_ptr(), h_lp.stride(0), _stream()))
    return h


def lstm_cell(gates, c, h, h_lp):
    """nn.LSTMCell gate update (generator.py:158-159 with rnn_type='LSTM'): gates f32 [B,4H] = W_ih x + b_ih + W_hh h + b_hh
    ([i;f;g;o]); ``c`` and ``h`` f32 [B,H] are updated IN PLACE, ``h_lp`` [B,H] (compute dtype) receives the operand copy."""
    lib = L.load()
    _require(gates, torch.float32, "gates")
    _require(c, torch.float32, "c")
    _require(h, torch.float32, "h")
    _require(h_lp, None, "h_lp")
    B, H = h.shape
    if gates.shape != (B, 4 * H) or c.shape != (B, H) or h_lp.shape != (B, H) or not (gates.is_contiguous() and
                                                                                      c.is_contiguous() and h.is_contiguous()):
        raise ValueError("lstm_cell: shape mismatch")
    L.check(lib.vqa_lstm_cell(gates.data_ptr(), B, H, dtype_code(h_lp.dtype), c.data_ptr(), h.data_ptr(), h_lp.data_ptr(),
                              h_lp.stride(0), _stream()))
    return h, c


def caption_decode_steps(x, proj, batches, att_mode, w_q, q_scale, q_bias, logit_w, logit_bias, gi_prev, w_att, w_hh, b_hh,
                         h, h0_lp, c=None):
    """The teacher-forced time loop of the caption head in one C call (vqa_caption_decode_steps;
    generator.py:99-111 + :168-181).  ``batches``: python list, batch_t per step (non-increasing).
    x [B,K,V], proj [B*K,Hd], gi_prev f32 [B, T*3Hd], h f32 [B,Hd] (updated in place), h0_lp [B,Hd].
    ``c`` f32 [B,Hd] (updated in place) selects the LSTMCell form: gate blocks are then 4Hd wide instead of 3Hd.
    Returns h_all [Σ batch_t, Hd] in the compute dtype (every new state, pack_padded_sequence order)."""
    lib = L.load()
    _require(x, None, "x")
    dt = x.dtype
    ng = 3 if c is None else 4
    if c is not None:
        _require(c, torch.float32, "c")
    for nm, t in (("proj", proj), ("w_q", w_q), ("w_att", w_att), ("w_hh", w_hh), ("h0_lp", h0_lp)):
        _require(t, dt, nm)
    for nm, t in (("q_scale", q_scale), ("q_bias", q_bias), ("logit_w", logit_w), ("gi_prev", gi_prev), ("b_hh", b_hh),
                  ("h", h)):
        if t is not None:
            _require(t, torch.float32, nm)
    B, K, V = x.shape
    Hd, T = h.shape[1], len(batches)
    if (proj.shape != (B * K, Hd) or gi_prev.shape != (B, T * ng * Hd) or w_att.shape != (ng * Hd, V) or
            w_hh.shape != (ng * Hd, Hd) or w_q.shape != (Hd, Hd) or h.shape != (B, Hd) or h0_lp.shape != (B, Hd) or
            (c is not None and (c.shape != (B, Hd) or not c.is_contiguous())) or
            not all(t.is_contiguous() for t in (x, proj, gi_prev, w_att, w_hh, w_q, h, h0_lp))):
        raise ValueError("caption_decode_steps: shape mismatch")
    code = dtype_code(dt)
    h_all = torch.empty((sum(batches), Hd), dtype=dt, device=x.device)
    ws_bytes = lib.vqa_caption_decode_workspace_bytes(B, K, V, Hd, code)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=x.device)
    a = L.CaptionDecodeArgs()
    a.B, a.K, a.V, a.Hd, a.T, a.dtype = B, K, V, Hd, T, code
    sched = (C.c_int * max(T, 1))(*batches)
    a.h_batches = sched
    a.d_x, a.d_proj, a.att_mode = x.data_ptr(), proj.data_ptr(), int(att_mode)
    a.d_wq, a.d_wq_scale, a.d_wq_bias = w_q.data_ptr(), _ptr(q_scale), _ptr(q_bias)
    a.d_logit_w, a.logit_bias, a.d_gi_prev = logit_w.data_ptr(), float(logit_bias), gi_prev.data_ptr()
    a.d_w_att, a.d_w_hh, a.d_b_hh = w_att.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr()
    a.d_h_all, a.d_h, a.d_h0_lp = h_all.data_ptr(), h.data_ptr(), h0_lp.data_ptr()
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
    a.cell, a.d_c = (0, None) if c is None else (1, c.data_ptr())
    L.check(lib.vqa_caption_decode_steps(C.byref(a), _stream()))
    return h_all


def attention_pool(parts, logit_bias, x, want_att=True, want_vsum=True, want_vatt=False):
    """softmax over K + weighted sums (attention.py:86, encoder.py:166, predictor.py:85).

    parts f32 [B*K, n_parts]; x [B,K,V].  Returns (att f32 [B,K], vsum [B,V], vatt [B,K,V]),
    None for outputs not requested."""
    lib = L.load()
    _require(parts, torch.float32, "parts")
    _require(x, None, "x")
    B, K, V = x.shape
    if parts.shape[0] != B * K:
        raise ValueError("attention_pool: parts rows != B*K")
    att = torch.empty((B, K), dtype=torch.float32, device=x.device) if want_att else None
    vsum = torch.empty((B, V), dtype=x.dtype, device=x.device) if want_vsum else None
    vatt = torch.empty((B, K, V), dtype=x.dtype, device=x.device) if want_vatt else None
    L.check(lib.vqa_attention_pool(_ptr(parts), parts.shape[1], float(logit_bias), _ptr(x), B, K, V,
                                   dtype_code(x.dtype), _ptr(att), _ptr(vsum), _ptr(vatt), _stream()))
    return att, vsum, vatt


def graph_attention(Y, att, labels, label_bias, ba, bb, K, want_out=True, want_vsum=False, want_alpha=False):
    """One CorrelatedGraphConv layer + ReLU after the wide projection (gcn.py:93-168,211-212), layout 0.

    Y [B*K, 4V] = x·[W0+W1; W2; Wa; Wb]ᵀ; att f32 [B,K] or None; labels u8 [B,K,K]."""
    lib = L.load()
    _require(Y, None, "Y")
    _require(labels, torch.uint8, "labels")
    _require(label_bias, torch.float32, "label_bias")
    _require(ba, torch.float32, "ba")
    _require(bb, torch.float32, "bb")
    if att is not None:
        _require(att, torch.float32, "att")
    V = Y.shape[1] // 4
    B = Y.shape[0] // K
    out = torch.empty((B, K, V), dtype=Y.dtype, device=Y.device) if want_out else None
    vsum = torch.empty((B, V), dtype=Y.dtype, device=Y.device) if want_vsum else None
    alpha = torch.empty((B, K, K), dtype=torch.float32, device=Y.device) if want_alpha else None
    a = L.GraphAttentionArgs()
    a.d_Y, a.ldy, a.d_att, a.d_labels = Y.data_ptr(), Y.stride(0), _ptr(att), labels.data_ptr()
    a.d_label_bias, a.num_labels, a.d_ba, a.d_bb = label_bias.data_ptr(), label_bias.shape[0], ba.data_ptr(), bb.data_ptr()
    a.B, a.K, a.V, a.dtype = B, K, V, dtype_code(Y.dtype)
    a.d_out, a.d_vsum, a.d_alpha = _ptr(out), _ptr(vsum), _ptr(alpha)
    L.check(lib.vqa_graph_attention(C.byref(a), _stream()))
    return out, vsum, alpha


def graph_attention_merged(Y, x, att, labels, wvec, c0, label_bias_lp, num_labels, K, want_out=True,
                           want_vsum=False, want_alpha=False):
    """Layout 1 (bf16, tcgen05): Y [B*K, 3V] = x·[W0+W1; W2; WbᵀWa]ᵀ, x [B*K, V] the raw features,
    wvec bf16 [16,V] (rows Waᵀbb, Wbᵀba), c0 = ba·bb, label_bias_lp bf16 [16,V]."""
    lib = L.load()
    _require(Y, torch.bfloat16, "Y")
    _require(x, torch.bfloat16, "x")
    _require(wvec, torch.bfloat16, "wvec")
    _require(label_bias_lp, torch.bfloat16, "label_bias_lp")
    _require(labels, torch.uint8, "labels")
    if att is not None:
        _require(att, torch.float32, "att")
    V = Y.shape[1] // 3
    B = Y.shape[0] // K
    out = torch.empty((B, K, V), dtype=Y.dtype, device=Y.device) if want_out else None
    vsum = torch.empty((B, V), dtype=Y.dtype, device=Y.device) if want_vsum else None
    alpha = torch.empty((B, K, K), dtype=torch.float32, device=Y.device) if want_alpha else None
    a = L.GraphAttentionArgs()
    a.d_Y, a.ldy, a.d_att, a.d_labels = Y.data_ptr(), Y.stride(0), _ptr(att), labels.data_ptr()
    a.num_labels = int(num_labels)
    a.B, a.K, a.V, a.dtype = B, K, V, L.VQA_BF16
    a.d_out, a.d_vsum, a.d_alpha = _ptr(out), _ptr(vsum), _ptr(alpha)
    a.layout, a.d_x, a.ldx, a.d_wvec, a.c0 = 1, x.data_ptr(), x.stride(0), wvec.data_ptr(), float(c0)
    a.d_label_bias_lp = label_bias_lp.data_ptr()
    L.check(lib.vqa_graph_attention(C.byref(a), _stream()))
    return out, vsum, alpha


def add_(dst, src):
    """dst += src in place (same shape and dtype, numel % 8 == 0): the sum of the relation branches (encoder.py:257,264)."""
    lib = L.load()
    _require(dst, None, "dst")
    _require(src, dst.dtype, "src")
    if dst.shape != src.shape or not (dst.is_contiguous() and src.is_contiguous()):
        raise ValueError("add_: shape mismatch")
    L.check(lib.vqa_add_inplace(dst.data_ptr(), src.data_ptr(), dst.numel(), dtype_code(dst.dtype), _stream()))
    return dst


def answer_scores(label, target, want_dense=True, want_sum=False):
    """one_hot(label) ⊙ target (wrapper.py:16-22) → (dense [B,A] or None, per-question score [B], sum [1] or None)."""
    lib = L.load()
    _require(label, torch.int64, "label")
    _require(target, torch.float32, "target")
    B, A = target.shape
    dense = torch.empty((B, A), dtype=torch.float32, device=target.device) if want_dense else None
    row = torch.empty((B,), dtype=torch.float32, device=target.device)
    total = torch.empty((1,), dtype=torch.float32, device=target.device) if want_sum else None
    L.check(lib.vqa_answer_scores(label.data_ptr(), target.data_ptr(), B, A, target.stride(0), _ptr(dense), row.data_ptr(),
                                  _ptr(total), _stream()))
    return dense, row, total


def argmax_rows(logits):
    """Lowest-index argmax per row (wrapper.py:14)."""
    lib = L.load()
    _require(logits, torch.float32, "logits")
    B, A = logits.shape
    out = torch.empty((B,), dtype=torch.int64, device=logits.device)
    L.check(lib.vqa_argmax_rows(_ptr(logits), B, A, logits.stride(0), _ptr(out), _stream()))
    return out


# ---- device guard -----------------------------------------------------------------------------------------------
# The library launches on the CURRENT device and the wrappers above take torch.cuda.current_stream() of the current
# device.  A model built on another device than the current one (the reference's --device / --decoder_device,
# main.py:88, wrapper.py:148-150) would otherwise launch on device A's stream with device B's pointers.  Every
# tensor-taking wrapper therefore runs under the device of its first CUDA tensor argument.
def _on_tensor_device(fn):
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for v in list(args) + list(kwargs.values()):
            if isinstance(v, torch.Tensor) and v.is_cuda:
                dev = v.device
                break
            if isinstance(v, (tuple, list, dict)):
                for u in (v.values() if isinstance(v, dict) else v):
                    if isinstance(u, torch.Tensor) and u.is_cuda:
                        dev = u.device
                        break
                if dev is not None:
                    break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


for _name in ("relation_labels", "cast_to_bf16", "cast_to_f32", "linear", "gru_last_state", "gru_sequence",
              "lstm_sequence", "caption_gate_scale", "seq_max", "softmax_mul", "attention_logits", "gru_cell",
              "lstm_cell", "caption_decode_steps", "attention_pool", "graph_attention", "graph_attention_merged",
              "add_", "answer_scores", "argmax_rows", "split_f32", "linear_split", "attention_pool_split",
              "gru_last_state_split"):
    globals()[_name] = _on_tensor_device(globals()[_name])
del _name
