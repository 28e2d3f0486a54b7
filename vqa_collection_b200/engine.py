"""Whole-path engine: prepared weights + one C call per forward.

``VQAEngine`` owns device copies of the weights in the layout the kernels want
(bf16 or f32, weight-norm scalars expanded to per-column scale vectors, W_q and
q_net concatenated, the four ReGAT maps concatenated) and a workspace, and runs
``vqa_forward`` (include/vqa_b200.h) — the B200 replacement for
``Wrapper.forward`` / ``forward_vqa`` (wrapper.py:64-74,113-118) with
encoder_type in {base, relation}, att_type in {'new', 'base'}, predictor 'base'.

Weights arrive under the reference's parameter names, e.g. from
``Wrapper.state_dict()`` plus the unregistered GCN tensors as ``gcn.{i}.*``.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops


def weight_norm_scale(v: torch.Tensor, g: torch.Tensor) -> float:
    """s = g/‖v‖_F evaluated with the same torch CPU op as the reference's
    weight_norm hook (modules.py:38; SURVEY.md F13/H9: the fp32 norm is only
    1e-5..1e-4 accurate, so it must be THIS op, not a better one)."""
    v = v.detach().to("cpu", torch.float32)
    g = g.detach().to("cpu", torch.float32)
    return float(g / torch.norm(v))


def _pad_cols(t: torch.Tensor, cols: int) -> torch.Tensor:
    if t.shape[1] == cols:
        return t
    out = torch.zeros((t.shape[0], cols), dtype=t.dtype)
    out[:, : t.shape[1]] = t
    return out


def host_cores_per_rank() -> int:
    """host cores this process may count on: all of them, divided by the ranks that share the box (torchrun)"""
    import os
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, n // max(1, local))


def host_pack_threads() -> int:
    """threads of the f32→bf16 packing pool of the host path (VQA_B200_PACK_THREADS overrides)"""
    import os
    v = int(os.environ.get("VQA_B200_PACK_THREADS", "0") or 0)
    return v if v > 0 else host_cores_per_rank()


def host_raw_chunk_period(cores: int = None) -> int:
    """Hybrid staging of the host path: every n-th chunk crosses PCIe as raw f32 (cast on the device) so that the host
    cores and the link finish together.  Measured on a 16-core slice per GPU: 3 of 4 chunks packed is best; with fewer
    cores per GPU more chunks go raw (0 = pack everything, 1 = send everything raw)."""
    cores = host_cores_per_rank() if cores is None else cores
    if cores >= 24:
        return 0
    if cores >= 12:
        return 4
    if cores >= 6:
        return 3
    if cores >= 3:
        return 2
    return 1


def pack_gru(w_ih, w_hh, b_ih, b_hh, units: int = 64):
    """Gate-interleaved GRU weights for the persistent fused kernel (vqa_gru_args):
    192-row block j = [r | z | n] rows of hidden units [64j, 64j+64); biases
    [b_ir+b_hr | b_iz+b_hz | b_in | b_hn].  Inputs in torch layout [3H, ·] / [3H]."""
    H = w_hh.shape[1]
    if H % units:
        return None
    j = torch.arange(H // units).view(-1, 1, 1)
    gate = torch.arange(3).view(1, -1, 1)
    u = torch.arange(units).view(1, 1, -1)
    perm = (gate * H + j * units + u).reshape(-1).to(w_ih.device)
    bias = torch.cat([b_ih[:H] + b_hh[:H], b_ih[H:2 * H] + b_hh[H:2 * H], b_ih[2 * H:], b_hh[2 * H:]])
    return w_ih[perm].contiguous(), w_hh[perm].contiguous(), bias.float().contiguous()


GI_TABLE_MAX_BYTES = 1 << 30          # vocabulary x 3H fp16 entries; beyond this the x-part GEMM stays


def _token_table_f32(emb, w_ih, bias, device):
    """emb [rows,E] · w_ih[3H,E]ᵀ + bias in f32 on the device, with the library's own fp32 GEMM (vqa_linear, FFMA)"""
    E_pad = (emb.shape[1] + 15) // 16 * 16
    to = lambda t: _pad_cols(t.detach().float().cpu(), E_pad).contiguous().to(device)
    return ops.linear(to(emb), to(w_ih), None, bias.detach().float().contiguous().to(device), out_dtype=torch.float32)


def gru_token_table(emb, w_ih, b_ih, b_hh, device):
    """Input half of the GRU gates per TOKEN (vqa_gru_args.d_gi_table): row v = W_ih·emb[v] + (b_ir+b_hr | b_iz+b_hz |
    b_in), fp16 [rows, 3H] stored as [rows, H/32, 3, 32].  modules.py:153 evaluates W_ih·x_t for every (sample, step); x_t = emb[token] takes one of
    `rows` values, so the product is a property of the weights.  Built from the f32 master weights (f32 matmul), i.e. it is
    closer to the reference than the bf16 x-part GEMM it replaces.  None when the table would be too large or an entry
    does not fit fp16."""
    H = w_ih.shape[0] // 3
    if emb.shape[0] * 3 * H * 2 > GI_TABLE_MAX_BYTES or torch.device(device).type != "cuda":
        return None                                          # (host-side layout checks prepare weights on the CPU)
    bias = b_ih.clone()
    bias[:2 * H] += b_hh[:2 * H]
    gi = _token_table_f32(emb, w_ih, bias, device)
    if H % 32 or not bool(torch.isfinite(gi).all()) or float(gi.abs().max()) > 6.0e4:
        return None
    # kernel layout: per 32-unit block the gates r | z | n side by side, so that the 3 x UNITS entries a CTA pair needs of a
    # row are ONE contiguous piece (a single bulk copy per row and step)
    rows = gi.shape[0]
    gi = gi.view(rows, 3, H // 32, 32).permute(0, 2, 1, 3).reshape(rows, 3 * H)
    return gi.to(torch.float16).contiguous()


def prepare_weights(W: dict, dtype: torch.dtype, device, relation: bool, gcn_layer: int = 0, K: int = 36,
                    split: bool = False) -> dict:
    """Reference-named tensors → device tensors in kernel layout.  ``split`` (precision 'fp32tc'): every GEMM weight
    becomes the fp16 plane pair [2, out, in] of VQA_F16X2 (ops.split_f32 of the f32 weight), everything else stays f32."""
    f32 = lambda t: t.detach().to("cpu", torch.float32)
    _dev = lambda t, dt=torch.float32: t.to(dt).contiguous().to(device)

    def dev(t, dt=torch.float32):
        if split and dt is dtype and t.dim() == 2:           # a GEMM operand in the compute format
            return ops.split_f32(_dev(t))
        return _dev(t, dt)
    if split:
        dtype = _SPLIT                                       # marker: dev(t, dtype) makes plane pairs
    P = {}
    emb = f32(W["encoder.embedding.weight"])
    E = emb.shape[1]
    E_pad = (E + 63) // 64 * 64
    P["E"], P["E_pad"], P["ntoken_rows"] = E, E_pad, emb.shape[0]
    r = "encoder.q_rnn.rnn."
    P["emb"] = _dev(_pad_cols(emb, E_pad), torch.float32 if split else dtype)
    P["w_ih"] = _dev(_pad_cols(f32(W[r + "weight_ih_l0"]), E_pad), torch.float32 if split else dtype)
    P["w_hh"] = dev(f32(W[r + "weight_hh_l0"]), dtype)
    P["b_ih"] = dev(f32(W[r + "bias_ih_l0"]))
    P["b_hh"] = dev(f32(W[r + "bias_hh_l0"]))
    H = P["w_hh"].shape[-1]
    if split:
        # gate-interleaved W_hh planes: 96-row block j = [r | z | n] rows of units [32j, 32j+32), one GEMM tile per block, so
        # the gate update can be the step GEMM's epilogue
        if H % 32 == 0:
            w_hh32 = f32(W[r + "weight_hh_l0"])
            P["wh_packed"] = ops.split_f32(_dev(pack_gru(w_hh32, w_hh32, P["b_ih"].cpu(), P["b_hh"].cpu(), units=32)[1]))
        # the input half of the gates per token, f32: W_ih·emb[v] + b_ih (vqa_gru_args.d_gi_table in the f16x2 mode)
        P["gi_table"] = _token_table_f32(emb, f32(W[r + "weight_ih_l0"]), P["b_ih"], device)
    if dtype == torch.bfloat16:
        packed = pack_gru(P["w_ih"], P["w_hh"], P["b_ih"], P["b_hh"])
        if packed is not None:
            P["wx_packed"], P["wh_packed"], P["bias_packed"] = packed
            tab = gru_token_table(emb, f32(W[r + "weight_ih_l0"]), f32(W[r + "bias_ih_l0"]), f32(W[r + "bias_hh_l0"]), device)
            if tab is not None:
                P["gi_table"] = tab

    def wn(prefix):
        v, g, b = f32(W[prefix + ".weight_v"]), f32(W[prefix + ".weight_g"]), f32(W[prefix + ".bias"])
        return v, weight_norm_scale(v, g), b

    vn, sn, bn = wn("encoder.q_net.main.0")
    P["att_concat"] = "encoder.attention.sequence.0.weight_v" in W
    if not P["att_concat"]:                      # MultiplyAttention (attention.py:54-86)
        v, s, b = wn("encoder.attention.W_v.main.0")
        P["Wv"], P["sv"], P["bv"] = dev(v, dtype), dev(torch.full((v.shape[0],), s)), dev(b)
        vq, sq, bq = wn("encoder.attention.W_q.main.0")
        P["Wqq"] = dev(torch.cat([vq, vn], 0), dtype)
        P["sqq"] = dev(torch.cat([torch.full((vq.shape[0],), sq), torch.full((vn.shape[0],), sn)]))
        P["bqq"] = dev(torch.cat([bq, bn]))
        vl, sl, bl = wn("encoder.attention.linear")
    else:                                        # ConcatAttention (attention.py:18-51): split W1 = [W1v | W1q]
        v, s, b = wn("encoder.attention.sequence.0")
        Vd = v.shape[1] - H
        P["Wv"], P["sv"], P["bv"] = dev(v[:, :Vd], dtype), dev(torch.full((v.shape[0],), s)), dev(torch.zeros(v.shape[0]))
        P["W1q"], P["b1"] = dev(v[:, Vd:], dtype), dev(b)
        P["Wqq"], P["sqq"], P["bqq"] = dev(vn, dtype), dev(torch.full((vn.shape[0],), sn)), dev(bn)
        vl, sl, bl = wn("encoder.attention.sequence.2")
    P["wlin"] = dev((vl * sl).reshape(-1))
    P["b_lin"] = float(bl.reshape(-1)[0])
    v, s, b = wn("predictor.v_net.main.0")
    P["Wvn"], P["svn"], P["bvn"] = dev(v, dtype), dev(torch.full((v.shape[0],), s)), dev(b)
    v, s, b = wn("predictor.classifier.main.0")
    P["Wc0"], P["sc0"], P["bc0"] = dev(v, dtype), dev(torch.full((v.shape[0],), s)), dev(b)
    v, s, b = wn("predictor.classifier.main.3")
    P["Wc1"], P["sc1"], P["bc1"] = dev(v, dtype), dev(torch.full((v.shape[0],), s)), dev(b)
    P["H"], P["V"], P["A"] = H, P["Wv"].shape[-1], P["Wc1"].shape[-2]
    if relation:
        p = f"gcn.{gcn_layer}."
        P.update(prepare_gcn_layer({k[len(p):]: t for k, t in W.items() if k.startswith(p)},
                                   torch.float32 if split else dtype, device, K, merged=False if split else None))
        if split:
            P["Wg"] = ops.split_f32(P["Wg"])
    return P


_SPLIT = object()


def gat_tc_supported(dtype, V: int, K: int = 36) -> bool:
    """shapes the tcgen05 graph-attention kernel (vqa_graph_attention layout 1) is built for"""
    return dtype == torch.bfloat16 and K == 36 and V % 128 == 0


def prepare_gcn_layer(Wl: dict, dtype, device, K: int = 36, merged=None) -> dict:
    """One CorrelatedGraphConv layer (gcn.py:55-67,113-117).

    fp32 (layout 0): concatenate [W0+W1 ; W2 ; Wa ; Wb] so a single pass over x feeds all four maps.
    bf16 (layout 1): the DotProduct maps are merged exactly, (Wa x_i + ba)·(Wb x_j + bb) =
    (WbᵀWa x_i)·x_j + x_i·(Waᵀbb) + x_j·(Wbᵀba) + ba·bb, so the wide GEMM has 3 maps
    [W0+W1 ; W2 ; WbᵀWa] and the rank-1 terms ride along as two extra MMA columns (`wvec`)."""
    f32 = lambda t: t.detach().to("cpu", torch.float32)
    dev = lambda t, dt=torch.float32: t.to(dt).contiguous().to(device)
    w01 = f32(Wl["weight.0.weight"]) + f32(Wl["weight.1.weight"])
    w2, wa, wb = f32(Wl["weight.2.weight"]), f32(Wl["dot_product.wa.weight"]), f32(Wl["dot_product.wb.weight"])
    ba, bb, lbias = f32(Wl["dot_product.wa.bias"]), f32(Wl["dot_product.wb.bias"]), f32(Wl["bias"])
    V, L = w2.shape[1], lbias.shape[0]
    P = {"num_labels": L}
    if merged is None:
        merged = gat_tc_supported(dtype, V, K) and w2.shape[0] == V and L <= 16
    if merged:
        wq = wb.double().t().matmul(wa.double()).float()
        wvec = torch.zeros((16, V))
        wvec[0] = wa.double().t().matmul(bb.double()).float()
        wvec[1] = wb.double().t().matmul(ba.double()).float()
        lb = torch.zeros((16, V))
        lb[:L] = lbias
        P.update({"Wg3": dev(torch.cat([w01, w2, wq], 0), dtype), "wvec": dev(wvec, dtype),
                  "gat_c0": float(ba.double().dot(bb.double())), "label_bias_lp": dev(lb, dtype)})
    else:
        P.update({"Wg": dev(torch.cat([w01, w2, wa, wb], 0), dtype), "label_bias": dev(lbias),
                  "ba": dev(ba), "bb": dev(bb)})
    return P


class VQAEngine:
    """B200 forward engine for the Up-Down (+ReGAT) VQA path."""

    def __init__(self, weights: dict, relation: bool = False, precision: str = "bf16",
                 device="cuda", num_objs: int = 36, overlap=None, side_sms: int = 0, side_tile_permille: int = 0,
                 gat_chase_sms=None):
        """Two-stream schedules of ``vqa_forward`` (bf16 only, see include/vqa_b200.h; measurements in DESIGN.md §3.6):
        ``overlap`` (B >= 512): the question encoder on a side stream on ``side_sms`` SMs while the question-independent
        projection of the region features runs on the other SMs.  ReGAT: 2 % faster than the serial order (the wide
        projection keeps 84 SMs busy while the latency-bound GRU runs) -> None = on; Up-Down: 5 % slower (the W_v projection
        must then be stored and reduced by a second kernel) -> None = off.  VQA_B200_OVERLAP=0/1 overrides both.
        ``gat_chase_sms`` (ReGAT): the graph attention on that many SMs BESIDE the wide projection, consuming Y row block
        by row block.  Correct but not faster (one SM of the graph attention moves 33 GB/s whatever the memory system
        does, so it needs > 32 SMs to keep up with the GEMM): None = off unless VQA_B200_GAT_CHASE=n."""
        import os
        self.lib = L.load()
        self.device = torch.device(device)
        if self.device.index is None and self.device.type == "cuda":
            self.device = torch.device("cuda", torch.cuda.current_device())
        if overlap is None:
            env = os.environ.get("VQA_B200_OVERLAP", "")
            overlap = (env == "1") if env in ("0", "1") else bool(relation)
        if gat_chase_sms is None:
            gat_chase_sms = int(os.environ.get("VQA_B200_GAT_CHASE", "0") or 0)
        self.overlap, self.side_sms, self.side_tile_permille = bool(overlap), int(side_sms), int(side_tile_permille)
        self.gat_chase_sms = int(gat_chase_sms) if relation else 0
        # 'fp32tc': fp32-class arithmetic on the tensor cores (VQA_F16X2: fp16 plane pairs, three tcgen05.mma per k-step);
        # the wire format stays f32, the resident format of the features is the plane pair [2,B,K,V]
        self.dtype = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp32tc": torch.float32}[precision]
        self.split = precision == "fp32tc"
        self.code = L.VQA_F16X2 if self.split else ops.dtype_code(self.dtype)
        self.precision = precision
        self.relation = bool(relation)
        self.K = num_objs
        # token-table form of the fused GRU (bf16): on unless VQA_B200_GRU_TABLE=0
        self.use_gi_table = os.environ.get("VQA_B200_GRU_TABLE", "1") != "0"
        if self.split:
            self.overlap, self.gat_chase_sms = False, 0
        with torch.cuda.device(self.device):
            self.P = prepare_weights(weights, self.dtype, self.device, self.relation, K=num_objs, split=self.split)
        self._ws = {}
        self.last_launches = 0

    # -- helpers ---------------------------------------------------------------
    def _args(self, B, T):
        P, a = self.P, L.ForwardArgs()
        a.B, a.K, a.V, a.H, a.A, a.T = B, self.K, P["V"], P["H"], P["A"], T
        a.E_pad, a.ntoken_rows = P["E_pad"], P["ntoken_rows"]
        a.num_labels = P.get("num_labels", 0)
        a.dtype, a.relation = self.code, int(self.relation)
        a.overlap, a.side_sms, a.side_tile_permille = int(self.overlap), self.side_sms, self.side_tile_permille
        a.gat_chase_sms = self.gat_chase_sms
        for name in ("emb", "w_ih", "b_ih", "w_hh", "b_hh", "Wv", "sv", "bv", "Wqq", "sqq", "bqq", "wlin",
                     "Wvn", "svn", "bvn", "Wc0", "sc0", "bc0", "Wc1", "sc1", "bc1"):
            setattr(a, "d_" + name, P[name].data_ptr())
        a.b_lin = P["b_lin"]
        if P["att_concat"]:
            a.att_concat, a.d_W1q, a.d_b1 = 1, P["W1q"].data_ptr(), P["b1"].data_ptr()
        for name in ("wx_packed", "wh_packed", "bias_packed"):
            if name in P:
                setattr(a, "d_" + name, P[name].data_ptr())
        if "gi_table" in P and (self.split or self.use_gi_table):
            a.d_gi_table = P["gi_table"].data_ptr()
        if self.relation:
            if "Wg3" in P:
                for name in ("Wg3", "wvec", "label_bias_lp"):
                    setattr(a, "d_" + name, P[name].data_ptr())
                a.gat_c0 = P["gat_c0"]
            else:
                for name in ("Wg", "label_bias", "ba", "bb"):
                    setattr(a, "d_" + name, P[name].data_ptr())
        return a

    def _workspace(self, a):
        need = self.lib.vqa_forward_workspace_bytes(C.byref(a))
        key = (a.B, a.T)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws, need

    def resident(self, img: torch.Tensor) -> torch.Tensor:
        """Wire-format f32 features → the resident compute dtype (bf16 cast kernel)."""
        if self.split:
            if img.dtype == torch.float16 and img.dim() == 4 and img.shape[0] == 2:
                return img
            if img.dtype == torch.float32 and img.dim() == 3:
                return ops.split_f32(img.contiguous())
            raise TypeError("an fp32tc engine takes f32 [B,K,V] features or their fp16 plane pair [2,B,K,V]")
        if img.dtype == self.dtype:
            return img
        if img.dtype == torch.float32 and self.dtype == torch.bfloat16:
            return ops.cast_to_bf16(img.contiguous())
        raise TypeError(f"img dtype {img.dtype} cannot feed a {self.precision} engine")

    # -- device-resident forward -------------------------------------------------
    def forward(self, img, tokens, labels=None, bbox=None, wh=None, want_v=False, want_q=False,
                want_alpha=False):
        """img [B,K,V] (CUDA, f32 or the engine dtype), tokens int64 [B,T] (CUDA);
        relation engines also need ``labels`` u8 [B,K,K] or ``bbox`` f32 [B,K,4] + ``wh``=(w,h).
        Returns dict(logits f32 [B,A], label int64 [B], att f32 [B,K], ...)."""
        if not (img.is_cuda and tokens.is_cuda):
            raise RuntimeError("VQAEngine.forward needs CUDA tensors (use forward_host for host buffers)")
        n_cast = int(img.dtype != (torch.float16 if self.split else self.dtype))
        img = self.resident(img).contiguous()
        tokens = tokens.contiguous()
        B, K, V = img.shape[-3:]
        if self.split and want_v:
            raise NotImplementedError("fp32tc engine: the encoder output 'v' is not produced (use precision='fp32')")
        if K != self.K or V != self.P["V"]:
            raise ValueError(f"img must be [B,{self.K},{self.P['V']}], got {tuple(img.shape)}")
        a = self._args(B, tokens.shape[1])
        ws, need = self._workspace(a)
        dev = self.device
        out = {
            "logits": torch.empty((B, a.A), dtype=torch.float32, device=dev),
            "label": torch.empty((B,), dtype=torch.int64, device=dev),
            "att": torch.empty((B, K), dtype=torch.float32, device=dev),
        }
        a.d_img, a.d_tokens = img.data_ptr(), tokens.data_ptr()
        a.d_workspace, a.workspace_bytes = ws.data_ptr(), need
        a.d_logits, a.d_label, a.d_att = out["logits"].data_ptr(), out["label"].data_ptr(), out["att"].data_ptr()
        if want_v:
            out["v"] = torch.empty((B, K, V), dtype=self.dtype, device=dev)
            a.d_v = out["v"].data_ptr()
        if want_q:
            out["q"] = torch.empty((B, a.H), dtype=torch.float32, device=dev)
            a.d_q = out["q"].data_ptr()
        if self.relation:
            if labels is not None:
                if labels.dtype != torch.uint8:
                    labels = labels.to(torch.uint8)          # loader format is float64 (dataset.py:102)
                labels = labels.contiguous()
                a.d_labels = labels.data_ptr()
                out["labels"] = labels
            elif bbox is not None:
                if wh is None:
                    raise ValueError("bbox needs wh=(img_w, img_h)")
                bbox = bbox.contiguous()
                out["labels"] = torch.empty((B, K, K), dtype=torch.uint8, device=dev)
                a.d_bbox, a.d_labels_out = bbox.data_ptr(), out["labels"].data_ptr()
                a.img_w, a.img_h = float(wh[0]), float(wh[1])
            else:
                raise ValueError("relation engine needs labels or bbox")
            if want_alpha:
                out["alpha"] = torch.empty((B, K, K), dtype=torch.float32, device=dev)
                a.d_alpha = out["alpha"].data_ptr()
        if B == 0:                                   # empty shard: nothing to launch, empty outputs
            self.last_launches = 0
            return out
        # the library launches on the CURRENT device: make that the engine's device, and use ITS current stream
        with torch.cuda.device(self.device):
            L.check(self.lib.vqa_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
            self.last_launches = self.lib.vqa_forward_last_launch_count() + n_cast
        return out

    def capture(self, img, tokens, labels=None, bbox=None, wh=None, warmup: int = 2, **want):
        """Capture ``forward`` on fixed (resident) inputs into a CUDA graph: returns ``(graph, out)``; every
        ``graph.replay()`` recomputes ``out`` from the CURRENT contents of ``img`` / ``tokens`` / ``labels``.  One
        graph launch instead of 8-12 kernel launches + tensor-map encodes per step: the steady-state form for a
        serving loop over resident batches (all pointers of the path are fixed, tensor maps are by-value parameters)."""
        if img.dtype != (torch.float16 if self.split else self.dtype):
            raise TypeError("capture() needs inputs already in the engine's resident format (use resident())")
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):          # the first, uncaptured calls set attributes / create the side stream
                    self.forward(img, tokens, labels=labels, bbox=bbox, wh=wh, **want)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                out = self.forward(img, tokens, labels=labels, bbox=bbox, wh=wh, **want)
        return g, out

    # -- host-buffer forward (the e2e path) ----------------------------------------
    def _next_host_ctx(self):
        """two host contexts, used alternately, so that batch n+1 can be staged while batch n computes"""
        import os
        ctxs = getattr(self, "_host_ctxs", None)
        if ctxs is None:
            ctxs = self._host_ctxs = []
            self._host_turn = 0
        if len(ctxs) < 2:
            ctx = C.c_void_p()
            with torch.cuda.device(self.device):
                L.check(self.lib.vqa_host_ctx_create(C.byref(ctx), host_pack_threads()))
            ctxs.append(ctx)
            return ctx
        self._host_turn ^= 1
        return ctxs[self._host_turn]

    def forward_host(self, img_h, tokens_h, labels_h=None, bbox_h=None, wh=None, chunk=64, pack_on_host=True,
                     raw_chunk_period=0):
        """synchronous form of forward_host_async: returns (label int64 [B] on host, bytes_h2d, bytes_d2h)"""
        return self.forward_host_async(img_h, tokens_h, labels_h, bbox_h, wh, chunk, pack_on_host, raw_chunk_period).result()

    def forward_host_async(self, img_h, tokens_h, labels_h=None, bbox_h=None, wh=None, chunk=64, pack_on_host=True,
                           raw_chunk_period=0):
        """Host buffers in the reference wire format → a handle whose ``result()`` gives the host answers
        (vqa_forward_host_submit / _wait).

        img_h f32 [B,K,V] (pinned recommended; or bf16 [B,K,V] when the caller keeps its feature cache in the
        resident format — then chunks go straight to HBM), tokens_h int64 [B,T] (+ labels_h u8 [B,K,K] or bbox_h f32
        [B,K,4] with wh).  bf16 engines: ``pack_on_host`` lets the host cores convert the features to bf16
        (bit-identical to the device cast) while the previous chunk is in flight, so PCIe carries 2 bytes per
        feature; otherwise f32 crosses PCIe and is cast on the device.  Up to two batches may be in flight
        (two staging contexts): submit batch n+1 before asking for batch n's result to overlap staging with
        compute."""
        wire_bf16 = img_h.dtype == torch.bfloat16          # a host-side feature cache already in the resident format
        if wire_bf16 and self.dtype != torch.bfloat16:
            raise TypeError("forward_host: bf16 host features need a bf16 engine")
        for name, t, dt in (("img_h", img_h, img_h.dtype if wire_bf16 else torch.float32), ("tokens_h", tokens_h, torch.int64)):
            if t.is_cuda or t.dtype != dt or not t.is_contiguous():
                raise TypeError(f"forward_host: {name} must be a contiguous CPU {dt} tensor")
        B, K, V = img_h.shape
        if K != self.K or V != self.P["V"]:
            raise ValueError(f"img must be [B,{self.K},{self.P['V']}], got {tuple(img_h.shape)}")
        ctx = self._next_host_ctx()
        if not hasattr(self, "_host_keep"):
            self._host_keep = {}
        a = self._args(B, tokens_h.shape[1])
        ws, need = self._workspace(a)
        dev = self.device
        logits = torch.empty((B, a.A), dtype=torch.float32, device=dev)
        att = torch.empty((B, K), dtype=torch.float32, device=dev)
        a.d_workspace, a.workspace_bytes = ws.data_ptr(), need
        a.d_logits, a.d_att = logits.data_ptr(), att.data_ptr()
        ha = L.ForwardHostArgs()
        ha.h_img, ha.h_tokens = img_h.data_ptr(), tokens_h.data_ptr()
        keep = []
        if self.relation:
            if labels_h is not None:
                if labels_h.dtype != torch.uint8:
                    labels_h = labels_h.to(torch.uint8)      # loader format is float64 (dataset.py:102)
                labels_h = labels_h.contiguous()
                keep.append(labels_h)
                ha.h_labels = labels_h.data_ptr()
            elif bbox_h is not None:
                if wh is None:
                    raise ValueError("bbox needs wh=(img_w, img_h)")
                bbox_h = bbox_h.float().contiguous()
                keep.append(bbox_h)
                ha.h_bbox = bbox_h.data_ptr()
                a.img_w, a.img_h = float(wh[0]), float(wh[1])
            else:
                raise ValueError("relation engine needs labels_h or bbox_h")
        label_h = torch.empty((B,), dtype=torch.int64)
        ha.fwd, ha.h_label = a, label_h.data_ptr()
        if int(raw_chunk_period) == 1:                   # "everything raw" (host_raw_chunk_period on a small core slice)
            pack_on_host, raw_chunk_period = False, 0
        ha.chunk_rows, ha.pack_on_host = int(chunk), int(bool(pack_on_host) and self.dtype == torch.bfloat16)
        ha.raw_chunk_period = int(raw_chunk_period) if ha.pack_on_host else 0
        ha.img_is_bf16 = int(wire_bf16)
        L.check(self.lib.vqa_forward_host_wait(ctx))             # a context carries one batch at a time
        # everything the C context points into (it keeps raw pointers to the host buffers and to h_label until
        # vqa_forward_host_wait) lives on the ENGINE, keyed by context, until the next wait on that context returns —
        # a caller that drops the handle without result() cannot free memory that is still being read or written
        self._host_keep[ctx.value] = (label_h, keep, img_h, tokens_h, ha, logits, att, ws)
        with torch.cuda.device(self.device):
            L.check(self.lib.vqa_forward_host_submit(ctx, C.byref(ha), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        n_chunks = (B + chunk - 1) // chunk
        self.last_launches = self.lib.vqa_forward_last_launch_count() + (
            (n_chunks if not ha.pack_on_host else (n_chunks // ha.raw_chunk_period if ha.raw_chunk_period else 0))
            if ((self.dtype == torch.bfloat16 and not wire_bf16) or self.split) else 0)
        self.last_host_outputs = {"logits": logits, "att": att}
        eng = self

        class _Pending:
            def result(self_inner):
                L.check(eng.lib.vqa_forward_host_wait(ctx))
                return label_h, int(ha.h2d_bytes), int(ha.d2h_bytes)
        return _Pending()

    def __del__(self):
        for ctx in getattr(self, "_host_ctxs", None) or []:
            try:
                self.lib.vqa_forward_host_wait(ctx)      # a submitted batch still reads / writes the kept host buffers
                self.lib.vqa_host_ctx_destroy(ctx)
            except Exception:
                pass
        self._host_ctxs = None
        self._host_keep = None
