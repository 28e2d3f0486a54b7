"""Training step of the Up-Down path on the C-ABI kernels (BASELINE config 4).

``updown_loss`` is what ``Wrapper.get_loss`` (wrapper.py:76-105) calls in place of
forward + instance_bce_with_logits: ONE C call (vqa_updown_train_step) runs the forward with
saved activations, the loss and the whole backward; the returned loss is attached to the
autograd graph by a torch.autograd.Function whose backward only hands the already computed
parameter gradients (scaled by grad_output) to autograd — so the reference's unchanged
train loop (loss.backward(); clip_grad_norm_; optimizer.step(), train.py:108-110) works as is.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops

# (module path, layer index) in the order vqa_train_args expects
WN_LAYERS = [
    "encoder.attention.W_v.main.0", "encoder.attention.W_q.main.0", "encoder.attention.linear",
    "encoder.q_net.main.0", "predictor.v_net.main.0", "predictor.classifier.main.0", "predictor.classifier.main.3",
]
GRU_PARAMS = ["encoder.embedding.weight", "encoder.q_rnn.rnn.weight_ih_l0", "encoder.q_rnn.rnn.weight_hh_l0",
              "encoder.q_rnn.rnn.bias_ih_l0", "encoder.q_rnn.rnn.bias_hh_l0"]


def param_names():
    """the 26 reference parameter names, in the order of UpDownTrainStep's *params"""
    names = list(GRU_PARAMS)
    for p in WN_LAYERS:
        names += [p + ".weight_v", p + ".weight_g", p + ".bias"]
    return names


_WS = {}
_FLAT = {}
_GROUP = None
_dp_disabled = False          # tests: run a step without the exchange inside an initialised process group


def set_process_group(group):
    """data-parallel group whose ranks average their gradients (default: the world group when
    torch.distributed is initialised with more than one rank)"""
    global _GROUP
    _GROUP = group


def _dp_active():
    import torch.distributed as dist
    return (not _dp_disabled) and dist.is_available() and dist.is_initialized() and dist.get_world_size(_GROUP) > 1


def _workspace(lib, a, device):
    need = lib.vqa_train_workspace_bytes(C.byref(a))
    key = (device, a.B, a.T, a.dtype)
    ws = _WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws, need


_BUCKET = {}


def _bucket_event(dev):
    """(head-done event, side stream) of a device; the event object exists in the driver (it has been recorded once), so
    its raw handle can be handed to the C call"""
    hit = _BUCKET.get(dev)
    if hit is None:
        with torch.cuda.device(dev):
            ev = torch.cuda.Event()
            ev.record()
            hit = _BUCKET[dev] = (ev, torch.cuda.Stream(dev))
    return hit


class UpDownTrainStep(torch.autograd.Function):
    """(img [B,K,V] compute dtype, tokens int64 [B,T], target f32 [B,A], p_att, p_cls, seed, *26 params)
    → (loss 0-d f32, logits f32 [B,A])"""

    @staticmethod
    def forward(ctx, img, tokens, target, p_att, p_cls, seed, *params):
        lib = L.load()
        ctx_params = params
        if len(params) != len(GRU_PARAMS) + 3 * len(WN_LAYERS):
            raise ValueError("UpDownTrainStep: expected 26 parameters")
        for t in (img, tokens, target) + tuple(params):
            if not t.is_cuda:
                raise RuntimeError("UpDownTrainStep needs CUDA tensors (no CPU fallback)")
        img, tokens, target = img.contiguous(), tokens.contiguous(), target.float().contiguous()
        params = [p.detach().float().contiguous() for p in params]
        dev = img.device
        B, K, V = img.shape
        emb, w_ih, w_hh = params[0], params[1], params[2]
        a = L.TrainArgs()
        a.B, a.K, a.V, a.T = B, K, V, tokens.shape[1]
        a.H, a.E, a.ntoken_rows = w_hh.shape[1], emb.shape[1], emb.shape[0]
        a.A = params[5 + 3 * 6].shape[0]
        a.dtype = ops.dtype_code(img.dtype)
        a.dropout_att, a.dropout_cls, a.seed = float(p_att), float(p_cls), int(seed) & (2 ** 64 - 1)
        a.d_img, a.d_tokens, a.d_target = img.data_ptr(), tokens.data_ptr(), target.data_ptr()
        from .parallel import FlatGradients, average_gradients_
        shapes = [tuple(p.shape) for p in params]
        fg = _FLAT.get(dev)
        # the gradients of the PREVIOUS step may still be alive as views of the flat buffer (backward hands them to
        # autograd without copies): reuse the buffer only when no parameter's .grad aliases it (i.e. after zero_grad)
        alias = fg is not None and any(p.grad is not None and p.grad.untyped_storage().data_ptr() ==
                                       fg.flat.untyped_storage().data_ptr() for p in ctx_params)
        # ... and only when no earlier get_loss still waits for its backward() (two losses summed before one backward)
        if fg is None or alias or getattr(fg, "pending", False) or not fg.matches(shapes, dev):
            fg = _FLAT[dev] = FlatGradients(shapes, dev)
        fg.pending = True
        grads = fg.views
        a.p_emb, a.p_w_ih, a.p_w_hh, a.p_b_ih, a.p_b_hh = (p.data_ptr() for p in params[:5])
        a.g_emb, a.g_w_ih, a.g_w_hh, a.g_b_ih, a.g_b_hh = (g.data_ptr() for g in grads[:5])
        for l in range(len(WN_LAYERS)):
            a.p_v[l], a.p_g[l], a.p_b[l] = (params[5 + 3 * l + j].data_ptr() for j in range(3))
            a.g_v[l], a.g_g[l], a.g_b[l] = (grads[5 + 3 * l + j].data_ptr() for j in range(3))
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        logits = torch.empty((B, a.A), dtype=torch.float32, device=dev)
        a.d_loss, a.d_logits = loss.data_ptr(), logits.data_ptr()
        ws, need = _workspace(lib, a, dev)
        a.d_workspace, a.workspace_bytes = ws.data_ptr(), need
        # Called under no_grad / with frozen parameters (evaluation through get_loss): nobody will run backward(), so the
        # gradients are dropped — no all-reduce is started (ranks that skip differently must not desynchronise) and
        # the flat buffer is free for the next step.
        wants_grad = any(ctx.needs_input_grad)
        dp = _dp_active() and wants_grad
        ev = side = None
        if dp:
            ev, side = _bucket_event(dev)
            a.ev_head_done = ev.cuda_event
        with torch.cuda.device(dev):                 # the library launches on the current device
            L.check(lib.vqa_updown_train_step(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        if not wants_grad:
            fg.pending = False
        # The exchange step of the data-parallel training path: average the shard gradients over ranks, in TWO buckets in
        # the order the backward pass finishes them.  Bucket 1 = the seven weight-normed layers (59 of 75.5 MB), final
        # before the back-propagation through time: its all-reduce is issued from a side stream that only waits for the
        # step's head-done event, so NCCL moves it over NVLink while the BPTT still runs.  Bucket 2 = GRU + embedding, after
        # the whole step.  backward() waits for both, so clip_grad_norm_ (train.py:109) sees the global gradient.
        ctx.work = None
        if dp:
            head_off = fg.offset_of(len(GRU_PARAMS))
            with torch.cuda.stream(side):
                side.wait_event(ev)
                w1 = average_gradients_(fg.flat[head_off:], _GROUP, async_op=True)
            w2 = average_gradients_(fg.flat[:head_off], _GROUP, async_op=True)
            ctx.work = (w1, w2)
        ctx.fg = fg if wants_grad else None
        ctx.mark_non_differentiable(logits)
        return loss.reshape(()), logits

    @staticmethod
    def backward(ctx, g_loss, _g_logits):
        fg = ctx.fg
        if fg is None:
            raise RuntimeError("UpDownTrainStep.backward ran twice (retain_graph / a second backward): the fused step "
                               "keeps ONE set of gradients per forward; call get_loss again")
        ctx.fg = None
        fg.pending = False
        for work in (ctx.work or ()):
            if work is not None:
                work.wait()
        # ONE pass scales the whole flat buffer by the incoming gradient (1 after loss.backward()); the parameters then
        # receive FRESH views of it, which autograd adopts as .grad without copying (26 multiplies + 26 copies less)
        fg.flat.mul_(g_loss)
        return (None,) * 6 + tuple(fg.fresh_views())


def updown_loss(model, img, tokens, target, seed=None):
    """loss (autograd-attached) and the detached predictions for a vqa_collection_b200 Wrapper with a
    BaseEncoder / MultiplyAttention / BasePredictor; dropout follows model.training."""
    named = dict(model.named_parameters())
    # a PretrainedWordEmbedding table is frozen and not a parameter (modules.py:166-199): its gradient is dropped by autograd
    named.setdefault("encoder.embedding.weight", model.encoder.embedding.weight)
    params = [named[n] for n in param_names()]
    enc, pred = model.encoder, model.predictor
    training = model.training
    p_att = enc.attention.dropout.p if training else 0.0
    p_cls = pred.classifier.dropout_p if training else 0.0
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())        # from torch's CPU generator: torch.manual_seed applies
    return UpDownTrainStep.apply(img, tokens, target, p_att, p_cls, seed, *params)
