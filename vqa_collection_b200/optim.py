"""Optimizer side of the training step on the C-ABI kernels (BASELINE config 4; csrc/optim.cu).

Drop-ins for the two calls the reference's loop makes after ``loss.backward()`` (train.py:109-110):

    nn.utils.clip_grad_norm_(model.parameters(), max_norm)   →  vqa_collection_b200.optim.clip_grad_norm_
    torch.optim.Adamax(params, lr=lr)  (train.py:54-57)      →  vqa_collection_b200.optim.Adamax

Both walk ALL parameter tensors in one launch per operation (pointer table in the kernel parameters) instead of
a few dozen small torch launches, and neither synchronises with the host.  ``Adamax`` keeps torch's
constructor signature, param-group handling and state layout (``step`` / ``exp_avg`` / ``exp_inf``), so
``state_dict()`` / ``load_state_dict()`` round-trip with ``torch.optim.Adamax`` and ``StepLR`` (train.py:58) works.
"""
import ctypes as C

import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_grad(p, g):
    if not (p.is_cuda and g.is_cuda):
        raise RuntimeError("vqa_collection_b200.optim: parameters and gradients must be CUDA tensors (no CPU fallback)")
    if p.dtype != torch.float32 or g.dtype != torch.float32:
        raise TypeError("vqa_collection_b200.optim: float32 parameters / gradients only")
    if g.is_sparse:
        raise RuntimeError("vqa_collection_b200.optim: sparse gradients are not supported")


def _table(entries):
    """entries: list of (p_ptr, g_ptr, m_ptr, u_ptr, n, lr) → ctypes array"""
    arr = (L.OptimTensor * len(entries))()
    for i, (p, g, m, u, n, lr) in enumerate(entries):
        a = arr[i]
        a.d_p, a.d_g, a.d_m, a.d_u, a.n, a.lr = p, g, m, u, n, lr
    return arr


_CLIP_WS = {}


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """nn.utils.clip_grad_norm_ (train.py:109) for the 2-norm: scales every ``.grad`` in place by
    min(1, max_norm / (total_norm + 1e-6)) and returns total_norm as a 0-d device tensor — three launches
    (Σg² partials over all tensors, finalize, scale), no host sync."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("vqa_collection_b200.optim.clip_grad_norm_: only norm_type=2 (train.py:109 default)")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    lib = L.load()
    dev = grads[0].device
    entries = []
    for g in grads:
        _check_grad(g, g)
        if not g.is_contiguous():
            raise RuntimeError("vqa_collection_b200.optim: gradients must be contiguous")
        entries.append((None, g.data_ptr(), None, None, g.numel(), 0.0))
    ws = _CLIP_WS.get(dev)
    if ws is None:
        ws = _CLIP_WS[dev] = torch.empty((lib.vqa_grad_clip_workspace_bytes() // 4 + 2,), dtype=torch.float32, device=dev)
    out = torch.empty((2,), dtype=torch.float32, device=dev)          # [total_norm, scale]
    with torch.cuda.device(dev):                                      # the library launches on the current device
        L.check(lib.vqa_grad_clip(_table(entries), len(entries), float(max_norm), 1, ws.data_ptr(), out.data_ptr(),
                                  out.data_ptr() + 4, _stream()))
    torch._C._increment_version(grads)                                # scaled in place behind autograd's back
    return out[0]


class Adamax(torch.optim.Optimizer):
    """torch.optim.Adamax (train.py:57) with the update of every parameter tensor in one kernel launch."""

    def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None):
        """``grad_scale`` (extension): 0-d device tensor multiplied into every gradient inside the update (e.g. the
        clip coefficient), instead of a separate pass over the gradients."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        live = [(gi, p) for gi, group in enumerate(self.param_groups) for p in group["params"] if p.grad is not None]
        if not live:
            return loss
        gs = None
        if grad_scale is not None:
            if not (grad_scale.is_cuda and grad_scale.dtype == torch.float32 and grad_scale.numel() == 1):
                raise TypeError("Adamax.step: grad_scale must be a 1-element float32 CUDA tensor")
            gs = grad_scale.data_ptr()
        # one pointer table per distinct (betas, eps, weight_decay, step): a single launch for the reference's param
        # groups, which differ in lr only (lr is per tensor in the table)
        buckets = {}
        for gi, p in live:
            group, g = self.param_groups[gi], p.grad
            st = self.state[p]
            if len(st) == 0:
                _check_grad(p, g)
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_inf"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if g.dtype != torch.float32 or not g.is_cuda or not (g.is_contiguous() and p.is_contiguous()):
                _check_grad(p, g)
                raise RuntimeError("vqa_collection_b200.optim.Adamax: parameters and gradients must be contiguous")
            st["step"] += 1
            b1, b2 = group["betas"]
            hk = (b1, b2, group["eps"], group["weight_decay"], int(st["step"]))
            buckets.setdefault(hk, []).append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                               st["exp_inf"].data_ptr(), p.numel(), group["lr"]))
        for (b1, b2, eps, wd, step), entries in buckets.items():
            arr = self._arrays.get(len(entries)) if hasattr(self, "_arrays") else None
            if arr is None:
                if not hasattr(self, "_arrays"):
                    self._arrays = {}
                arr = self._arrays[len(entries)] = (L.OptimTensor * len(entries))()
            for i, e in enumerate(entries):
                a = arr[i]
                a.d_p, a.d_g, a.d_m, a.d_u, a.n, a.lr = e
            with torch.cuda.device(live[0][1].device):                # the library launches on the current device
                L.check(lib.vqa_adamax_step(arr, len(entries), b1, b2, eps, wd, step, gs, _stream()))
        # The kernel wrote the parameters through raw pointers: tell autograd.  The weight caches of the forward path
        # (Wrapper.engine(), PreparedCache) key on Tensor._version — without the bump, evaluate() after a training epoch
        # would keep serving the bf16 copies and weight-norm scales of the OLD weights.
        torch._C._increment_version([p for _, p in live])
        return loss
