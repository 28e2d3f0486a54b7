"""How does tcgen05 (kind::f16, fp32 accumulator) round?  bf16-exact inputs, K = 2048, against float64.
Positive operands make a round-toward-zero accumulator show up as a systematic negative error of the order of
(K / 16 roundings) x 2^-24; round-to-nearest leaves a zero-mean error of ~ sqrt(K / 16) x 2^-25."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from vqa_collection_b200 import ops

g = torch.Generator().manual_seed(1)
for K in (256, 2048, 8192):
    for signed in (False, True):
        A = torch.rand((256, K), generator=g)
        W = torch.rand((256, K), generator=g)
        if signed:
            A, W = A - 0.5, W - 0.5
        A, W = A.to(torch.bfloat16), W.to(torch.bfloat16)
        out = ops.linear(A.cuda(), W.cuda(), None, None, relu=False, out_dtype=torch.float32).double().cpu()
        ref = A.double() @ W.double().t()
        f32 = (A.float() @ W.float().t()).double()            # CPU fp32 accumulation, for scale
        scale = ref.abs().max()
        err = (out - ref) / scale
        e32 = (f32 - ref) / scale
        print(f"K={K:5d} signed={signed}: tcgen05 mean {err.mean():+.3e} max|.| {err.abs().max():.3e}   "
              f"cpu-fp32 mean {e32.mean():+.3e} max|.| {e32.abs().max():.3e}", flush=True)
