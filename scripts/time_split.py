"""fp32-class split GEMM (VQA_F16X2, 3 tcgen05.mma per k-step) against the bf16 GEMM on the shapes of the path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqa_collection_b200 import ops

def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

g = torch.Generator().manual_seed(0)
for name, M, N, K, logit in (("W_v", 36864, 1024, 2048, True), ("wide", 36864, 6144, 2048, False), ("Wqq", 1024, 2048, 1024, False),
                             ("v_net", 1024, 1024, 2048, False), ("cls0", 1024, 2048, 1024, False), ("cls1", 1024, 3129, 2048, False),
                             ("gru_h", 1024, 3072, 1024, False)):
    A = torch.rand((M, K), generator=g).cuda(); W = (torch.randn((N, K), generator=g) / K ** 0.5).cuda()
    lw = torch.randn((N,), generator=g).cuda() if logit else None
    A2, W2 = ops.split_f32(A), ops.split_f32(W)
    Ab, Wb = A.bfloat16(), W.bfloat16()
    t_split = timeit(lambda: ops.linear_split(A2, W2, relu=True, logit_w=lw))
    t_bf16 = timeit(lambda: ops.linear(Ab, Wb, relu=True, logit_w=lw, out_dtype=torch.float32))
    fl = 2.0 * M * N * K
    print(f"{name:6s} {M}x{N}x{K}: split {t_split:8.1f} us ({3 * fl / t_split / 1e6:7.1f} TF/s of MMA work)   bf16 {t_bf16:8.1f} us ({fl / t_bf16 / 1e6:7.1f} TF/s)   ratio {t_split / t_bf16:.2f}", flush=True)
    del A, W, A2, W2, Ab, Wb
