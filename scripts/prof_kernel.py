"""Launch one hot kernel a few times (for `ncu --set full`): python scripts/prof_kernel.py wv|pool|relation"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vqa_collection_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "wv"
B, K, V, H = 1024, 36, 2048, 1024
g = torch.Generator().manual_seed(0)
dev = "cuda"
if which == "wv":
    x = torch.rand((B * K, V), generator=g).to(torch.bfloat16).to(dev)
    W = (torch.randn((H, V), generator=g) / V ** 0.5).to(torch.bfloat16).to(dev)
    s, b, wl = torch.ones(H, device=dev), torch.zeros(H, device=dev), torch.randn(H, device=dev)
    qq = torch.rand((B, 2 * H), device=dev)
    for _ in range(5):
        ops.linear(x, W, s, b, relu=True, mul=qq, mul_row_div=K, logit_w=wl)
elif which == "pool":
    x = torch.rand((B, K, V), generator=g).to(torch.bfloat16).to(dev)
    parts = torch.randn((B * K, 4), device=dev)
    for _ in range(5):
        ops.attention_pool(parts, 0.1, x, True, True, False)
elif which == "relation":
    from oracle import vqa_oracle as O
    boxes = torch.from_numpy(O.make_boxes(1 << 18, 36, 1)).to(dev)
    for _ in range(5):
        ops.relation_labels(boxes, 640, 480)
torch.cuda.synchronize()
print("done", which)
