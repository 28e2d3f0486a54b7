"""Launch one hot kernel a few times (for `ncu --set full`):
python scripts/prof_kernel.py wv|pool|relation|gat|wide|gru|cls1|wv_split|gru_split|pool_split"""
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
elif which in ("gat", "wide", "gru", "cls1"):
    from oracle import vqa_oracle as O
    from vqa_collection_b200.engine import prepare_weights
    cfg = O.FULL_REGAT
    P = prepare_weights(O.make_weights(cfg, 1111), torch.bfloat16, dev, True)
    x2 = torch.rand((B * K, V), generator=g).to(torch.bfloat16).to(dev)
    if which == "wide":
        for _ in range(4):
            ops.linear(x2, P["Wg3"])
    elif which == "gat":
        Y = ops.linear(x2, P["Wg3"])
        att = torch.softmax(torch.randn((B, K), device=dev), 1)
        labels = ops.relation_labels(torch.from_numpy(O.make_boxes(B, K, 3)).to(dev), 640, 480)
        for _ in range(4):
            ops.graph_attention_merged(Y, x2, att, labels, P["wvec"], P["gat_c0"], P["label_bias_lp"], P["num_labels"], K,
                                       False, True, False)
    elif which == "gru":
        q = torch.randint(0, cfg.ntoken, (B, 14), generator=g).to(dev)
        packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
        for _ in range(4):
            ops.gru_last_state(q, P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=packed)
    else:
        hid = torch.rand((B, 2 * H), generator=g).to(torch.bfloat16).to(dev)
        for _ in range(4):
            ops.linear(hid, P["Wc1"], P["sc1"], P["bc1"], relu=True, out_dtype=torch.float32)
elif which == "wv_split":
    # fp32-class mode: the W_v projection on fp16 plane pairs (three tcgen05.mma per k-step, CTA pairs)
    x2 = ops.split_f32(torch.rand((B * K, V), generator=g).to(dev))
    W2 = ops.split_f32((torch.randn((H, V), generator=g) / V ** 0.5).to(dev))
    s, b, wl = torch.ones(H, device=dev), torch.zeros(H, device=dev), torch.randn(H, device=dev)
    qq = torch.rand((B, 2 * H), device=dev)
    for _ in range(4):
        ops.linear_split(x2, W2, s, b, relu=True, mul=qq, mul_row_div=K, logit_w=wl)
elif which == "pool_split":
    x2 = ops.split_f32(torch.rand((B, K, V), generator=g).to(dev))
    parts = torch.randn((B * K, 4), device=dev)
    for _ in range(4):
        ops.attention_pool_split(parts, 0.1, x2)
elif which == "gru_split":
    # fp32-class question encoder through the engine's C call (gate kernel + the fused GEMM / gate-update step kernel)
    from oracle import vqa_oracle as O
    from vqa_collection_b200.engine import VQAEngine
    cfg = O.FULL
    eng = VQAEngine(O.make_weights(cfg, 1111), relation=False, precision="fp32tc")
    P = eng.P
    q = torch.randint(0, cfg.ntoken, (B, 14), generator=g).to(dev)
    for _ in range(3):
        ops.gru_last_state_split(q, P["gi_table"], P["w_hh"], P["b_hh"], P["wh_packed"])
torch.cuda.synchronize()
print("done", which)
