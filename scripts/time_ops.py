"""CUDA-event timings of the individual ops at BASELINE size (B=1024): python scripts/time_ops.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200 import ops
from vqa_collection_b200.engine import prepare_weights

dev = "cuda"
B = int(os.environ.get("B", 1024))
cfg = O.FULL_REGAT
W = O.make_weights(cfg, 1111)
P = prepare_weights(W, torch.bfloat16, dev, True)
g = torch.Generator().manual_seed(0)


def timeit(name, fn, reps=50, flops=None, bytes_=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{name:44s} {us:9.2f} us{extra}", flush=True)
    return us


packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
for T in (1, 2, 4, 14):
    q = torch.randint(0, cfg.ntoken, (B, T), generator=g).to(dev)
    fl = 2.0 * B * 3 * 1024 * (320 * T + 1024 * (T - 1))
    timeit(f"gru persistent T={T}", lambda: ops.gru_last_state(q, P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=packed), flops=fl)
q = torch.randint(0, cfg.ntoken, (B, 14), generator=g).to(dev)
timeit("gru generic T=14", lambda: ops.gru_last_state(q, P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"]), reps=10)

x = torch.rand((B, 36, 2048), generator=g).to(torch.bfloat16).to(dev)
x2 = x.view(B * 36, 2048)
qq = torch.rand((B, 2048), device=dev)
h_lp = torch.rand((B, 1024), device=dev).to(torch.bfloat16)
timeit("qq gemm [B,1024]x[2048,1024]", lambda: ops.linear(h_lp, P["Wqq"], P["sqq"], P["bqq"], relu=True, out_dtype=torch.float32), flops=2.0 * B * 2048 * 1024)
timeit("W_v gemm + logit", lambda: ops.linear(x2, P["Wv"], P["sv"], P["bv"], relu=True, mul=qq, mul_row_div=36, logit_w=P["wlin"]), flops=2.0 * B * 36 * 1024 * 2048)
parts = torch.randn((B * 36, 4), device=dev)
timeit("attention_pool (att+vsum)", lambda: ops.attention_pool(parts, 0.1, x, True, True, False), bytes_=B * 36 * 2048 * 2)
timeit("attention_pool (att only)", lambda: ops.attention_pool(parts, 0.1, x, True, False, False))
vsum = torch.rand((B, 2048), device=dev).to(torch.bfloat16)
timeit("v_net gemm [B,2048]x[1024,2048]", lambda: ops.linear(vsum, P["Wvn"], P["svn"], P["bvn"], relu=True, mul=qq), flops=2.0 * B * 1024 * 2048)
joint = torch.rand((B, 1024), device=dev).to(torch.bfloat16)
timeit("cls0 gemm [B,1024]x[2048,1024]", lambda: ops.linear(joint, P["Wc0"], P["sc0"], P["bc0"], relu=True), flops=2.0 * B * 2048 * 1024)
hid = torch.rand((B, 2048), device=dev).to(torch.bfloat16)
timeit("cls1 gemm [B,2048]x[3129,2048] f32 out", lambda: ops.linear(hid, P["Wc1"], P["sc1"], P["bc1"], relu=True, out_dtype=torch.float32), flops=2.0 * B * 3129 * 2048)
logits = torch.rand((B, 3129), device=dev)
timeit("argmax", lambda: ops.argmax_rows(logits))
timeit("wide ReGAT gemm [B*36,2048]x[6144,2048]", lambda: ops.linear(x2, P["Wg3"]), reps=10, flops=2.0 * B * 36 * 6144 * 2048)
Y = ops.linear(x2, P["Wg3"])
att = torch.softmax(torch.randn((B, 36), device=dev), 1)
labels = ops.relation_labels(torch.from_numpy(O.make_boxes(B, 36, 3)).to(dev), 640, 480)
gat_bytes = B * 36 * 2048 * 2 * 4 + B * 2048 * 2        # Q, x, P, S in; vsum out
timeit("graph_attention tcgen05 (vsum only)", lambda: ops.graph_attention_merged(
    Y, x2, att, labels, P["wvec"], P["gat_c0"], P["label_bias_lp"], P["num_labels"], 36, False, True, False),
    reps=20, bytes_=gat_bytes)
timeit("graph_attention tcgen05 (out+vsum+alpha)", lambda: ops.graph_attention_merged(
    Y, x2, att, labels, P["wvec"], P["gat_c0"], P["label_bias_lp"], P["num_labels"], 36, True, True, True),
    reps=20, bytes_=gat_bytes + B * 36 * 2048 * 2)
boxes = torch.from_numpy(O.make_boxes(B, 36, 3)).to(dev)
timeit("relation_labels B=1024", lambda: ops.relation_labels(boxes, 640, 480), bytes_=B * 1872)
big = torch.from_numpy(O.make_boxes(1 << 18, 36, 4)).to(dev)
timeit("relation_labels B=262144", lambda: ops.relation_labels(big, 640, 480), reps=10, bytes_=(1 << 18) * 1872)
xf = torch.rand((B, 36, 2048), device=dev)
timeit("cast f32->bf16 (B*36*2048)", lambda: ops.cast_to_bf16(xf), bytes_=B * 36 * 2048 * 6)

# ---- optimizer side of the training step (csrc/optim.cu): all 26 parameter tensors of the Up-Down model in one launch
from vqa_collection_b200 import optim as fused
Wf = O.make_weights(O.FULL, 1111)
params = [torch.nn.Parameter(v.clone().float().to(dev)) for v in Wf.values()]
n_par = sum(p.numel() for p in params)
for p in params:
    p.grad = torch.randn_like(p) * 1e-3
opt = fused.Adamax(params, lr=0.002)
opt.step()
timeit(f"adamax, {len(params)} tensors / {n_par / 1e6:.1f} M parameters", lambda: opt.step(), bytes_=n_par * 28)
timeit("clip_grad_norm_ (norm + finalize + scale)", lambda: fused.clip_grad_norm_(params, 0.25), bytes_=n_par * 12)
opt2 = torch.optim.Adamax(params, lr=0.002)
opt2.step()
timeit("torch.optim.Adamax (same tensors)", lambda: opt2.step(), bytes_=n_par * 28)
timeit("torch clip_grad_norm_ (same tensors)", lambda: torch.nn.utils.clip_grad_norm_(params, 0.25), bytes_=n_par * 12)
# caption head: one step's attention logits from the stored projection (B=128 captions, Hd=512)
proj = torch.rand((1024 * 36, 512), device=dev).to(torch.bfloat16)
qd = torch.rand((1024, 512), device=dev)
wd_ = torch.rand((512,), device=dev)
timeit("attention_logits B=1024 Hd=512", lambda: ops.attention_logits(proj, qd, wd_, 36, 0), bytes_=1024 * 36 * 512 * 2)
