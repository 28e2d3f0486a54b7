import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from oracle import vqa_oracle as O
from vqa_collection_b200 import ops
from vqa_collection_b200.engine import prepare_weights
B = int(os.environ.get("B", 1024)); T = int(os.environ.get("T", 14))
cfg = O.FULL
P = prepare_weights(O.make_weights(cfg, 1111), torch.bfloat16, "cuda", False)
q = torch.randint(0, cfg.ntoken, (B, T)).cuda()
packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
f = lambda: ops.gru_last_state(q, P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=packed)
for _ in range(3): f()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): f()
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "gru" in e.key:
        print(e.key[:60], f"{e.device_time_total / e.count:.1f} us x{e.count}")
