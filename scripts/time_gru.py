import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200 import ops
from vqa_collection_b200.engine import prepare_weights
dev = "cuda"; B = int(os.environ.get("B", 1024))
cfg = O.FULL
W = O.make_weights(cfg, 1111)
P = prepare_weights(W, torch.bfloat16, dev, False)
g = torch.Generator().manual_seed(0)
packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
for T in ((14,) if os.environ.get('ONLY14') else (1, 2, 14)):
    q = torch.randint(0, int(os.environ.get("TOKMAX", cfg.ntoken)), (B, T), generator=g).to(dev)
    tab = P.get("gi_table") if os.environ.get("TABLE", "1") != "0" else None
    f = lambda: ops.gru_last_state(q, P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=packed, gi_table=tab)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): f()
    e1.record(); torch.cuda.synchronize()
    print(f"table={tab is not None} debug={os.environ.get('VQA_B200_GRU_DEBUG','0')} B={B} T={T}: {e0.elapsed_time(e1)/30*1e3:.1f} us", flush=True)
