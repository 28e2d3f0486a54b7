"""fp32-class (fp32tc) question encoder alone: gate kernel + the fused GEMM / gate-update step kernel, 13 steps in one launch"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200 import ops
from vqa_collection_b200.engine import VQAEngine
B = int(os.environ.get("B", 1024))
cfg = O.FULL
eng = VQAEngine(O.make_weights(cfg, 1111), relation=False, precision="fp32tc")
P = eng.P
g = torch.Generator().manual_seed(0)
q = torch.randint(0, cfg.ntoken, (B, 14), generator=g).cuda()
f = lambda: ops.gru_last_state_split(q, P["gi_table"], P["w_hh"], P["b_hh"], P["wh_packed"])
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
print(f"debug={os.environ.get('VQA_B200_GRUS_DEBUG','0')} persist={os.environ.get('VQA_B200_GRU_SPLIT_PERSIST','1')} B={B}: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
