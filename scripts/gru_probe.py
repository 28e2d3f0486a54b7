"""time the persistent GRU alone (B = 1024, T = 14), optionally with VQA_B200_GRU_DEBUG experiments / VQA_B200_GRU_CFG"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200 import ops
from vqa_collection_b200.engine import prepare_weights
cfg = O.FULL
P = prepare_weights(O.make_weights(cfg, 1111), torch.bfloat16, "cuda", False)
q = O.make_batch(cfg, 1024, 123)["q"].cuda()
packed = (P["wx_packed"], P["wh_packed"], P["bias_packed"])
run = lambda: ops.gru_last_state(q, P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=packed)
for _ in range(5):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
print(f"cfg={os.environ.get('VQA_B200_GRU_CFG', 'auto')} debug={os.environ.get('VQA_B200_GRU_DEBUG', '0')}: "
      f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us (gather + memset + GRU)", flush=True)
