"""CUDA-event timing of BASELINE config 5 (q-cap joint forward, B=512, 20-token captions) through the module API."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200.modules.wrapper import set_model
from vqa_collection_b200 import _lib

B = int(os.environ.get("B", 512))
cfg = O.FULL_QCAP
W = O.make_weights(cfg, 1111)
m = set_model(encoder_type="base", predictor_type="q-cap", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
              embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, decoder_hidden_dim=0, rnn_layer=1, ans_dim=cfg.ans_dim,
              cls_layer=2, c_len=cfg.c_len, device="cuda", dropout=0.2, neg_slope=cfg.neg_slope, rnn_type="GRU",
              att_type="new")
m.load_state_dict(W, strict=True)
m.eval()
batch = O.make_batch(cfg, B, 7)
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
dev["img"] = dev["img"].to(torch.bfloat16)


def step():
    with torch.no_grad():
        return m.forward_vqa(dict(dev))


for _ in range(3):
    step()
torch.cuda.synchronize()
lib = _lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"q-cap forward_vqa B={B}: {ms*1e3:.1f} us/step = {B/ms*1e3:.0f} questions/s", flush=True)
import json
print(json.dumps({"metric": "Q-Relevant caption-embedding + Up-Down joint forward questions/sec (BASELINE config 5)",
                  "value": B / ms * 1e3, "unit": "questions/s", "n_gpus": 1, "ms_per_step": ms, "dtype": "bf16",
                  "data": "synthetic", "config": {"workload": "q-cap predictor, 20-token captions, module-level API "
                  "(Wrapper.forward_vqa), bf16 features resident", "batch": B}}), flush=True)
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
