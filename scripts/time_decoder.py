"""CUDA-event timing of the caption head: Wrapper.forward with decoder_type='base' (the main.py
default model: Up-Down encoder + VQA predictor + BaseDecoder teacher-forced over 20-token captions), module API,
next to the CPU oracle port on a bounded sample.  B=128 is the reference's default batch size (main.py:60)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200.modules.wrapper import set_model

B = int(os.environ.get("B", 128))
cfg = O.FULL_DECODER
W = O.make_weights(cfg, 1111)
m = set_model(encoder_type="base", predictor_type="base", decoder_type="base", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
              embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, decoder_hidden_dim=cfg.decoder_hidden_dim, rnn_layer=1,
              ans_dim=cfg.ans_dim, cls_layer=2, c_len=cfg.c_len, device="cuda", dropout=0.2, rnn_type="GRU", att_type="new")
m.load_state_dict(W, strict=True)
m.eval()
batch = O.make_decoder_batch(cfg, B, 7)
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
dev["img"] = dev["img"].to(torch.bfloat16)
dev["cap_len"] = batch["cap_len"]                      # the lengths are read on the host (sort + batch_t schedule)
words = int((batch["cap_len"] - 1).sum())


def step():
    with torch.no_grad():
        return m(dict(dev))


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
# CPU oracle port on a bounded sample of the same workload
nb = min(B, 16)
small = {k: (v[:nb] if torch.is_tensor(v) else v) for k, v in batch.items()}
with torch.no_grad():
    t0 = time.perf_counter()
    _, enc = O.forward(small, W, cfg)
    O.base_decoder_forward(enc, W, cfg)
    cpu_s = time.perf_counter() - t0
print(json.dumps({"metric": "VQA + caption-head joint forward questions/sec (decoder_type='base', teacher forced)",
                  "value": B / ms * 1e3, "unit": "questions/s", "n_gpus": 1, "ms_per_step": ms, "dtype": "bf16",
                  "data": "synthetic", "caption_words_per_step": words,
                  "config": {"workload": "Up-Down encoder + BasePredictor + BaseDecoder (Hd=512, ntoken=20000, ragged "
                             "caption lengths 2..20), module-level API (Wrapper.forward), bf16 features resident", "batch": B},
                  "cpu_baseline": {"value": nb / cpu_s, "unit": "questions/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"{nb} questions, one pass, fp32 torch-CPU oracle port"}}), flush=True)
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
