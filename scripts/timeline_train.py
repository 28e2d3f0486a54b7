"""Kernel timeline of one fused training step (vqa_updown_train_step), aggregated by kernel name."""
import json, os, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import vqa_collection_b200 as pkg
from oracle import vqa_oracle as O
from vqa_collection_b200.modules.wrapper import set_model
from vqa_collection_b200 import training
B = int(os.environ.get("B", 512))
cfg = O.FULL
m = set_model(encoder_type="base", predictor_type="base", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
              embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, rnn_layer=1, ans_dim=cfg.ans_dim, cls_layer=2, c_len=20,
              device="cuda", dropout=0.2, rnn_type="GRU", att_type="new", conv_layer=1, conv_type="corr")
m.load_state_dict(O.make_weights(cfg, 1111), strict=True)
m.train()
b = O.make_batch(cfg, B, 7)
img, q, a = b["img"].to(torch.bfloat16).cuda(), b["q"].cuda(), b["a"].float().cuda()
step = lambda: training.updown_loss(m, img, q, a, seed=1)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace_train.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
n = len(ev) // 3
one = ev[n:2 * n]
span = one[-1]["ts"] + one[-1]["dur"] - one[0]["ts"]
agg = collections.OrderedDict()
busy = 0.0
for e in one:
    k = e["name"][:64]
    c = agg.setdefault(k, [0, 0.0])
    c[0] += 1; c[1] += e["dur"]; busy += e["dur"]
print(f"B={B}: {n} launches, span {span:.0f} us, busy {busy:.0f} us")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{d:8.1f} us  x{c:3d}  {k}")
