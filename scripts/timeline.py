"""Kernel timeline of one forward step (CUPTI via torch.profiler): per-kernel start/duration and the idle gaps
between consecutive kernels on the stream.  python scripts/timeline.py [updown|regat]"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from oracle import vqa_oracle as O
from vqa_collection_b200.engine import VQAEngine
from vqa_collection_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "updown"
relation = wl == "regat"
cfg = O.FULL_REGAT if relation else O.FULL
B = int(os.environ.get("B", 1024))
eng = VQAEngine(O.make_weights(cfg, 1111), relation=relation, precision=os.environ.get("PRECISION", "bf16"), device=torch.device("cuda"))
g = torch.Generator().manual_seed(3)
img = eng.resident(torch.rand((B, 36, 2048), generator=g).cuda())
tok = torch.randint(0, cfg.ntoken, (B, 14), generator=g).cuda()
lab = ops.relation_labels(torch.from_numpy(O.make_boxes(B, 36, 5)).cuda(), 640, 480) if relation else None
step = lambda: eng.forward(img, tok, labels=lab)
for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
n = len(ev) // 4
one = ev[2 * n:3 * n]                       # third step
t0 = one[0]["ts"]
prev_end = None
busy = 0.0
for e in one:
    gap = (e["ts"] - prev_end) if prev_end is not None else 0.0
    print(f"{e['ts'] - t0:9.1f} us  dur {e['dur']:8.1f}  gap {gap:6.1f}  {e['name'][:70]}")
    prev_end = e["ts"] + e["dur"]
    busy += e["dur"]
span = one[-1]["ts"] + one[-1]["dur"] - t0
nxt = ev[3 * n]["ts"] - prev_end
print(f"step span {span:.1f} us, busy {busy:.1f} us, idle inside {span - busy:.1f} us, gap to next step {nxt:.1f} us")
