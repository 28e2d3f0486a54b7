import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import vqa_oracle as O
import torch.nn.functional as F
torch.manual_seed(0)
def split(x):
    hi = x.to(torch.bfloat16).float(); lo = (x - hi).to(torch.bfloat16).float(); return hi, lo
orig_linear = F.linear
def lin3(x, w, b=None):
    xh, xl = split(x); wh, wl = split(w)
    y = orig_linear(xh, wh) + orig_linear(xl, wh) + orig_linear(xh, wl)
    return y if b is None else y + b
for cfg, B in ((O.FULL, 64),):
    W = O.make_weights(cfg, 1111); batch = O.make_batch(cfg, B, 4321)
    with torch.no_grad():
        ref, enc = O.forward(batch, W, cfg)
        Wd = {k: v.double() for k, v in W.items()}
        bd = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
        truth, _ = O.forward(bd, Wd, cfg)
        F.linear = lin3
        O.F.linear = lin3
        got, enc3 = O.forward(batch, W, cfg)
        F.linear = orig_linear; O.F.linear = orig_linear
    s = ref.abs().max()
    print("fp32 vs f64   :", float((ref.double() - truth).abs().max() / truth.abs().max()))
    print("bf16x3 vs f64 :", float((got.double() - truth).abs().max() / truth.abs().max()))
    print("bf16x3 vs fp32:", float((got - ref).abs().max() / s), "labels equal", int((got.argmax(1) == ref.argmax(1)).sum()), "/", B)
    print("att err", float((enc3["v_att"] - enc["v_att"]).abs().max() / enc["v_att"].abs().max()))
