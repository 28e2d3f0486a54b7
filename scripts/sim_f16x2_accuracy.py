"""CPU simulation of the fp32-class tensor-core mode: every GEMM operand split into fp16 hi + fp16 (residual x 2^11)
(Ootomo & Yokota), three products hi·hi + 2^-11 (hi·lo' + lo'·hi), fp32 accumulation; the GRU's recurrent GEMM included
(oracle's nn.GRU replaced by an explicit loop over F.linear).  Compare with scripts/sim_bf16x3_accuracy.py."""
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import vqa_oracle as O
import torch.nn.functional as F
torch.manual_seed(0)
RZ = len(sys.argv) > 1 and sys.argv[1] == "rz"
def split(x):
    hi = x.to(torch.float16).float(); lo = ((x - hi) * 2048.0).to(torch.float16).float(); return hi, lo
orig_linear = F.linear
def lin3(x, w, b=None):
    xh, xl = split(x); wh, wl = split(w)
    y = orig_linear(xh, wh) + (orig_linear(xl, wh) + orig_linear(xh, wl)) * (1.0 / 2048.0)
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
    print("f16x2 vs f64  :", float((got.double() - truth).abs().max() / truth.abs().max()))
    print("f16x2 vs fp32 :", float((got - ref).abs().max() / s), "labels equal", int((got.argmax(1) == ref.argmax(1)).sum()), "/", B)
    print("att err", float((enc3["v_att"] - enc["v_att"]).abs().max() / enc["v_att"].abs().max()))
