"""per-parameter gradient error of vqa_updown_train_step vs the oracle: python scripts/train_diag.py [small|full] [B] [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import vqa_oracle as O
from test_gpu_train import run_step, grad_err
cfg = O.SMALL if sys.argv[1] == "small" else O.FULL
B = int(sys.argv[2]); precision = sys.argv[3]
W = O.make_weights(cfg, 1111); batch = O.make_batch(cfg, B, 5001)
ref_loss, ref_logits, ref = O.loss_and_grads(batch, W, cfg)
loss, logits, grads = run_step(cfg, W, batch, precision)
print("loss", float(loss), float(ref_loss), "logits err", float((logits - ref_logits).abs().max() / ref_logits.abs().max()))
for n, g in grads.items():
    print(f"{n:48s} {grad_err(g, ref[n]):.3e}   refmax {float(ref[n].abs().max()):.3e}")
