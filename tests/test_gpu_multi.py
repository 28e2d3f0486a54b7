"""Multi-GPU (one process per GPU, NCCL over NVLink) checks of the data-parallel paths.  Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import vqa_collection_b200 as pkg
from oracle import vqa_oracle as O
from vqa_collection_b200.engine import VQAEngine
from vqa_collection_b200.parallel import shard_batch, gather_rows
from vqa_collection_b200 import training

# ---- forward: sharded == unsharded, no data-path collective (SURVEY.md 8e)
cfg = O.FULL
W = O.make_weights(cfg, 1111)
batch = O.make_batch(cfg, 256, 99)
eng = VQAEngine(W, relation=False, precision="bf16", device=dev)
full = eng.forward(batch["img"].to(dev), batch["q"].to(dev))
loc = shard_batch(batch, world, rank)
part = eng.forward(loc["img"].to(dev), loc["q"].to(dev))
labels = gather_rows(part["label"], 256)
assert torch.equal(labels, full["label"]), "sharded labels differ"
logits = gather_rows(part["logits"], 256)
assert torch.equal(logits, full["logits"]), "sharded logits differ"

# ---- training: all-reduced shard gradients == single-process gradients of the concatenated batch
names = training.param_names()
tb = O.make_batch(cfg, 64 * world, 5)


def grads_of(b, dp):
    params = [W[n].to(dev).requires_grad_(True) for n in names]
    training.set_process_group(None)
    training._dp_disabled = not dp
    loss, _ = training.UpDownTrainStep.apply(b["img"].float().to(dev), b["q"].to(dev), b["a"].float().to(dev), 0.0, 0.0, 1, *params)
    loss.backward()
    return [p.grad.clone() for p in params]


single = grads_of(tb, False)                       # whole batch on every rank, no exchange
dp = grads_of(shard_batch(tb, world, rank), True)  # shard per rank + NCCL average
gmax = max(float(g.abs().max()) for g in single)
for n, a, b in zip(names, dp, single):
    err = float((a - b).abs().max())
    assert err <= 2e-4 * max(float(b.abs().max()), 1e-3 * gmax), (n, err, float(b.abs().max()))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_gpu_sharded_forward_and_gradient_allreduce(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == n


def test_model_on_a_device_other_than_the_current_one():
    """the reference's --device / --decoder_device (main.py:88, wrapper.py:148-150): a model built on cuda:1 while the
    current device is 0 must launch on cuda:1's stream with cuda:1's kernel attributes — same answers as on cuda:0"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, ROOT)
    from oracle import vqa_oracle as O
    from vqa_collection_b200.engine import VQAEngine
    from vqa_collection_b200 import ops
    torch.cuda.set_device(0)
    for relation, B in ((False, 600), (True, 130)):
        cfg = O.FULL_REGAT if relation else O.FULL
        W = O.make_weights(cfg, 1111)
        batch = O.make_batch(cfg, B, 77)
        outs = []
        for dev in ("cuda:0", "cuda:1"):
            eng = VQAEngine(W, relation=relation, precision="bf16", device=dev)
            kw = dict(labels=batch["graph"].to(torch.uint8).to(dev)) if relation else {}
            out = eng.forward(batch["img"].to(dev), batch["q"].to(dev), **kw)
            assert out["logits"].device == torch.device(dev)
            outs.append((out["logits"].cpu(), out["label"].cpu()))
        assert torch.cuda.current_device() == 0
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # tensor-level wrappers follow their tensors' device too
    boxes = torch.from_numpy(O.make_boxes(8, 36, 1)).to("cuda:1")
    lab1 = ops.relation_labels(boxes, 640, 480)
    assert lab1.device == torch.device("cuda:1")
    assert torch.equal(lab1.cpu(), ops.relation_labels(boxes.to("cuda:0"), 640, 480).cpu())
