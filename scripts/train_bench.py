"""Training-step throughput (BASELINE config 4: Up-Down, global batch 512, NCCL gradient all-reduce).

    python scripts/train_bench.py [--batch 512] [--steps 30]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/train_bench.py

Strong scaling: the global batch is fixed, each rank trains on batch/N questions; one step = train.py:103-111
(get_loss → backward → clip_grad_norm_ → Adamax.step → zero_grad).  Prints one JSON line on rank 0.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--torch-optim", action="store_true", help="torch.optim.Adamax + nn.utils.clip_grad_norm_ (the reference's "
                "own calls) instead of the library's one-launch drop-ins (vqa_collection_b200.optim)")
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
import vqa_collection_b200 as pkg
from oracle import vqa_oracle as O
from vqa_collection_b200.modules.wrapper import set_model
from vqa_collection_b200.parallel import shard_batch
pkg.set_precision(args.precision)
cfg = O.FULL
W = O.make_weights(cfg, 1111)
m = set_model(encoder_type="base", predictor_type="base", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
              embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, rnn_layer=1, ans_dim=cfg.ans_dim, cls_layer=2, c_len=20,
              device=str(dev), dropout=0.2, rnn_type="GRU", att_type="new", conv_layer=1, conv_type="corr")
m.load_state_dict(W, strict=True)
from vqa_collection_b200 import optim as fused
Adamax = torch.optim.Adamax if args.torch_optim else fused.Adamax
clip_grad_norm_ = torch.nn.utils.clip_grad_norm_ if args.torch_optim else fused.clip_grad_norm_
opt = Adamax([{'params': m.encoder.parameters()}, {'params': m.predictor.parameters(), 'lr': 0.002}], lr=0.002)
full = O.make_batch(cfg, args.batch, 7)
b = shard_batch(full, world, rank)
batch = {"img": b["img"].to(torch.bfloat16 if args.precision == "bf16" else torch.float32).to(dev), "q": b["q"].to(dev), "a": b["a"].float().to(dev)}
m.train()


def step():
    loss, writes = m.get_loss(batch)
    loss.backward()
    clip_grad_norm_(m.parameters(), 0.25)
    opt.step()
    opt.zero_grad()
    return loss


for _ in range(args.warmup):
    step()
# do the parameters' .grad alias the flat gradient buffer (handed over without copies)?
_l, _w = m.get_loss(batch)
_l.backward()
_flat = training_flat = None
from vqa_collection_b200 import training as _tr
_fg = _tr._FLAT.get(dev)
grads_alias_flat = bool(_fg is not None and all(p.grad is not None and p.grad.untyped_storage().data_ptr() ==
                                                _fg.flat.untyped_storage().data_ptr() for p in m.parameters()))
opt.zero_grad()
torch.cuda.synchronize()
if dist: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
# the fused fwd+bwd call alone (no .item() syncs, no optimizer)
from vqa_collection_b200 import training
torch.cuda.synchronize()
e0.record()
for _ in range(args.steps):
    l, _ = training.updown_loss(m, batch["img"], batch["q"], batch["a"], seed=1)
e1.record()
torch.cuda.synchronize()
ms_core = e0.elapsed_time(e1) / args.steps
if dist:
    t = torch.tensor([ms, ms_core], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_core = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({"metric": "Up-Down VQA training questions/sec (global batch %d)" % args.batch, "value": args.batch / (ms / 1e3),
                      "unit": "questions/s", "n_gpus": world, "ms_per_step": ms, "ms_fwd_bwd_allreduce": ms_core, "scaling": "strong",
                      "dtype": args.precision, "loss": float(loss), "grads_alias_flat": grads_alias_flat, "config": {"workload": "Up-Down VQA training step batch 512 with NCCL gradient allreduce",
                      "global_batch": args.batch, "per_gpu_batch": batch["img"].shape[0], "optimizer": ("torch.optim.Adamax + nn.utils.clip_grad_norm_(0.25)" if args.torch_optim else
                                    "vqa_collection_b200.optim.Adamax + clip_grad_norm_(0.25): one launch per operation") + " (train.py:108-111)"}}), flush=True)
if dist: dist.destroy_process_group()
