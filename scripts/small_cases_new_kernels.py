"""Small-shape run of the kernels added in round 2's second session (fp32tc forward incl. the cooperative GRU-step kernel,
split GEMM tile shapes, token-table bf16 GRU, relation labels for several K) with their errors against the oracle.  Written as
a compute-sanitizer target; the tool is closed on this pool, so scripts/gpu_r2b_san.sh's plain run is what there is."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import vqa_oracle as O
from vqa_collection_b200 import ops
from vqa_collection_b200.engine import VQAEngine, prepare_weights

for cfg, B in ((O.SMALL, 37), (O.SMALL_REGAT, 129)):
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, B, 5)
    with torch.no_grad():
        ref, _ = O.forward(batch, W, cfg)
    for precision in ("fp32tc", "bf16"):
        eng = VQAEngine(W, relation=cfg.relation, precision=precision)
        kw = dict(bbox=batch["bbox"].cuda(), wh=batch["wh"]) if cfg.relation else {}
        out = eng.forward(batch["img"].cuda(), batch["q"].cuda(), **kw)
        err = float((out["logits"].cpu() - ref).abs().max() / ref.abs().max())
        print(precision, "relation" if cfg.relation else "updown", B, "err", err, flush=True)
g = torch.Generator().manual_seed(1)
for M, N, K in ((300, 200, 128), (260, 3129, 256), (1024, 1024, 256)):
    A = torch.rand((M, K), generator=g); Wt = torch.randn((N, K), generator=g) / K ** 0.5
    out = ops.linear_split(ops.split_f32(A.cuda()), ops.split_f32(Wt.cuda()), relu=True, want_argmax=True)[0]
    print("linear_split", M, N, K, float((out.cpu() - torch.relu(A @ Wt.t())).abs().max()), flush=True)
cfg = O.FULL
P = prepare_weights(O.make_weights(cfg, 1111), torch.bfloat16, "cuda", False)
q = torch.randint(0, cfg.ntoken, (130, 3), generator=g).cuda()
h = ops.gru_last_state(q, P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"], packed=(P["wx_packed"], P["wh_packed"], P["bias_packed"]),
                       gi_table=P["gi_table"])
print("gru table", float(h.abs().max()), flush=True)
for B, K in ((5, 36), (3, 1), (4, 45), (2, 64)):
    boxes = O.make_boxes(B, K, 3, 640, 480, grid=True)
    lab = ops.relation_labels(torch.from_numpy(boxes).cuda(), 640, 480).cpu().numpy()
    print("relation", B, K, bool(np.array_equal(lab, O.relation_graph_batch(boxes, 640, 480))), flush=True)
torch.cuda.synchronize()
print("done")
