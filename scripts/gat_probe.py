"""time the graph-attention kernel alone (B = 1024), optionally with VQA_B200_GAT_DEBUG experiments and a CTA cap"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200 import _lib as L, ops
from vqa_collection_b200.engine import prepare_weights
cfg = O.FULL_REGAT
P = prepare_weights(O.make_weights(cfg, 1111), torch.bfloat16, "cuda", True)
B, K, V = 1024, 36, 2048
g = torch.Generator().manual_seed(3)
x = torch.rand((B * K, V), generator=g).to(torch.bfloat16).cuda()
Y = ops.linear(x, P["Wg3"])
att = torch.rand((B, K), generator=g).cuda()
lab = ops.relation_labels(torch.from_numpy(O.make_boxes(B, K, 5)).cuda(), 640, 480)
lib = L.load()
vsum = torch.empty((B, V), dtype=torch.bfloat16, device="cuda")
def run(ctas):
    a = L.GraphAttentionArgs()
    a.d_Y, a.ldy, a.d_att, a.d_labels = Y.data_ptr(), Y.stride(0), att.data_ptr(), lab.data_ptr()
    a.num_labels, a.B, a.K, a.V, a.dtype = P["num_labels"], B, K, V, L.VQA_BF16
    a.d_vsum = vsum.data_ptr()
    a.layout, a.d_x, a.ldx, a.d_wvec, a.c0 = 1, x.data_ptr(), V, P["wvec"].data_ptr(), float(P["gat_c0"])
    a.d_label_bias_lp = P["label_bias_lp"].data_ptr()
    a.cta_limit = ctas
    L.check(lib.vqa_graph_attention(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for ctas in (0, 16):
    for _ in range(3):
        run(ctas)
    torch.cuda.synchronize()
    e0.record()
    n = 10 if ctas == 0 else 3
    for _ in range(n):
        run(ctas)
    e1.record()
    torch.cuda.synchronize()
    print(f"debug={os.environ.get('VQA_B200_GAT_DEBUG', '0')} ctas={ctas or 148}: {e0.elapsed_time(e1) / n * 1e3:.1f} us", flush=True)
