"""e2e host path tuning: chunk size / pack threads.  python scripts/e2e_probe.py"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vqa_oracle as O
from vqa_collection_b200.engine import VQAEngine
cfg = O.FULL
eng = VQAEngine(O.make_weights(cfg, 1111), relation=False, precision="bf16")
B = 1024
img = torch.rand((B, 36, 2048)).pin_memory(); tok = torch.randint(0, cfg.ntoken, (B, 14)).pin_memory()
lib = eng.lib
# raw pack rate into pinned memory
lib.vqa_packpool_create.restype = ctypes.c_void_p
lib.vqa_packpool_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
dst = torch.empty((B, 36, 2048), dtype=torch.bfloat16).pin_memory()
for th in (8, 10, 12, 14, 16):
    pool = ctypes.c_void_p(lib.vqa_packpool_create(th))
    for n_chunks in (1, 16):
        per = img.numel() // n_chunks
        for _ in range(2):
            for c in range(n_chunks): lib.vqa_packpool_run(pool, img.data_ptr() + c * per * 4, dst.data_ptr() + c * per * 2, per)
        t = time.perf_counter()
        for _ in range(5):
            for c in range(n_chunks): lib.vqa_packpool_run(pool, img.data_ptr() + c * per * 4, dst.data_ptr() + c * per * 2, per)
        print(f"pack threads={th} chunks={n_chunks}: {(time.perf_counter()-t)/5*1e3:.2f} ms", flush=True)
def run(**kw):
    for _ in range(3): eng.forward_host(img, tok, **kw)
    torch.cuda.synchronize(); t = time.perf_counter()
    pend = None
    for _ in range(12):
        nxt = eng.forward_host_async(img, tok, **kw)
        if pend: pend.result()
        pend = nxt
    pend.result()
    return (time.perf_counter() - t) / 12 * 1e3
for chunk in (32, 64):
    for period in (0, 2, 3, 4, 5, 6, 8):
        print(f"pipelined forward_host chunk={chunk} raw_period={period}: {run(chunk=chunk, raw_chunk_period=period):.2f} ms", flush=True)
print(f"f32 over PCIe: {run(pack_on_host=False):.2f} ms")
