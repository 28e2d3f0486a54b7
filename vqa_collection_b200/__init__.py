"""vqa_collection_b200 — B200-native (sm_100a) VQA forward hot path of Jayie/vqa-collection.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/vqa_b200.h)
  _lib.py    ctypes binding (no CPU fallback: raises if the library is not built)
  ops.py     tensor-level wrappers over the C ABI
  engine.py  prepared weights + one C call per forward (the measured path)
  modules/   the reference's module API (attention / gcn / predictor / encoder / wrapper)
  util/      relation_graph drop-in (util/relation.py)
  parallel.py  data-parallel batch sharding (one process per GPU, no forward collective)
"""
_PRECISION = "bf16"


def set_precision(p: str):
    """'bf16' (tcgen05 tensor cores, default), 'fp32' (FFMA, 1e-5 parity mode) or 'fp32tc' (fp32-class arithmetic on the
    tensor cores for the whole-forward engine — fp16 plane pairs, three tcgen05.mma per k-step; module-level calls outside
    the engine run the 'fp32' kernels)."""
    global _PRECISION
    if p not in ("bf16", "fp32", "fp32tc"):
        raise ValueError("precision must be 'bf16', 'fp32' or 'fp32tc'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def compute_dtype():
    import torch
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32
