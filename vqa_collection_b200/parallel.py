"""Data-parallel sharding of the forward path: one process per GPU, weights replicated,
the batch split into contiguous row ranges, NO collective on the data path (SURVEY.md
§8e).  Results are gathered only when the caller wants them on every rank (the
reference's evaluate() just accumulates scalars, train.py:186-189)."""
import torch


def shard_bounds(n_rows: int, world: int, rank: int):
    """rows [lo, hi) of rank `rank`; the first n_rows % world ranks get one extra row"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: dict, world: int, rank: int) -> dict:
    """slice every per-sample entry of a batch dict (dataset.py:96-104 wire format)"""
    n = None
    for k in ("img", "q", "a"):
        if k in batch:
            n = batch[k].shape[0]
            break
    if n is None:
        raise KeyError("batch has none of img / q / a")
    lo, hi = shard_bounds(n, world, rank)
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == n:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def gather_rows(local: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """all-gather the per-rank row blocks of a result back into [n_rows, ...] on every rank
    (ragged shards are padded to the largest block)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_rows, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)


def average_gradients_(flat: torch.Tensor, group=None, async_op: bool = False):
    """In-place mean over ranks of a flat gradient buffer — the one exchange step of the training
    path (all-reduce before clip_grad_norm_, train.py:109; the loss is a batch mean,
    wrapper.py:27, so shard gradients are AVERAGED).  NCCL reduces with AVG in one pass over
    NVLink/NVSwitch; gloo (CPU tests) has no AVG: SUM then divide.  Returns the async work handle
    (or None)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return None
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    return None


class FlatGradients:
    """Gradient tensors laid out as views of ONE flat f32 buffer so the data-parallel exchange is a
    single all-reduce (75.5 MB for the Up-Down model) instead of one per parameter."""

    def __init__(self, shapes, device):
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += (n + 63) // 64 * 64                       # 256-byte aligned views
        self.flat = torch.zeros((off,), dtype=torch.float32, device=device)
        self._layout = list(zip(offs, sizes, shapes))
        self.views = [self.flat[o:o + n].view(s) for o, n, s in self._layout]

    def offset_of(self, i: int) -> int:
        """element offset of tensor i in the flat buffer (buckets are contiguous ranges of it)"""
        return self._layout[i][0] if i < len(self._layout) else self.flat.numel()

    def fresh_views(self):
        """new tensor objects over the same memory (nobody else holds them, so autograd may adopt them as .grad)"""
        return [self.flat[o:o + n].view(s) for o, n, s in self._layout]

    def matches(self, shapes, device):
        return (len(shapes) == len(self.views) and self.flat.device == torch.device(device)
                and all(tuple(v.shape) == tuple(s) for v, s in zip(self.views, shapes)))


def reduce_score(local_score_sum: torch.Tensor, group=None) -> torch.Tensor:
    """sum of per-rank VQA scores (what evaluate() accumulates, train.py:186-189)"""
    import torch.distributed as dist
    t = local_score_sum.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
