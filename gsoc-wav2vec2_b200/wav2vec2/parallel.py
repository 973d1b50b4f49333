"""Data-parallel host logic (one process per GPU, torch.distributed).

The reference's only parallelism is synchronous data parallelism (tf.distribute Mirrored/TPU strategy,
src/main.py:141-156): global_batch = replicas x per_device, every utterance is independent in the forward
pass, and training sums gradients across replicas once per step with the loss pre-divided by the global
batch (src/main.py:196-200, losses.py:45).  Inference therefore needs NO collective; the only collective of
the path is one all-reduce(SUM) over a flat gradient buffer.
"""
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-even split of the batch dimension: rank r owns [lo, hi)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(batch.shape[0], rank, world)
    return batch[lo:hi]


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed quantities are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def flatten_grads(grads: Dict[str, torch.Tensor], order: List[str]) -> Tuple[torch.Tensor, List[Tuple[str, int, torch.Size]]]:
    """One flat fp32 buffer (C1 in SURVEY 2c: a single all-reduce message) + the layout to undo it."""
    layout, chunks, off = [], [], 0
    for name in order:
        g = grads[name].reshape(-1).float()
        layout.append((name, off, grads[name].shape))
        chunks.append(g)
        off += g.numel()
    return torch.cat(chunks) if chunks else torch.zeros(0), layout


def unflatten_grads(flat: torch.Tensor, layout) -> Dict[str, torch.Tensor]:
    out = {}
    for name, off, shape in layout:
        n = int(torch.Size(shape).numel())
        out[name] = flat[off: off + n].view(shape)
    return out


def allreduce_gradients(grads: Dict[str, torch.Tensor], order: List[str]) -> Dict[str, torch.Tensor]:
    """SUM all-reduce of every gradient in ONE collective (loss is pre-scaled by 1/global_batch)."""
    flat, layout = flatten_grads(grads, order)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return unflatten_grads(flat, layout)
