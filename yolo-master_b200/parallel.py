"""Multi-GPU plumbing for the inference path: replicas only, no data-path collective.

Every operator of the forward is per image (routing, GroupNorm, top-k, NMS), so a batch shards by images with nothing
exchanged in the timed loop.  What remains is start-up and bookkeeping, done with torch.distributed (NCCL over NVLink on the
GPU box, Gloo in the CPU tests):
  * `broadcast_module_state` — rank `src` owns the checkpoint; parameters and buffers are broadcast once;
  * `shard_range`            — contiguous split of a global batch over ranks (the reference's DDP sampler order);
  * `max_over_ranks`         — device-timed step times are reduced with MAX (the slowest rank defines the step);
  * `gather_detections`      — optional epilogue: all-gather of the per-rank (b, 300, 6) results.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_module_state(module: torch.nn.Module, src: int = 0) -> int:
    """Broadcast every parameter and buffer from `src`.  Returns the number of tensors sent."""
    n = 0
    if not is_dist():
        return n
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)
            t.add_(0)        # an in-place op on the tensor itself bumps `_version`: `.data` writes do not, and the weight-pack caches of
            n += 1           # the modules (PackCache / cached_f32 / cached_pack) key on (data_ptr, _version, device)
    for m in module.modules():      # belt and braces: drop derived packs and captured graphs built before the broadcast
        for k in ("_ym_pack", "_ym_packs", "_ym_f32"):
            m.__dict__.pop(k, None)
        if isinstance(m.__dict__.get("_graphs"), dict):
            m.__dict__["_graphs"].clear()
    return n


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the images rank `rank` owns; the first `global_batch % world` ranks take one extra image."""
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_detections(local: torch.Tensor) -> torch.Tensor:
    """All-gather equally sized per-rank detection tensors along dim 0."""
    if not is_dist():
        return local
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, local.contiguous())
    return torch.cat(parts, 0)
