"""Multi-GPU host plumbing: contiguous stream blocks per rank and a host-side event gather.

Streams are independent (reference: one DemodTask + one MessageReceiver per stream, src/demod.rs:25-40,
src/recv.rs:33-60), so there is no collective on the data path.  torch.distributed is used only to bring
the already-decoded events of every rank to rank 0 (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

from ._lib import EVENT_DTYPE


def stream_block(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block owned by `rank`: stream s lives on rank floor(s * world / n_total)."""
    first = -(-rank * n_total // world)
    last = -(-(rank + 1) * n_total // world)
    return first, last - first


def owner_of(stream: int, n_total: int, world: int) -> int:
    return stream * world // n_total


def globalise(events: np.ndarray, first_stream: int) -> np.ndarray:
    """Rewrite context-local stream indices to global ones."""
    out = events.copy()
    out["stream"] += np.uint32(first_stream)
    return out


def gather_events(events: np.ndarray, dst: int = 0, group=None):
    """Gather per-rank event arrays (global stream ids) on `dst`; returns the (stream, sample)-ordered
    concatenation there and None elsewhere.  Works with any backend (tensors live on the CPU for gloo and on
    the current CUDA device for NCCL)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    raw = torch.from_numpy(np.ascontiguousarray(events).view(np.uint8).copy()).to(dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([raw.numel()], dtype=torch.int64, device=dev), group=group)
    cap = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros(max(cap, 1), dtype=torch.uint8, device=dev)
    padded[: raw.numel()] = raw
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    if rank != dst:
        return None
    parts = [b[: int(s.item())].cpu().numpy().view(EVENT_DTYPE) for b, s in zip(bufs, sizes)]
    ev = np.concatenate(parts) if parts else np.zeros(0, dtype=EVENT_DTYPE)
    return ev[np.lexsort((ev["sample"], ev["stream"]))]
