"""Multi-GPU sharding of the embedding path: every image / text is an independent unit, so a batch
or a corpus is split into contiguous row ranges `[i*B/N, (i+1)*B/N)`, one engine replica per GPU (the scaled-up
form of the reference's `duplicate()`, src/vision.rs:87-91), and the per-shard embeddings are simply copied back.
There is no collective on the data path; `gather_rows` only assembles the result on one rank when a caller wants
the whole matrix in one place (e.g. `rank_images` over a sharded corpus, src/clip.rs:136-170)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ranges: the first `n % world` ranks get one extra row."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def counter_images(start: int, stop: int, size: int, seed: int = 0) -> np.ndarray:
    """Counter-based synthetic corpus (BASELINE config 5): image i is a pure function of (seed, i), so any index can
    be regenerated on any rank or on the CPU without holding the 44 GB corpus (splitmix64 per 8-byte lane)."""
    n = stop - start
    words = (size * size * 3 + 7) // 8
    idx = (np.arange(start, stop, dtype=np.uint64)[:, None] * np.uint64(words) + np.arange(words, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):  # arithmetic is modulo 2^64 by design
        z = idx + np.uint64((seed * 0x9E3779B97F4A7C15 + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.view(np.uint8).reshape(n, words * 8)[:, :size * size * 3].reshape(n, size, size, 3)


def embed_sharded(embed_fn: Callable[[int, int], np.ndarray], n: int, rank: int, world: int) -> Tuple[int, np.ndarray]:
    """Runs `embed_fn(start, stop)` on this rank's range; returns (start, rows)."""
    start, stop = shard_range(n, rank, world)
    if stop == start:
        return start, np.zeros((0, 0), dtype=np.float32)
    return start, np.ascontiguousarray(embed_fn(start, stop), dtype=np.float32)


def gather_rows(local: np.ndarray, n: int, dim: int, rank: int, world: int, dst: int = 0,
                device: Optional[str] = None) -> Optional[np.ndarray]:
    """Assembles the [n, dim] matrix on `dst` from per-rank contiguous shards (torch.distributed gather; works with
    gloo on CPU and nccl on GPUs).  Returns None on the other ranks."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist

    base, extra = divmod(n, world)
    max_rows = base + (1 if extra else 0)
    dev = torch.device(device) if device else torch.device("cpu")
    buf = torch.zeros((max_rows, dim), dtype=torch.float32, device=dev)
    if local.size:
        buf[:local.shape[0]] = torch.from_numpy(local).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst)
    if rank != dst:
        return None
    out = np.empty((n, dim), dtype=np.float32)
    for r in range(world):
        s, e = shard_range(n, r, world)
        out[s:e] = parts[r][:e - s].cpu().numpy()
    return out
