"""shard.py -- row sharding of one weight tensor across GPUs.

The path shards along the output rows: tiles of bm/bits rows are independent given the (tiny,
replicated) activation row -- exactly how the reference spreads tiles over CPU threads
(3rdparty/llama.cpp/ggml/src/ggml.c:12636-12691).  Rank r owns a contiguous range of whole
reference tiles; every rank runs its own preprocessor (cheaper than broadcasting the LUT); the
output slices are disjoint, so the collective is an all-gather (no reduction, bit exact).

Host-side logic only (partition arithmetic + the torch.distributed call); used by bench.py under
torchrun and by the gloo tests on CPU.
"""
from __future__ import annotations

from typing import List, Tuple


def row_partition(m_rows: int, tile_rows: int, world: int) -> List[Tuple[int, int]]:
    """[(row0, rows)] per rank: whole tiles, remainders spread over the first ranks
    (e.g. 11008 rows, 64-row tiles, 8 ranks -> 172 tiles -> four ranks take 22, four take 21)."""
    if m_rows % tile_rows:
        raise ValueError("rows must be a multiple of the tile (bm/bits)")
    tiles = m_rows // tile_rows
    base, rem = divmod(tiles, world)
    out, start = [], 0
    for r in range(world):
        t = base + (1 if r < rem else 0)
        out.append((start * tile_rows, t * tile_rows))
        start += t
    return out


def gather_rows(local_out, m_rows: int, parts: List[Tuple[int, int]], group=None):
    """All-gather the per-rank output slices [N, rows_r] into [N, m_rows] on every rank.
    Uneven slices are padded to the largest one (a single all_gather_into_tensor over NCCL / gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = local_out.shape[0]
    mx = max(r for _, r in parts)
    buf = torch.zeros((n, mx), dtype=local_out.dtype, device=local_out.device)
    buf[:, : local_out.shape[1]] = local_out
    gathered = torch.empty((world, n, mx), dtype=local_out.dtype, device=local_out.device)
    if local_out.is_cuda:
        dist.all_gather_into_tensor(gathered.view(-1), buf.view(-1), group=group)
    else:
        dist.all_gather(list(gathered.unbind(0)), buf, group=group)
    full = torch.empty((n, m_rows), dtype=local_out.dtype, device=local_out.device)
    for r, (row0, rows) in enumerate(parts):
        full[:, row0:row0 + rows] = gathered[r, :, :rows]
    return full
