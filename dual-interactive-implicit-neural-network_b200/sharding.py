"""Row-tile sharding of the HR query grid across the GPUs of one box.

Every HR query is independent given the (replicated) LR feature map and weights -- the reference itself already splits
queries into independent column strips (batched_step, diinn.py:149-160) -- so rank r decodes HR rows
[row0_r, row1_r) with the same kernels and NCCL is used only to assemble the output image. Per-pixel arithmetic does
not depend on the partition, hence the assembled image is bit-identical to the 1-GPU one.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def row_partition(n_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [row0,row1) per rank; the first n_rows % world ranks get one extra row
    (1356 rows on 8 ranks -> 170,170,170,170,169,169,169,169). Ranks beyond n_rows get empty ranges."""
    base, extra = divmod(n_rows, world)
    out, r0 = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((r0, r0 + n))
        r0 += n
    return out


def decode_sharded(decoder, x: torch.Tensor, size, group: Optional[dist.ProcessGroup] = None,
                   gather: str = "all") -> torch.Tensor:
    """Each rank decodes its row tile of the (B,3,H_up,W_up) output, then the tiles are assembled with NCCL
    (gloo in the CPU tests, with a stand-in decoder).

    gather="all": every rank returns the full image (all_gather of padded equal-size row tiles);
    gather="none": returns only this rank's (B,3,rows,W_up) tile.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    H_up, W_up = int(size[0]), int(size[1])
    parts = row_partition(H_up, world)
    r0, r1 = parts[rank]
    B = x.shape[0]
    max_rows = max(b - a for a, b in parts)
    # padded tile so that all_gather_into_tensor sees equal shapes even when H_up % world != 0
    tile = torch.zeros((B, 3, max_rows, W_up), dtype=x.dtype, device=x.device)
    if r1 > r0:
        tile[:, :, : r1 - r0] = decoder.forward_rows(x, (H_up, W_up), r0, r1)
    if gather == "none" or world == 1:
        return tile[:, :, : r1 - r0] if gather == "none" else tile[:, :, :H_up]
    # output laid out as the dim-0 concatenation of the per-rank tiles (the form both NCCL and gloo accept)
    flat = torch.empty((world * B, 3, max_rows, W_up), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, tile, group=group)
    gathered = flat.view(world, B, 3, max_rows, W_up)
    # one concatenation kernel re-interleaves the rank-major tiles into NCHW and drops the padding rows
    return torch.cat([gathered[r, :, :, : b - a] for r, (a, b) in enumerate(parts) if b > a], dim=2)
