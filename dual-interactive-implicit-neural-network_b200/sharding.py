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


def row_partition(n_rows: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """Contiguous, balanced [row0,row1) per rank; the first n_rows % world ranks get one extra row
    (1356 rows on 8 ranks -> 170,170,170,170,169,169,169,169). Ranks beyond n_rows get empty ranges.
    align > 1: every boundary is a multiple of `align` rows (the image is dealt out in units of `align` rows:
    1356 rows, 8 ranks, align 4 -> 172,172,172,168,168,168,168,168)."""
    align = max(int(align), 1)
    units = -(-n_rows // align)
    base, extra = divmod(units, world)
    out, u0 = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((min(u0 * align, n_rows), min((u0 + n) * align, n_rows)))
        u0 += n
    return out


def row_align(H: int, H_up: int) -> int:
    """Row granularity of a sharded decode. On an integer scale factor s (<= 16) tile boundaries on multiples of s keep every
    8-row patch of stage B on whole LR cells, so a row tile's patches touch as few cells as the full image's (x4: 2 x 8 = 16,
    one select MMA per half slot and room for the phase table; an unaligned tile touches 3 x 8 = 24 and computes layer 0's
    sines). Any partition is bit-identical to the full decode -- this only picks the fast one."""
    s = H_up // H if H > 0 and H_up % H == 0 else 1
    return s if 1 <= s <= 16 else 1


def tile_partition(H: int, H_up: int, world: int) -> List[Tuple[int, int]]:
    """The row tiles decode_sharded_fused / decode_sharded(gather="none") give the ranks."""
    return row_partition(H_up, world, row_align(H, H_up))


def band_partition(n_rows: int, world: int, bands: int, align: int = 1) -> Tuple[int, List[List[Tuple[int, int]]]]:
    """Block-cyclic row bands for pipelined assembly: the image is cut into `bands` super-bands of world*sub rows, and
    inside super-band k rank r owns rows [(k*world + r)*sub, +sub) (clipped to n_rows). Returns (sub, per-rank lists).
    With bands=1 this is the plain contiguous split with equal (padded) tiles. align: `sub` is rounded up to a multiple of it
    (see row_align)."""
    sub = -(-n_rows // (world * bands))
    sub = -(-sub // max(int(align), 1)) * max(int(align), 1)
    out = []
    for r in range(world):
        mine = []
        for k in range(bands):
            a = (k * world + r) * sub
            mine.append((min(a, n_rows), min(a + sub, n_rows)))
        out.append(mine)
    return sub, out


_side_streams = {}
_symm_images = {}
_symm_turn = {}
last_fused_mode = None  # how the last decode_sharded_fused call reached the peers (for reports)


def _out_dtype(decoder, x: torch.Tensor) -> torch.dtype:
    fn = getattr(decoder, "_out_dtype", None)   # FusedImplicitDecoder: uint8 with the quantising output transform
    return fn(x) if fn is not None else x.dtype


def broadcast_features(x: torch.Tensor, src: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Encoder hand-off for the multi-GPU decode: run the encoder on ONE rank and broadcast its
    feature map over NVLink instead of running it on all of them; every rank passes a tensor of the right shape / dtype /
    memory format (contents ignored except on `src`). Channels-last tensors are sent through their dense NHWC view."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dense = x if x.is_contiguous() else x.permute(0, 2, 3, 1)
        if not dense.is_contiguous():
            raise ValueError("broadcast_features needs a contiguous or channels-last feature map")
        dist.broadcast(dense, src=src, group=group)
    return x


def _symmetric_image(shape, dtype, device, group, slot=0):
    """A (B,3,H_up,W_up) image buffer allocated in symmetric memory and mapped into every rank of `group`
    (torch.distributed._symmetric_memory): returns (local tensor, handle with .buffer_ptrs / .multicast_ptr / .barrier).
    `slot` distinguishes the two buffers decode_sharded_fused alternates between."""
    import torch.distributed._symmetric_memory as symm_mem
    pg = group if group is not None else dist.group.WORLD
    key = (tuple(shape), dtype, device.index, pg.group_name, slot)
    if key not in _symm_images:
        buf = symm_mem.empty(*shape, dtype=dtype, device=device)
        hdl = symm_mem.rendezvous(buf, pg.group_name)
        _symm_images[key] = (buf, hdl)
    return _symm_images[key]


def decode_sharded_fused(decoder, x: torch.Tensor, size, group: Optional[dist.ProcessGroup] = None,
                         multicast="auto", clone: bool = True, feat_src: Optional[int] = None) -> torch.Tensor:
    """Fused decode + assembly: the stage-B kernel of every rank stores each RGB value of its row tile straight into
    the image buffers of ALL ranks over NVLink (one `multimem.st` through the NVSwitch multicast mapping when the
    fabric offers it, else one peer store per rank), so the transfer rides under the math tile by tile and no
    collective is launched. ONE symmetric-memory barrier per call (completion: every rank's stores have landed
    everywhere). The barrier that used to protect buffer reuse is gone: calls alternate between two symmetric buffers,
    and a rank can only enter call k+1 -- which overwrites the buffer of call k-1 -- after the completion barrier of call
    k, which every rank joins after (in stream order) its own reads of call k-1's image.

    Returns the assembled (B,3,H_up,W_up) image on every rank. clone=False returns the symmetric buffer itself, which
    stays valid until the SECOND next call with the same shape. feat_src: see decode_sharded.
    multicast: True / False / "auto" -- measured at 8 GPUs the single `multimem.st` wins on small row tiles (c3: 0.293 vs
    0.308 ms) and the eight peer stores on large ones (c4: 2.52 vs 2.64 ms); "auto" switches at 2 Mpx per rank."""
    if feat_src is not None:
        broadcast_features(x, feat_src, group)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    H_up, W_up = int(size[0]), int(size[1])
    B = x.shape[0]
    odt = _out_dtype(decoder, x)   # uint8 when the decoder's eval glue quantises: a quarter of the bytes over NVLink
    pg = group if group is not None else dist.group.WORLD
    tkey = ((B, 3, H_up, W_up), odt, x.device.index, pg.group_name)
    slot = _symm_turn.get(tkey, 0)
    _symm_turn[tkey] = slot ^ 1
    buf, hdl = _symmetric_image((B, 3, H_up, W_up), odt, x.device, group, slot)
    r0, r1 = tile_partition(x.shape[2], H_up, world)[rank]
    if multicast == "auto":
        multicast = B * (r1 - r0) * W_up < (1 << 21)
    mc = int(hdl.multicast_ptr) if (multicast and odt == torch.float32) else 0  # 0 when the fabric has no multicast
    global last_fused_mode
    last_fused_mode = "nvswitch-multicast multimem.st" if mc else f"{world} peer stores per value"
    if r1 > r0:
        decoder.forward_rows(x, (H_up, W_up), r0, r1, out=buf, peer_ptrs=list(hdl.buffer_ptrs), multicast_ptr=mc)
    hdl.barrier(channel=slot)  # every rank's stores have landed everywhere (and its reads of the other buffer are done)
    return buf.clone() if clone else buf



def decode_sharded(decoder, x: torch.Tensor, size, group: Optional[dist.ProcessGroup] = None, gather: str = "all",
                   bands: Optional[int] = None, feat_src: Optional[int] = None) -> torch.Tensor:
    """One rank per GPU: every rank decodes its HR row bands of the (B,3,H_up,W_up) image **in place** into a (row-
    padded) full-size buffer, and NCCL all-gathers each band in place, channel by channel, on a side stream while the
    next band is being decoded. No copy, no data-path collective; returns the assembled image on every rank (a view of
    the padded buffer; rows are contiguous, the channel stride is H_pad*W_up).

    gather="none": returns only this rank's first band tile (B,3,rows,W_up) (used by tests / e2e).
    bands: number of bands per rank (pipelining depth); default 1 below ~1 Mpx per rank, else 4.
    feat_src: rank whose `x` is the real feature map (broadcast first, see broadcast_features); None = already replicated.
    """
    if feat_src is not None:
        broadcast_features(x, feat_src, group)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    H_up, W_up = int(size[0]), int(size[1])
    B = x.shape[0]
    if gather == "none":
        r0, r1 = tile_partition(x.shape[2], H_up, world)[rank]
        if r1 > r0:
            return decoder.forward_rows(x, (H_up, W_up), r0, r1)
        return torch.empty((B, 3, 0, W_up), dtype=_out_dtype(decoder, x), device=x.device)
    if world == 1:
        return decoder.forward_rows(x, (H_up, W_up), 0, H_up)
    if bands is None:
        bands = 4 if (H_up // world) * W_up * B >= (1 << 21) else 1
    sub, parts = band_partition(H_up, world, bands, row_align(x.shape[2], H_up))
    H_pad = world * bands * sub
    out = torch.empty((B, 3, H_pad, W_up), dtype=_out_dtype(decoder, x), device=x.device)
    on_gpu = x.is_cuda
    in_place = dist.get_backend(group) == "nccl"
    if on_gpu:
        key = (x.device.index, id(group))
        side = _side_streams.setdefault(key, torch.cuda.Stream(device=x.device))
        main = torch.cuda.current_stream(x.device)
        side.wait_stream(main)  # `out` was allocated on the main stream
    for k, (a, b) in enumerate(parts[rank]):
        if b > a:
            decoder.forward_rows(x, (H_up, W_up), a, b, out=out)
        base = k * world * sub
        mine = (k * world + rank) * sub

        def _gather():
            for bi in range(B):
                for c in range(3):
                    dst = out[bi, c, base:base + world * sub]
                    src = out[bi, c, mine:mine + sub]
                    dist.all_gather_into_tensor(dst, src if in_place else src.clone(), group=group)

        if on_gpu:
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                _gather()
        else:
            _gather()
    if on_gpu:
        main.wait_stream(side)
    return out[:, :, :H_up]
