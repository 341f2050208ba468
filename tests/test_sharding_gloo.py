"""CPU, world_size 2 over gloo: the row-tile partition + assembly logic of diinn_b200.decode_sharded (the N>1 path of
bench.py) with a stand-in decoder, so the host-side logic is covered without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import diinn_b200


class StandInDecoder:
    """forward_rows of a decoder whose output depends only on (b, c, row, col): any partition must assemble to the
    same image, like the real kernels (per-pixel arithmetic independent of the tiling)."""

    def forward_rows(self, x, size, r0, r1, out=None):
        B = x.shape[0]
        H_up, W_up = size
        full = self.full(B, H_up, W_up, x.dtype)
        if out is not None:  # in-place into a (row-padded) full-size buffer, like FusedImplicitDecoder.forward_rows
            out[:, :, r0:r1] = full[:, :, r0:r1]
            return out
        return full[:, :, r0:r1].contiguous()

    @staticmethod
    def full(B, H_up, W_up, dtype):
        b = torch.arange(B).view(B, 1, 1, 1)
        c = torch.arange(3).view(1, 3, 1, 1)
        r = torch.arange(H_up).view(1, 1, H_up, 1)
        w = torch.arange(W_up).view(1, 1, 1, W_up)
        return (1000.0 * b + 100.0 * c + r + 0.001 * w).to(dtype)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sizes, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for (B, H_up, W_up) in sizes:
            x = torch.zeros(B, 64, 4, 4)
            for bands in (None, 1, 3):
                full = diinn_b200.decode_sharded(StandInDecoder(), x, (H_up, W_up), bands=bands)
                ok &= bool(torch.equal(full, StandInDecoder.full(B, H_up, W_up, x.dtype)))
            # encoder hand-off: only rank 0 holds the real feature map; channels-last goes through its dense NHWC view
            for fmt in (torch.contiguous_format, torch.channels_last):
                real = torch.arange(B * 64 * 16, dtype=torch.float32).view(B, 64, 4, 4).contiguous(memory_format=fmt)
                xb = real.clone() if rank == 0 else torch.zeros_like(real)
                full = diinn_b200.decode_sharded(StandInDecoder(), xb, (H_up, W_up), feat_src=0)
                ok &= bool(torch.equal(xb, real)) and bool(torch.equal(full, StandInDecoder.full(B, H_up, W_up, x.dtype)))
            tile = diinn_b200.decode_sharded(StandInDecoder(), x, (H_up, W_up), gather="none")
            r0, r1 = diinn_b200.tile_partition(x.shape[2], H_up, world)[rank]
            ok &= bool(torch.equal(tile, StandInDecoder.full(B, H_up, W_up, x.dtype)[:, :, r0:r1]))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_decode_sharded_world2_gloo():
    world = 2
    sizes = [(1, 8, 5), (2, 7, 3), (1, 1, 4), (1, 1356, 6)]  # even, uneven, fewer rows than ranks, DIV2K row count
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), sizes, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def test_band_partition_covers_every_row_once():
    for n_rows, world, bands in [(1356, 8, 1), (4320, 8, 4), (7, 2, 3), (1, 2, 1), (100, 3, 5)]:
        sub, parts = diinn_b200.band_partition(n_rows, world, bands)
        seen = []
        for r in range(world):
            assert len(parts[r]) == bands
            for k, (a, b) in enumerate(parts[r]):
                assert b - a <= sub and (b == a or a == (k * world + r) * sub)
                seen += list(range(a, b))
        assert sorted(seen) == list(range(n_rows))


def test_single_process_passthrough():
    x = torch.zeros(1, 64, 4, 4)
    out = diinn_b200.decode_sharded(StandInDecoder(), x, (9, 5))
    assert torch.equal(out, StandInDecoder.full(1, 9, 5, x.dtype))
