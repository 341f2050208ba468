"""torchrun --nproc-per-node N tools/time_sharded.py [config]: where does a sharded decode spend its time?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="fp16"), synth.make_weights(seed=0)).to(dev)
B, H, W, H_up, W_up = synth.CONFIGS[name]
x = torch.from_numpy(synth.make_feat(1, B, H, W)).to(dev)
r0, r1 = diinn_b200.tile_partition(H, H_up, world)[rank]


def timed(fn, n=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


with torch.no_grad():
    t_tile = timed(lambda: dec.forward_rows(x, (H_up, W_up), r0, r1))
    res = {}
    for bands in (1, 4):
        res[bands] = timed(lambda: diinn_b200.decode_sharded(dec, x, (H_up, W_up), bands=bands))
    parts = diinn_b200.tile_partition(H, H_up, world)
    if len({b - a for a, b in parts}) == 1:   # the bare all-gather timing needs equal tiles (c3 at 8 ranks is 170/169 rows)
        buf = torch.empty((B, 3, world * (r1 - r0), W_up), device=dev)
        t_gather = timed(lambda: [dist.all_gather_into_tensor(buf[0, c], buf[0, c, rank * (r1 - r0):(rank + 1) * (r1 - r0)]) for c in range(3)])
    else:
        t_gather = float("nan")
    fused = {}
    for mc in (False, True):
        try:
            fused[mc] = timed(lambda: diinn_b200.decode_sharded_fused(dec, x, (H_up, W_up), multicast=mc, clone=False))
        except Exception as e:
            fused[mc] = f"{type(e).__name__}: {e}"
if rank == 0:
    print(f"{name} world={world}: fused peer-store {fused[False]} ms, fused multicast {fused[True]} ms")
    print(f"{name} world={world}: own-tile decode {t_tile:.3f} ms | sharded bands->ms {res} | 3 in-place all-gathers alone {t_gather:.3f} ms")
dist.destroy_process_group()
