"""Per-decode device time of every decoder wiring (mode 1..4 x init_q) on one GPU: python tools/time_wirings.py [config]
Mode 4 is also split into stage B (q_3 dump) and the 3x3 conv; init_q lines carry the launch count of one decode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, H, W, H_up, W_up = synth.CONFIGS[name]
x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    for init_q in (False, True):
        for mode in (3, 4, 2, 1):
            w = synth.make_weights(seed=0, mode=mode, init_q=init_q)
            dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q, precision="fp16"), w).cuda()
            dec(x, (H_up, W_up))
            n0 = dec.launch_count()
            dec(x, (H_up, W_up))
            launches = dec.launch_count() - n0
            ms = timed(lambda: dec(x, (H_up, W_up)), n=5 if init_q else 10)
            px = B * H_up * W_up
            print(f"{name} mode {mode} init_q={int(init_q)}: {ms:.3f} ms/decode"
                  f" = {px / ms / 1e3:.0f} Mpx/s, {launches} launches", flush=True)
            del dec
    if name in ("c1", "c2x2"):
        for mode, init_q in ((3, True), (4, False)):
            w = synth.make_weights(seed=0, mode=mode, init_q=init_q)
            dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q, precision="fp32"), w).cuda()
            ms = timed(lambda: dec(x, (H_up, W_up)), n=3)
            print(f"{name} mode {mode} init_q={int(init_q)} fp32 path: {ms:.3f} ms/decode", flush=True)
