"""Timings of the secondary paths (modes 1/2, eval glue, PSNR) on one GPU. python tools/time_misc.py [config]"""
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
    for mode in (3, 2, 1):
        dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=mode, precision="fp16"),
                                            synth.make_weights(seed=0, mode=mode)).cuda()
        dec.set_profiling(True)
        ms = timed(lambda: dec(x, (H_up, W_up)))
        kt = dec.kernel_times()
        n = max(kt["decodes"], 1)
        print(f"{name} mode {mode}: {ms:.3f} ms/decode  (layout {kt['layout_ms'] / n:.3f}, stage A + LR chain {kt['stage_a_ms'] / n:.3f}, "
              f"stage B {kt['stage_b_ms'] / n:.3f})")
        dec.set_profiling(False)
    dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="fp16"), synth.make_weights(seed=0)).cuda()
    base = timed(lambda: dec(x, (H_up, W_up)))
    dec.set_output_transform(sub=0.5, div=0.5, clamp=(0, 1))
    t1 = timed(lambda: dec(x, (H_up, W_up)))
    dec.set_output_transform(sub=0.5, div=0.5, clamp=(0, 1), uint8=True)
    t2 = timed(lambda: dec(x, (H_up, W_up)))
    feat = x.cpu().pin_memory()
    t3 = timed(lambda: dec.decode_host(feat, (H_up, W_up)))
    dec.set_output_transform()
    t4 = timed(lambda: dec.decode_host(feat, (H_up, W_up)))
    print(f"{name} eval glue: plain {base:.3f} ms | +denorm+clamp {t1:.3f} ms | +uint8 {t2:.3f} ms | host entry uint8 {t3:.3f} ms vs fp32 {t4:.3f} ms")
    sr = dec(x, (H_up, W_up))
    hr = torch.rand_like(sr)
    tp = timed(lambda: dec.calc_psnr(sr, hr, dataset="div2k", scale=4))
    nbytes = 2 * sr.numel() * 4
    print(f"{name} calc_psnr: {tp:.3f} ms incl. the host sync = {nbytes / tp / 1e6:.0f} GB/s over {nbytes / 1e6:.0f} MB")
