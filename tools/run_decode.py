"""Run N decodes of one BASELINE config (for ncu / quick timing). python tools/run_decode.py c3 bf16 5 [mode [init_q]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
init_q = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
B, H, W, H_up, W_up = synth.CONFIGS[name]
dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q, precision=precision),
                                    synth.make_weights(seed=0, mode=mode, init_q=init_q)).cuda()
x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
with torch.no_grad():
    dec(x, (H_up, W_up))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = dec(x, (H_up, W_up))
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{name} {precision} mode {mode} init_q={int(init_q)}: {ms:.3f} ms/decode, {B * H_up * W_up / ms / 1e3:.1f} Mpx/s, launches {dec.launch_count()}")
