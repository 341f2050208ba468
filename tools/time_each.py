"""Per-iteration device times of one config (is the distribution bimodal?). python tools/time_each.py c3 30"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
B, H, W, H_up, W_up = synth.CONFIGS[name]
dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="fp16"), synth.make_weights(seed=0)).cuda()
x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
out = torch.empty((B, 3, H_up, W_up), device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
with torch.no_grad():
    dec.forward_rows(x, (H_up, W_up), 0, H_up, out=out)
    torch.cuda.synchronize()
    os.environ.pop("DIINN_DEBUG_OCC", None)
    ev[0].record()
    for i in range(n):
        dec.forward_rows(x, (H_up, W_up), 0, H_up, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
print(name, "ms per decode:", " ".join(f"{t:.2f}" for t in ts))
