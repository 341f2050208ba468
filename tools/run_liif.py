"""One LIIF-proper decode (FusedLIIFQuery.forward on a BASELINE grid) for ncu / sanitizer / quick timing:
python tools/run_liif.py c1 fp16 3 [ensemble=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ens = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
B, H, W, H_up, W_up = synth.CONFIGS[name]
m = diinn_b200.load_liif_imnet(diinn_b200.FusedLIIFQuery(local_ensemble=ens, precision=precision),
                               {"imnet." + k: v for k, v in synth.make_liif_weights(1).items()}).cuda()
x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
with torch.no_grad():
    coord, cell = m.make_coord_and_cell(x, (H_up, W_up))
    m.query_rgb(x, coord, cell)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = m.query_rgb(x, coord, cell)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"LIIF {name} {precision} ensemble={int(ens)}: {ms:.3f} ms, {B * H_up * W_up / ms / 1e3:.1f} Mquery/s, checksum {float(out.double().sum()):.6f}")
