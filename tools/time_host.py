"""Time the host entry (diinn_decode_host: H2D + decode + D2H) for several band splits. python tools/time_host.py [cfg]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, H, W, H_up, W_up = synth.CONFIGS[name]
dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="fp16"),
                                    synth.make_weights(seed=0)).cuda()
feat = torch.from_numpy(synth.make_feat(1, B, H, W)).pin_memory()
out = torch.empty((B, 3, H_up, W_up), dtype=torch.float32).pin_memory()
x = feat.cuda()
with torch.no_grad():
    for _ in range(3):
        dec(x, (H_up, W_up))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        dec(x, (H_up, W_up))
    torch.cuda.synchronize()
    print(f"{name}: device-resident decode {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
    for spec in sys.argv[2:] or ["", "1", "1,1,1,1", "1,5,5,4,1", "1,3,4,4,3,1", "1,7,8,7,1", "2,5,5,3,1", "1,2,4,4,3,1,1"]:
        if spec:
            os.environ["DIINN_HOST_BANDS"] = spec
        else:
            os.environ.pop("DIINN_HOST_BANDS", None)
        for _ in range(3):
            dec.decode_host(feat, (H_up, W_up), out_host=out)
        t0 = time.perf_counter()
        for _ in range(20):
            dec.decode_host(feat, (H_up, W_up), out_host=out)
        ms = (time.perf_counter() - t0) / 20 * 1e3
        print(f"  bands {spec or 'default':16s}: {ms:.3f} ms  ({B * H_up * W_up / ms / 1e3:.1f} Mpx/s)  checksum {float(out.double().sum()):.4f}")
