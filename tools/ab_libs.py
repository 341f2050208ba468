"""Same-box A/B of library builds (DIINN_B200_LIB side libraries of tools/ablate_stage_b.py): per-decode device times of one
config, min / median / mean per leg, legs alternating.   python tools/ab_libs.py c3 fp16 3 0 100 ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def worker(name, prec, n):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import diinn_b200
    from diinn_b200 import synth
    B, H, W, H_up, W_up = synth.CONFIGS[name]
    dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=prec), synth.make_weights(seed=0)).cuda()
    x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
    out = torch.empty((B, 3, H_up, W_up), device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    with torch.no_grad():
        for _ in range(3):
            dec.forward_rows(x, (H_up, W_up), 0, H_up, out=out)
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(n):
            dec.forward_rows(x, (H_up, W_up), 0, H_up, out=out)
            ev[i + 1].record()
        torch.cuda.synchronize()
    ts = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
    print(f"RES min {ts.min():.3f} med {np.median(ts):.3f} mean {ts.mean():.3f} max {ts.max():.3f} first10 {ts[:10].mean():.3f}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "worker":
        worker(sys.argv[2], sys.argv[3], int(sys.argv[4]))
    else:
        import ablate_stage_b as A
        name, prec, rounds = sys.argv[1], sys.argv[2], int(sys.argv[3])
        masks = [int(a) for a in sys.argv[4:]]
        for r in range(rounds):
            for m in masks:
                env = dict(os.environ, DIINN_B200_LIB=A.lib_for(m))
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "worker", name, prec, os.environ.get("AB_N", "40")], env=env,
                                   capture_output=True, text=True, timeout=300)
                line = [ln for ln in p.stdout.splitlines() if ln.startswith("RES")]
                print(f"lib {m:3d} round {r}: {line[0] if line else p.stderr[-300:]}", flush=True)
