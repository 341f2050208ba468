"""First thing to run on a GPU box after touching the kernels: one small decode per precision / wiring against the oracle
(each guarded by the caller's `timeout`), then c3 timings. Prints one line per check; exits non-zero on a failed check."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth  # noqa: E402
from oracle import diinn_oracle as orc  # noqa: E402

TOL = {"fp32": 2e-6, "fp32_simt": 2e-6, "fp16": 4e-5, "bf16": 2e-4}
bad = 0
B, H, W, size = 2, 16, 20, (37, 51)
feat = synth.make_feat(4, B, H, W)
x = torch.from_numpy(feat).cuda()
with torch.no_grad():
    for mode in (3, 1, 2, 4):
        w = synth.make_weights(seed=mode, mode=mode)
        ref = orc.decoder_forward(w, feat, size, mode=mode)
        for prec in ("fp16", "bf16", "fp32", "fp32_simt"):
            dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=mode, precision=prec), w).cuda()
            t0 = time.time()
            out = dec(x, size)
            torch.cuda.synchronize()
            err = float(np.abs(out.cpu().numpy() - ref).max())
            tol = TOL["bf16"] if (mode == 4 and prec == "fp16") else TOL[prec]
            ok = err <= tol
            bad += not ok
            print(f"mode {mode} {prec:9s} err {err:.3e} (tol {tol:.0e}) {'OK' if ok else 'FAIL'}  [{time.time() - t0:.2f}s]", flush=True)
    # the select-MMA variant of stage B (CTA-pair patches of <= 30 LR cells: x4 -> K_sel 32, x12 -> K_sel 16), odd sizes
    w = synth.make_weights(seed=0)
    for (b_, h_, w_, hu_, wu_) in ((1, 24, 24, 96, 96), (2, 19, 23, 77, 93), (1, 9, 11, 108, 132), (1, 48, 48, 192, 192)):
        f_ = synth.make_feat(5, b_, h_, w_)
        ref_ = orc.decoder_forward(w, f_, (hu_, wu_))
        x_ = torch.from_numpy(f_).cuda()
        for prec in ("fp16", "bf16"):
            dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=prec), w).cuda()
            out = dec(x_, (hu_, wu_))
            torch.cuda.synchronize()
            err = float(np.abs(out.cpu().numpy() - ref_).max())
            ok = err <= TOL[prec]
            bad += not ok
            print(f"sel {h_}x{w_}->{hu_}x{wu_} B={b_} {prec}: err {err:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    # stress weights: absolute errors
    w = synth.make_weights(seed=0, k_gain=3.0, q_gain=10.0)
    ref = orc.decoder_forward(w, feat, size)
    for prec in ("fp16", "bf16", "fp32"):
        dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=prec), w).cuda()
        err = float(np.abs(dec(x, size).cpu().numpy() - ref).max())
        print(f"stress k3 q10 {prec:6s} err {err:.3e}", flush=True)
    # c3 timings
    Bc, Hc, Wc, Hu, Wu = synth.CONFIGS["c3"]
    xc = torch.from_numpy(synth.make_feat(1, Bc, Hc, Wc)).cuda()
    w0 = synth.make_weights(seed=0)
    outs = {}
    for prec in ("fp16", "bf16", "fp32"):
        dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=prec), w0).cuda()
        for _ in range(3):
            outs[prec] = dec(xc, (Hu, Wu))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dec(xc, (Hu, Wu))
        e1.record()
        torch.cuda.synchronize()
        print(f"c3 {prec}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    print(f"c3 fp16 vs fp32 max diff {float((outs['fp16'] - outs['fp32']).abs().max()):.3e}; bf16 vs fp32 "
          f"{float((outs['bf16'] - outs['fp32']).abs().max()):.3e}", flush=True)
print("QUICK_OK" if not bad else f"QUICK_FAIL ({bad})", flush=True)
sys.exit(1 if bad else 0)
