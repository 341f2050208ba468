"""Same-box A/B of stage B's phase table (integer scale factors): default vs DIINN_NO_TAB=1 (canonical coordinates, sines
computed per pixel -- must be BIT-IDENTICAL to the table) vs DIINN_NO_CANON=1 (the reference's per-pixel fp32 coordinates,
i.e. the kernel before this change). The switches are read once per process, hence one subprocess per leg, alternating.

    python tools/ab_canon.py [rounds]
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("c3", "fp16"), ("c3", "bf16"), ("c2x4", "fp16"), ("c2x2", "fp16"), ("c2x3", "fp16"), ("c1", "fp16"), ("c4", "fp16")]


def worker():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import diinn_b200
    from diinn_b200 import synth
    from oracle import diinn_oracle as orc

    res = {}
    w0 = synth.make_weights(seed=0)
    with torch.no_grad():
        # small integer-scale shapes against the oracle (x4 select, x2 / x3 classic, 2x8 phases, modes 1 / 2 / 4)
        for (b, h, w, s_h, s_w) in ((1, 24, 24, 4, 4), (2, 13, 17, 2, 2), (1, 11, 9, 3, 3), (1, 10, 12, 2, 8), (1, 12, 10, 4, 3)):
            f = synth.make_feat(5, b, h, w)
            size = (h * s_h, w * s_w)
            x = torch.from_numpy(f).cuda()
            for mode in (3, 1, 4):
                wm = synth.make_weights(seed=mode, mode=mode) if mode != 3 else w0
                ref = orc.decoder_forward(wm, f, size, mode=mode)
                for prec in ("fp16", "bf16"):
                    dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=mode, precision=prec), wm).cuda()
                    out = dec(x, size)
                    key = f"small {h}x{w} x{s_h}x{s_w} B{b} m{mode} {prec}"
                    res[key] = {"err": float(np.abs(out.cpu().numpy() - ref).max()),
                                "hash": hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]}
                    # row tiles (some of which take K_sel = 32 and must compute) are bit-identical to the full decode
                    rows = [(0, 5), (5, 6), (6, 19), (19, size[0])]
                    tiles = torch.cat([dec.forward_rows(x, size, a, b_) for a, b_ in rows if b_ > a], dim=2)
                    res[key]["tiles_equal"] = bool(torch.equal(tiles, out))
        ws = synth.make_weights(seed=0, k_gain=3.0, q_gain=10.0)
        f = synth.make_feat(4, 1, 16, 20)
        ref = orc.decoder_forward(ws, f, (64, 80))
        for prec in ("fp16", "bf16"):
            dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=prec), ws).cuda()
            out = dec(torch.from_numpy(f).cuda(), (64, 80))
            res[f"stress x4 {prec}"] = {"err": float(np.abs(out.cpu().numpy() - ref).max()),
                                         "hash": hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]}
        for name, prec in CASES:
            B, H, W, H_up, W_up = synth.CONFIGS[name]
            dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=prec), w0).cuda()
            x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
            out = torch.empty((B, 3, H_up, W_up), device="cuda")
            n = 3 if name == "c4" else 10
            for _ in range(3):
                dec.forward_rows(x, (H_up, W_up), 0, H_up, out=out)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    dec.forward_rows(x, (H_up, W_up), 0, H_up, out=out)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / n)
            res[f"{name} {prec}"] = {"ms": best, "hash": hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]}
    print("RESULT " + json.dumps(res), flush=True)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    legs = {"table": {}, "no_tab": {"DIINN_NO_TAB": "1"}, "no_canon": {"DIINN_NO_CANON": "1"}}
    runs = {k: [] for k in legs}
    for _ in range(rounds):
        for leg, extra in legs.items():
            env = dict(os.environ, **extra)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=env, capture_output=True, text=True,
                               timeout=900)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
            if p.returncode != 0 or not line:
                print(f"[{leg}] FAILED rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}", flush=True)
                continue
            runs[leg].append(json.loads(line[0][7:]))
    bad = 0
    keys = list(runs["table"][0].keys()) if runs["table"] else []
    for k in keys:
        row = []
        for leg in legs:
            rs = [r[k] for r in runs[leg] if k in r]
            if not rs:
                row.append(f"{leg}: -")
                continue
            if "ms" in rs[0]:
                row.append(f"{leg}: " + "/".join(f"{r['ms']:.3f}" for r in rs) + f" ms #{rs[0]['hash'][:6]}")
            else:
                extra = "" if rs[0].get("tiles_equal", True) else " TILES-DIFFER"
                bad += 0 if rs[0].get("tiles_equal", True) else 1
                row.append(f"{leg}: err {rs[0]['err']:.2e} #{rs[0]['hash'][:6]}{extra}")
        same = (runs["no_tab"] and runs["table"] and runs["table"][0][k]["hash"] == runs["no_tab"][0][k]["hash"])
        bad += 0 if same else 1
        print(f"{k:34s} " + " | ".join(row) + ("" if same else "  TABLE != COMPUTE"), flush=True)
    print("AB_CANON_OK" if not bad else f"AB_CANON_FAIL ({bad})", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    worker() if len(sys.argv) > 1 and sys.argv[1] == "worker" else main()
