"""GPU bring-up: runs each check in its own subprocess with a timeout so that a hung kernel in one stage cannot
take the others (or the box) down. Usage on the GPU box:  python tools/gpu_bringup.py [stage ...]
Writes gpurun_out/bringup.log."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = {}


def stage(fn):
    STAGES[fn.__name__] = fn
    return fn


def _setup():
    import numpy as np
    import torch
    import diinn_b200
    from diinn_b200 import synth
    from oracle import diinn_oracle as orc
    return np, torch, diinn_b200, synth, orc


def _selftest(cg):
    np, torch, diinn_b200, synth, orc = _setup()
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    for (M, N, K) in [(256, 256, 64), (256, 256, 256), (512, 512, 576)]:
        g = torch.Generator(device="cpu").manual_seed(M + N + K)
        A = (torch.randn(M, K, generator=g)).to(torch.bfloat16).cuda()
        B = (torch.randn(N, K, generator=g)).to(torch.bfloat16).cuda()
        D = dec.debug_umma_gemm(A, B, cta_group=cg)
        torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        err = (D - ref).abs().max().item()
        print(f"selftest cg={cg} M={M} N={N} K={K}: max err {err:.3e} (ref absmax {ref.abs().max().item():.2f})")
        if err > 1e-2:
            bad = (D - ref).abs() > 1e-2
            rows = bad.any(1).nonzero().flatten()[:8].tolist()
            cols = bad.any(0).nonzero().flatten()[:8].tolist()
            print("  MISMATCH rows", rows, "cols", cols, "frac bad", bad.float().mean().item())


@stage
def selftest_f16():
    np, torch, diinn_b200, synth, orc = _setup()
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    for cg in (11, 12):
        for (M, N, K) in [(256, 256, 64), (512, 512, 256)]:
            g = torch.Generator(device="cpu").manual_seed(M + N + K)
            A = (torch.randn(M, K, generator=g) * 0.25).to(torch.float16).cuda()
            B = (torch.randn(N, K, generator=g) * 0.25).to(torch.float16).cuda()
            D = dec.debug_umma_gemm(A.view(torch.bfloat16), B.view(torch.bfloat16), cta_group=cg)
            torch.cuda.synchronize()
            ref = A.float() @ B.float().t()
            err = (D - ref).abs().max().item()
            print(f"selftest fp16-acc cg={cg} M={M} N={N} K={K}: max err {err:.3e} (ref absmax {ref.abs().max().item():.2f})")
            if err > 0.05:
                bad = (D - ref).abs() > 0.05
                print("  MISMATCH frac", bad.float().mean().item(), "D[0,:8]", D[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())


@stage
def selftest1():
    _selftest(1)


@stage
def selftest2():
    _selftest(2)


def _case(name):
    np, torch, diinn_b200, synth, orc = _setup()
    g = np.load(os.path.join(ROOT, "tests", "golden", "decoder.npz"))
    seed, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in g[f"{name}.meta"])
    kg, qg = (float(v) for v in g[f"{name}.gains"])
    weights = synth.make_weights(seed=seed, k_gain=kg, q_gain=qg)
    feat = synth.make_feat(fseed, B, H, W)
    return weights, feat, (H_up, W_up), g[f"{name}.out"]


def _decoder(weights, precision):
    np, torch, diinn_b200, synth, orc = _setup()
    dec = diinn_b200.FusedImplicitDecoder(mode=3, precision=precision)
    diinn_b200.load_numpy_weights(dec, weights)
    return dec.cuda()


@stage
def gather():
    np, torch, diinn_b200, synth, orc = _setup()
    g = np.load(os.path.join(ROOT, "tests", "golden", "posenc.npz"))
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    for name in ["c1", "c3", "c4", "odd2", "down"]:
        H, W, H_up, W_up = (int(v) for v in g[f"{name}.shape"])
        ih, iw, rh, rw = dec.debug_gather(H, W, H_up, W_up, "cuda")
        ok = (np.array_equal(ih.cpu().numpy(), g[f"{name}.ih"]) and np.array_equal(iw.cpu().numpy(), g[f"{name}.iw"])
              and np.array_equal(rh.cpu().numpy().view(np.uint32), g[f"{name}.rel_h"].view(np.uint32))
              and np.array_equal(rw.cpu().numpy().view(np.uint32), g[f"{name}.rel_w"].view(np.uint32)))
        print(f"gather {name}: bit-exact={ok}")


@stage
def fp32():
    np, torch, diinn_b200, synth, orc = _setup()
    for name in ["c1", "odd2", "x1_batch", "frac", "stress"]:
        weights, feat, size, ref = _case(name)
        dec = _decoder(weights, "fp32")
        with torch.no_grad():
            out = dec(torch.from_numpy(feat).cuda(), size)
        torch.cuda.synchronize()
        print(f"fp32 {name}: max-abs err vs reference golden {np.abs(out.cpu().numpy() - ref).max():.3e}")


def _stage_a(cg):
    os.environ["DIINN_CTA_GROUP_A"] = str(cg)
    np, torch, diinn_b200, synth, orc = _setup()
    weights, feat, size, ref = _case("odd2")
    x = torch.from_numpy(feat).cuda()
    P32 = _decoder(weights, "fp32").debug_stage_a(x)
    P16 = _decoder(weights, "bf16").debug_stage_a(x)
    torch.cuda.synchronize()
    d = (P32 - P16).abs()
    print(f"stage A cg={cg}: |P_fp32| max {P32.abs().max().item():.3f}; bf16-vs-fp32 max {d.max().item():.3e} "
          f"mean {d.mean().item():.3e}")
    if d.max().item() > 0.05:
        bad = d > 0.05
        print("  bad rows", bad.any(1).nonzero().flatten()[:10].tolist(), "bad cols",
              bad.any(0).nonzero().flatten()[:10].tolist(), "frac", bad.float().mean().item())


@stage
def stage_a1():
    _stage_a(1)


@stage
def stage_a2():
    _stage_a(2)


def _bf16(cg):
    os.environ["DIINN_CTA_GROUP"] = str(cg)
    os.environ["DIINN_CTA_GROUP_A"] = str(cg)
    np, torch, diinn_b200, synth, orc = _setup()
    for name in ["x1_batch", "c1", "odd2", "frac", "stress"]:
        weights, feat, size, ref = _case(name)
        dec = _decoder(weights, "bf16")
        with torch.no_grad():
            out = dec(torch.from_numpy(feat).cuda(), size)
        torch.cuda.synchronize()
        e = np.abs(out.cpu().numpy() - ref)
        print(f"bf16 cg={cg} {name}: max-abs err {e.max():.3e} mean {e.mean():.3e} (|ref| max {np.abs(ref).max():.3f})")


@stage
def fp16():
    np, torch, diinn_b200, synth, orc = _setup()
    for name in ["x1_batch", "c1", "odd2", "frac", "stress"]:
        weights, feat, size, ref = _case(name)
        dec = _decoder(weights, "fp16")
        with torch.no_grad():
            out = dec(torch.from_numpy(feat).cuda(), size)
        torch.cuda.synchronize()
        e = np.abs(out.cpu().numpy() - ref)
        print(f"fp16 {name}: max-abs err {e.max():.3e} mean {e.mean():.3e} (|ref| max {np.abs(ref).max():.3f})")
    weights = synth.make_weights(seed=0)
    for prec in ("fp16", "bf16"):
        for name in ["c2x2", "c2x4", "c3", "c4"]:
            B, H, W, H_up, W_up = synth.CONFIGS[name]
            x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
            dec = _decoder(weights, prec)
            with torch.no_grad():
                for _ in range(3):
                    dec(x, (H_up, W_up))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    dec(x, (H_up, W_up))
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"timing {prec} {name}: {ms:.3f} ms  {B * H_up * W_up / ms / 1e3:.1f} Mpx/s")


@stage
def bf16_cg1():
    _bf16(1)


@stage
def bf16_cg2():
    _bf16(2)


@stage
def timing():
    np, torch, diinn_b200, synth, orc = _setup()
    weights = synth.make_weights(seed=0)
    for cg in (2, 1):
        os.environ["DIINN_CTA_GROUP"] = str(cg)
        for name in ["c2x2", "c2x4", "c3"]:
            B, H, W, H_up, W_up = synth.CONFIGS[name]
            x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
            dec = _decoder(weights, "bf16")
            with torch.no_grad():
                for _ in range(3):
                    dec(x, (H_up, W_up))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    dec(x, (H_up, W_up))
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            px = B * H_up * W_up
            print(f"timing cg={cg} {name}: {ms:.3f} ms  {px / ms / 1e3:.1f} Mpx/s")
        break  # cta_group is latched per process on first use


@stage
def mma_rate():
    np, torch, diinn_b200, synth, orc = _setup()
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    M, N, K = 8192, 2048, 4096
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    for cg in (1, 2):
        for _ in range(2):
            dec.debug_umma_gemm(A, B, cta_group=cg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dec.debug_umma_gemm(A, B, cta_group=cg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"mma_rate cg={cg}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    C = A @ B.t()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(10):
        C = A @ B.t()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(f"cuBLAS same shape: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


@stage
def mma_pace():
    import ctypes as C
    np, torch, diinn_b200, synth, orc = _setup()
    from diinn_b200 import _lib
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    lib, h = dec._ensure_handle(torch.device("cuda:0"))
    for per_q in (2, 4):
        for x32 in (0, 128):
            mma_iters = 20000
            buf = torch.zeros(74, device="cuda")
            noise = 32 | 8 | (per_q << 8) | x32
            for _ in range(2):
                _lib.check(lib, h, lib.diinn_debug_umma_pace(h, 2, 256, mma_iters, 148, C.c_void_p(buf.data_ptr()), noise, None))
            torch.cuda.synchronize()
            v = buf.cpu().numpy()
            print(f"ldtm_pace warps/quarter={per_q} mode={x32}: {v.mean():.1f} clk per load instruction per warp "
                  f"(mode 0: 16 columns per instruction, mode 128: 32 columns of 16-bit data packed)")
    for cg in ():
        for noise in (0, 8, 24):
            n_cols, n_ctas = 256, 148
            buf = torch.zeros(n_ctas // cg, device="cuda")
            for _ in range(2):
                _lib.check(lib, h, lib.diinn_debug_umma_pace(h, cg, n_cols, 65536, n_ctas, C.c_void_p(buf.data_ptr()), noise, None))
            torch.cuda.synchronize()
            v = buf.cpu().numpy()
            print(f"mma_pace cg={cg} N={n_cols} ctas={n_ctas} noise={noise}: {v.mean():.1f} clk/MMA (min {v.min():.1f} max {v.max():.1f})")


def main():
    names = sys.argv[1:] or list(STAGES)
    if len(names) == 1 and names[0].startswith("--run="):
        STAGES[names[0][6:]]()
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "a")
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), f"--run={n}"], capture_output=True, text=True,
                               timeout=int(os.environ.get("BRINGUP_TIMEOUT", "150")))
            msg = f"=== {n}: rc={p.returncode} ({time.time() - t0:.1f}s)\n{p.stdout}{p.stderr[-3000:]}"
        except subprocess.TimeoutExpired as e:
            msg = f"=== {n}: TIMEOUT after {time.time() - t0:.1f}s\n{(e.stdout or b'').decode() if isinstance(e.stdout, bytes) else (e.stdout or '')}"
        print(msg, flush=True)
        log.write(msg + "\n")
        log.flush()


if __name__ == "__main__":
    main()
