"""Timing ablations of the stage-B epilogue: builds side libraries with -DDIINN_ABL=<mask> (results are wrong by
construction, only the time is meaningful) and times c3 with each.

    python tools/ablate_stage_b.py build 1 2 4 8 15      (here, no GPU needed)
    python tools/ablate_stage_b.py run 0 1 2 4 8 15      (on the GPU box)
    python tools/ablate_stage_b.py build-src 100 /tmp/old_stage_b.cu   (A/B another stage_b source as "mask" 100;
                                                                        box-to-box variance is ~10 %, so only
                                                                        numbers from the SAME gpurun call compare)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dual-interactive-implicit-neural-network_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402


def lib_for(mask):
    return B.LIB if mask == 0 else os.path.join(PKG, "build", f"libdiinn_b200_abl{mask}.so")


def build(masks):
    B.build()
    procs = []
    for m in masks:
        obj = os.path.join(PKG, "build", f"stage_b_umma_abl{m}.o")
        procs.append((m, obj, subprocess.Popen([B._nvcc(), *B.NVCC_FLAGS, *(os.environ.get("ABL_DEFS", f"-DDIINN_ABL={m}").split()), "-c",
                                                os.path.join(B.CSRC, "stage_b_umma.cu"), "-o", obj])))
    for m, obj, p in procs:
        assert p.wait() == 0
        objs = [os.path.join(PKG, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "stage_b_umma.cu"] + [obj]
        subprocess.check_call([B._nvcc(), "-shared", "-o", lib_for(m), *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
        print("built", lib_for(m))


def build_src(args):
    tag, path = int(args[0]), args[1]
    B.build()
    tmp = os.path.join(B.CSRC, "_stage_b_variant.cu")
    with open(path) as f, open(tmp, "w") as g:
        g.write(f.read())
    try:
        obj = os.path.join(PKG, "build", f"stage_b_umma_abl{tag}.o")
        subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, "-c", tmp, "-o", obj])
    finally:
        os.remove(tmp)
    objs = [os.path.join(PKG, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "stage_b_umma.cu"] + [obj]
    subprocess.check_call([B._nvcc(), "-shared", "-o", lib_for(tag), *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    print("built", lib_for(tag))


def run(masks):
    for m in masks:
        env = dict(os.environ, DIINN_B200_LIB=lib_for(m))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_decode.py"), "c3", os.environ.get("ABL_PRECISION", "bf16"), os.environ.get("ABL_ITERS", "10")], env=env,
                             capture_output=True, text=True, timeout=120)
        print(f"ABL mask {m:2d}: {out.stdout.strip() or out.stderr.strip()[-300:]}")


if __name__ == "__main__":
    if sys.argv[1] == "build-src":
        build_src(sys.argv[2:])
    else:
        {"build": build, "run": run}[sys.argv[1]]([int(a) for a in sys.argv[2:]])
