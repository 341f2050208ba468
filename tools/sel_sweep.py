"""Bring-up of the select-MMA variant: one small x4 and one x12 decode under several B_sel descriptor hypotheses
(DIINN_SEL_LBO / _SBO / _KSTEP), each in its own subprocess with a timeout, against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import diinn_b200
from diinn_b200 import synth
from oracle import diinn_oracle as orc
w = synth.make_weights(seed=0)
for (b_, h_, w_, hu_, wu_) in ((1, 24, 24, 96, 96), (1, 9, 11, 108, 132)):
    f_ = synth.make_feat(5, b_, h_, w_)
    ref_ = orc.decoder_forward(w, f_, (hu_, wu_))
    dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="fp16"), w).cuda()
    with torch.no_grad():
        out = dec(torch.from_numpy(f_).cuda(), (hu_, wu_))
    torch.cuda.synchronize()
    e = np.abs(out.cpu().numpy() - ref_)
    print(f"  {h_}x{w_}->{hu_}x{wu_}: max err {e.max():.3e} mean {e.mean():.3e}", flush=True)
''' % ROOT

configs = [
    {"DIINN_NO_SEL": "1"},
    {},
    {"DIINN_SEL_LBO": "1024", "DIINN_SEL_SBO": "4096"},
    {"DIINN_SEL_LBO": "1024", "DIINN_SEL_SBO": "2048"},
    {"DIINN_SEL_KSTEP": "4096"},
]
for cfg in configs:
    env = dict(os.environ, **cfg)
    print("config", cfg or "default", flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
        print(r.stdout.rstrip() or "  (no output)", flush=True)
        if r.returncode != 0:
            print("  rc", r.returncode, r.stderr[-600:], flush=True)
    except subprocess.TimeoutExpired:
        print("  TIMEOUT (hang)", flush=True)
