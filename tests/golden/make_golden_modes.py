"""Golden fixtures for decoder modes 1 and 2, produced by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_modes.py
Outputs of the unmodified ``ImplicitDecoder(mode=1|2, init_q=False)`` (/root/reference/src/models/components/
diinn.py:57-72,116-131) on torch CPU fp32 for bit-reproducible synthetic weights / features (diinn_b200.synth).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from src.models.components.diinn import ImplicitDecoder  # noqa: E402  (the reference)
from diinn_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)

CASES = {  # name: (B, H, W, H_up, W_up, weight kwargs, bsize)
    "small": (1, 24, 20, 71, 63, {}, None),
    "c1": (1, 48, 48, 192, 192, {}, None),
    "batch_bsize": (2, 17, 23, 40, 51, {}, 700),
    "stress": (1, 24, 20, 60, 50, dict(k_gain=1.5, q_gain=10.0), None),
}


def main():
    out = {}
    for mode in (1, 2):
        for name, (B, H, W, H_up, W_up, wkw, bsize) in CASES.items():
            w = synth.make_weights(seed=mode, mode=mode, **wkw)
            dec = ImplicitDecoder(mode=mode, init_q=False).eval()
            dec.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w.items()}, strict=True)
            feat = synth.make_feat(10 + mode, B, H, W)
            y = dec(torch.from_numpy(feat), (H_up, W_up), bsize).numpy()
            key = f"m{mode}.{name}"
            out[f"{key}.out"] = y
            out[f"{key}.meta"] = np.array([mode, 10 + mode, B, H, W, H_up, W_up, -1 if bsize is None else bsize], dtype=np.int64)
            out[f"{key}.gains"] = np.array([wkw.get("k_gain", 1.0), wkw.get("q_gain", 1.0)])
            print(key, y.shape, float(np.abs(y).max()))
    np.savez_compressed(os.path.join(HERE, "modes.npz"), **out)
    print("wrote modes.npz")


if __name__ == "__main__":
    main()
