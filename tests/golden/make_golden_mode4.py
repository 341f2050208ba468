"""Golden fixtures for decoder mode 4, produced by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_mode4.py
Outputs of the unmodified ``ImplicitDecoder(mode=4, init_q=False)`` (/root/reference/src/models/components/
diinn.py:81-90,140-147: the mode-3 stack with a 3x3 reflect-padded last conv) on torch CPU fp32 for bit-reproducible
synthetic weights / features (diinn_b200.synth). The ``bsize`` cases pin that batched_step (diinn.py:149-160) convolves
every column strip on its own -- in mode 4 the reference's result depends on bsize.
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
    "batch_bsize": (2, 17, 23, 40, 51, {}, 700),          # strips of 17 columns: 17 | 17 | 17
    "strips_uneven": (1, 12, 10, 30, 26, {}, 240),         # strips of 8 columns: 8 | 8 | 8 | 2
    "tiny": (1, 3, 2, 2, 2, {}, None),                     # the smallest output reflect padding accepts
    "stress": (1, 24, 20, 60, 50, dict(k_gain=1.5, q_gain=10.0, last_gain=3.0), None),
}


def main():
    out = {}
    mode = 4
    for name, (B, H, W, H_up, W_up, wkw, bsize) in CASES.items():
        w = synth.make_weights(seed=mode, mode=mode, **wkw)
        dec = ImplicitDecoder(mode=mode, init_q=False).eval()
        dec.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w.items()}, strict=True)
        feat = synth.make_feat(10 + mode, B, H, W)
        y = dec(torch.from_numpy(feat), (H_up, W_up), bsize).numpy()
        key = f"m{mode}.{name}"
        out[f"{key}.out"] = y
        out[f"{key}.meta"] = np.array([mode, 10 + mode, B, H, W, H_up, W_up, -1 if bsize is None else bsize], dtype=np.int64)
        out[f"{key}.gains"] = np.array([wkw.get("k_gain", 1.0), wkw.get("q_gain", 1.0), wkw.get("last_gain", 1.0)])
        print(key, y.shape, float(np.abs(y).max()))
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "mode4.npz"), **out)
    print("wrote mode4.npz")


if __name__ == "__main__":
    main()
