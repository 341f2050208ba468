"""Golden fixtures for init_q=True, produced by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_initq.py
Outputs of the unmodified ``ImplicitDecoder(mode=1..4, init_q=True)`` (/root/reference/src/models/components/
diinn.py:48-51,113-115: first_layer = Conv2d(3,576,1)+sin gates the unfolded features, Q.0 reads the 576-wide gate) on
torch CPU fp32 for bit-reproducible synthetic weights / features (diinn_b200.synth).
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

CASES = {  # name: (modes, B, H, W, H_up, W_up, weight kwargs, bsize)
    "small": ((1, 2, 3, 4), 1, 24, 20, 71, 63, {}, None),
    "batch_bsize": ((1, 2, 3, 4), 2, 17, 23, 40, 51, {}, 700),
    "c1": ((3,), 1, 48, 48, 192, 192, {}, None),
    "stress": ((2, 3), 1, 24, 20, 60, 50, dict(k_gain=1.5, q_gain=6.0, first_gain=3.0), None),
}


def main():
    out = {}
    for name, (modes, B, H, W, H_up, W_up, wkw, bsize) in CASES.items():
        for mode in modes:
            w = synth.make_weights(seed=20 + mode, mode=mode, init_q=True, **wkw)
            dec = ImplicitDecoder(mode=mode, init_q=True).eval()
            dec.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w.items()}, strict=True)
            feat = synth.make_feat(30 + mode, B, H, W)
            y = dec(torch.from_numpy(feat), (H_up, W_up), bsize).numpy()
            key = f"iq{mode}.{name}"
            out[f"{key}.out"] = y
            out[f"{key}.meta"] = np.array([mode, 30 + mode, B, H, W, H_up, W_up, -1 if bsize is None else bsize], dtype=np.int64)
            out[f"{key}.gains"] = np.array([wkw.get("k_gain", 1.0), wkw.get("q_gain", 1.0), wkw.get("first_gain", 1.0)])
            print(key, y.shape, float(np.abs(y).max()), float(np.abs(y - w["last_layer.bias"].reshape(1, 3, 1, 1)).max()))
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "initq.npz"), **out)
    print("wrote initq.npz")


if __name__ == "__main__":
    main()
