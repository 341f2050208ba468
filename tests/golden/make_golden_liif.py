"""Golden vectors for LIIF-proper decoding by RUNNING THE REFERENCE LIIF.query_rgb with its own imnet.

Build container only (needs /root/reference):   python tests/golden/make_golden_liif.py
The unmodified ``LIIF`` class (/root/reference/src/models/components/liif.py:9-158; feat_unfold=True, cell_decode=True,
local_ensemble True and False) is instantiated, its imnet (mlp.py) loaded with diinn_b200.synth.make_liif_weights, and
query_rgb / batched_predict run on torch CPU fp32 for sampled coordinates and for the regular grid of LIIF.forward
(make_coord_and_cell). Inputs are regenerated from seeds by the tests; only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from src.models.components.liif import LIIF  # noqa: E402  (the reference)
import diinn_b200  # noqa: E402,F401
from diinn_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)

# name -> (B, H, W, Q or (H_up, W_up), weight seed, gain, feature seed)
CASES = {
    "sampled": (2, 12, 17, 700, 5, 1.0, 21),
    "sampled_gain": (1, 9, 8, 500, 6, 4.0, 22),
    "grid_x3": (1, 10, 13, (30, 39), 5, 1.0, 23),
    "grid_odd": (2, 7, 9, (16, 25), 7, 2.0, 24),
}


def case_inputs(name):
    B, H, W, q, wseed, gain, fseed = CASES[name]
    weights = synth.make_liif_weights(wseed, gain)
    feat = synth.make_feat(fseed, B, H, W)
    return weights, feat, q


def main():
    out = {}
    for name in CASES:
        weights, feat, q = case_inputs(name)
        B = feat.shape[0]
        for ens in (True, False):
            model = LIIF(local_ensemble=ens).eval()
            model.imnet.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in weights.items()}, strict=True)
            x = torch.from_numpy(feat)
            if isinstance(q, tuple):
                coord, cell = model.make_coord_and_cell(x, q)
                pred = model.batched_predict(x, coord, cell, 257)      # ragged chunks on purpose
                out[f"{name}.coord"] = coord[0].numpy().copy()
                out[f"{name}.cell"] = cell[0].numpy().copy()
            else:
                c, ce = synth.make_query(31, B, q, cell_hw=(2.0 / 37, 2.0 / 53))
                # a few coordinates on and beyond the clamp / cell borders
                c[:, :4, 0] = np.float32([-1.0, 1.0, 0.0, -0.999999])
                c[:, :4, 1] = np.float32([1.0, -1.0, 0.0, 0.999999])
                pred = model.query_rgb(x, torch.from_numpy(c), torch.from_numpy(ce))
                out[f"{name}.coord"] = c
                out[f"{name}.cell"] = ce
            out[f"{name}.ens{int(ens)}"] = pred.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "liif.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
