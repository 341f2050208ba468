"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE decoder.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Inputs come from diinn_b200.synth (bit-reproducible anywhere), outputs from the unmodified reference
``ImplicitDecoder(mode=3, init_q=False)`` (/root/reference/src/models/components/diinn.py:39-173) on
torch CPU fp32. The reference itself ships no golden vectors so these are the pin.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from src.models.components.diinn import ImplicitDecoder  # noqa: E402  (the reference)
import diinn_b200  # noqa: E402,F401
from diinn_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)


def ref_decoder(weights):
    dec = ImplicitDecoder(mode=3, init_q=False).eval()
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    dec.load_state_dict(sd, strict=True)
    return dec


def posenc_case(H, W, H_up, W_up):
    dec = ImplicitDecoder(mode=3, init_q=False)
    # the reference only needs x.shape/x.device here; an expanded 1-element tensor keeps c4 cheap
    x = torch.zeros(1).expand(1, 1, H, W)
    rel = dec._make_pos_encoding(x, (H_up, W_up))            # (1,2,H_up,W_up)
    rel_h = rel[0, 0, :, 0].numpy().copy()
    rel_w = rel[0, 1, 0, :].numpy().copy()
    assert bool((rel[0, 0] == rel[0, 0, :, :1]).all()) and bool((rel[0, 1] == rel[0, 1, :1, :]).all())
    ih = F.interpolate(torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(1, 1, H, W),
                       size=(H_up, W_up), mode="nearest-exact")[0, 0, :, 0].long().numpy()
    iw = F.interpolate(torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(1, 1, H, W),
                       size=(H_up, W_up), mode="nearest-exact")[0, 0, 0, :].long().numpy()
    return dict(rel_h=rel_h, rel_w=rel_w, ih=ih.astype(np.int32), iw=iw.astype(np.int32))


def main():
    # ---- 1. coordinates / indices for all BASELINE shapes + odd, non-integer-scale shapes -------------
    shapes = {k: v[1:] for k, v in synth.CONFIGS.items()}
    shapes.update({"odd1": (48, 48, 151, 151), "odd2": (37, 53, 100, 211), "odd3": (7, 5, 23, 9),
                   "down": (48, 48, 31, 17)})
    pos = {}
    for name, (H, W, H_up, W_up) in shapes.items():
        for k, v in posenc_case(H, W, H_up, W_up).items():
            pos[f"{name}.{k}"] = v
        pos[f"{name}.shape"] = np.array([H, W, H_up, W_up], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "posenc.npz"), **pos)

    # ---- 2. decoder outputs -----------------------------------------------------------------------------
    out = {}
    cases = {
        # name: (weight kwargs, feat seed, B, H, W, H_up, W_up, bsize)
        "c1": (dict(seed=0), 1, 1, 48, 48, 192, 192, None),
        "c1_bsize": (dict(seed=0), 1, 1, 48, 48, 192, 192, 30000),
        "odd2": (dict(seed=0), 2, 1, 37, 53, 100, 211, None),
        "x1_batch": (dict(seed=0), 3, 3, 12, 12, 12, 12, None),
        "frac": (dict(seed=0), 4, 2, 16, 20, 37, 51, None),
        "stress": (dict(seed=5, k_gain=3.0, q_gain=10.0), 6, 1, 24, 24, 72, 96, None),
    }
    for name, (wkw, fseed, B, H, W, H_up, W_up, bsize) in cases.items():
        weights = synth.make_weights(**wkw)
        feat = synth.make_feat(fseed, B, H, W)
        dec = ref_decoder(weights)
        y = dec(torch.from_numpy(feat), [H_up, W_up], bsize).numpy()
        out[f"{name}.out"] = y.astype(np.float32)
        out[f"{name}.meta"] = np.array([wkw.get("seed", 0), fseed, B, H, W, H_up, W_up, -1 if bsize is None else bsize],
                                       dtype=np.int64)
        out[f"{name}.gains"] = np.array([wkw.get("k_gain", 1.0), wkw.get("q_gain", 1.0)], dtype=np.float64)
        print(name, y.shape, float(np.abs(y).max()))

    # ---- 3. per-layer taps of step() on a few pixels of c1 (diinn.py:132-139 re-traced on the reference
    #         module's own layers) ----------------------------------------------------------------------
    weights = synth.make_weights(seed=0)
    feat = torch.from_numpy(synth.make_feat(1, 1, 48, 48))
    dec = ref_decoder(weights)
    size = (192, 192)
    rel = dec._make_pos_encoding(feat, size)
    ratio = feat.new_tensor([(48 * 48) / (192 * 192)]).view(1, -1, 1, 1).expand(1, -1, *size)
    syn = torch.cat([rel, ratio], dim=1)
    x = F.interpolate(F.unfold(feat, 3, padding=1).view(1, 576, 48, 48), size=size, mode="nearest-exact")
    rows = slice(93, 95)
    xs, ss = x[:, :, rows, :64], syn[:, :, rows, :64]
    k = dec.K[0](xs)
    q = k * dec.Q[0](ss)
    out["taps.k0"], out["taps.q0"] = k.numpy(), q.numpy()
    for i in range(1, 4):
        k = dec.K[i](torch.cat([q, xs], dim=1))
        q = k * dec.Q[i](q)
        out[f"taps.k{i}"], out[f"taps.q{i}"] = k.numpy(), q.numpy()
    out["taps.out"] = dec.last_layer(q).numpy()
    out["taps.rows"] = np.array([93, 95, 0, 64], dtype=np.int64)

    # ---- 4. "next" row: LIIF's local-ensemble query machinery (liif.py:59-127, unmodified) around the reference DIINN
    #         step: LIIF.imnet is replaced by an adapter that feeds ImplicitDecoder.step -----------------------------
    from src.models.components.liif import LIIF

    class ImnetAdapter(torch.nn.Module):
        def __init__(self, dec):
            super().__init__()
            self.dec = dec

        def forward(self, inp):  # (N, 580) = [q_feat 576 | rel_coord 2 | rel_cell 2]  (liif.py:105-111)
            n = inp.shape[0]
            x = inp[:, :576].t().reshape(1, 576, n, 1)
            ratio = inp[:, 578] * inp[:, 579] * 0.25
            syn = torch.stack([inp[:, 576], inp[:, 577], ratio], 0).reshape(1, 3, n, 1)
            return self.dec.step(x, syn)[0, :, :, 0].t()

    for name, (wkw, fseed, B, H, W, Q, cellhw) in {
        "ens": (dict(seed=0), 12, 2, 24, 20, 3000, (2.0 / 53, 2.0 / 47)),
        "ens_stress": (dict(seed=5, k_gain=3.0, q_gain=10.0), 13, 1, 16, 16, 2000, (2.0 / 64, 2.0 / 64)),
    }.items():
        weights = synth.make_weights(**wkw)
        feat = synth.make_feat(fseed, B, H, W)
        coord, cell = synth.make_query(fseed + 1, B, Q, cellhw)
        # include exact grid centres and the image border among the queries
        coord[:, :50, 0] = np.linspace(-1, 1, 50, dtype=np.float32)
        coord[:, 50:100, 1] = np.linspace(-1, 1, 50, dtype=np.float32)
        liif = LIIF().eval()
        liif.imnet = ImnetAdapter(ref_decoder(weights))
        y = liif.query_rgb(torch.from_numpy(feat), torch.from_numpy(coord), torch.from_numpy(cell)).numpy()
        out[f"{name}.out"] = y.astype(np.float32)
        out[f"{name}.meta"] = np.array([wkw.get("seed", 0), fseed, B, H, W, Q], dtype=np.int64)
        out[f"{name}.gains"] = np.array([wkw.get("k_gain", 1.0), wkw.get("q_gain", 1.0)], dtype=np.float64)
        out[f"{name}.cell"] = np.array(cellhw, dtype=np.float64)
        out[f"{name}.coord"] = coord
        print(name, y.shape, float(np.abs(y).max()))

    np.savez_compressed(os.path.join(HERE, "decoder.npz"), **out)
    np.savez(os.path.join(HERE, "provenance.npz"),
             torch_version=np.array(torch.__version__), numpy_version=np.array(np.__version__))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
