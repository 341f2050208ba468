"""CPU, build container only: the oracle against the LIVE reference (imported from /root/reference) on random shapes and
every (mode, init_q) wiring -- a wider net than the committed golden vectors, which pin the same functions on fixed cases.
Skipped where the reference tree is absent (the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

from diinn_b200 import synth
from oracle import diinn_oracle as orc

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs /root/reference")


@pytest.fixture(scope="module")
def ref_decoder_cls():
    sys.path.insert(0, "/root/reference")
    try:
        from src.models.components.diinn import ImplicitDecoder
    finally:
        sys.path.remove("/root/reference")
    return ImplicitDecoder


def _random_shapes(n, seed, hi=400):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        H, W = int(rng.integers(1, hi)), int(rng.integers(1, hi))
        if rng.random() < 0.5:   # unrelated output size (any aspect, down-scaling included)
            yield H, W, int(rng.integers(1, 3 * hi)), int(rng.integers(1, 3 * hi))
        else:                    # one non-integer scale for both axes, as demo2.py uses the decoder
            s = rng.uniform(0.3, 12.0)
            yield H, W, max(1, int(round(H * s))), max(1, int(round(W * s)))


def test_coordinates_bit_exact_on_random_shapes(ref_decoder_cls):
    """_make_pos_encoding (diinn.py:94-110) == oracle.make_pos_encoding, bit for bit, on 150 random shapes"""
    dec = ref_decoder_cls(mode=3)
    for H, W, H_up, W_up in _random_shapes(150, seed=0):
        ref = dec._make_pos_encoding(torch.zeros(1, 1, H, W), (H_up, W_up)).numpy()[0]
        assert np.array_equal(ref, orc.make_pos_encoding(H, W, H_up, W_up)), (H, W, H_up, W_up)


@pytest.mark.parametrize("init_q", [False, True])
@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_forward_on_random_shapes_every_wiring(ref_decoder_cls, mode, init_q):
    """ImplicitDecoder(mode, init_q).forward == oracle.decoder_forward within fp32 rounding on random small shapes, with
    and without bsize (mode 4: bsize changes the reference's result, and the oracle's with it)"""
    torch.set_grad_enabled(False)
    w = synth.make_weights(seed=40 + mode, mode=mode, init_q=init_q)
    dec = ref_decoder_cls(mode=mode, init_q=init_q).eval()
    dec.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w.items()}, strict=True)
    rng = np.random.default_rng(100 + mode)
    for i, (H, W, H_up, W_up) in enumerate(_random_shapes(6, seed=mode + 10 * init_q, hi=24)):
        if mode == 4 and (H_up < 2 or W_up < 2):
            continue   # reflect padding needs two pixels
        B = 1 + i % 2
        feat = synth.make_feat(50 + i, B, H, W)
        bsize = None
        if i % 3 == 2:
            strip = int(rng.integers(2, max(3, W_up)))
            if W_up % strip != 1:   # torch rejects a one-column strip under reflect padding
                bsize = strip * H_up
        ref = dec(torch.from_numpy(feat), (H_up, W_up), bsize).numpy()
        got = orc.decoder_forward(w, feat, (H_up, W_up), mode=mode, bsize=bsize)
        assert got.shape == ref.shape
        assert float(np.abs(got - ref).max()) <= 2e-6, (mode, init_q, H, W, H_up, W_up, bsize)
    torch.set_grad_enabled(True)


def test_local_ensemble_on_random_queries(ref_decoder_cls):
    """LIIF.query_rgb (liif.py:59-127, unmodified, its imnet replaced by an adapter around ImplicitDecoder.step) ==
    oracle.query_ensemble on random feature-map shapes, random coordinates (borders and exact cell centres included) and
    random cell sizes"""
    torch.set_grad_enabled(False)
    sys.path.insert(0, "/root/reference")
    try:
        from src.models.components.liif import LIIF
    finally:
        sys.path.remove("/root/reference")

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

    w = synth.make_weights(seed=9, k_gain=1.5, q_gain=4.0)
    dec = ref_decoder_cls(mode=3, init_q=False).eval()
    dec.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w.items()}, strict=True)
    liif = LIIF().eval()
    liif.imnet = ImnetAdapter(dec)
    rng = np.random.default_rng(5)
    for i in range(5):
        B, H, W, Q = 1 + i % 2, int(rng.integers(2, 30)), int(rng.integers(2, 30)), 400
        feat = synth.make_feat(60 + i, B, H, W)
        coord, cell = synth.make_query(70 + i, B, Q, (2.0 / float(rng.integers(H, 6 * H)), 2.0 / float(rng.integers(W, 6 * W))))
        coord[:, :40, 0] = np.linspace(-1, 1, 40, dtype=np.float32)                     # the image border ...
        coord[:, 40:80, 1] = (-1 + (2 * np.arange(40) + 1) / 40).astype(np.float32)     # ... and cell centres of a 40-grid
        ref = liif.query_rgb(torch.from_numpy(feat), torch.from_numpy(coord), torch.from_numpy(cell)).numpy()
        got = orc.query_ensemble(w, feat, coord, cell)
        assert float(np.abs(got - ref).max()) <= 2e-6, (i, B, H, W)
    torch.set_grad_enabled(True)


def test_liif_proper_on_random_queries(ref_decoder_cls):
    """the UNMODIFIED reference LIIF (its own imnet, liif.py:9-127 + mlp.py) == oracle.liif_query_rgb on random feature-map
    shapes, random coordinates (borders / exact cell centres / out-of-range values included), random cells, with and
    without the local ensemble"""
    torch.set_grad_enabled(False)
    sys.path.insert(0, "/root/reference")
    try:
        from src.models.components.liif import LIIF
    finally:
        sys.path.remove("/root/reference")
    rng = np.random.default_rng(8)
    for i in range(6):
        ens = i % 2 == 0
        w = synth.make_liif_weights(20 + i, gain=1.0 + i)
        liif = LIIF(local_ensemble=ens).eval()
        liif.imnet.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in w.items()}, strict=True)
        B, H, W, Q = 1 + i % 2, int(rng.integers(1, 24)), int(rng.integers(1, 24)), 300
        feat = synth.make_feat(80 + i, B, H, W)
        coord, cell = synth.make_query(90 + i, B, Q, (2.0 / float(rng.integers(H, 6 * H)), 2.0 / float(rng.integers(W, 6 * W))))
        coord[:, :40, 0] = np.linspace(-1.05, 1.05, 40, dtype=np.float32)
        coord[:, 40:80, 1] = (-1 + (2 * np.arange(40) + 1) / 40).astype(np.float32)
        ref = liif.query_rgb(torch.from_numpy(feat), torch.from_numpy(coord), torch.from_numpy(cell)).numpy()
        got = orc.liif_query_rgb(w, feat, coord, cell, local_ensemble=ens)
        scale = max(1.0, float(np.abs(ref).max()))
        assert float(np.abs(got - ref).max()) <= 4e-6 * scale, (i, B, H, W, ens)
    torch.set_grad_enabled(True)
