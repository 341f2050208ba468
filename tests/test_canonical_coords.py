"""The one documented deviation of the 16-bit tensor paths from the reference's arithmetic, bounded on the CPU with the oracle:
on integer scale factors (<= 16 phases) the relative coordinates are the exact value of the pixel's phase, (2p + 1)/s - 1,
instead of the reference's two rounded fp32 grids (diinn.py:94-110) -- see oracle.canonical_rel_axis, stage_b_umma.cu
(Work::canon) and DESIGN.md section 4.1d. Gather indices are untouched. The fp32-precision paths never do this."""
import numpy as np
import pytest

from diinn_b200 import synth
from oracle import diinn_oracle as orc


@pytest.mark.parametrize("n_in,s", [(48, 4), (256, 2), (256, 3), (256, 4), (339, 4), (510, 4), (1000, 2), (37, 16)])
def test_distance_to_the_reference_coordinates(n_in, s):
    idx, ref = orc.rel_axis(n_in, n_in * s)
    idx_c, can = orc.canonical_rel_axis(n_in, n_in * s)
    assert np.array_equal(idx, idx_c) and np.array_equal(idx, np.arange(n_in * s) // s)   # nearest-exact IS floor(j / s)
    # the reference's value is the exact one plus the rounding of two coordinates in [-1, 1] (<= 2^-24 each, and of their
    # difference), scaled by n_in: a few ulps of 1.0 times the axis length
    assert float(np.abs(can - ref).max()) <= 4 * 2.0 ** -23 * n_in
    exact = (2 * (np.arange(n_in * s) % s) + 1) / s - 1.0
    assert float(np.abs(can - exact).max()) <= 2.0 ** -23                                   # the closed form is exact to an ulp
    assert float(np.abs(ref - exact).max()) >= float(np.abs(can - exact).max())             # ... and never further than the reference


@pytest.mark.parametrize("H,W,s,rows", [(24, 24, 4, None), (50, 40, 3, (20, 60)), (64, 64, 2, (10, 40)), (339, 510, 4, (1000, 1008))])
def test_effect_on_the_output(H, W, s, rows):
    """fp32 oracle with either set of coordinates. Default-init weights: nothing (<= 1e-8). Gain-scaled weights (K x3, Q x10,
    the set on which fp16 operands cost 2-4e-3): what the reference's own coordinate noise is worth there -- 2.5e-5 on a 24 x 24
    map, 8e-4 at the c3 geometry (the noise grows with the axis length) -- inside the 1e-2 contract of the 16-bit paths with
    room for their operand error, and one reason more why precision="auto" calibrates against the fp32 path."""
    feat = synth.make_feat(5, 1, H, W)
    size = (H * s, W * s)
    for gains, tol in (((1.0, 1.0), 1e-8), ((3.0, 10.0), 2e-3)):
        w = synth.make_weights(seed=0, k_gain=gains[0], q_gain=gains[1])
        ref = orc.decoder_forward(w, feat, size, rows=rows)
        can = orc.decoder_forward(w, feat, size, rows=rows, canonical_rel=True)
        assert float(np.abs(ref - can).max()) <= tol, (gains, float(np.abs(ref - can).max()))
