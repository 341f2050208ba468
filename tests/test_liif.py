"""LIIF-proper decoding (second half): ``FusedLIIFQuery`` = LIIF.query_rgb with its own imnet
(liif.py:59-127, mlp.py) on the DIINN kernels. CPU part: module tree / error behaviour. GPU part (pytest -m gpu): parity
against the oracle and against the golden vectors the unmodified reference produced (tests/golden/make_golden_liif.py)."""
import numpy as np
import pytest
import torch

import diinn_b200
from diinn_b200 import synth
from oracle import diinn_oracle as orc
from conftest import LIIF_CASES, liif_case

# max-abs against the reference's own output, relative to max(1, |out|_max): the fp32 tensor path at fp32 level, fp16 / bf16
# operands at their operand noise (measured on the B200: see the assertions' messages when they fire)
TOL = {"fp32": 2e-5, "fp16": 4e-3, "bf16": 3e-2}


def _module(weights, precision, ens=True):
    m = diinn_b200.FusedLIIFQuery(local_ensemble=ens, precision=precision)
    return diinn_b200.load_liif_imnet(m, {"imnet." + k: v for k, v in weights.items()})


# ---------------------------------------------------------------------------------------------------------------- CPU
def test_module_tree_mirrors_reference_imnet():
    m = diinn_b200.FusedLIIFQuery()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == {"imnet." + k: v for k, v in synth.LIIF_IMNET_SHAPES.items()}       # mlp.py:9-15 via liif.py:26
    w = synth.make_liif_weights(3)
    _module(w, "fp16")
    with pytest.raises(RuntimeError):
        diinn_b200.load_liif_imnet(m, {"imnet.layers.0.weight": np.zeros((256, 580), np.float32)})  # strict
    for kw in (dict(feat_unfold=False), dict(cell_decode=False)):
        with pytest.raises(NotImplementedError):
            diinn_b200.FusedLIIFQuery(**kw)
    with pytest.raises(ValueError):
        diinn_b200.FusedLIIFQuery(precision="fp32_simt")


def test_same_default_init_as_reference_mlp():
    """same layer sequence => same RNG consumption as MLP(580, 3, [256]*4) (mlp.py): seeded inits reproduce"""
    torch.manual_seed(11)
    a = diinn_b200.FusedLIIFQuery().state_dict()
    torch.manual_seed(11)
    lin = [torch.nn.Linear(i, o) for i, o in ((580, 256), (256, 256), (256, 256), (256, 256), (256, 3))]
    for j, l in enumerate(lin):
        assert torch.equal(a[f"imnet.layers.{2 * j}.weight"], l.weight) and torch.equal(a[f"imnet.layers.{2 * j}.bias"], l.bias)


def test_make_coord_matches_oracle_bit_exact():
    m = diinn_b200.FusedLIIFQuery()
    coord, cell = m.make_coord_and_cell(torch.zeros(2, 64, 3, 3), (13, 22))
    assert coord.shape == (2, 13 * 22, 2) and cell.shape == coord.shape
    c = coord[1].view(13, 22, 2).numpy()
    assert np.array_equal(c[:, 0, 0].view(np.uint32), orc.liif_make_coord(13).view(np.uint32))
    assert np.array_equal(c[0, :, 1].view(np.uint32), orc.liif_make_coord(22).view(np.uint32))
    assert float(cell[0, 0, 0]) == np.float32(2 / 13) and float(cell[0, 0, 1]) == np.float32(2 / 22)
    assert m.reshape_pred(torch.zeros(2, 13 * 22, 3), (13, 22)).shape == (2, 3, 13, 22)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    m = diinn_b200.FusedLIIFQuery()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m.query_rgb(torch.zeros(1, 64, 4, 4), torch.zeros(1, 5, 2), torch.zeros(1, 5, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.imnet(torch.zeros(1, 580))


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("ens", [1, 0])
@pytest.mark.parametrize("name", list(LIIF_CASES))
def test_query_rgb_matches_reference_golden(golden_liif, name, ens, precision):
    weights, feat, coord, cell, ref = liif_case(golden_liif, name)
    m = _module(weights, precision, bool(ens)).cuda()
    with torch.no_grad():
        out = m.query_rgb(torch.from_numpy(feat).cuda(), torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
    out = out.cpu().numpy()
    scale = max(1.0, float(np.abs(ref[ens]).max()))
    err = float(np.abs(out - ref[ens]).max())
    assert err <= TOL[precision] * scale, (name, ens, precision, err, scale)
    # and against the fp64 oracle (the golden file is fp32 arithmetic)
    o64 = orc.liif_query_rgb(weights, feat, coord, cell, local_ensemble=bool(ens), fp64=True)
    assert float(np.abs(out - o64).max()) <= TOL[precision] * scale


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_forward_grid_and_batched_predict(golden_liif, precision):
    """LIIF.forward from the encoder output on (liif.py:151-158): grid coordinates made on the device, ragged bsize chunks,
    reshape -- equal to the reference's batched_predict output, and bit-identical with and without chunking"""
    weights, feat, coord, cell, ref = liif_case(golden_liif, "grid_x3")
    m = _module(weights, precision, True).cuda()
    x = torch.from_numpy(feat).cuda()
    with torch.no_grad():
        img = m(x, (30, 39))
        img_b = m(x, (30, 39), bsize=257)
    assert img.shape == (1, 3, 30, 39) and torch.equal(img, img_b)
    want = np.transpose(ref[1].reshape(1, 30, 39, 3), (0, 3, 1, 2))
    assert float(np.abs(img.cpu().numpy() - want).max()) <= TOL[precision]


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_large_random_queries_and_bf16_io():
    """more queries than one wave of CTA pairs (ragged last tile), both lookups; bf16 feature maps in / bf16 out"""
    weights = synth.make_liif_weights(9)
    feat = synth.make_feat(41, 1, 24, 31)
    coord, cell = synth.make_query(43, 1, 148 * 128 + 77, cell_hw=(2.0 / 96, 2.0 / 124))
    for ens in (True, False):
        want = orc.liif_query_rgb(weights, feat, coord, cell, local_ensemble=ens)
        m = _module(weights, "fp16", ens).cuda()
        with torch.no_grad():
            out = m.query_rgb(torch.from_numpy(feat).cuda(), torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
            outb = m.query_rgb(torch.from_numpy(feat).cuda().bfloat16(), torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
        assert float(np.abs(out.cpu().numpy() - want).max()) <= TOL["fp16"]
        assert outb.dtype == torch.bfloat16
        assert float(np.abs(outb.float().cpu().numpy() - want).max()) <= 2e-2


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_handle_switches_between_liif_and_diinn_weights():
    """C ABI: a LIIF handle refuses the grid decode (LIIF.forward queries coordinates) and the CUDA-core path; loading DIINN
    weights again restores the decoder, bit-identical to a fresh handle"""
    import ctypes as C
    from diinn_b200 import _lib
    lib = _lib.load()
    dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="fp16"), synth.make_weights(0)).cuda()
    x = torch.from_numpy(synth.make_feat(3, 1, 12, 12)).cuda()
    with torch.no_grad():
        base = dec(x, (30, 30)).clone()
        lib_, h = dec._ensure_handle(x.device)
        lw = {k: torch.from_numpy(v).cuda() for k, v in synth.make_liif_weights(2).items()}
        w = _lib.LiifWeightsF32()
        for i in range(5):
            w.weight[i], w.bias[i] = lw[f"layers.{2 * i}.weight"].data_ptr(), lw[f"layers.{2 * i}.bias"].data_ptr()
        w.on_device = 1
        assert lib.diinn_set_weights_liif(h, C.byref(w), None) == 0
        with pytest.raises(_lib.DiinnError, match="LIIF"):
            dec(x, (30, 30))
        dec.precision = "fp32_simt"
        coord, cell = synth.make_query(5, 1, 50)
        with pytest.raises(_lib.DiinnError, match="tensor paths"):
            dec.query(x, torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
        dec.precision = "fp16"
        dec._packed_versions = None            # force diinn_set_weights
        assert torch.equal(dec(x, (30, 30)), base)
    bad = diinn_b200.FusedImplicitDecoder(mode=1).cuda()
    _, hb = bad._ensure_handle(x.device)
    assert lib.diinn_set_weights_liif(hb, C.byref(w), None) == -4      # hosted by a mode-3 handle only


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 1, 5), (2, 3, 1), (1, 2, 2)])
def test_tiny_feature_maps(shape):
    """degenerate LR grids (a single cell, single rows / columns): every shifted lookup clamps into the map"""
    B, H, W = shape
    weights = synth.make_liif_weights(4)
    feat = synth.make_feat(51, B, H, W)
    coord, cell = synth.make_query(53, B, 200, cell_hw=(2.0 / (3 * H), 2.0 / (3 * W)))
    coord[:, :2] = np.float32([[-1.0, -1.0], [1.0, 1.0]])
    for ens in (True, False):
        want = orc.liif_query_rgb(weights, feat, coord, cell, local_ensemble=ens)
        m = _module(weights, "fp32", ens).cuda()
        with torch.no_grad():
            out = m.query_rgb(torch.from_numpy(feat).cuda(), torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
        assert float(np.abs(out.cpu().numpy() - want).max()) <= TOL["fp32"], (shape, ens)
