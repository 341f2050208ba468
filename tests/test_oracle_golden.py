"""CPU: pin the oracle (oracle/diinn_oracle.py) against the fixtures produced by the reference itself
(tests/golden/make_golden.py ran /root/reference/src/models/components/diinn.py on torch CPU fp32)."""
import numpy as np
import pytest

import diinn_b200  # noqa: F401
from diinn_b200 import synth
from oracle import diinn_oracle as orc

POS_CASES = ["c1", "c2x2", "c2x3", "c2x4", "c3", "c4", "c5", "odd1", "odd2", "odd3", "down"]


@pytest.mark.parametrize("name", POS_CASES)
def test_index_and_rel_coords_bit_exact(golden_posenc, name):
    H, W, H_up, W_up = (int(v) for v in golden_posenc[f"{name}.shape"])
    ih, rh = orc.rel_axis(H, H_up)
    iw, rw = orc.rel_axis(W, W_up)
    assert np.array_equal(ih, golden_posenc[f"{name}.ih"])
    assert np.array_equal(iw, golden_posenc[f"{name}.iw"])
    # bit-exact: compare the raw fp32 bit patterns
    assert np.array_equal(rh.view(np.uint32), golden_posenc[f"{name}.rel_h"].view(np.uint32))
    assert np.array_equal(rw.view(np.uint32), golden_posenc[f"{name}.rel_w"].view(np.uint32))


def _case(golden_decoder, name):
    seed, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in golden_decoder[f"{name}.meta"])
    kg, qg = (float(v) for v in golden_decoder[f"{name}.gains"])
    weights = synth.make_weights(seed=seed, k_gain=kg, q_gain=qg)
    feat = synth.make_feat(fseed, B, H, W)
    return weights, feat, (H_up, W_up), golden_decoder[f"{name}.out"]


@pytest.mark.parametrize("name,tol", [("c1", 2e-6), ("c1_bsize", 2e-6), ("odd2", 2e-6), ("x1_batch", 2e-6),
                                      ("frac", 2e-6), ("stress", 2e-5)])
def test_decoder_forward_matches_reference(golden_decoder, name, tol):
    weights, feat, size, ref = _case(golden_decoder, name)
    out = orc.decoder_forward(weights, feat, size)
    assert out.shape == ref.shape and out.dtype == np.float32
    assert float(np.abs(out - ref).max()) <= tol


def test_fp64_oracle_brackets_reference(golden_decoder):
    weights, feat, size, ref = _case(golden_decoder, "c1")
    out64 = orc.decoder_forward(weights, feat, size, fp64=True)
    assert float(np.abs(out64 - ref).max()) <= 1e-6


def test_row_band_equals_full(golden_decoder):
    weights, feat, size, ref = _case(golden_decoder, "odd2")
    band = orc.decoder_forward(weights, feat, size, rows=(37, 61))
    assert float(np.abs(band - ref[:, :, 37:61]).max()) <= 2e-6


def test_layer_taps(golden_decoder):
    weights = synth.make_weights(seed=0)
    feat = synth.make_feat(1, 1, 48, 48)
    r0, r1, c0, c1 = (int(v) for v in golden_decoder["taps.rows"])
    ih, rh = orc.rel_axis(48, 192)
    iw, rw = orc.rel_axis(48, 192)
    u = orc.unfold3x3(feat)[0].transpose(1, 2, 0)
    x = u[ih[r0:r1]][:, iw[c0:c1]].reshape(-1, 576)
    syn = np.empty((r1 - r0, c1 - c0, 3), np.float32)
    syn[..., 0] = rh[r0:r1, None]
    syn[..., 1] = rw[None, c0:c1]
    syn[..., 2] = orc.ratio_value(48, 48, 192, 192)
    taps = {}
    out = orc.step_mode3(weights, x, syn.reshape(-1, 3), taps=taps)
    for key in ["k0", "q0", "k1", "q1", "k2", "q2", "k3", "q3"]:
        ref = golden_decoder[f"taps.{key}"][0].transpose(1, 2, 0).reshape(-1, 256)
        assert float(np.abs(taps[key] - ref).max()) <= 3e-6, key
    ref = golden_decoder["taps.out"][0].transpose(1, 2, 0).reshape(-1, 3)
    assert float(np.abs(out - ref).max()) <= 2e-6


def test_query_on_regular_grid_equals_forward(golden_decoder):
    """The (feat, coord, cell) superset entry reproduces forward() on the HR grid (integer and odd scales)."""
    for name in ["x1_batch", "frac"]:
        weights, feat, (H_up, W_up), ref = _case(golden_decoder, name)
        B, _, H, W = feat.shape
        ch, cw = orc.grid_coords(H_up, W_up)
        coord = np.stack(np.meshgrid(ch, cw, indexing="ij"), -1).reshape(1, -1, 2).repeat(B, 0)
        # indices must agree exactly with nearest-exact
        ih, _ = orc.query_index_rel(ch, H)
        iw, _ = orc.query_index_rel(cw, W)
        assert np.array_equal(ih, orc.nearest_exact_index(H, H_up))
        assert np.array_equal(iw, orc.nearest_exact_index(W, W_up))
        cell = np.empty_like(coord)
        cell[..., 0], cell[..., 1] = 2.0 / H_up, 2.0 / W_up
        out = orc.query(weights, feat, coord, cell)
        out = out.reshape(B, H_up, W_up, 3).transpose(0, 3, 1, 2)
        assert float(np.abs(out - ref).max()) <= 5e-6


def _ens_case(golden_decoder, name):
    seed, fseed, B, H, W, Q = (int(v) for v in golden_decoder[f"{name}.meta"])
    kg, qg = (float(v) for v in golden_decoder[f"{name}.gains"])
    weights = synth.make_weights(seed=seed, k_gain=kg, q_gain=qg)
    feat = synth.make_feat(fseed, B, H, W)
    coord = golden_decoder[f"{name}.coord"]
    cell = np.empty_like(coord)
    cell[..., 0], cell[..., 1] = (np.float32(v) for v in golden_decoder[f"{name}.cell"])
    return weights, feat, coord, cell, golden_decoder[f"{name}.out"]


@pytest.mark.parametrize("name,tol", [("ens", 2e-6), ("ens_stress", 2e-5)])
def test_local_ensemble_matches_reference_liif_machinery(golden_decoder, name, tol):
    """'next' row 1: LIIF.query_rgb's 4-neighbour lookup + area blend (run unmodified in make_golden.py with the DIINN
    step as its imnet) vs the oracle's restatement."""
    weights, feat, coord, cell, ref = _ens_case(golden_decoder, name)
    out = orc.query_ensemble(weights, feat, coord, cell)
    assert float(np.abs(out - ref).max()) <= tol


def test_torch_cpu_port_matches(golden_decoder):
    weights, feat, size, ref = _case(golden_decoder, "frac")
    out = orc.decoder_forward_torch_cpu(weights, feat, size)
    assert float(np.abs(out - ref).max()) <= 2e-6
    out_b = orc.decoder_forward_torch_cpu(weights, feat, size, bsize=300)
    assert float(np.abs(out_b - ref).max()) <= 2e-6


def test_synth_is_stable():
    """The synthetic generator is the other half of every golden pin: freeze a few values."""
    w = synth.make_weights(seed=0)
    assert w["K.1.0.weight"].shape == (256, 832, 1, 1) and w["last_layer.weight"].shape == (3, 256, 1, 1)
    f = synth.make_feat(1, 1, 4, 4)
    chk = float(np.float64(f.astype(np.float64).sum()))
    assert abs(float(f.std()) - 0.34) < 0.05
    assert chk == pytest.approx(float(synth.make_feat(1, 1, 4, 4).astype(np.float64).sum()), abs=0)


# ---- eval glue: fixtures produced by the reference's own calc_psnr / torchvision.save_image
def test_calc_psnr_matches_reference(golden_eval):
    g = golden_eval
    name = {0: None, 1: "benchmark", 2: "div2k"}
    for ds, sc, rr, want in zip(g["psnr.dataset"], g["psnr.scale"], g["psnr.rgb_range"], g["psnr.value"]):
        got = orc.calc_psnr(g["sr"], g["hr"], rgb_range=float(rr), dataset=name[int(ds)], scale=int(sc))
        assert abs(got - float(want)) <= 2e-4, (ds, sc, rr, got, want)   # reference sums in fp32, oracle in fp64
    got = orc.calc_psnr(g["sr"][:, :1], g["hr"][:, :1], dataset="benchmark", scale=3)
    assert abs(got - float(g["psnr.gray1"])) <= 2e-4


def test_denorm_clamp_and_u8_bit_exact(golden_eval):
    g = golden_eval
    den = orc.denorm_clamp(g["pred"], 0.5, 0.5)
    assert np.array_equal(den, g["denorm"])
    assert np.array_equal(orc.quantize_u8(den), g["u8"])


# ---- decoder modes 1 / 2: fixtures produced by the reference ImplicitDecoder(mode=1|2)
MODE_CASES = [f"m{m}.{n}" for m in (1, 2) for n in ("small", "c1", "batch_bsize", "stress")]


@pytest.mark.parametrize("key", MODE_CASES)
def test_modes_1_2_match_reference(golden_modes, key):
    mode, fseed, B, H, W, H_up, W_up, _ = (int(v) for v in golden_modes[f"{key}.meta"])
    kg, qg = (float(v) for v in golden_modes[f"{key}.gains"])
    w = synth.make_weights(seed=mode, mode=mode, k_gain=kg, q_gain=qg)
    got = orc.decoder_forward(w, synth.make_feat(fseed, B, H, W), (H_up, W_up), mode=mode)
    want = golden_modes[f"{key}.out"]
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= (2e-6 if kg == 1.0 else 2e-5)


# ---- decoder mode 4: fixtures produced by the reference ImplicitDecoder(mode=4); the bsize cases
# pin the per-strip reflect padding of batched_step (diinn.py:149-160)
MODE4_CASES = ["small", "c1", "batch_bsize", "strips_uneven", "tiny", "stress"]


def _mode4_case(g, name):
    mode, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in g[f"m4.{name}.meta"])
    kg, qg, lg = (float(v) for v in g[f"m4.{name}.gains"])
    w = synth.make_weights(seed=mode, mode=mode, k_gain=kg, q_gain=qg, last_gain=lg)
    return w, synth.make_feat(fseed, B, H, W), (H_up, W_up), (None if bsize < 0 else bsize), g[f"m4.{name}.out"]


@pytest.mark.parametrize("name", MODE4_CASES)
def test_mode4_matches_reference(golden_mode4, name):
    w, feat, size, bsize, want = _mode4_case(golden_mode4, name)
    assert w["last_layer.weight"].shape == (3, 256, 3, 3)
    got = orc.decoder_forward(w, feat, size, mode=4, bsize=bsize)
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= (2e-6 if name != "stress" else 5e-5)


def test_mode4_bsize_changes_the_result_and_bands_do_not(golden_mode4):
    """the strips are part of the reference's mode-4 semantics: ignoring bsize gives a different image near strip borders;
    a row band, by contrast, is a pure slice of the full result"""
    w, feat, size, bsize, want = _mode4_case(golden_mode4, "strips_uneven")
    whole = orc.decoder_forward(w, feat, size, mode=4)
    assert float(np.abs(whole - want).max()) > 1e-4
    strip = bsize // size[0]
    interior = [c for c in range(size[1]) if 0 < c % strip < strip - 1 and c != size[1] - 1]
    assert float(np.abs(whole[..., interior] - want[..., interior]).max()) <= 2e-6
    band = orc.decoder_forward(w, feat, size, rows=(7, 19), mode=4, bsize=bsize)
    assert np.array_equal(band, orc.decoder_forward(w, feat, size, mode=4, bsize=bsize)[:, :, 7:19])


# ---- init_q=True: fixtures produced by the reference ImplicitDecoder(mode=1..4, init_q=True)
INITQ_CASES = ([f"iq{m}.{n}" for n in ("small", "batch_bsize") for m in (1, 2, 3, 4)]
               + ["iq3.c1", "iq2.stress", "iq3.stress"])


def _initq_case(g, key):
    mode, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in g[f"{key}.meta"])
    kg, qg, fg = (float(v) for v in g[f"{key}.gains"])
    w = synth.make_weights(seed=20 + mode, mode=mode, init_q=True, k_gain=kg, q_gain=qg, first_gain=fg)
    return mode, w, synth.make_feat(fseed, B, H, W), (H_up, W_up), (None if bsize < 0 else bsize), g[f"{key}.out"]


@pytest.mark.parametrize("key", INITQ_CASES)
def test_init_q_matches_reference(golden_initq, key):
    mode, w, feat, size, bsize, want = _initq_case(golden_initq, key)
    assert w["Q.0.0.weight"].shape == (256, 576, 1, 1) and w["first_layer.0.weight"].shape == (576, 3, 1, 1)
    got = orc.decoder_forward(w, feat, size, mode=mode, bsize=bsize)
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= (2e-6 if not key.endswith("stress") else 2e-5)


def test_init_q_weights_leave_the_other_streams_alone():
    """synth appends first_layer last: every tensor shared with init_q=False keeps its values (Q.0 changes fan-in)"""
    a, b = synth.make_weights(seed=3, mode=3), synth.make_weights(seed=3, mode=3, init_q=True)
    assert set(b) - set(a) == {"first_layer.0.weight", "first_layer.0.bias"}
    assert all(np.array_equal(a[k], b[k]) for k in a if not k.startswith("Q.0.0."))


# ---- LIIF-proper decoding: the unmodified reference LIIF.query_rgb with its own imnet
@pytest.mark.parametrize("name,tol", [("sampled", 2e-6), ("sampled_gain", 2e-5), ("grid_x3", 2e-6), ("grid_odd", 5e-6)])
@pytest.mark.parametrize("ens", [1, 0])
def test_liif_query_rgb_matches_reference(golden_liif, name, tol, ens):
    from conftest import liif_case, LIIF_CASES
    import ast, os, re   # the generator and the tests must describe the same cases (read, not imported: it needs the reference)
    src = open(os.path.join(os.path.dirname(__file__), "golden", "make_golden_liif.py")).read()
    assert ast.literal_eval(re.search(r"^CASES = (\{.*?^\})", src, re.S | re.M).group(1)) == LIIF_CASES
    weights, feat, coord, cell, ref = liif_case(golden_liif, name)
    out = orc.liif_query_rgb(weights, feat, coord, cell, local_ensemble=bool(ens))
    assert out.shape == ref[ens].shape and out.dtype == np.float32
    assert float(np.abs(out - ref[ens]).max()) <= tol * max(1.0, float(np.abs(ref[ens]).max()))
    # the fp64 evaluation of the same restatement agrees too: the pin is not an artefact of matching rounding
    out64 = orc.liif_query_rgb(weights, feat, coord, cell, local_ensemble=bool(ens), fp64=True)
    assert float(np.abs(out64 - ref[ens]).max()) <= 4 * tol * max(1.0, float(np.abs(ref[ens]).max()))


def test_liif_grid_coordinates_bit_exact(golden_liif):
    """LIIF.make_coord_and_cell (liif.py:32-57), as stored by the generator, equals the oracle's axis centres bit for bit."""
    for name, (H_up, W_up) in (("grid_x3", (30, 39)), ("grid_odd", (16, 25))):
        c = golden_liif[f"{name}.coord"].reshape(H_up, W_up, 2)
        assert np.array_equal(c[:, 0, 0].view(np.uint32), orc.liif_make_coord(H_up).view(np.uint32))
        assert np.array_equal(c[0, :, 1].view(np.uint32), orc.liif_make_coord(W_up).view(np.uint32))
        ce = golden_liif[f"{name}.cell"]
        assert np.all(ce[:, 0] == np.float32(2 / H_up)) and np.all(ce[:, 1] == np.float32(2 / W_up))
