"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C ABI (libdiinn_b200.so via
ctypes); the checker is the CPU oracle (oracle/diinn_oracle.py) and the golden vectors the reference itself produced
(tests/golden). Tolerances are BASELINE.json's: gather indices / relative coordinates bit-exact, fp32 path <= 1e-4
max-abs, 16-bit-operand paths <= 1e-2 max-abs and < 0.01 dB PSNR delta.

Precisions (decoder.py): "fp32" = the fp32-PRECISION tensor path (fp16 hi+lo split, three MMAs per product), "fp16" = fp16
operands (the default), "bf16" = bf16 operands, "fp32_simt" = exact fp32 FMA on CUDA cores (cross-check)."""
import os

import numpy as np
import pytest
import torch

import diinn_b200
from diinn_b200 import synth
from oracle import diinn_oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")]

TOL = {"fp32": 1e-4, "fp32_simt": 1e-4, "bf16": 1e-2, "fp16": 1e-2}
# regression guards well inside the contract (measured: fp32 paths ~5e-8, fp16 ~3e-6, bf16 ~3e-5 on the default-init weights)
TIGHT = {"fp32": 2e-6, "fp32_simt": 2e-6, "bf16": 2e-4, "fp16": 4e-5}
ALL = ["fp32", "fp32_simt", "bf16", "fp16"]
POS_CASES = ["c1", "c2x2", "c2x3", "c2x4", "c3", "c4", "c5", "odd1", "odd2", "odd3", "down"]
GOLDEN_CASES = ["c1", "c1_bsize", "odd2", "x1_batch", "frac"]


def _decoder(weights, precision):
    dec = diinn_b200.FusedImplicitDecoder(mode=3, init_q=False, precision=precision)
    return diinn_b200.load_numpy_weights(dec, weights).cuda()


def _case(golden_decoder, name):
    seed, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in golden_decoder[f"{name}.meta"])
    kg, qg = (float(v) for v in golden_decoder[f"{name}.gains"])
    weights = synth.make_weights(seed=seed, k_gain=kg, q_gain=qg)
    return weights, synth.make_feat(fseed, B, H, W), (H_up, W_up), golden_decoder[f"{name}.out"], (None if bsize < 0 else bsize)


def _psnr_delta(a, ref, seed=2):
    hr = synth.uniform(seed, 3, ref.shape, 0.0, 1.0)
    return abs(orc.calc_psnr(a, hr) - orc.calc_psnr(ref, hr))


@pytest.fixture(scope="module")
def w0():
    return synth.make_weights(seed=0)


@pytest.fixture(autouse=True)
def _no_grad():
    """the decoder is forward-only and raises when autograd could expect a backward pass (test_error_behaviour)"""
    with torch.no_grad():
        yield


# ---------------------------------------------------------------------------------------------------------
# tcgen05 plumbing
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("shape", [(256, 256, 64), (512, 512, 576)])
def test_umma_selftest(cg, shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(7)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    D = diinn_b200.FusedImplicitDecoder(mode=3).cuda().debug_umma_gemm(A, B, cta_group=cg)
    ref = A.float() @ B.float().t()
    assert float((D - ref).abs().max()) <= 1e-3


# ---------------------------------------------------------------------------------------------------------
# a1: indices and relative coordinates, bit-exact
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", POS_CASES)
def test_gather_bit_exact(golden_posenc, name):
    H, W, H_up, W_up = (int(v) for v in golden_posenc[f"{name}.shape"])
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    ih, iw, rh, rw = dec.debug_gather(H, W, H_up, W_up, "cuda")
    assert np.array_equal(ih.cpu().numpy(), golden_posenc[f"{name}.ih"])
    assert np.array_equal(iw.cpu().numpy(), golden_posenc[f"{name}.iw"])
    assert np.array_equal(rh.cpu().numpy().view(np.uint32), golden_posenc[f"{name}.rel_h"].view(np.uint32))
    assert np.array_equal(rw.cpu().numpy().view(np.uint32), golden_posenc[f"{name}.rel_w"].view(np.uint32))


@pytest.mark.parametrize("shape", [(48, 48, 192, 192), (37, 53, 100, 211), (339, 510, 1356, 2040)])
def test_query_gather_reproduces_grid(shape):
    """The (feat, coord, cell) entry fed with the HR grid's cell centres yields the nearest-exact indices bit for bit."""
    H, W, H_up, W_up = shape
    ch, cw = orc.grid_coords(H_up, W_up)
    coord = np.stack(np.meshgrid(ch, cw, indexing="ij"), -1).reshape(1, -1, 2)
    cell = np.empty_like(coord)
    cell[..., 0], cell[..., 1] = 2.0 / H_up, 2.0 / W_up
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    idx, rel, ratio = dec.debug_query_gather(H, W, torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
    ih, rh = orc.rel_axis(H, H_up)
    iw, rw = orc.rel_axis(W, W_up)
    ref_idx = (ih[:, None] * W + iw[None, :]).reshape(-1)
    assert np.array_equal(idx.cpu().numpy().reshape(-1), ref_idx)
    qi, qr = orc.query_index_rel(coord[0, :, 0], H)
    assert np.array_equal(rel.cpu().numpy()[0, :, 0].view(np.uint32), qr.view(np.uint32))
    assert abs(float(ratio[0, 0]) - float(orc.ratio_value(H, W, H_up, W_up))) <= 3e-7 * float(ratio[0, 0])  # one or two fp32 roundings


def test_query_gather_random_coords():
    H, W, B, Q = 48, 40, 3, 5000
    coord, cell = synth.make_query(3, B, Q)
    dec = diinn_b200.FusedImplicitDecoder(mode=3).cuda()
    idx, rel, ratio = dec.debug_query_gather(H, W, torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
    for b in range(B):
        ih, rh = orc.query_index_rel(coord[b, :, 0], H)
        iw, rw = orc.query_index_rel(coord[b, :, 1], W)
        assert np.array_equal(idx[b].cpu().numpy(), (ih * W + iw).astype(np.int32))
        assert np.array_equal(rel[b, :, 0].cpu().numpy().view(np.uint32), rh.view(np.uint32))
        assert np.array_equal(rel[b, :, 1].cpu().numpy().view(np.uint32), rw.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------
# a3/a4: stage A (hoisted LR-resolution pre-activations) and the full decode against the reference's outputs
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("fp32_simt", 2e-5), ("bf16", 2e-2), ("fp16", 3e-3)])
def test_stage_a_matches_oracle(w0, precision, tol):
    feat = synth.make_feat(9, 2, 21, 37)
    u = orc.unfold3x3(feat).transpose(0, 2, 3, 1).reshape(-1, 576).astype(np.float64)
    ref = np.empty((u.shape[0], 1024))
    ref[:, :256] = np.maximum(u @ w0["K.0.0.weight"].reshape(256, 576).T.astype(np.float64) + w0["K.0.0.bias"], 0)
    for i in range(1, 4):
        wk = w0[f"K.{i}.0.weight"].reshape(256, 832)[:, 256:].astype(np.float64)
        ref[:, 256 * i:256 * (i + 1)] = u @ wk.T + w0[f"K.{i}.0.bias"]
    P = _decoder(w0, precision).debug_stage_a(torch.from_numpy(feat).cuda()).cpu().numpy()
    assert float(np.abs(P - ref).max()) <= tol


@pytest.mark.parametrize("precision", ALL)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_decode_matches_reference_golden(golden_decoder, precision, name):
    weights, feat, size, ref, bsize = _case(golden_decoder, name)
    dec = _decoder(weights, precision)
    with torch.no_grad():
        out = dec(torch.from_numpy(feat).cuda(), list(size), bsize)
    assert out.shape == ref.shape and out.dtype == torch.float32 and out.is_contiguous()
    out = out.cpu().numpy()
    err = float(np.abs(out - ref).max())
    assert err <= TOL[precision], err
    assert err <= TIGHT[precision], err
    assert _psnr_delta(out, ref) < 0.01


@pytest.mark.parametrize("precision,abs_tol", [("fp32", 1e-4), ("fp32_simt", 1e-4), ("fp16", 1e-2), ("bf16", 2.5e-2)])
def test_decode_stress_weights(golden_decoder, precision, abs_tol):
    """Gain-scaled weights (K x3, Q x10: activations O(0.3..1) instead of being dominated by last_layer.bias, SURVEY section
    4 item 8), ABSOLUTE tolerances: the fp32-precision paths keep 1e-4 and the default 16-bit path (fp16 operands) keeps the
    1e-2 contract (measured 2e-3). bf16 operands do NOT (measured 1.6e-2): that is the documented envelope of "bf16"
    (include/diinn_b200.h) and the reason it is not the default."""
    weights, feat, size, ref, _ = _case(golden_decoder, "stress")
    with torch.no_grad():
        out = _decoder(weights, precision)(torch.from_numpy(feat).cuda(), size).cpu().numpy()
    err = float(np.abs(out - ref).max())
    assert err <= abs_tol, err
    if precision != "bf16":
        assert _psnr_delta(out, ref) < 0.01


def test_default_precision_is_the_one_that_meets_the_contract():
    assert diinn_b200.FusedImplicitDecoder(mode=3).precision == "fp16"


def test_siren_strength_weights_route_to_the_fp32_path():
    """SURVEY's stronger stress set (K x4, Q x30): a chaotic network -- the fp32 reference itself is 2.8e-4 away from an fp64
    evaluation (measured with the reference module in the build container) -- where no 16-bit operand format can hold 1e-2
    (fp16 0.2, bf16 1.1). precision="auto" detects that on a calibration crop and routes to the fp32-precision tensor path,
    which stays within a few 1e-3 of the fp64 oracle there (the tensor core adds every MMA's result to its fp32 accumulator
    with truncation, which this network amplifies ~1000x; the exact-FMA "fp32_simt" path lands where the reference does)."""
    w = synth.make_weights(seed=0, k_gain=4.0, q_gain=30.0)
    feat = synth.make_feat(4, 1, 12, 14)
    size = (47, 55)
    ref64 = orc.decoder_forward(w, feat, size, fp64=True)
    x = torch.from_numpy(feat).cuda()
    auto = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="auto"), w).cuda()
    with torch.no_grad():
        out = auto(x, size).cpu().numpy()
        out16 = _decoder(w, "fp16")(x, size).cpu().numpy()
    assert auto._auto_choice[0] == "fp32" and auto._auto_choice[1] > 5e-3
    assert float(np.abs(out - ref64).max()) <= 5e-3
    assert float(np.abs(out16 - ref64).max()) > 1e-2           # what the router avoided
    with torch.no_grad():
        simt = _decoder(w, "fp32_simt")(x, size).cpu().numpy()
    assert float(np.abs(simt - ref64).max()) <= 6e-4           # the reference's own distance to fp64 here is 2.8e-4
    # benign weights stay on the fast path
    w0_ = synth.make_weights(seed=0)
    auto0 = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="auto"), w0_).cuda()
    with torch.no_grad():
        o = auto0(x, size).cpu().numpy()
    assert auto0._auto_choice[0] == "fp16"
    assert float(np.abs(o - orc.decoder_forward(w0_, feat, size)).max()) <= TIGHT["fp16"]


def _canonical_rel(n_in, n_up):
    """Integer scale s = n_up / n_in: every pixel of phase p = j - s * floor(j / s) sits at (2p + 1)/s - 1 in exact arithmetic;
    the kernel uses fl(fl((2p + 1)/s) - 1) (stage_b_umma.cu: canon_rel)."""
    s = n_up // n_in
    p = np.arange(n_up) % s
    return ((2 * p + 1).astype(np.float32) / np.float32(s) + np.float32(-1.0)).astype(np.float32)


@pytest.mark.parametrize("name", ["c1", "c2x3", "c3", "odd1", "odd2", "odd3", "down"])
def test_fused_kernel_indices_and_coordinates_bit_exact(golden_posenc, w0, name):
    """(ih, iw, rel_h, rel_w) as make_row() INSIDE the fused stage-B kernel derives them for every output pixel
    (diinn_debug_set_tap), against the reference's _make_pos_encoding fixtures. Gather indices: bit for bit, always. Relative
    coordinates: bit for bit, except on INTEGER scale factors with <= 16 phases (c1, c2x3, c3), where the 16-bit paths use the
    exact closed form of the phase -- the value the reference's fp32 coordinate grids scatter around by their own rounding
    (measured 5e-5 on c3; bound: a few ulps of a [-1, 1] coordinate times n_in) -- so that layer 0's sines form a table."""
    H, W, H_up, W_up = (int(v) for v in golden_posenc[f"{name}.shape"])
    dec = _decoder(w0, "fp16")
    x = torch.from_numpy(synth.make_feat(3, 1, H, W)).cuda()
    ih, iw, rh, rw = (t.cpu().numpy() for t in dec.debug_rows(x, (H_up, W_up)))
    assert np.array_equal(ih, np.broadcast_to(golden_posenc[f"{name}.ih"][:, None], (H_up, W_up)))
    assert np.array_equal(iw, np.broadcast_to(golden_posenc[f"{name}.iw"][None, :], (H_up, W_up)))
    ref_h, ref_w = golden_posenc[f"{name}.rel_h"], golden_posenc[f"{name}.rel_w"]
    canon = H_up % H == 0 and W_up % W == 0 and (H_up // H) * (W_up // W) <= 16
    assert canon == (name in ("c1", "c2x3", "c3"))
    if canon:
        can_h, can_w = _canonical_rel(H, H_up), _canonical_rel(W, W_up)
        assert float(np.abs(can_h - ref_h).max()) <= 4 * 2.0 ** -23 * H and float(np.abs(can_w - ref_w).max()) <= 4 * 2.0 ** -23 * W
        ref_h, ref_w = can_h, can_w
    assert np.array_equal(rh.view(np.uint32), np.broadcast_to(ref_h.view(np.uint32)[:, None], (H_up, W_up)))
    assert np.array_equal(rw.view(np.uint32), np.broadcast_to(ref_w.view(np.uint32)[None, :], (H_up, W_up)))


def test_phase_table_is_bit_identical_to_computing_and_close_to_per_pixel_coordinates():
    """Integer scale factors: layer 0's sines come out of a per-CTA phase table. DIINN_NO_TAB=1 computes them per pixel from
    the same canonical coordinates and must not change a bit (row tiles whose patches need K_sel = 32 have no room for the
    table and compute); DIINN_NO_CANON=1 uses the reference's per-pixel fp32 coordinates (the kernel before the table) and
    may differ by the coordinate noise only. The switches are read once per process, hence subprocesses."""
    import subprocess
    import sys
    code = (
        "import sys, torch, numpy as np; sys.path.insert(0, %r)\n"
        "import diinn_b200; from diinn_b200 import synth\n"
        "outs = []\n"
        "for mode in (3, 2, 4):\n"
        "    w = synth.make_weights(seed=mode, mode=mode)\n"
        "    for (B, H, W, sh, sw) in ((1, 24, 24, 4, 4), (2, 13, 17, 2, 2), (1, 11, 9, 3, 3), (1, 10, 12, 2, 8), (1, 7, 9, 1, 1)):\n"
        "        x = torch.from_numpy(synth.make_feat(6, B, H, W)).cuda()\n"
        "        for prec in ('fp16', 'bf16'):\n"
        "            d = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=mode, precision=prec), w).cuda()\n"
        "            with torch.no_grad():\n"
        "                o = d(x, (H * sh, W * sw))\n"
        "                t = torch.cat([d.forward_rows(x, (H * sh, W * sw), a, b) for a, b in ((0, 5), (5, 6), (6, H * sh))], dim=2)\n"
        "            assert torch.equal(o, t), (mode, H, W, sh, sw, prec)\n"
        "            outs.append(o.float().cpu().numpy().ravel())\n"
        "np.save(sys.argv[1], np.concatenate(outs))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    got = {}
    with tempfile.TemporaryDirectory() as td:
        for leg, env in (("table", {}), ("no_tab", {"DIINN_NO_TAB": "1"}), ("no_canon", {"DIINN_NO_CANON": "1"})):
            path = os.path.join(td, leg + ".npy")
            r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True, text=True,
                               timeout=600)
            assert r.returncode == 0, (leg, r.stderr[-2000:])
            got[leg] = np.load(path)
    assert np.array_equal(got["table"].view(np.uint32), got["no_tab"].view(np.uint32))
    # (a coordinate that moves by 5e-5 can flip the 16-bit rounding of a q_0 value: bounded by the operand noise of the format)
    assert float(np.abs(got["table"] - got["no_canon"]).max()) <= 1e-4


def test_bf16_io(golden_decoder):
    """bf16 feature map in, bf16 image out (config c2 'bf16'): compare with the fp32 reference output."""
    weights, feat, size, ref, _ = _case(golden_decoder, "c1")
    with torch.no_grad():
        out = _decoder(weights, "bf16")(torch.from_numpy(feat).cuda().to(torch.bfloat16), size)
    assert out.dtype == torch.bfloat16
    out = out.float().cpu().numpy()
    assert float(np.abs(out - ref).max()) <= 1e-2
    assert _psnr_delta(out, ref) < 0.01


@pytest.mark.parametrize("precision", ALL)
def test_bsize_is_pure_scheduling(w0, precision):
    x = torch.from_numpy(synth.make_feat(5, 1, 20, 24)).cuda()
    dec = _decoder(w0, precision)
    with torch.no_grad():
        assert torch.equal(dec(x, (60, 96)), dec(x, (60, 96), 30000))


# ---------------------------------------------------------------------------------------------------------
# sharding: row tiles are bit-identical to the full decode
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ALL)
@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_tiles_bit_identical(w0, precision, world):
    x = torch.from_numpy(synth.make_feat(6, 2, 23, 31)).cuda()
    size = (91, 125)
    dec = _decoder(w0, precision)
    with torch.no_grad():
        full = dec(x, size)
        tiled = torch.empty_like(full)
        for r0, r1 in diinn_b200.row_partition(size[0], world):
            if r1 > r0:
                dec.forward_rows(x, size, r0, r1, out=tiled)
        parts = torch.cat([dec.forward_rows(x, size, r0, r1) for r0, r1 in diinn_b200.row_partition(size[0], world)
                           if r1 > r0], dim=2)
    assert torch.equal(full, tiled)
    assert torch.equal(full, parts)


def test_patch_shapes_and_stage_a_splits_are_bit_identical():
    """A launch may pick another stage-B patch shape (8x16 / 4x32 / 16x8 / 2x64) or split stage A's N-blocks over more work
    items to fill its last wave; neither may change a bit of the image. The shapes are pinned through the environment
    (read once per process, hence subprocesses), on a x4 decode (select-MMA variant) and a x2.3 decode (classic variant)."""
    import hashlib
    import subprocess
    import sys
    code = (
        "import sys, hashlib, torch; sys.path.insert(0, %r)\n"
        "import diinn_b200; from diinn_b200 import synth\n"
        "w = synth.make_weights(seed=0); h = hashlib.sha256()\n"
        "for (B, H, W, hu, wu) in ((2, 23, 31, 91, 125), (1, 16, 20, 37, 51)):\n"
        "    x = torch.from_numpy(synth.make_feat(6, B, H, W)).cuda()\n"
        "    for prec in ('fp16', 'bf16', 'fp32'):\n"
        "        d = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=prec), w).cuda()\n"
        "        with torch.no_grad():\n"
        "            h.update(d(x, (hu, wu)).cpu().numpy().tobytes())\n"
        "print('HASH', h.hexdigest())\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seen = {}
    for env in ({}, {"DIINN_PATCH_W_LOG2": "3"}, {"DIINN_PATCH_W_LOG2": "5"}, {"DIINN_PATCH_W_LOG2": "6"},
                {"DIINN_STAGE_A_NSPLIT": "1"}, {"DIINN_STAGE_A_NSPLIT": "2"}, {"DIINN_STAGE_A_NSPLIT": "4"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        seen[str(env)] = [ln for ln in r.stdout.splitlines() if ln.startswith("HASH")][0]
    assert len(set(seen.values())) == 1, seen


def test_decode_multi_writes_every_peer_buffer(w0):
    """diinn_decode_multi (fused multi-GPU assembly) on one GPU: two 'peer' image buffers both receive the row tile,
    bit-identical to a plain decode; rows outside the tile stay untouched."""
    x = torch.from_numpy(synth.make_feat(6, 1, 23, 31)).cuda()
    size = (91, 125)
    dec = _decoder(w0, "bf16")
    with torch.no_grad():
        full = dec(x, size)
        a = torch.full_like(full, -7.0)
        b = torch.full_like(full, -7.0)
        dec.forward_rows(x, size, 20, 61, out=a, peer_ptrs=[a.data_ptr(), b.data_ptr()])
    for buf in (a, b):
        assert torch.equal(buf[:, :, 20:61], full[:, :, 20:61])
        assert bool((buf[:, :, :20] == -7.0).all()) and bool((buf[:, :, 61:] == -7.0).all())


# ---------------------------------------------------------------------------------------------------------
# the (feat, coord, cell) superset entry
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ALL)
def test_query_random_coords_vs_oracle(w0, precision):
    """Config c5, sampled form: 16 patches of 48x48, 2304 sampled coords each."""
    B, H, W, Q = 16, 48, 48, 2304
    feat = synth.make_feat(8, B, H, W)
    coord, cell = synth.make_query(3, B, Q)
    ref = orc.query(w0, feat, coord, cell)
    with torch.no_grad():
        out = _decoder(w0, precision).query(torch.from_numpy(feat).cuda(), torch.from_numpy(coord).cuda(),
                                           torch.from_numpy(cell).cuda())
    assert out.shape == (B, Q, 3)
    err = float(np.abs(out.cpu().numpy() - ref).max())
    assert err <= TOL[precision] and err <= TIGHT[precision], err


@pytest.mark.parametrize("precision", ALL)
def test_query_on_grid_equals_forward(w0, precision):
    B, H, W, H_up, W_up = 2, 16, 20, 37, 51
    feat = synth.make_feat(4, B, H, W)
    ch, cw = orc.grid_coords(H_up, W_up)
    coord = np.stack(np.meshgrid(ch, cw, indexing="ij"), -1).reshape(1, -1, 2).repeat(B, 0)
    cell = np.empty_like(coord)
    cell[..., 0], cell[..., 1] = 2.0 / H_up, 2.0 / W_up
    dec = _decoder(w0, precision)
    x = torch.from_numpy(feat).cuda()
    with torch.no_grad():
        grid = dec(x, (H_up, W_up))
        q = dec.query(x, torch.from_numpy(coord).cuda(), torch.from_numpy(cell).cuda())
    q = q.reshape(B, H_up, W_up, 3).permute(0, 3, 1, 2)
    # same indices and relative coordinates; only `ratio` is formed differently (one fp32 rounding)
    assert float((grid - q).abs().max()) <= 1e-6


@pytest.mark.parametrize("precision", ALL)
@pytest.mark.parametrize("name", ["ens", "ens_stress"])
def test_local_ensemble_query(golden_decoder, precision, name):
    """4-neighbour local ensemble + area blend fused into the last epilogue, against the
    reference's own LIIF.query_rgb machinery (golden) wrapped around the DIINN step."""
    seed, fseed, B, H, W, Q = (int(v) for v in golden_decoder[f"{name}.meta"])
    kg, qg = (float(v) for v in golden_decoder[f"{name}.gains"])
    weights = synth.make_weights(seed=seed, k_gain=kg, q_gain=qg)
    feat = synth.make_feat(fseed, B, H, W)
    coord = golden_decoder[f"{name}.coord"]
    cell = np.empty_like(coord)
    cell[..., 0], cell[..., 1] = (np.float32(v) for v in golden_decoder[f"{name}.cell"])
    ref = golden_decoder[f"{name}.out"]
    with torch.no_grad():
        out = _decoder(weights, precision).query(torch.from_numpy(feat).cuda(), torch.from_numpy(coord).cuda(),
                                                 torch.from_numpy(cell).cuda(), local_ensemble=True).cpu().numpy()
    err = float(np.abs(out - ref).max())
    if name == "ens":
        assert err <= TIGHT[precision], err
    else:
        assert err <= {"fp32": 1e-4, "fp32_simt": 1e-4, "fp16": 1e-2}.get(precision, 5e-2 * float(np.abs(ref).max())), err


def test_c5_grid_form(w0):
    """Config c5, grid form: 16 patches 48x48 decoded on the x1 grid (2304 queries per patch, ratio = 1)."""
    B, H, W, H_up, W_up = synth.CONFIGS["c5"]
    feat = synth.make_feat(8, B, H, W)
    ref = orc.decoder_forward(w0, feat, (H_up, W_up))
    x = torch.from_numpy(feat).cuda()
    with torch.no_grad():
        for precision in ALL:
            out = _decoder(w0, precision)(x, (H_up, W_up)).cpu().numpy()
            err = float(np.abs(out - ref).max())
            assert err <= TIGHT[precision], (precision, err)


# ---------------------------------------------------------------------------------------------------------
# host-buffer entry (what bench.py's e2e times) and the call-site contract
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ALL)
def test_decode_host_equals_device(w0, precision):
    feat = torch.from_numpy(synth.make_feat(2, 1, 37, 53))
    dec = _decoder(w0, precision)
    with torch.no_grad():
        dev = dec(feat.cuda(), (100, 211)).cpu()
        host = dec.decode_host(feat.pin_memory(), (100, 211))
        band = dec.decode_host(feat.pin_memory(), (100, 211), 40, 77)
    assert torch.equal(dev, host)
    assert torch.equal(dev[:, :, 40:77], band)


def test_swap_decoder_call_site(w0):
    """DIINN.forward (diinn.py:16-19) / SRLitModule.forward (sr_module.py:104-105) stand-ins keep working after the
    decoder is swapped: model(lr, size) with a python list size, as demo2.py:40 calls it."""

    class Net(torch.nn.Module):
        def __init__(self, dec):
            super().__init__()
            self.encoder = torch.nn.Conv2d(3, 64, 3, padding=1)
            self.decoder = dec

        def forward(self, x, size, bsize=None):
            return self.decoder(self.encoder(x), size, bsize)

    class Lit(torch.nn.Module):
        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, x, size, eval_bsize=None):
            return self.net(x, size, eval_bsize)

    ref_like = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="fp32"), w0)
    model = Lit(Net(ref_like)).cuda()
    diinn_b200.swap_decoder(model, precision="bf16")
    lr = torch.rand(1, 3, 24, 24, device="cuda")
    with torch.no_grad():
        sr = model(lr, [96, 96])
        feat = model.net.encoder(lr)
    ref = orc.decoder_forward(w0, feat.cpu().numpy(), (96, 96))
    assert sr.shape == (1, 3, 96, 96)
    assert float(np.abs(sr.cpu().numpy() - ref).max()) <= 1e-2


def test_error_behaviour(w0):
    dec = _decoder(w0, "bf16")
    with torch.no_grad():
        with pytest.raises(ValueError):
            dec(torch.zeros(1, 32, 8, 8, device="cuda"), (16, 16))
        with pytest.raises(TypeError):
            dec(torch.zeros(1, 64, 8, 8, device="cuda", dtype=torch.float16), (16, 16))
        with pytest.raises(RuntimeError):
            dec(torch.zeros(1, 64, 8, 8), (16, 16))  # CPU tensor: no fallback
        with pytest.raises(diinn_b200._lib.DiinnError):
            dec.forward_rows(torch.zeros(1, 64, 8, 8, device="cuda"), (16, 16), 5, 3)
    with torch.enable_grad():
        x = torch.zeros(1, 64, 8, 8, device="cuda", requires_grad=True)
        with pytest.raises(RuntimeError, match="forward-only"):
            dec(x, (16, 16))
        # trainable parameters + autograd on: a frozen / detached encoder output must not slip through either
        with pytest.raises(RuntimeError, match="forward-only"):
            dec(torch.zeros(1, 64, 8, 8, device="cuda"), (16, 16))
        dec.requires_grad_(False)
        assert dec(torch.zeros(1, 64, 8, 8, device="cuda"), (16, 16)).shape == (1, 3, 16, 16)


def test_weight_update_is_picked_up(w0):
    dec = _decoder(w0, "fp32")
    x = torch.from_numpy(synth.make_feat(1, 1, 8, 8)).cuda()
    with torch.no_grad():
        a = dec(x, (16, 16)).clone()
        dec.last_layer.bias.add_(1.0)
        b = dec(x, (16, 16))
    assert float((b - a - 1.0).abs().max()) <= 1e-6


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: band oracle + cross-path + size-independent properties
# ---------------------------------------------------------------------------------------------------------
def _band_check(weights, feat, size, out_np, bands, tol):
    for r0, r1 in bands:
        ref = orc.decoder_forward(weights, feat, size, rows=(r0, r1))
        err = float(np.abs(out_np[:, :, r0:r1] - ref).max())
        assert err <= tol, (r0, r1, err)


@pytest.mark.parametrize("name", ["c2x2", "c2x3", "c2x4"])
def test_config_c2(w0, name):
    B, H, W, H_up, W_up = synth.CONFIGS[name]
    feat = synth.make_feat(1, B, H, W)
    x = torch.from_numpy(feat).cuda()
    with torch.no_grad():
        o32 = _decoder(w0, "fp32")(x, (H_up, W_up))
        o16 = _decoder(w0, "bf16")(x, (H_up, W_up))
        of16 = _decoder(w0, "fp16")(x, (H_up, W_up))
        o16b = _decoder(w0, "fp16")(x.to(torch.bfloat16), (H_up, W_up)).float()
    o32n = o32.cpu().numpy()
    assert _psnr_delta(o16.cpu().numpy(), o32n) < 0.01 and _psnr_delta(o16b.cpu().numpy(), o32n) < 0.01
    # every path against the ORACLE on eight 2-row bands spread over the image (first / last rows included)
    bands = [(r, r + 2) for r in np.linspace(0, H_up - 2, 8).astype(int)]
    _band_check(w0, feat, (H_up, W_up), o32n, bands, TIGHT["fp32"])
    _band_check(w0, feat, (H_up, W_up), of16.cpu().numpy(), bands, TIGHT["fp16"])
    _band_check(w0, feat, (H_up, W_up), o16.cpu().numpy(), bands, TIGHT["bf16"])
    _band_check(w0, feat, (H_up, W_up), o16b.cpu().numpy(), bands, 1e-2)


def test_config_c3_div2k(w0):
    B, H, W, H_up, W_up = synth.CONFIGS["c3"]
    feat = synth.make_feat(1, B, H, W)
    x = torch.from_numpy(feat).cuda()
    with torch.no_grad():
        dec16 = _decoder(w0, "fp16")
        o16 = dec16(x, (H_up, W_up))
        o32 = _decoder(w0, "fp32")(x, (H_up, W_up))
        obf = _decoder(w0, "bf16")(x, (H_up, W_up))
        # 8-rank row tiling (170,170,170,170,169,169,169,169) is bit-identical to the single decode
        tiled = torch.empty_like(o16)
        for r0, r1 in diinn_b200.row_partition(H_up, 8):
            dec16.forward_rows(x, (H_up, W_up), r0, r1, out=tiled)
        # and so are the tiles the sharded decodes actually use: boundaries on multiples of the scale factor (172 / 168 rows;
        # these take the 4x32 patch shape, one select MMA and the phase table, the unaligned ones above K_sel = 32 and no
        # table), for 2, 4 and 8 ranks and for bf16 operands
        aligned = torch.empty_like(o16)
        parts = diinn_b200.tile_partition(H, H_up, 8)
        assert [b - a for a, b in parts] == [172] * 3 + [168] * 5
        for world in (8, 4, 2):
            aligned.fill_(-7.0)
            for r0, r1 in diinn_b200.tile_partition(H, H_up, world):
                dec16.forward_rows(x, (H_up, W_up), r0, r1, out=aligned)
            assert torch.equal(o16, aligned), world
        dbf = _decoder(w0, "bf16")
        for r0, r1 in parts:
            dbf.forward_rows(x, (H_up, W_up), r0, r1, out=aligned)
        assert torch.equal(obf, aligned)
    assert torch.equal(o16, tiled)
    # host entry at full size: 4 pipelined row bands (upload / decode / download on three streams)
    host = dec16.decode_host(torch.from_numpy(feat).pin_memory(), (H_up, W_up))
    assert torch.equal(o16.cpu(), host)
    band = dec16.decode_host(torch.from_numpy(feat).pin_memory(), (H_up, W_up), 170, 1187)
    assert torch.equal(o16[:, :, 170:1187].cpu(), band)
    assert _psnr_delta(o16.cpu().numpy(), o32.cpu().numpy()) < 0.01
    # every path against the ORACLE: first / last rows and all seven boundaries of the 8-rank row partition
    bands = [(0, 1)] + [(r1 - 1, r1 + 1) for _, r1 in diinn_b200.row_partition(H_up, 8)[:-1]] + [(1355, 1356)]
    _band_check(w0, feat, (H_up, W_up), o32.cpu().numpy(), bands, TIGHT["fp32"])
    _band_check(w0, feat, (H_up, W_up), o16.cpu().numpy(), bands, TIGHT["fp16"])
    _band_check(w0, feat, (H_up, W_up), obf.cpu().numpy(), bands, TIGHT["bf16"])


def test_config_c4_8k(w0):
    B, H, W, H_up, W_up = synth.CONFIGS["c4"]
    feat = synth.make_feat(1, B, H, W)
    x = torch.from_numpy(feat).cuda()
    dec16 = _decoder(w0, "fp16")
    with torch.no_grad():
        o16 = dec16(x, (H_up, W_up))
        # shard invariance on two of the eight 540-row tiles
        for r0, r1 in (diinn_b200.row_partition(H_up, 8)[i] for i in (0, 5)):
            assert torch.equal(o16[:, :, r0:r1], dec16.forward_rows(x, (H_up, W_up), r0, r1))
        # the fp32-precision path as a FULL 8K decode
        o32 = _decoder(w0, "fp32")(x, (H_up, W_up))
    o16n = o16.cpu().numpy()
    assert np.isfinite(o16n).all()
    # against the ORACLE: first / last rows and all seven boundaries of the 8-rank row partition
    bands = [(0, 1)] + [(r1 - 1, r1 + 1) for _, r1 in diinn_b200.row_partition(H_up, 8)[:-1]] + [(4319, 4320)]
    _band_check(w0, feat, (H_up, W_up), o16n, bands, TIGHT["fp16"])
    _band_check(w0, feat, (H_up, W_up), o32.cpu().numpy(), bands, TIGHT["fp32"])


# ---------------------------------------------------------------------------------------------------------
# eval glue fused into the store and PSNR on the device
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_output_transform_denorm_clamp_u8(w0, precision):
    """decode + `(pred*div+sub).clamp_(0,1)` + save_image's uint8 quantisation in one kernel == the same three steps
    applied by the oracle to the plain decode of the SAME kernel (bit-exact: the glue is two rounded fp32 ops)."""
    wts = synth.make_weights(seed=0, k_gain=3.0, q_gain=10.0, last_gain=8.0)   # outputs spill over both clamp ends
    dec = _decoder(wts, precision)
    x = torch.from_numpy(synth.make_feat(4, 1, 24, 20)).cuda()
    size = (71, 63)
    plain = dec(x, size).cpu().numpy()
    assert plain.min() < -1.0 and plain.max() > 1.0
    dec.set_output_transform(sub=0.5, div=0.5, clamp=(0, 1))
    den = dec(x, size)
    assert den.dtype == torch.float32
    want = orc.denorm_clamp(plain, 0.5, 0.5)
    assert np.array_equal(den.cpu().numpy(), want)
    dec.set_output_transform(sub=0.5, div=0.5, clamp=(0, 1), uint8=True)
    u8 = dec(x, size)
    assert u8.dtype == torch.uint8 and u8.shape == (1, 3, 71, 63)
    assert np.array_equal(u8.cpu().numpy(), orc.quantize_u8(want))
    # row tiles into a full uint8 buffer, the host entry and the query entry take the same glue
    full = torch.zeros((1, 3, 71, 63), dtype=torch.uint8, device="cuda")
    dec.forward_rows(x, size, 0, 30, out=full)
    dec.forward_rows(x, size, 30, 71, out=full)
    assert torch.equal(full, u8)
    host = dec.decode_host(x.cpu().pin_memory(), size)
    assert host.dtype == torch.uint8 and torch.equal(host, u8.cpu())
    dec.set_output_transform()
    assert np.array_equal(dec(x, size).cpu().numpy(), plain)


def test_output_transform_golden(golden_eval):
    """the reference's own numbers: torch `(t*div+sub).clamp_(0,1)` and the PNG torchvision.save_image wrote"""
    g = golden_eval
    dec = _decoder(synth.make_weights(seed=0), "bf16")
    # a decoder whose output IS the fixture's `pred` is not constructible; check the glue through the oracle pin instead:
    assert np.array_equal(orc.quantize_u8(orc.denorm_clamp(g["pred"])), g["u8"])
    # and the device PSNR against the reference's calc_psnr on the fixture tensors
    sr, hr = torch.from_numpy(g["sr"]).cuda(), torch.from_numpy(g["hr"]).cuda()
    name = {0: None, 1: "benchmark", 2: "div2k"}
    for ds, sc, rr, want in zip(g["psnr.dataset"], g["psnr.scale"], g["psnr.rgb_range"], g["psnr.value"]):
        got = dec.calc_psnr(sr, hr, dataset=name[int(ds)], scale=int(sc), rgb_range=float(rr))
        assert abs(got - float(want)) <= 2e-4, (ds, sc, rr, got, want)
    got = dec.calc_psnr(sr[:, :1].contiguous(), hr[:, :1].contiguous(), dataset="benchmark", scale=3)
    assert abs(got - float(g["psnr.gray1"])) <= 2e-4
    got16 = dec.calc_psnr(sr.bfloat16(), hr.bfloat16())
    want16 = orc.calc_psnr(sr.bfloat16().float().cpu().numpy(), hr.bfloat16().float().cpu().numpy())
    assert abs(got16 - want16) <= 1e-6
    assert np.isnan(dec.calc_psnr(sr, hr, dataset="benchmark", scale=0))


def test_psnr_full_size_c3():
    """33 MB images: device PSNR == oracle PSNR (fp64 means on both sides)"""
    B, H, W, H_up, W_up = synth.CONFIGS["c3"]
    hr = synth.uniform(2, 3, (1, 3, H_up, W_up), 0.0, 1.0)
    sr = (hr + synth.uniform(5, 1, hr.shape, -0.01, 0.01)).astype(np.float32)
    dec = _decoder(synth.make_weights(seed=0), "bf16")
    for ds in (None, "div2k"):
        got = dec.calc_psnr(torch.from_numpy(sr).cuda(), torch.from_numpy(hr).cuda(), dataset=ds, scale=4)
        assert abs(got - orc.calc_psnr(sr, hr, dataset=ds, scale=4)) <= 1e-6


# ---------------------------------------------------------------------------------------------------------
# decoder modes 1 / 2: K chain evaluated per LR pixel, stage B with zero K rows
# ---------------------------------------------------------------------------------------------------------
MODE_CASES = [f"m{m}.{n}" for m in (1, 2) for n in ("small", "c1", "batch_bsize", "stress")]


def _mode_case(golden_modes, key):
    mode, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in golden_modes[f"{key}.meta"])
    kg, qg = (float(v) for v in golden_modes[f"{key}.gains"])
    w = synth.make_weights(seed=mode, mode=mode, k_gain=kg, q_gain=qg)
    dec = lambda precision: diinn_b200.load_numpy_weights(  # noqa: E731
        diinn_b200.FusedImplicitDecoder(mode=mode, init_q=False, precision=precision), w).cuda()
    return mode, w, dec, synth.make_feat(fseed, B, H, W), (H_up, W_up), golden_modes[f"{key}.out"], (None if bsize < 0 else bsize)


@pytest.mark.parametrize("precision", ALL)
@pytest.mark.parametrize("key", MODE_CASES)
def test_modes_1_2_match_reference(golden_modes, key, precision):
    mode, w, dec, feat, size, want, bsize = _mode_case(golden_modes, key)
    got = dec(precision)(torch.from_numpy(feat).cuda(), size, bsize).cpu().numpy()
    err = float(np.abs(got - want).max())
    stress = key.endswith("stress")
    assert err <= (2.5e-2 if stress and precision == "bf16" else TOL[precision]), (key, precision, err)
    if not stress:
        assert err <= TIGHT[precision], (key, precision, err)
        assert _psnr_delta(got, want) < 0.01


@pytest.mark.parametrize("mode", [1, 2])
def test_modes_1_2_row_tiles_and_query(golden_modes, mode):
    """row tiles bit-identical to the full decode, host entry == device entry, query on the grid == forward"""
    _, w, dec, feat, size, want, _ = _mode_case(golden_modes, f"m{mode}.small")
    d = dec("fp16")
    x = torch.from_numpy(feat).cuda()
    full = d(x, size)
    tiles = torch.cat([d.forward_rows(x, size, a, b) for a, b in ((0, 19), (19, 50), (50, size[0]))], dim=2)
    assert torch.equal(tiles, full)
    assert torch.equal(d.decode_host(x.cpu().pin_memory(), size), full.cpu())
    # (feat, coord, cell) entry on the regular grid
    H_up, W_up = size
    ch = torch.from_numpy(orc.axis_centres(H_up)).cuda()
    cw = torch.from_numpy(orc.axis_centres(W_up)).cuda()
    coord = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).reshape(1, -1, 2)
    cell = torch.tensor([2.0 / H_up, 2.0 / W_up], device="cuda").expand(1, coord.shape[1], 2)
    q = d.query(x, coord, cell).reshape(1, H_up, W_up, 3).permute(0, 3, 1, 2)
    assert float((q - full).abs().max()) <= 1e-6
    # fp32 paths against the fp64 oracle; the fp32-precision tensor path tiles bit-identically too
    ref64 = orc.decoder_forward(w, feat, size, fp64=True, mode=mode)
    for precision in ("fp32", "fp32_simt"):
        d32 = dec(precision)
        got32 = d32(x, size)
        assert float(np.abs(got32.cpu().numpy() - ref64).max()) <= 2e-6, precision
        assert torch.equal(torch.cat([d32.forward_rows(x, size, a, b) for a, b in ((0, 19), (19, size[0]))], dim=2), got32)


def test_mode_swap_decoder_keeps_mode():
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.decoder = diinn_b200.FusedImplicitDecoder(mode=1, precision="fp32")
    m = diinn_b200.swap_decoder(Net().cuda(), precision="bf16")
    assert m.decoder.mode == 1 and m.decoder.K[1][0].weight.shape[1] == 256
    x = torch.from_numpy(synth.make_feat(3, 1, 16, 16)).cuda()
    assert m.decoder(x, (32, 32)).shape == (1, 3, 32, 32)


def test_decode_is_cuda_graph_capturable(w0):
    """the C ABI promises no allocation / no sync inside decode: capture one decode, replay it on new feature values"""
    dec = _decoder(w0, "bf16")
    B, H, W, size = 1, 40, 36, (97, 120)
    x = torch.from_numpy(synth.make_feat(21, B, H, W)).cuda()
    out = torch.empty((B, 3, size[0], size[1]), device="cuda")
    dec.forward_rows(x, size, 0, size[0], out=out)          # warm-up: packs weights, sizes the cached workspace
    want_a = out.clone()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            dec.forward_rows(x, size, 0, size[0], out=out)
    torch.cuda.current_stream().wait_stream(side)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want_a)
    x.copy_(torch.from_numpy(synth.make_feat(22, B, H, W)))  # same buffers, new values
    g.replay()
    torch.cuda.synchronize()
    want_b = dec(x, size)
    assert torch.equal(out, want_b) and not torch.equal(want_a, want_b)


def test_two_handles_on_one_device_are_independent(w0):
    """no shared global state between handles: interleaved decodes of two decoders with different weights"""
    a, b = _decoder(w0, "bf16"), _decoder(synth.make_weights(seed=5), "bf16")
    x = torch.from_numpy(synth.make_feat(23, 1, 24, 24)).cuda()
    ya, yb = a(x, (48, 48)), b(x, (48, 48))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        ya2 = a(x, (48, 48))
    with torch.cuda.stream(s2):
        yb2 = b(x, (48, 48))
    torch.cuda.synchronize()
    assert torch.equal(ya, ya2) and torch.equal(yb, yb2) and not torch.equal(ya, yb)


def test_channels_last_bf16_features_are_read_in_place(w0):
    """encoder hand-off: a bf16 channels-last feature map IS stage A's TMA layout -- no layout
    pass, results bit-identical to the same values handed over as NCHW bf16"""
    dec = _decoder(w0, "bf16")
    x = torch.from_numpy(synth.make_feat(31, 2, 37, 45)).cuda().bfloat16()
    xcl = x.contiguous(memory_format=torch.channels_last)
    assert not xcl.is_contiguous() and dec._io_dtype(xcl) == 2 and dec._io_dtype(x) == 1
    size = (101, 150)
    dec(x, size)                               # first call also packs the weights
    n0 = dec.launch_count()
    want = dec(x, size)
    n1 = dec.launch_count()
    got = dec(xcl, size)
    n2 = dec.launch_count()
    assert got.dtype == torch.bfloat16 and torch.equal(got, want)
    assert (n2 - n1) == (n1 - n0) - 1          # one kernel fewer: the NCHW -> NHWC pass
    # ... and against the ORACLE (on the bf16-rounded feature values; the image is rounded to bf16 on the way out)
    ref = orc.decoder_forward(w0, x.float().cpu().numpy(), size)
    assert float(np.abs(got.float().cpu().numpy() - ref).max()) <= 1e-3
    got16 = _decoder(w0, "fp16")(xcl, size)     # fp16 stage B behind the in-place bf16 stage A
    assert float(np.abs(got16.float().cpu().numpy() - ref).max()) <= 1e-3
    assert torch.equal(dec.forward_rows(xcl, size, 33, 77), want[:, :, 33:77])
    coord, cell = (torch.from_numpy(v).cuda() for v in synth.make_query(3, 2, 500))
    assert torch.equal(dec.query(xcl, coord, cell), dec.query(x, coord, cell))
    # the fp32 CUDA-core path takes the tensor too (through a plain .contiguous())
    d32 = _decoder(w0, "fp32")
    assert torch.equal(d32(xcl, size), d32(x, size))


@pytest.mark.parametrize("shape", [(1, 1, 1, 1, 1), (1, 1, 1, 7, 5), (2, 2, 3, 1, 1), (1, 3, 2, 200, 3), (1, 17, 1, 3, 250),
                                   (3, 5, 5, 5, 5)])
def test_degenerate_shapes(w0, shape):
    """single-pixel feature maps / outputs, extreme aspect ratios, down-scaling: partial tiles everywhere"""
    B, H, W, H_up, W_up = shape
    feat = synth.make_feat(40 + H + W, B, H, W)
    want = orc.decoder_forward(w0, feat, (H_up, W_up))
    x = torch.from_numpy(feat).cuda()
    for precision in ("fp32", "bf16"):
        got = _decoder(w0, precision)(x, torch.Size((H_up, W_up))).cpu().numpy()
        assert got.shape == (B, 3, H_up, W_up)
        assert float(np.abs(got - want).max()) <= TIGHT[precision], (shape, precision)


# ---------------------------------------------------------------------------------------------------------
# decoder mode 4: mode-3 stack, 3x3 reflect-padded last conv over HR pixels (csrc/mode4.cu)
# ---------------------------------------------------------------------------------------------------------
MODE4_CASES = ["small", "c1", "batch_bsize", "strips_uneven", "tiny", "stress"]


def _mode4_case(g, name):
    mode, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in g[f"m4.{name}.meta"])
    kg, qg, lg = (float(v) for v in g[f"m4.{name}.gains"])
    w = synth.make_weights(seed=mode, mode=mode, k_gain=kg, q_gain=qg, last_gain=lg)
    dec = lambda precision: diinn_b200.load_numpy_weights(  # noqa: E731
        diinn_b200.FusedImplicitDecoder(mode=4, init_q=False, precision=precision), w).cuda()
    return w, dec, synth.make_feat(fseed, B, H, W), (H_up, W_up), (None if bsize < 0 else bsize), g[f"m4.{name}.out"]


@pytest.mark.parametrize("precision", ALL)
@pytest.mark.parametrize("name", MODE4_CASES)
def test_mode4_matches_reference(golden_mode4, name, precision):
    """against the outputs of the reference ImplicitDecoder(mode=4), bsize strips included"""
    w, dec, feat, size, bsize, want = _mode4_case(golden_mode4, name)
    got = dec(precision)(torch.from_numpy(feat).cuda(), size, bsize).cpu().numpy()
    assert got.shape == want.shape
    err = float(np.abs(got - want).max())
    assert err <= (2.5e-2 if name == "stress" and precision == "bf16" else TOL[precision]), (name, precision, err)
    if name != "stress":
        # (the 16-bit paths dump q_3 as bf16 for the tensor-core projection: bf16-level accuracy whatever the operands)
        assert err <= (TIGHT["bf16"] if precision == "fp16" else TIGHT[precision]), (name, precision, err)
        assert _psnr_delta(got, want) < 0.01


def test_mode4_row_tiles_host_entry_and_fp64(golden_mode4):
    """row tiles (halo rows recomputed per tile) and the banded host entry are bit-identical to the full decode, with and
    without bsize strips; the fp32 path against the fp64 oracle; bf16 I/O; the eval glue rides on the conv's store"""
    w, dec, feat, size, _, want = _mode4_case(golden_mode4, "small")
    x = torch.from_numpy(feat).cuda()
    H_up, W_up = size
    for precision in ("bf16", "fp16", "fp32", "fp32_simt"):
        d = dec(precision)
        for bsize in (None, H_up * 10):
            full = d(x, size, bsize)
            tiles = torch.cat([d.forward_rows(x, size, a, b, bsize=bsize)
                               for a, b in ((0, 1), (1, 19), (19, 50), (50, H_up - 1), (H_up - 1, H_up))], dim=2)
            assert torch.equal(tiles, full), (precision, bsize)
            ref = orc.decoder_forward(w, feat, size, mode=4, bsize=bsize)
            assert float(np.abs(full.cpu().numpy() - ref).max()) <= (TIGHT["bf16"] if precision == "fp16" else TIGHT[precision])
        assert torch.equal(d.decode_host(x.cpu().pin_memory(), size), d(x, size).cpu())
    ref64 = orc.decoder_forward(w, feat, size, fp64=True, mode=4)
    assert float(np.abs(dec("fp32")(x, size).cpu().numpy() - ref64).max()) <= 2e-6
    d = dec("bf16")
    y16 = d(x.bfloat16(), size)
    assert y16.dtype == torch.bfloat16 and float((y16.float().cpu() - torch.from_numpy(want)).abs().max()) <= 1e-3
    d.set_output_transform(sub=0.5, div=0.5, clamp=(0.0, 1.0), uint8=True)
    u8 = d(x, size)
    d.set_output_transform()
    assert u8.dtype == torch.uint8
    assert np.abs(u8.cpu().numpy().astype(np.int32) - orc.quantize_u8(orc.denorm_clamp(want)).astype(np.int32)).max() <= 1


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_mode4_host_entry_bands_and_row_tiles(precision):
    """Mode 4 also evaluates q_3 on one HR halo row each side of a band; at x4 a band edge on a multiple of 4 makes that halo
    row read an LR row the band itself does not. The banded host entry (3 bands from 2^17 px, 5 from 2^21) and host row tiles
    must upload those rows: every call on a FRESH handle (stale device copies of earlier calls would mask a missing row)."""
    w = synth.make_weights(seed=4, mode=4)
    mk = lambda: diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=4, precision=precision), w).cuda()  # noqa: E731
    H, W = 96, 160
    size = (4 * H, 4 * W)                       # 245 760 px: three bands with edges at rows 128, 256
    feat = synth.make_feat(41, 1, H, W)
    x = torch.from_numpy(feat).cuda()
    want = mk()(x, size).cpu()
    host = torch.from_numpy(feat).pin_memory()
    assert torch.equal(mk().decode_host(host, size), want)
    for r0, r1 in ((128, 256), (4, 380), (0, 8), (376, 384), (131, 133)):
        assert torch.equal(mk().decode_host(host, size, r0, r1), want[:, :, r0:r1]), (r0, r1)
    os.environ["DIINN_HOST_BANDS"] = "1,1,1,1,1,1,1"
    try:
        assert torch.equal(mk().decode_host(host, size), want)
    finally:
        del os.environ["DIINN_HOST_BANDS"]


def test_mode4_contract_errors(golden_mode4):
    w, dec, feat, size, _, _ = _mode4_case(golden_mode4, "small")
    d = dec("bf16")
    x = torch.from_numpy(feat).cuda()
    H_up, W_up = size
    with pytest.raises(diinn_b200._lib.DiinnError):          # strips of one column: reflect padding rejects them
        d(x, size, H_up)
    with pytest.raises(diinn_b200._lib.DiinnError):          # last strip one column wide (63 = 31 * 2 + 1)
        d(x, size, H_up * 2)
    with pytest.raises(diinn_b200._lib.DiinnError):          # bsize < H_up: the reference never terminates
        d(x, size, H_up - 1)
    with pytest.raises(diinn_b200._lib.DiinnError):          # 1-pixel-wide output
        d(x, (8, 1))
    coord, cell = (torch.from_numpy(v).cuda() for v in synth.make_query(3, 1, 64))
    with pytest.raises(NotImplementedError):                 # the 3x3 conv needs the HR grid
        d.query(x, coord, cell)
    assert d(x, size, None).shape == (1, 3, H_up, W_up)      # the handle is still usable, bsize reset to None


# ---------------------------------------------------------------------------------------------------------
# init_q=True: sine gate on the unfolded features, per-HR-pixel x-facing GEMMs (csrc/init_q.cu)
# ---------------------------------------------------------------------------------------------------------
INITQ_CASES = ([f"iq{m}.{n}" for n in ("small", "batch_bsize") for m in (1, 2, 3, 4)]
               + ["iq3.c1", "iq2.stress", "iq3.stress"])


def _initq_case(g, key):
    mode, fseed, B, H, W, H_up, W_up, bsize = (int(v) for v in g[f"{key}.meta"])
    kg, qg, fg = (float(v) for v in g[f"{key}.gains"])
    w = synth.make_weights(seed=20 + mode, mode=mode, init_q=True, k_gain=kg, q_gain=qg, first_gain=fg)
    dec = lambda precision: diinn_b200.load_numpy_weights(  # noqa: E731
        diinn_b200.FusedImplicitDecoder(mode=mode, init_q=True, precision=precision), w).cuda()
    return mode, w, dec, synth.make_feat(fseed, B, H, W), (H_up, W_up), (None if bsize < 0 else bsize), g[f"{key}.out"]


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("key", INITQ_CASES)
def test_init_q_matches_reference(golden_initq, key, precision):
    """against the outputs of the reference ImplicitDecoder(mode, init_q=True), every mode"""
    mode, w, dec, feat, size, bsize, want = _initq_case(golden_initq, key)
    got = dec(precision)(torch.from_numpy(feat).cuda(), size, bsize).cpu().numpy()
    assert got.shape == want.shape
    err = float(np.abs(got - want).max())
    assert err <= (2.5e-2 if key.endswith("stress") and precision != "fp32" else TOL[precision]), (key, precision, err)
    if not key.endswith("stress"):
        # (the gate and its per-pixel GEMMs take bf16 operands in both 16-bit modes)
        assert err <= (TIGHT["bf16"] if precision == "fp16" else TIGHT[precision]), (key, precision, err)
        assert _psnr_delta(got, want) < 0.01


@pytest.mark.parametrize("mode", [2, 3, 4])
def test_init_q_row_tiles_chunks_and_io(golden_initq, mode):
    """row tiles and the host entry are bit-identical to the full decode; a decode spanning several tensor-path chunks
    (c2x2-sized output) agrees with the fp32 path; bf16 / channels-last feature maps; fp16acc; the fp64 oracle"""
    _, w, dec, feat, size, _, want = _initq_case(golden_initq, f"iq{mode}.small")
    x = torch.from_numpy(feat).cuda()
    H_up = size[0]
    for precision in ("bf16", "fp32"):
        d = dec(precision)
        full = d(x, size)
        tiles = torch.cat([d.forward_rows(x, size, a, b) for a, b in ((0, 19), (19, 50), (50, H_up))], dim=2)
        assert torch.equal(tiles, full), precision
        assert torch.equal(d.decode_host(x.cpu().pin_memory(), size), full.cpu())
    ref64 = orc.decoder_forward(w, feat, size, fp64=True, mode=mode)
    assert float(np.abs(dec("fp32")(x, size).cpu().numpy() - ref64).max()) <= 2e-6
    assert float(np.abs(dec("fp16")(x, size).cpu().numpy() - want).max()) <= TIGHT["bf16"]
    d = dec("bf16")
    x16 = x.bfloat16()
    y16 = d(x16, size)
    assert y16.dtype == torch.bfloat16 and float((y16.float().cpu() - torch.from_numpy(want)).abs().max()) <= 1e-3
    assert torch.equal(d(x16.contiguous(memory_format=torch.channels_last), size), y16)
    # several chunks (tensor path: 37 888 pixels each; fp32 path: 32 768)
    big = torch.from_numpy(synth.make_feat(77, 1, 64, 64)).cuda()
    a, b = d(big, (300, 400)), dec("fp32")(big, (300, 400))
    assert float((a - b).abs().max()) <= TIGHT["bf16"]
    coord, cell = (torch.from_numpy(v).cuda() for v in synth.make_query(3, 1, 64))
    with pytest.raises(NotImplementedError):
        d.query(x, coord, cell)


# ---------------------------------------------------------------------------------------------------------
# the secondary wirings at the BASELINE full sizes (c3, c4): band oracles at the first / last rows and at shard boundaries
# ---------------------------------------------------------------------------------------------------------
def _wiring(mode, init_q, precision, seed=0):
    w = synth.make_weights(seed=seed, mode=mode, init_q=init_q)
    return w, diinn_b200.load_numpy_weights(
        diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q, precision=precision), w).cuda()


@pytest.mark.parametrize("mode,init_q", [(4, False), (3, True), (4, True)])
def test_config_c3_secondary_wirings(mode, init_q):
    """full c3 decode on the tensor path: bands against the oracle, the 8-rank row tiling bit-identical, and the fp32 path
    (band decodes) in agreement"""
    B, H, W, H_up, W_up = synth.CONFIGS["c3"]
    feat = synth.make_feat(1, B, H, W)
    x = torch.from_numpy(feat).cuda()
    size = (H_up, W_up)
    w, dec16 = _wiring(mode, init_q, "bf16")
    with torch.no_grad():
        o16 = dec16(x, size)
        tiled = torch.empty_like(o16)
        for r0, r1 in diinn_b200.row_partition(H_up, 8):
            dec16.forward_rows(x, size, r0, r1, out=tiled)
        assert torch.equal(o16, tiled)
        _, dec32 = _wiring(mode, init_q, "fp32")
        bands = [(0, 2), (169, 171), (677, 679), (1354, 1356)]
        o16n = o16.cpu().numpy()
        assert np.isfinite(o16n).all()
        for r0, r1 in bands:
            ref = orc.decoder_forward(w, feat, size, rows=(r0, r1), mode=mode)
            assert float(np.abs(o16n[:, :, r0:r1] - ref).max()) <= TIGHT["bf16"], (mode, init_q, r0, r1)
            o32 = dec32.forward_rows(x, size, r0, r1).cpu().numpy()
            assert float(np.abs(o32 - ref).max()) <= TIGHT["fp32"], (mode, init_q, r0, r1)


@pytest.mark.parametrize("mode,init_q", [(4, False), (3, True)])
def test_config_c4_secondary_wirings(mode, init_q):
    """8K x12: mode 4 as a full decode (20 GB q_3 dump + projections), init_q on the shard-boundary and border bands
    (a full init_q decode of 33 Mpx is 0.25 s of GPU time and adds nothing the bands do not cover)"""
    B, H, W, H_up, W_up = synth.CONFIGS["c4"]
    feat = synth.make_feat(1, B, H, W)
    x = torch.from_numpy(feat).cuda()
    size = (H_up, W_up)
    w, dec16 = _wiring(mode, init_q, "bf16")
    bands = [(0, 1), (539, 541), (4319, 4320)]
    with torch.no_grad():
        full = dec16(x, size) if mode == 4 else None
        for r0, r1 in bands:
            got = dec16.forward_rows(x, size, r0, r1)
            if full is not None:
                assert torch.equal(full[:, :, r0:r1], got)
            ref = orc.decoder_forward(w, feat, size, rows=(r0, r1), mode=mode)
            assert float(np.abs(got.cpu().numpy() - ref).max()) <= TIGHT["bf16"], (mode, init_q, r0, r1)
        if full is None:   # a multi-chunk init_q tile: 24 rows of 7 680 pixels = three 8-row chunks
            t = dec16.forward_rows(x, size, 532, 556)
            assert torch.equal(t[:, :, 7:9], dec16.forward_rows(x, size, 539, 541))
