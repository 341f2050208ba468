"""CPU: the C-ABI shared library loads and exports every symbol include/diinn_b200.h declares; the host-side mirror
of the reference interface keeps the reference's names, shapes and error behaviour. No compute calls (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import diinn_b200
from diinn_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def test_header_symbols_are_exported():
    _ensure_built()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    # product surface and debug taps / probes live in separate headers; both are exported by the one library
    for fname, symbols in (("diinn_b200.h", _lib.SYMBOLS), ("diinn_b200_debug.h", _lib.DEBUG_SYMBOLS)):
        header = open(os.path.join(ROOT, "include", fname)).read()
        declared = set(re.findall(r"\b(diinn_[a-z_0-9]+)\s*\(", header))
        declared -= {"diinn_status", "diinn_compute", "diinn_io_dtype"}
        assert declared == set(symbols), (fname, declared ^ set(symbols))
        for name in declared:
            assert hasattr(lib, name), name
    assert not [n for n in _lib.SYMBOLS if "debug" in n]      # nothing debug-flavoured on the product surface


def test_version_and_create_errors_without_gpu():
    _ensure_built()
    lib = _lib.load()
    assert b"sm_100a" in lib.diinn_version()
    h = ctypes.c_void_p()
    for bad in (_lib.Config(64, 256, 4, 5, 0, 0), _lib.Config(64, 256, 4, 3, 2, 0), _lib.Config(32, 256, 4, 3, 0, 0)):
        assert lib.diinn_create(ctypes.byref(h), ctypes.byref(bad)) == -4   # no such mode / init_q flag / width
        assert b"mode in {1,2,3,4}" in lib.diinn_last_error(None)
    assert lib.diinn_set_bsize(None, 0) == -1
    if not torch.cuda.is_available():
        ok = _lib.Config(64, 256, 4, 3, 0, 0)
        assert lib.diinn_create(ctypes.byref(h), ctypes.byref(ok)) == -8  # no CPU fallback
        assert not h.value
    assert lib.diinn_workspace_bytes(None, 1, 48, 48, 192, 192, 0, 192, 1) > 48 * 48 * 1024 * 4
    assert lib.diinn_workspace_bytes(None, 1, 48, 48, 192, 192, 10, 5, 1) == 0


def test_module_mirrors_reference_state_dict():
    dec = diinn_b200.FusedImplicitDecoder(mode=3, init_q=False)
    shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    assert shapes == synth.weight_shapes()
    assert sorted(shapes) == sorted(synth.weight_names())
    # strict round trip with reference-layout tensors
    w = synth.make_weights(seed=3)
    diinn_b200.load_numpy_weights(dec, w)
    back = dec.state_dict()
    for k, v in w.items():
        assert np.array_equal(back[k].numpy(), v)
    with pytest.raises(RuntimeError):
        dec.load_state_dict({k: torch.zeros(1) for k in shapes}, strict=True)


def test_same_default_init_as_reference_layout():
    """Same module tree => same RNG consumption as ImplicitDecoder(mode=3): seeded inits are reproducible."""
    torch.manual_seed(0)
    a = diinn_b200.FusedImplicitDecoder(mode=3).state_dict()
    torch.manual_seed(0)
    b = diinn_b200.FusedImplicitDecoder(mode=3).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert sum(v.numel() for v in a.values()) == 986627


def test_unsupported_wirings_raise():
    for kw in (dict(mode=5), dict(mode=0, init_q=True), dict(mode=3, in_channels=32), dict(mode=3, hidden_dims=[128] * 4)):
        with pytest.raises(NotImplementedError):
            diinn_b200.FusedImplicitDecoder(**kw)


@pytest.mark.parametrize("init_q", [False, True])
@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_every_wiring_mirrors_the_reference_module_tree(mode, init_q):
    """state_dict keys / shapes per (mode, init_q) as diinn.py:46-92 builds them (mode 1: K.i takes k alone; mode 4: 3x3
    last conv; init_q: first_layer and a 576-wide Q.0), first_layer registered first like the reference"""
    dec = diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q)
    shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    assert shapes == synth.weight_shapes(mode=mode, init_q=init_q)
    assert shapes["Q.0.0.weight"] == ((256, 576, 1, 1) if init_q else (256, 3, 1, 1))
    assert (list(shapes)[0] == "first_layer.0.weight") == init_q
    assert shapes["K.1.0.weight"] == ((256, 256, 1, 1) if mode == 1 else (256, 832, 1, 1))
    assert shapes["last_layer.weight"] == ((3, 256, 3, 3) if mode == 4 else (3, 256, 1, 1))
    if mode == 4:
        assert dec.last_layer.padding_mode == "reflect" and dec.last_layer.padding == (1, 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    dec = diinn_b200.FusedImplicitDecoder(mode=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(torch.zeros(1, 64, 4, 4), (8, 8))


def test_swap_decoder_keeps_call_site():
    class RefLikeDecoder(torch.nn.Module):  # parameter container with the reference's key layout
        def __init__(self):
            super().__init__()
            self.mode, self.init_q = 3, False
            inner = diinn_b200.FusedImplicitDecoder(mode=3)
            self.K, self.Q, self.last_layer = inner.K, inner.Q, inner.last_layer

    class Net(torch.nn.Module):  # stand-in for DIINN (diinn.py:8-19)
        def __init__(self):
            super().__init__()
            self.encoder = torch.nn.Identity()
            self.decoder = RefLikeDecoder()

        def forward(self, x, size, bsize=None):
            return self.decoder(self.encoder(x), size, bsize)

    class Lit(torch.nn.Module):  # stand-in for SRLitModule (sr_module.py:93,104-105)
        def __init__(self):
            super().__init__()
            self.net = Net()

        def forward(self, x, size, eval_bsize=None):
            return self.net(x, size, eval_bsize)

    m = Lit()
    before = {k: v.clone() for k, v in m.net.decoder.state_dict().items()}
    diinn_b200.swap_decoder(m, precision="fp32")
    assert isinstance(m.net.decoder, diinn_b200.FusedImplicitDecoder)
    after = m.net.decoder.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("mode,init_q", [(1, False), (3, False), (4, False), (2, True), (4, True)])
def test_swap_decoder_on_the_real_reference_module(mode, init_q):
    """the actual reference DIINN-style container with ImplicitDecoder(mode, init_q): swap_decoder keeps every parameter
    (strict state_dict round trip, first_layer and the 3x3 last conv included) and the constructor flags"""
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from src.models.components.diinn import ImplicitDecoder
    finally:
        sys.path.remove("/root/reference")

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = torch.nn.Identity()
            self.decoder = ImplicitDecoder(mode=mode, init_q=init_q)

    torch.manual_seed(7)
    net = Net()
    before = {k: v.clone() for k, v in net.decoder.state_dict().items()}
    diinn_b200.swap_decoder(net, precision="bf16")
    dec = net.decoder
    assert isinstance(dec, diinn_b200.FusedImplicitDecoder) and dec.mode == mode and dec.init_q == init_q
    after = dec.state_dict()
    assert list(after) == list(before)                       # same keys in the same order
    assert all(torch.equal(before[k], after[k]) for k in before)
    # same default init for the same seed: the module tree consumes the RNG like the reference's
    torch.manual_seed(7)
    fresh = diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q).state_dict()
    assert all(torch.equal(before[k], fresh[k]) for k in before)


def test_row_partition():
    parts = diinn_b200.row_partition(1356, 8)
    assert [b - a for a, b in parts] == [170, 170, 170, 170, 169, 169, 169, 169]
    assert parts[0][0] == 0 and parts[-1][1] == 1356
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert diinn_b200.row_partition(4320, 8) == [(540 * i, 540 * (i + 1)) for i in range(8)]
    assert diinn_b200.row_partition(3, 8)[3:] == [(3, 3)] * 5
    # aligned tiles: every boundary on a multiple of the (integer) scale factor, still contiguous, balanced in units
    al = diinn_b200.tile_partition(339, 1356, 8)
    assert [b - a for a, b in al] == [172, 172, 172, 168, 168, 168, 168, 168] and al[-1][1] == 1356
    assert all(a % 4 == 0 for a, _ in al) and all(al[i][1] == al[i + 1][0] for i in range(7))
    assert diinn_b200.tile_partition(360, 4320, 8) == [(540 * i, 540 * (i + 1)) for i in range(8)]   # x12: already aligned
    assert diinn_b200.tile_partition(16, 37, 3) == diinn_b200.row_partition(37, 3)                  # non-integer scale
    assert diinn_b200.tile_partition(4, 1356, 2) == diinn_b200.row_partition(1356, 2)               # x339: no alignment
    assert diinn_b200.row_partition(10, 4, 4) == [(0, 4), (4, 8), (8, 10), (10, 10)]
    sub, parts = diinn_b200.band_partition(1356, 8, 1, 4)
    assert sub == 172 and parts[0] == [(0, 172)] and parts[7] == [(1204, 1356)]


@pytest.mark.skipif(not __import__("shutil").which("gcc"), reason="needs gcc")
def test_header_is_c99_and_links_from_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/diinn_b200.h must compile as C99 (-pedantic) and link against the .so"""
    import subprocess
    _ensure_built()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-ldiinn_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr      # DIINN_OK with a B200, DIINN_ERR_UNSUPPORTED_DEVICE (-8) without
    assert "sm_100a" in r.stdout
