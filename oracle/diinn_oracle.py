"""CPU ORACLE for the DIINN query decoder -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module, and only as the checker (or as the timed CPU baseline). The shipped path
(``dual-interactive-implicit-neural-network_b200``) never imports it and has no CPU fallback.

What it restates (numpy, fp32 unless ``fp64=True``): the reference ``ImplicitDecoder`` with ``mode=3,
init_q=False`` -- /root/reference/src/models/components/diinn.py:39-173 -- whose arithmetic lives in
PyTorch/ATen (not vendored in the reference tree; the reference pins no torch version, README.md:59-61):

* ``nearest_exact_index``  <- F.interpolate(mode='nearest-exact'), ATen/native/UpSample.h
  ``nearest_exact_idx`` = min(floorf((dst + 0.5) * scale), in - 1), scale = (float)in / out.
* ``axis_centres`` / ``rel_axis`` / ``make_pos_encoding``  <- diinn.py:94-110.
* ``syn_input``            <- diinn.py:165-167.
* ``unfold3x3``            <- F.unfold(x, 3, padding=1).view(B, C*9, H, W), diinn.py:168.
* ``step_mode3``           <- diinn.py:132-139 with K/Q/last_layer built at diinn.py:73-80,92 (and the mode 1 / 2 / 4
                              wirings, diinn.py:116-131,140-147).
* ``last_conv3x3_reflect`` <- mode 4's last_layer, Conv2d(256, 3, 3, padding=1, padding_mode='reflect'), diinn.py:89-90.
* ``decoder_forward``      <- diinn.py:163-173 (bsize chunking, diinn.py:149-160, is pure scheduling).
* ``liif_query_rgb``       <- LIIF.query_rgb with its own imnet MLP(580,3,[256]*4), liif.py:59-127 + mlp.py:5-20.
* ``query``                <- the (feat, coord, cell) superset entry of SURVEY.md section 8(b); on a regular
                              grid it reproduces ``decoder_forward`` (tested).

PARITY PIN: the reference has no tests/golden vectors for this path, so the oracle is
pinned against outputs of the reference itself, produced in the build container by importing
/root/reference (tests/golden/make_golden.py) and committed under tests/golden/*.npz;
tests/test_oracle_golden.py re-checks every fixture on CPU.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------------
# a1: coordinates and nearest-exact indices (diinn.py:94-110)
# --------------------------------------------------------------------------------------------------
def nearest_exact_index(n_in: int, n_out: int) -> np.ndarray:
    """Source index of every destination index for mode='nearest-exact' along one axis.

    ATen: scale = (float)n_in / n_out; idx = min((int64)floorf((dst + 0.5) * scale), n_in - 1), the product
    being formed in double on CPU and rounded to float by floorf's argument conversion (identical to the
    CUDA kernel's float product because both factors have <= 24 significant bits)."""
    scale = F32(n_in) / F32(n_out)
    dst = np.arange(n_out, dtype=np.float64) + 0.5
    prod = (dst * np.float64(scale)).astype(F32)
    idx = np.floor(prod).astype(np.int64)
    return np.minimum(idx, n_in - 1)


def axis_centres(n: int) -> np.ndarray:
    """-1 + 1/n + 2/n * arange(n).float(): python-double scalars rounded to fp32, mul then add, no FMA."""
    a = F32(-1.0 + 1.0 / n)
    b = F32(2.0 / n)
    i = np.arange(n, dtype=F32)
    return (a + (b * i).astype(F32)).astype(F32)


def rel_axis(n_in: int, n_out: int):
    """(idx, rel) along one axis: rel[j] = fl(fl(c_up[j] - c_in[idx[j]]) * fl(n_in))  (diinn.py:106-108)."""
    idx = nearest_exact_index(n_in, n_out)
    c_in = axis_centres(n_in)
    c_up = axis_centres(n_out)
    rel = ((c_up - c_in[idx]).astype(F32) * F32(n_in)).astype(F32)
    return idx, rel


def canonical_rel_axis(n_in: int, n_out: int):
    """NOT the reference's arithmetic -- the yardstick for one documented deviation of the product. On an integer scale
    factor s = n_out / n_in the HR sample j = s*i + p sits at (2p + 1)/s - 1 of its source cell in exact arithmetic; the
    reference's rel (rel_axis above) is that value plus the rounding of its two fp32 coordinate grids (diinn.py:98-108). The
    16-bit tensor paths use the closed form, fl(fl((2p + 1)/s) - 1), so that Q.0's sines form a table (stage_b_umma.cu:
    canon_rel); tests bound its distance to the reference's values and its effect on the output."""
    assert n_out % n_in == 0
    s = n_out // n_in
    idx = nearest_exact_index(n_in, n_out)
    p = np.arange(n_out, dtype=np.int64) - idx * s
    rel = ((2 * p + 1).astype(F32) / F32(s) + F32(-1.0)).astype(F32)
    return idx, rel


def make_pos_encoding(H: int, W: int, H_up: int, W_up: int) -> np.ndarray:
    """(2, H_up, W_up) fp32: channel 0 = rel_h (varies along rows), channel 1 = rel_w."""
    _, rh = rel_axis(H, H_up)
    _, rw = rel_axis(W, W_up)
    out = np.empty((2, H_up, W_up), dtype=F32)
    out[0] = rh[:, None]
    out[1] = rw[None, :]
    return out


def ratio_value(H: int, W: int, H_up: int, W_up: int) -> np.float32:
    """x.new_tensor([(H*W)/(H_up*W_up)]): python double rounded to fp32 (diinn.py:166)."""
    return F32((H * W) / (H_up * W_up))


def syn_input(H: int, W: int, H_up: int, W_up: int) -> np.ndarray:
    """(3, H_up, W_up): [rel_h, rel_w, ratio] (diinn.py:165-167)."""
    out = np.empty((3, H_up, W_up), dtype=F32)
    out[:2] = make_pos_encoding(H, W, H_up, W_up)
    out[2] = ratio_value(H, W, H_up, W_up)
    return out


# --------------------------------------------------------------------------------------------------
# a3: 3x3 unfold (diinn.py:168)
# --------------------------------------------------------------------------------------------------
def unfold3x3(feat: np.ndarray) -> np.ndarray:
    """(B,C,H,W) -> (B,C*9,H,W); channel c*9 + kh*3 + kw = feat[c, h+kh-1, w+kw-1], zero outside."""
    B, C, H, W = feat.shape
    pad = np.zeros((B, C, H + 2, W + 2), dtype=feat.dtype)
    pad[:, :, 1:-1, 1:-1] = feat
    out = np.empty((B, C, 9, H, W), dtype=feat.dtype)
    for kh in range(3):
        for kw in range(3):
            out[:, :, kh * 3 + kw] = pad[:, :, kh:kh + H, kw:kw + W]
    return out.reshape(B, C * 9, H, W)


# --------------------------------------------------------------------------------------------------
# a4: the dual-interactive MLP, mode 3 (diinn.py:132-139)
# --------------------------------------------------------------------------------------------------
def _w2d(w: np.ndarray) -> np.ndarray:
    return w.reshape(w.shape[0], w.shape[1])


def step_mode3(weights: dict, x: np.ndarray, syn: np.ndarray, fp64: bool = False, taps: dict | None = None,
               mode: int = 3):
    """x: (N,576) gathered unfolded features, syn: (N,3) -> (N,3) RGB.

    k = relu(K0 x); q = k * sin(Q0 syn); for i in 1..3: k = relu(K_i [q,x]); q = k * sin(Q_i q);
    out = last(q).  1x1 Conv2d == per-pixel affine map.
    mode 2 (diinn.py:124-131): K_i takes [k,x] instead of [q,x]; mode 1 (diinn.py:116-123): K_i takes k alone.
    init_q=True is recognised by the presence of first_layer in `weights`."""
    dt = np.float64 if fp64 else F32
    x = x.astype(dt)
    syn = syn.astype(dt)
    W = {k: v.astype(dt) for k, v in weights.items()}
    n_layers = sum(1 for k in W if k.startswith("K.") and k.endswith("weight"))
    if "first_layer.0.weight" in W:   # init_q=True (diinn.py:48-51,113-115): sine gate on x, Q.0 reads the gate
        syn = np.sin(syn @ _w2d(W["first_layer.0.weight"]).T + W["first_layer.0.bias"])
        x = syn * x
    k = np.maximum(x @ _w2d(W["K.0.0.weight"]).T + W["K.0.0.bias"], 0)
    q = k * np.sin(syn @ _w2d(W["Q.0.0.weight"]).T + W["Q.0.0.bias"])
    if taps is not None:
        taps["k0"], taps["q0"] = k, q
    for i in range(1, n_layers):
        qx = k if mode == 1 else np.concatenate([k if mode == 2 else q, x], axis=1)
        k = np.maximum(qx @ _w2d(W[f"K.{i}.0.weight"]).T + W[f"K.{i}.0.bias"], 0)
        q = k * np.sin(q @ _w2d(W[f"Q.{i}.0.weight"]).T + W[f"Q.{i}.0.bias"])
        if taps is not None:
            taps[f"k{i}"], taps[f"q{i}"] = k, q
    if mode == 4:   # the 3x3 last conv couples HR pixels: the caller (decoder_forward) applies it on the assembled q_3
        return q.astype(dt)
    out = q @ _w2d(W["last_layer.weight"]).T + W["last_layer.bias"]
    return out.astype(dt)


def _reflect1(i: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """torch 'reflect' padding by one element on [lo, hi): lo-1 -> lo+1, hi -> hi-2."""
    return np.where(i < lo, 2 * lo - i, np.where(i >= hi, 2 * hi - 2 - i, i))


def last_conv3x3_reflect(weights: dict, q3: np.ndarray, fp64: bool = False, rows=None, row_base: int = 0,
                         H_up: int | None = None) -> np.ndarray:
    """mode 4 (diinn.py:89-90,146): q3 (B,R,W_up,256) -> (B,3,rows,W_up) through Conv2d(256,3,3,padding=1,
    padding_mode='reflect'): out[c,y,x] = b[c] + sum_{ky,kx,f} W[c,f,ky,kx] q3[refl(y+ky-1), refl(x+kx-1), f].

    q3 holds HR rows [row_base, row_base + R) of an image H_up rows high (default: the whole image) and `rows`=(r0,r1)
    selects the output rows; rows reflect at the IMAGE borders, so a band needs one halo row per side inside q3."""
    dt = np.float64 if fp64 else F32
    Wl = weights["last_layer.weight"].astype(dt)          # (3,256,3,3)
    B, R, Wd, _ = q3.shape
    H_up = R if H_up is None else H_up
    r0, r1 = (row_base, row_base + R) if rows is None else rows
    ys, xs = np.arange(r0, r1), np.arange(Wd)
    q3 = q3.astype(dt)
    out = np.zeros((B, 3, r1 - r0, Wd), dtype=dt)
    for ky in range(3):
        yy = _reflect1(ys + ky - 1, 0, H_up) - row_base
        for kx in range(3):
            xx = _reflect1(xs + kx - 1, 0, Wd)
            out += np.einsum("bhwf,cf->bchw", q3[:, yy][:, :, xx], Wl[:, :, ky, kx]).astype(dt)
    return (out + weights["last_layer.bias"].astype(dt).reshape(1, 3, 1, 1)).astype(dt)


# --------------------------------------------------------------------------------------------------
# forward (diinn.py:163-173) and the row-band form used for sharding / large configs
# --------------------------------------------------------------------------------------------------
def decoder_forward(weights: dict, feat: np.ndarray, size, rows=None, fp64: bool = False,
                    chunk: int = 1 << 16, mode: int = 3, bsize=None, canonical_rel: bool = False) -> np.ndarray:
    """(B,64,H,W), size=(H_up,W_up) -> (B,3,rows,W_up); rows=(r0,r1) restricts to an HR row band.
    canonical_rel (modes 1-3, integer scale factors): canonical_rel_axis instead of the reference's rel -- a deviation study.

    bsize matters in mode 4 only: batched_step (diinn.py:149-160) applies `step`, hence the reflect-padded 3x3 last
    conv, to column strips of bsize // H_up columns one at a time."""
    B, C, H, W = feat.shape
    H_up, W_up = int(size[0]), int(size[1])
    r0, r1 = (0, H_up) if rows is None else (int(rows[0]), int(rows[1]))
    if mode == 4:
        lo, hi = max(r0 - 1, 0), min(r1 + 1, H_up)                          # the band plus the rows its conv reads
        q3 = _q3_grid(weights, feat, (H_up, W_up), fp64, chunk, rows=(lo, hi))   # (B,hi-lo,W_up,256)
        strip = W_up if bsize is None else int(bsize) // H_up
        if strip < 1:
            raise ValueError("bsize < H_up: batched_step makes no progress (diinn.py:155)")
        outs = [last_conv3x3_reflect(weights, q3[:, :, a:a + strip], fp64=fp64, rows=(r0, r1), row_base=lo, H_up=H_up)
                for a in range(0, W_up, strip)]
        return np.concatenate(outs, axis=-1)
    ih, rh = (canonical_rel_axis if canonical_rel else rel_axis)(H, H_up)
    iw, rw = (canonical_rel_axis if canonical_rel else rel_axis)(W, W_up)
    ratio = ratio_value(H, W, H_up, W_up)
    u = unfold3x3(feat)                                   # (B,576,H,W)
    u = np.ascontiguousarray(u.transpose(0, 2, 3, 1))     # (B,H,W,576)
    nr = r1 - r0
    out = np.empty((B, 3, nr, W_up), dtype=np.float64 if fp64 else F32)
    rows_per_chunk = max(1, chunk // W_up)
    for b in range(B):
        for a in range(r0, r1, rows_per_chunk):
            e = min(a + rows_per_chunk, r1)
            x = u[b][ih[a:e]][:, iw].reshape(-1, C * 9)
            syn = np.empty((e - a, W_up, 3), dtype=F32)
            syn[..., 0] = rh[a:e, None]
            syn[..., 1] = rw[None, :]
            syn[..., 2] = ratio
            y = step_mode3(weights, x, syn.reshape(-1, 3), fp64=fp64, mode=mode)
            out[b, :, a - r0:e - r0] = y.reshape(e - a, W_up, 3).transpose(2, 0, 1)
    return out


def _q3_grid(weights: dict, feat: np.ndarray, size, fp64: bool, chunk: int, rows=None) -> np.ndarray:
    """mode 4: q_3 of the HR pixels of rows [rows[0], rows[1]) (default: all), (B,R,W_up,256) -- the input of the 3x3 last
    conv (diinn.py:141-146)."""
    B, C, H, W = feat.shape
    H_up, W_up = size
    g0, g1 = (0, H_up) if rows is None else rows
    ih, rh = rel_axis(H, H_up)
    iw, rw = rel_axis(W, W_up)
    ratio = ratio_value(H, W, H_up, W_up)
    u = np.ascontiguousarray(unfold3x3(feat).transpose(0, 2, 3, 1))
    n_hidden = weights["Q.3.0.weight"].shape[0]
    q3 = np.empty((B, g1 - g0, W_up, n_hidden), dtype=np.float64 if fp64 else F32)
    rows_per_chunk = max(1, chunk // W_up)
    for b in range(B):
        for a in range(g0, g1, rows_per_chunk):
            e = min(a + rows_per_chunk, g1)
            syn = np.empty((e - a, W_up, 3), dtype=F32)
            syn[..., 0] = rh[a:e, None]
            syn[..., 1] = rw[None, :]
            syn[..., 2] = ratio
            x = u[b][ih[a:e]][:, iw].reshape(-1, C * 9)
            q3[b, a - g0:e - g0] = step_mode3(weights, x, syn.reshape(-1, 3), fp64=fp64, mode=4).reshape(e - a, W_up, -1)
    return q3


# --------------------------------------------------------------------------------------------------
# superset entry: query(feat, coord, cell) with DIINN semantics
# --------------------------------------------------------------------------------------------------
def query_index_rel(coord_axis: np.ndarray, n: int):
    """Per-axis: idx = clamp(floor((c+1)*n/2), 0, n-1); rel = fl(fl(c - centre[idx]) * n).

    All in fp32: t = fl(fl(c + 1) * fl(n * 0.5)); idx = floorf(t)."""
    c = coord_axis.astype(F32)
    t = ((c + F32(1.0)).astype(F32) * F32(n * 0.5)).astype(F32)
    idx = np.clip(np.floor(t).astype(np.int64), 0, n - 1)
    centre = axis_centres(n)
    rel = ((c - centre[idx]).astype(F32) * F32(n)).astype(F32)
    return idx, rel


def query(weights: dict, feat: np.ndarray, coord: np.ndarray, cell: np.ndarray, fp64: bool = False,
          mode: int = 3) -> np.ndarray:
    """feat (B,64,H,W), coord (B,Q,2) as (h,w) in [-1,1], cell (B,Q,2) -> (B,Q,3).

    ratio = fl(fl(fl(cell_h * cell_w) * fl(H*W)) * 0.25) -- for a regular grid cell=(2/H_up, 2/W_up)
    this is (H*W)/(H_up*W_up) up to fp32 rounding."""
    B, C, H, W = feat.shape
    u = np.ascontiguousarray(unfold3x3(feat).transpose(0, 2, 3, 1))
    out = np.empty(coord.shape[:2] + (3,), dtype=np.float64 if fp64 else F32)
    for b in range(B):
        ih, rh = query_index_rel(coord[b, :, 0], H)
        iw, rw = query_index_rel(coord[b, :, 1], W)
        ce = cell[b].astype(F32)
        ratio = (((ce[:, 0] * ce[:, 1]).astype(F32) * F32(H * W)).astype(F32) * F32(0.25)).astype(F32)
        syn = np.stack([rh, rw, ratio], axis=1).astype(F32)
        out[b] = step_mode3(weights, u[b][ih, iw], syn, fp64=fp64, mode=mode)
    return out


# --------------------------------------------------------------------------------------------------
# "next" row 1: LIIF-style 4-neighbour local ensemble + area blend around the DIINN step.
# Restates LIIF.query_rgb, /root/reference/src/models/components/liif.py:59-127, with the imnet replaced by
# ImplicitDecoder.step (diinn.py:132-139) fed (rel_h, rel_w, ratio) instead of the concatenated 580-vector.
# --------------------------------------------------------------------------------------------------
def ensemble_index_rel(coord_axis: np.ndarray, n: int, v: int):
    """One axis, one shift direction v in {-1,+1} (liif.py:88-104):
      c_  = clamp(fl(c + fp32(v/n + 1e-6)), fp32(-1+1e-6), fp32(1-1e-6))           shifted lookup coordinate
      idx = nearbyint(fl(fl(fl(c_ + 1) * n) - 1) / 2)                               grid_sample(nearest, align_corners=False)
      rel = fl(fl(c - centre[idx]) * n)                                             relative to the ORIGINAL coordinate"""
    c = coord_axis.astype(F32)
    shift = F32(v * (2.0 / n / 2.0) + 1e-6)
    c_ = np.clip((c + shift).astype(F32), F32(-1 + 1e-6), F32(1 - 1e-6)).astype(F32)
    t = ((((c_ + F32(1.0)).astype(F32) * F32(n)).astype(F32) - F32(1.0)).astype(F32) * F32(0.5)).astype(F32)
    idx = np.clip(np.rint(t).astype(np.int64), 0, n - 1)
    rel = ((c - axis_centres(n)[idx]).astype(F32) * F32(n)).astype(F32)
    return idx, rel


def query_ensemble(weights: dict, feat: np.ndarray, coord: np.ndarray, cell: np.ndarray, fp64: bool = False):
    """feat (B,64,H,W), coord/cell (B,Q,2) -> (B,Q,3): sum_v pred_v * area_{3-v} / sum(area), v = (vx,vy) in
    [(-1,-1), (-1,1), (1,-1), (1,1)], area_v = |rel_h * rel_w| + 1e-9 (liif.py:117-127)."""
    B, C, H, W = feat.shape
    u = np.ascontiguousarray(unfold3x3(feat).transpose(0, 2, 3, 1))
    dt = np.float64 if fp64 else F32
    out = np.zeros(coord.shape[:2] + (3,), dtype=dt)
    for b in range(B):
        ce = cell[b].astype(F32)
        ratio = (((ce[:, 0] * ce[:, 1]).astype(F32) * F32(H * W)).astype(F32) * F32(0.25)).astype(F32)
        preds, areas = [], []
        for vx in (-1, 1):
            for vy in (-1, 1):
                ih, rh = ensemble_index_rel(coord[b, :, 0], H, vx)
                iw, rw = ensemble_index_rel(coord[b, :, 1], W, vy)
                syn = np.stack([rh, rw, ratio], axis=1).astype(F32)
                preds.append(step_mode3(weights, u[b][ih, iw], syn, fp64=fp64))
                areas.append((np.abs((rh * rw).astype(F32)) + F32(1e-9)).astype(F32))
        tot = ((areas[0] + areas[1]).astype(F32) + areas[2]).astype(F32) + areas[3]
        areas = [areas[3], areas[2], areas[1], areas[0]]
        acc = np.zeros_like(preds[0])
        for p_, a_ in zip(preds, areas):
            acc = acc + p_ * (a_ / tot).astype(dt)[:, None]
        out[b] = acc
    return out


# --------------------------------------------------------------------------------------------------
# "next" row 1, second half: LIIF-proper decoding -- LIIF.query_rgb with its own imnet = MLP(580, 3, [256]*4)
# (/root/reference/src/models/components/liif.py:59-127, mlp.py:5-20), feat_unfold=True, cell_decode=True.
# Pinned by tests/golden/liif.npz (the unmodified reference run by tests/golden/make_golden_liif.py).
# --------------------------------------------------------------------------------------------------
def liif_imnet(weights: dict, inp: np.ndarray, fp64: bool = False) -> np.ndarray:
    """MLP.forward (mlp.py:17-20): Linear + ReLU four times, then Linear. weights: 'layers.{0,2,4,6,8}.{weight,bias}'."""
    dt = np.float64 if fp64 else F32
    x = inp.astype(dt)
    for i in (0, 2, 4, 6):
        x = np.maximum(x @ weights[f"layers.{i}.weight"].astype(dt).T + weights[f"layers.{i}.bias"].astype(dt), 0)
    return (x @ weights["layers.8.weight"].astype(dt).T + weights["layers.8.bias"].astype(dt)).astype(dt)


def liif_query_rgb(weights: dict, feat: np.ndarray, coord: np.ndarray, cell: np.ndarray, local_ensemble: bool = True,
                   fp64: bool = False) -> np.ndarray:
    """feat (B,64,H,W), coord / cell (B,Q,2) as (h,w) -> (B,Q,3)  (liif.py:59-127).

    Per shift (vx, vy) in [(-1,-1), (-1,1), (1,-1), (1,1)] (one unshifted pass without the ensemble, liif.py:71-77):
      lookup index / rel_coord as ``ensemble_index_rel`` (liif.py:88-104; v = 0: no shift and eps_shift = 0, same clamp),
      rel_cell = cell * (H, W) in fp32 (liif.py:107-110), inp = [unfold(feat)[idx] | rel_coord | rel_cell] (580 wide),
      pred = imnet(inp), area = |rel_h * rel_w| + 1e-9; the areas are swapped diagonally and normalised (liif.py:117-127)."""
    B, C, H, W = feat.shape
    u = np.ascontiguousarray(unfold3x3(feat).transpose(0, 2, 3, 1))
    dt = np.float64 if fp64 else F32
    out = np.zeros(coord.shape[:2] + (3,), dtype=dt)
    shifts = [(-1, -1), (-1, 1), (1, -1), (1, 1)] if local_ensemble else [(0, 0)]
    for b in range(B):
        ce = cell[b].astype(F32)
        rel_cell = np.stack([(ce[:, 0] * F32(H)).astype(F32), (ce[:, 1] * F32(W)).astype(F32)], axis=1)
        preds, areas = [], []
        for vx, vy in shifts:
            ih, rh = _liif_index_rel(coord[b, :, 0], H, vx)
            iw, rw = _liif_index_rel(coord[b, :, 1], W, vy)
            inp = np.concatenate([u[b][ih, iw], np.stack([rh, rw], axis=1), rel_cell], axis=1).astype(F32)
            preds.append(liif_imnet(weights, inp, fp64=fp64))
            areas.append((np.abs((rh * rw).astype(F32)) + F32(1e-9)).astype(F32))
        tot = areas[0]
        for a_ in areas[1:]:
            tot = (tot + a_).astype(F32)
        if local_ensemble:
            areas = [areas[3], areas[2], areas[1], areas[0]]
        acc = np.zeros_like(preds[0])
        for p_, a_ in zip(preds, areas):
            acc = acc + p_ * (a_ / tot).astype(dt)[:, None]
        out[b] = acc
    return out


def _liif_index_rel(coord_axis: np.ndarray, n: int, v: int):
    """ensemble_index_rel for v in {-1, 0, +1}; v = 0 is local_ensemble=False: shift 0 AND eps_shift 0 (liif.py:76)."""
    if v != 0:
        return ensemble_index_rel(coord_axis, n, v)
    c = coord_axis.astype(F32)
    c_ = np.clip(c, F32(-1 + 1e-6), F32(1 - 1e-6)).astype(F32)
    t = ((((c_ + F32(1.0)).astype(F32) * F32(n)).astype(F32) - F32(1.0)).astype(F32) * F32(0.5)).astype(F32)
    idx = np.clip(np.rint(t).astype(np.int64), 0, n - 1)
    rel = ((c - axis_centres(n)[idx]).astype(F32) * F32(n)).astype(F32)
    return idx, rel


def liif_make_coord(n: int) -> np.ndarray:
    """One axis of LIIF.make_coord (liif.py:36-42): fl(fp32(-1 + 1/n) + fl(fp32(2/n) * i)) -- the same numbers as axis_centres."""
    return axis_centres(n)


def grid_coords(H_up: int, W_up: int):
    """Cell-centre coords of the HR grid, as the reference builds them (diinn.py:101-102)."""
    return axis_centres(H_up), axis_centres(W_up)


def calc_psnr(sr: np.ndarray, hr: np.ndarray, rgb_range: float = 1.0, dataset=None, scale: int = 1) -> float:
    """PSNR as in sr_module.py:21-38: dataset None = every pixel and channel; 'benchmark' = shave `scale` pixels and,
    for C > 1, convert the difference to luma with (65.738, 129.057, 25.064)/256 in fp32; 'div2k' = shave scale + 6.
    (`[shave:-shave]` with shave == 0 is an empty slice there -> nan, mirrored.) Mean in fp64."""
    diff = ((sr.astype(F32) - hr.astype(F32)) / F32(rgb_range)).astype(F32)
    if dataset is not None:
        if dataset == "benchmark":
            shave = int(scale)
            if diff.shape[1] > 1:
                conv = (np.array([65.738, 129.057, 25.064], dtype=F32) / F32(256)).reshape(1, 3, 1, 1)
                prod = (diff * conv).astype(F32)
                diff = ((prod[:, 0] + prod[:, 1]).astype(F32) + prod[:, 2]).astype(F32)
        elif dataset == "div2k":
            shave = int(scale) + 6
        else:
            raise NotImplementedError(dataset)
        if shave == 0:
            return float("nan")
        diff = diff[..., shave:-shave, shave:-shave]
    mse = float(np.mean(diff.astype(np.float64) ** 2))
    return float(-10.0 * np.log10(mse))


def denorm_clamp(pred: np.ndarray, sub: float = 0.5, div: float = 0.5, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """`(pred_hr * self.div + self.sub).clamp_(0, 1)` (sr_module.py:123): two rounded fp32 ops, then the clamp."""
    v = (pred.astype(F32) * F32(div)).astype(F32)
    v = (v + F32(sub)).astype(F32)
    return np.clip(v, F32(lo), F32(hi)).astype(F32)


def quantize_u8(img: np.ndarray) -> np.ndarray:
    """torchvision.utils.save_image's quantisation (reached from demo2.py:41): mul(255).add_(0.5).clamp_(0,255).to(uint8)."""
    v = (img.astype(F32) * F32(255)).astype(F32)
    v = (v + F32(0.5)).astype(F32)
    return np.clip(v, F32(0), F32(255)).astype(np.uint8)


# --------------------------------------------------------------------------------------------------
# torch-CPU timing port: the same algorithm (un-hoisted, materialising the (B,576,H_up,W_up) tensor like
# diinn.py:168 does) on PyTorch CPU kernels with all host threads -- used ONLY as bench.py's cpu_baseline
# / --impl reference arm, because /root/reference does not exist on the GPU box.
# --------------------------------------------------------------------------------------------------
def decoder_forward_torch_cpu(weights: dict, feat, size, bsize=None, rows=None, device="cpu", return_tensor=False):
    """rows=(r0,r1) restricts the work to an HR row band (the bounded sample bench.py times); everything else is the
    reference's algorithm at full arithmetic cost per pixel (1.97 MFLOP/px, nothing hoisted).
    device="cuda" runs the SAME eager op sequence on the GPU (bench.py's informative `eager_gpu_reference`: what the
    reference's own PyTorch path costs on the box, SURVEY.md section 8(d) "the real same-box bar")."""
    import torch
    import torch.nn.functional as TF

    with torch.no_grad():
        x = torch.as_tensor(feat, dtype=torch.float32).to(device)
        B, C, H, W = x.shape
        H_up, W_up = int(size[0]), int(size[1])
        r0, r1 = (0, H_up) if rows is None else (int(rows[0]), int(rows[1]))
        _, rh = rel_axis(H, H_up)
        _, rw = rel_axis(W, W_up)
        syn_np = np.empty((3, r1 - r0, W_up), dtype=F32)
        syn_np[0] = rh[r0:r1, None]
        syn_np[1] = rw[None, :]
        syn_np[2] = ratio_value(H, W, H_up, W_up)
        syn = torch.from_numpy(syn_np).to(device).unsqueeze(0).expand(B, -1, -1, -1)
        ih = torch.from_numpy(nearest_exact_index(H, H_up)[r0:r1]).to(device)
        iw = torch.from_numpy(nearest_exact_index(W, W_up)).to(device)
        u = TF.unfold(x, 3, padding=1).view(B, C * 9, H, W)
        xu = u[:, :, ih][:, :, :, iw]                      # nearest-exact gather -> (B,576,H_up,W_up)
        Wt = {k: torch.as_tensor(v).to(device) for k, v in weights.items()}
        fin = (lambda t: t) if return_tensor else (lambda t: t.cpu().numpy())

        def step(xs, ss):
            k = torch.relu(TF.conv2d(xs, Wt["K.0.0.weight"], Wt["K.0.0.bias"]))
            q = k * torch.sin(TF.conv2d(ss, Wt["Q.0.0.weight"], Wt["Q.0.0.bias"]))
            for i in range(1, 4):
                k = torch.relu(TF.conv2d(torch.cat([q, xs], 1), Wt[f"K.{i}.0.weight"], Wt[f"K.{i}.0.bias"]))
                q = k * torch.sin(TF.conv2d(q, Wt[f"Q.{i}.0.weight"], Wt[f"Q.{i}.0.bias"]))
            return TF.conv2d(q, Wt["last_layer.weight"], Wt["last_layer.bias"])

        if bsize is None:
            return fin(step(xu, syn))
        strip = max(1, bsize // (r1 - r0))
        outs = [step(xu[..., a:a + strip], syn[..., a:a + strip]) for a in range(0, W_up, strip)]
        return fin(torch.cat(outs, -1))
