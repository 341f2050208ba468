"""LIIF-proper decoding on the same two kernels ("next" row 1).

Mirror of the query half of /root/reference/src/models/components/liif.py (``LIIF.query_rgb`` :59-127, ``batched_predict``
:129-140, ``make_coord`` :32-45, ``make_coord_and_cell`` :47-57, ``reshape_pred`` :142-146) with its imnet
``MLP(580, 3, [256]*4)`` (mlp.py:5-15). The encoder stays the caller's: every entry takes the feature map.

How it maps onto the DIINN kernels (csrc/api.cu, diinn_set_weights_liif): the 576 feature columns of the first Linear are
applied once per LR cell by stage A, the four coordinate columns and the three hidden ReLU layers per query by stage B, the
4-neighbour ensemble and the area blend inside stage B's last epilogue.
"""
import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .decoder import _PRECISIONS, _ptr, _stream, FusedImplicitDecoder


class MLP(nn.Module):
    """Same module tree (state_dict keys ``layers.{0,2,4,6,8}.*``, default init, RNG consumption) as mlp.py:5-15. Holds the
    parameters only: the arithmetic runs in the CUDA library."""

    def __init__(self, in_dim: int, out_dim: int, hidden_list):
        super().__init__()
        layers, lastv = [], in_dim
        for hidden in hidden_list:
            layers += [nn.Linear(lastv, hidden), nn.ReLU()]
            lastv = hidden
        layers.append(nn.Linear(lastv, out_dim))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        raise RuntimeError("the imnet of FusedLIIFQuery is a parameter container; call query_rgb() (no CPU fallback)")


class FusedLIIFQuery(nn.Module):
    """``LIIF`` without its encoder: ``query_rgb(feat, coord, cell)`` and friends on the B200 library.

    Only the reference's default wiring is implemented -- feat_unfold=True, cell_decode=True (a 580-wide imnet input);
    local_ensemble may be True (default) or False. precision: "fp16" (default) | "bf16" | "fp32" (fp16 hi+lo split
    operands), as for FusedImplicitDecoder. Forward-only."""

    def __init__(self, local_ensemble: bool = True, feat_unfold: bool = True, cell_decode: bool = True,
                 precision: str = "fp16"):
        super().__init__()
        if not feat_unfold or not cell_decode:
            raise NotImplementedError("only feat_unfold=True, cell_decode=True (the reference's defaults, liif.py:11) is implemented")
        if precision not in ("fp16", "bf16", "fp32"):
            raise ValueError("precision must be one of ['fp16', 'bf16', 'fp32']")
        self.local_ensemble, self.feat_unfold, self.cell_decode = bool(local_ensemble), True, True
        self.precision = precision
        self.imnet = MLP(64 * 9 + 2 + 2, 3, [256, 256, 256, 256])
        self._handle = None
        self._handle_device = None
        self._packed_versions = None
        self._workspace = None

    # ------------------------------------------------------------------ handle / weights
    def _ref_tensors(self):
        lin = [self.imnet.layers[i] for i in (0, 2, 4, 6, 8)]
        return [l.weight for l in lin] + [l.bias for l in lin]

    def _ensure_handle(self, device: torch.device):
        lib = _lib.load()
        if device.type != "cuda":
            raise RuntimeError("FusedLIIFQuery runs on CUDA (sm_100a) only; there is no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_device != idx:
            self.release()
            cfg = _lib.Config(64, 256, 4, 3, 0, idx)
            h = C.c_void_p()
            _lib.check(lib, None, lib.diinn_create(C.byref(h), C.byref(cfg)))
            self._handle, self._handle_device, self._packed_versions = h, idx, None
        tensors = self._ref_tensors()
        versions = tuple((t.data_ptr(), t._version) for t in tensors)
        if versions != self._packed_versions:
            for t in tensors:
                if t.device.type != "cuda" or t.device.index != idx:
                    raise RuntimeError("imnet parameters must live on the same CUDA device as the input")
            ws = [t.detach().to(torch.float32).contiguous() for t in tensors]
            w = _lib.LiifWeightsF32()
            for i in range(5):
                w.weight[i], w.bias[i] = ws[i].data_ptr(), ws[5 + i].data_ptr()
            w.on_device = 1
            _lib.check(lib, self._handle, lib.diinn_set_weights_liif(self._handle, C.byref(w), _stream(device)))
            self._packed_versions = versions
        return lib, self._handle

    def release(self):
        if self._handle is not None:
            _lib.load().diinn_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # ------------------------------------------------------------------ the reference's helpers (liif.py:32-57,142-146)
    @staticmethod
    def make_coord(size, device, ranges=None, flatten=True):
        """Coordinates at grid centres; the arithmetic order of liif.py:36-42 (python-double scalars, fp32 tensor ops)."""
        seqs = []
        for i, n in enumerate(size):
            v0, v1 = (-1, 1) if ranges is None else ranges[i]
            r = (v1 - v0) / (2 * n)
            seqs.append(v0 + r + (2 * r) * torch.arange(n, device=device).float())
        ret = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1)
        return ret.view(-1, ret.shape[-1]) if flatten else ret

    def make_coord_and_cell(self, inp: torch.Tensor, size):
        hr_coord = self.make_coord(size, inp.device)
        cell = torch.ones_like(hr_coord)
        cell[:, 0] *= 2 / size[-2]
        cell[:, 1] *= 2 / size[-1]
        B = inp.shape[0]
        return hr_coord.unsqueeze(0).expand(B, -1, -1).contiguous(), cell.unsqueeze(0).expand(B, -1, -1).contiguous()

    @staticmethod
    def reshape_pred(pred: torch.Tensor, size):
        return pred.view(pred.shape[0], *size, 3).permute(0, 3, 1, 2).contiguous()

    # ------------------------------------------------------------------ the hot path
    def query_rgb(self, feat: torch.Tensor, coord: torch.Tensor, cell: Optional[torch.Tensor] = None) -> torch.Tensor:
        """LIIF.query_rgb (liif.py:59-127): feat (B,64,H,W), coord / cell (B,Q,2) as (h,w) -> (B,Q,3)."""
        if feat.dim() != 4 or feat.shape[1] != 64:
            raise ValueError(f"expected a (B,64,H,W) feature map, got {tuple(feat.shape)}")
        if feat.device.type != "cuda":
            raise RuntimeError("FusedLIIFQuery runs on CUDA (sm_100a) only; there is no CPU fallback")
        if torch.is_grad_enabled() and (feat.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("FusedLIIFQuery is forward-only (no backward pass): call it under torch.no_grad()")
        if cell is None:
            raise ValueError("cell_decode=True needs the cell tensor (liif.py:107-111)")
        lib, h = self._ensure_handle(feat.device)
        comp = _PRECISIONS[self.precision]
        if not (FusedImplicitDecoder._is_nhwc_bf16(feat) and comp in (_lib.COMPUTE_BF16, _lib.COMPUTE_FP16)):
            feat = feat.contiguous()
        B, Cc, H, W = feat.shape
        Q = coord.shape[1]
        coord = coord.to(torch.float32).contiguous()
        cell = cell.to(torch.float32).contiguous()
        if coord.shape != (B, Q, 2) or cell.shape != (B, Q, 2):
            raise ValueError("coord and cell must be (B,Q,2)")
        io = FusedImplicitDecoder._io_dtype(feat)
        nbytes = lib.diinn_query_workspace_bytes(h, B, H, W, Q * (4 if self.local_ensemble else 1), comp)
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != feat.device:
            self._workspace = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=feat.device)
        ws = self._workspace
        out = torch.empty((B, Q, 3), dtype=torch.float32 if feat.dtype == torch.float32 else torch.bfloat16, device=feat.device)
        fn = lib.diinn_query_ensemble if self.local_ensemble else lib.diinn_query
        _lib.check(lib, h, fn(h, _ptr(feat), B, Cc, H, W, _ptr(coord), _ptr(cell), Q, _ptr(out), _ptr(ws), ws.numel(), io,
                              comp, _stream(feat.device)))
        return out

    def batched_predict(self, feat, coord, cell, bsize):
        """liif.py:129-140. Chunks of bsize queries, concatenated; every query is independent, so the result is the same
        as one call (kept for call-site compatibility)."""
        with torch.no_grad():
            n, ql, preds = coord.shape[1], 0, []
            while ql < n:
                qr = min(ql + bsize, n)
                preds.append(self.query_rgb(feat, coord[:, ql:qr, :], cell[:, ql:qr, :]))
                ql = qr
            return torch.cat(preds, dim=1)

    def forward(self, feat: torch.Tensor, size, bsize: Optional[int] = None) -> torch.Tensor:
        """LIIF.forward (liif.py:151-158) from the encoder's output on: feat (B,64,H,W) -> (B,3,*size)."""
        coord, cell = self.make_coord_and_cell(feat, size)
        pred = self.batched_predict(feat, coord, cell, bsize) if bsize is not None else self.query_rgb(feat, coord, cell)
        return self.reshape_pred(pred, size)


def load_liif_imnet(module: FusedLIIFQuery, state_dict: dict, prefix: str = "imnet.") -> FusedLIIFQuery:
    """Copy the ``imnet.layers.*`` entries of a reference LIIF state_dict (torch tensors or numpy arrays) into the module."""
    sd = {}
    for k, v in state_dict.items():
        if k.startswith(prefix):
            sd[k[len(prefix):]] = v if isinstance(v, torch.Tensor) else torch.from_numpy(v.copy())
    module.imnet.load_state_dict(sd, strict=True)
    return module
