"""Golden fixtures for the eval glue either side of the decoder.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_eval.py
``calc_psnr`` is EXECUTED FROM THE REFERENCE'S SOURCE: src/models/sr_module.py cannot be imported here (it needs
pytorch_lightning), so the function definition (sr_module.py:21-38) is cut out of the file with ``ast`` and exec'd
unchanged. The de-normalise + clamp is the reference's own expression (sr_module.py:123) applied with torch; the uint8
quantisation is torchvision.utils.save_image's (``mul(255).add_(0.5).clamp_(0, 255).to(uint8)``, reached from
demo2.py:41): the fixture is the PNG that call writes, decoded again.
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from diinn_b200 import synth  # noqa: E402

SRC = "/root/reference/src/models/sr_module.py"


def reference_calc_psnr():
    tree = ast.parse(open(SRC).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "calc_psnr")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    return ns["calc_psnr"], (fn.lineno, fn.end_lineno)


def main():
    calc_psnr, lines = reference_calc_psnr()
    out = {"calc_psnr.lines": np.array(lines)}
    B, H, W = 2, 40, 56
    hr = synth.uniform(7, 0, (B, 3, H, W), 0.0, 1.0)
    sr = (hr + synth.uniform(8, 0, (B, 3, H, W), -0.025, 0.025)).astype(np.float32)
    out["sr"], out["hr"] = sr, hr
    t_sr, t_hr = torch.from_numpy(sr), torch.from_numpy(hr)
    cases = [(None, 1, 1.0), (None, 4, 255.0), ("benchmark", 2, 1.0), ("benchmark", 4, 1.0), ("div2k", 4, 1.0),
             ("div2k", 2, 1.0)]
    vals = []
    for ds, sc, rr in cases:
        vals.append(float(calc_psnr(t_sr, t_hr, dataset=ds, scale=sc, rgb_range=rr)))
        print(ds, sc, rr, vals[-1])
    out["psnr.dataset"] = np.array([{None: 0, "benchmark": 1, "div2k": 2}[c[0]] for c in cases])
    out["psnr.scale"] = np.array([c[1] for c in cases])
    out["psnr.rgb_range"] = np.array([c[2] for c in cases], dtype=np.float32)
    out["psnr.value"] = np.array(vals, dtype=np.float64)
    # single-channel 'benchmark' (no luma conversion)
    out["psnr.gray1"] = np.array(float(calc_psnr(t_sr[:, :1], t_hr[:, :1], dataset="benchmark", scale=3)))
    # de-normalise + clamp (sr_module.py:123) and uint8 quantisation (torchvision save_image)
    pred = synth.uniform(9, 0, (1, 3, 33, 47), -1.5, 1.5)   # spills over both clamp ends
    t = torch.from_numpy(pred)
    sub, div = 0.5, 0.5
    den = (t * div + sub).clamp_(0, 1)
    out["pred"], out["denorm"] = pred, den.numpy()
    import io
    import torchvision
    from PIL import Image
    buf = io.BytesIO()
    torchvision.utils.save_image(den, buf, format="png")   # demo2.py:41 writes the PNG exactly like this
    buf.seek(0)
    out["u8"] = np.asarray(Image.open(buf)).transpose(2, 0, 1)[None].copy()   # back to (1,3,H,W) uint8
    assert np.array_equal(out["u8"], den.clone().mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8).numpy())
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)
    print("wrote eval.npz")


if __name__ == "__main__":
    main()
