"""Timeline of the fused stage-B kernel (DIINN_TRACE=1). python tools/trace_stage_b.py [config]"""
import ctypes as C
import os
import sys

os.environ["DIINN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth, _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, H, W, H_up, W_up = synth.CONFIGS[name]
dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision="bf16"),
                                    synth.make_weights(seed=0)).cuda()
x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
with torch.no_grad():
    dec(x, (H_up, W_up))
    dec(x, (H_up, W_up))
torch.cuda.synchronize()
buf = np.zeros(2048, dtype=np.int64)
lib = _lib.load()
_lib.check(lib, dec._handle, lib.diinn_debug_read_trace(dec._handle, buf.ctypes.data_as(C.c_void_p), 2048))
tr = buf.reshape(8, 256)
t0 = tr[tr > 0].min()
for t in range(1, 5):
    print(f"--- tile {t} (clk relative to first event {t0})")
    for layer in range(3):
        for q in range(4):
            m = tr[t, layer * 32 + q * 8: layer * 32 + q * 8 + 6] - t0
            e = tr[t, 128 + layer * 32 + q * 8: 128 + layer * 32 + q * 8 + 4] - t0
            print(f" L{layer + 1}.q{q} MMA slot_free {m[0]:7d} | kc-pair ready {m[1]:7d} {m[2]:7d} issued {m[5]:7d} "
                  f"|| EPI wait {e[0]:7d} full {e[1]:7d} slot_freed {e[2]:7d} done {e[3]:7d}")
