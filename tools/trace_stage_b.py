"""Timeline of the fused stage-B kernel (DIINN_TRACE=1). python tools/trace_stage_b.py [config]"""
import ctypes as C
import os
import sys

os.environ["DIINN_TRACE"] = "1"
# the timeline points are compiled out of the product library: build (here, no GPU needed) a side library first with
#   ABL_DEFS="-DDIINN_TRACE_BUILD=1" python tools/ablate_stage_b.py build 60      [add -DDIINN_FINE_TRACE=1 for per-step points]
_side = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dual-interactive-implicit-neural-network_b200",
                     "build", "libdiinn_b200_abl60.so")
if "DIINN_B200_LIB" not in os.environ and os.path.exists(_side):
    os.environ["DIINN_B200_LIB"] = _side
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import diinn_b200  # noqa: E402
from diinn_b200 import synth, _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, H, W, H_up, W_up = synth.CONFIGS[name]
dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=(sys.argv[2] if len(sys.argv) > 2 else "fp16")),
                                    synth.make_weights(seed=0)).cuda()
x = torch.from_numpy(synth.make_feat(1, B, H, W)).cuda()
with torch.no_grad():
    dec(x, (H_up, W_up))
    dec(x, (H_up, W_up))
torch.cuda.synchronize()
buf = np.zeros(1024, dtype=np.int64)
lib = _lib.load()
_lib.check(lib, dec._handle, lib.diinn_debug_read_trace(dec._handle, buf.ctypes.data_as(C.c_void_p), 1024))
tr = buf.reshape(8, 128)
t0 = tr[tr > 0].min()
for t in range(1, 5):
    print(f"--- tile {t} (clk relative to first event {t0})")
    for layer in range(3):
        for h in range(2):
            m = tr[t, layer * 20 + h * 10: layer * 20 + h * 10 + 10] - t0
            e = tr[t, 64 + layer * 10 + h * 5: 64 + layer * 10 + h * 5 + 5] - t0
            sw = tr[t, 112 + layer * 4 + h * 2: 114 + layer * 4 + h * 2] - t0
            print(f" L{layer + 1}.h{h} sel: chunk0 ready {sw[0]:7d} B_sel landed {sw[1]:7d}")
            print(f" L{layer + 1}.h{h} MMA slot_free {m[0]:7d} | act ready / issued kc0 {m[1]:7d}/{m[2]:7d} kc1 {m[3]:7d}/{m[4]:7d} "
                  f"kc2 {m[5]:7d}/{m[6]:7d} kc3 {m[7]:7d}/{m[8]:7d} issued {m[9]:7d} || EPI wait {e[0]:7d} full {e[1]:7d} "
                  f"c0 {e[2]:7d} c1 {e[3]:7d} freed {e[4]:7d}")
    # producer side (leader CTA): when it entered the tile's loop, and per half slot when the
    # B_sel stage came free (= the previous half slot's select MMA retired) and when the first weight stage came free
    pr = tr[t, 96:112] - t0
    print(f" producer: tile loop entered {pr[13]:7d} | last weight stage of L3.h1 requested {pr[14]:7d}")
    print(" producer: per half slot  B_sel stage free / kc0 stage free: "
          + "  ".join(f"L{lh // 2 + 1}.h{lh % 2} {pr[lh * 2 + 1]:7d}/{pr[lh * 2]:7d}" for lh in range(6)))
    for h in range(2) if os.environ.get("DIINN_FINE") else []:
        f = tr[t, 96 + h * 8: 96 + h * 8 + 8] - t0
        print(f" L2.h{h} fine (needs -DDIINN_FINE_TRACE=1): ld0 landed {f[0]} math0 done {f[1]} ld1 landed {f[2]} math1 done {f[3]}")
