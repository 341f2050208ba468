"""Which stage-B kernel a launch takes, decided on the host (csrc/stage_b_umma.cu: make_plan) and probed here WITHOUT a GPU
through diinn_debug_plan_stage_b. The decisions that change the last bits of the image -- select-MMA variant (fp16 P added by
the tensor core) and canonical relative coordinates -- must depend on the image geometry and the operand format only, never
on the row range of a launch: row tiles have to be bit-identical to the full decode (SURVEY.md section 4 test 7)."""
import ctypes as C
import random

import numpy as np
import pytest

import diinn_b200
from diinn_b200 import _lib, synth

FP32, BF16, FP16 = 0, 1, 2
KEYS = ("sel", "tab", "canon", "s_h", "s_w", "pw_log2", "ksel", "box_r", "box_c", "n_work", "tiles_y", "n_txp")


def plan(H, W, H_up, W_up, row0=0, row1=None, compute=FP16, B=1, mode=3, sms=148):
    lib = _lib.load()
    out = np.zeros(12, dtype=np.int32)
    rc = lib.diinn_debug_plan_stage_b(sms, mode, B, H, W, H_up, W_up, row0, H_up if row1 is None else row1, compute,
                                      out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return dict(zip(KEYS, (int(v) for v in out)))


def test_benchmark_configurations():
    _, H, W, H_up, W_up = synth.CONFIGS["c3"]
    p = plan(H, W, H_up, W_up)
    assert (p["sel"], p["tab"], p["canon"], p["s_h"], p["s_w"]) == (1, 1, 1, 4, 4)
    assert (p["pw_log2"], p["ksel"], p["box_r"], p["box_c"]) == (4, 16, 2, 8)          # 8x16 patches: 2 x 8 LR cells per pair
    assert p["n_work"] == 170 * 64
    for name, want in (("c1", (1, 1, 1)), ("c2x4", (1, 1, 1)), ("c2x2", (0, 1, 1)), ("c2x3", (0, 1, 1)), ("c5", (0, 1, 1))):
        _, h, w, hu, wu = synth.CONFIGS[name]
        q = plan(h, w, hu, wu, B=synth.CONFIGS[name][0])
        assert (q["sel"], q["tab"], q["canon"]) == want, name
    _, h, w, hu, wu = synth.CONFIGS["c4"]                                              # x12: 144 phases, no table
    q = plan(h, w, hu, wu)
    assert (q["sel"], q["tab"], q["canon"], q["ksel"]) == (1, 0, 0, 16)
    # the fp32-precision path never takes canonical coordinates or the select variant; bf16 operands do
    assert (plan(H, W, H_up, W_up, compute=FP32)["canon"], plan(H, W, H_up, W_up, compute=FP32)["sel"]) == (0, 0)
    assert plan(H, W, H_up, W_up, compute=BF16)["tab"] == 1
    # non-integer scale factors: the reference's per-pixel coordinates
    assert plan(16, 20, 37, 51)["canon"] == 0 and plan(23, 31, 91, 125)["canon"] == 0
    # 2 x 8 = 16 phases still fit the table, 3 x 6 do not
    assert plan(10, 12, 20, 96)["tab"] == 1 and plan(10, 12, 30, 72)["canon"] == 0


def test_row_tiles_of_an_eight_way_split():
    _, H, W, H_up, W_up = synth.CONFIGS["c3"]
    full = plan(H, W, H_up, W_up)
    # tiles on multiples of the scale factor (tile_partition): 4x32 patches = one LR row x 16 cells, 19 waves on 74 CTA pairs
    for r0, r1 in diinn_b200.tile_partition(H, H_up, 8):
        p = plan(H, W, H_up, W_up, r0, r1)
        assert (p["sel"], p["canon"], p["tab"], p["ksel"]) == (1, 1, 1, 16), (r0, r1)
        assert -(-p["n_work"] // 74) <= 19
    p = plan(H, W, H_up, W_up, 172, 344)
    assert (p["pw_log2"], p["box_r"], p["box_c"], p["n_work"]) == (5, 1, 16, 43 * 32)
    # an unaligned tile straddles LR rows: same bit-relevant decisions, but two select MMAs and no room for the table
    p = plan(H, W, H_up, W_up, 170, 340)
    assert (p["sel"], p["canon"]) == (full["sel"], full["canon"]) and (p["tab"], p["ksel"]) == (0, 32)


def test_bit_relevant_decisions_do_not_depend_on_the_row_range():
    rng = random.Random(7)
    for _ in range(300):
        H, W = rng.randint(2, 60), rng.randint(2, 60)
        if rng.random() < 0.5:
            H_up, W_up = H * rng.randint(1, 13), W * rng.randint(1, 13)
        else:
            H_up, W_up = rng.randint(1, 400), rng.randint(1, 400)
        compute = rng.choice((FP32, BF16, FP16))
        mode = rng.choice((1, 2, 3, 4))
        full = plan(H, W, H_up, W_up, compute=compute, mode=mode)
        for _ in range(4):
            a = rng.randint(0, H_up - 1)
            b = rng.randint(a + 1, H_up)
            p = plan(H, W, H_up, W_up, a, b, compute=compute, mode=mode)
            assert (p["sel"], p["canon"], p["s_h"], p["s_w"]) == (full["sel"], full["canon"], full["s_h"], full["s_w"])
            if p["sel"]:
                assert p["box_r"] * p["box_c"] <= p["ksel"] <= 32
            if p["tab"]:
                assert p["canon"] and p["s_h"] * p["s_w"] <= 16 and (not p["sel"] or p["ksel"] == 16)
            ph, pw = 128 >> p["pw_log2"], 1 << p["pw_log2"]
            assert p["tiles_y"] == -(-(b - a) // ph) and p["n_txp"] == -(-(-(-W_up // pw)) // 2)


def test_bad_arguments():
    lib = _lib.load()
    out = np.zeros(12, dtype=np.int32)
    assert lib.diinn_debug_plan_stage_b(148, 3, 1, 8, 8, 32, 32, 5, 5, FP16, out.ctypes.data_as(C.c_void_p)) != 0
    assert lib.diinn_debug_plan_stage_b(148, 3, 1, 8, 8, 32, 32, 0, 32, 3, out.ctypes.data_as(C.c_void_p)) != 0
    assert lib.diinn_debug_plan_stage_b(148, 3, 1, 8, 8, 32, 32, 0, 32, FP16, None) != 0
