"""Deterministic synthetic inputs for the DIINN query decoder (weights, feature maps, query coords).

There is no network here for datasets or checkpoints, so every test / bench input is synthetic.
The generator is a counter-based integer hash evaluated with numpy uint64 arithmetic only (no libm
calls), so the same (seed, shape) yields bit-identical float32 arrays on any host: the golden vectors in
``tests/golden`` were produced from these arrays by the reference ``ImplicitDecoder``
(/root/reference/src/models/components/diinn.py:39-173) and can be re-checked on the GPU box where the
reference tree does not exist.

Shapes / names of the weight set are the ``state_dict`` of the reference decoder for mode=3,
init_q=False (diinn.py:73-80,92): K.i.0.{weight,bias}, Q.i.0.{weight,bias}, last_layer.{weight,bias}.
Default scale mimics nn.Conv2d's default init, U(-1/sqrt(fan_in), 1/sqrt(fan_in)).
"""
from __future__ import annotations

import numpy as np

IN_CHANNELS = 64
HIDDEN = 256
N_LAYERS = 4
UNFOLD = IN_CHANNELS * 9  # 576

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, stream: int, n: int) -> np.ndarray:
    """n float64 values in [0,1), exactly representable (53-bit), deterministic everywhere."""
    with np.errstate(over="ignore"):
        base = _mix(np.array([seed], dtype=np.uint64) * np.uint64(0x632BE59BD9B4E019)
                    + np.uint64(stream) * np.uint64(0xD1342543DE82EF95))
        ctr = np.arange(n, dtype=np.uint64)
        bits = _mix(ctr * np.uint64(0x2545F4914F6CDD1D) + base)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(seed: int, stream: int, shape, lo: float, hi: float) -> np.ndarray:
    n = int(np.prod(shape))
    u = uniform01(seed, stream, n)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def normalish(seed: int, stream: int, shape, std: float) -> np.ndarray:
    """Approximately N(0, std^2): centred sum of 4 uniforms (Irwin-Hall), adds/muls only."""
    n = int(np.prod(shape))
    u = uniform01(seed, stream, 4 * n).reshape(4, n)
    z = (u[0] + u[1] + u[2] + u[3] - 2.0) * np.sqrt(3.0)  # unit variance; sqrt(3.0) is one exact constant
    return (z * std).astype(np.float32).reshape(shape)


def weight_names(n_layers: int = N_LAYERS):
    names = []
    for i in range(n_layers):
        names += [f"K.{i}.0.weight", f"K.{i}.0.bias", f"Q.{i}.0.weight", f"Q.{i}.0.bias"]
    names += ["last_layer.weight", "last_layer.bias"]
    return names


def weight_shapes(in_channels: int = IN_CHANNELS, hidden: int = HIDDEN, n_layers: int = N_LAYERS, mode: int = 3,
                  init_q: bool = False):
    """name -> shape, identical to the reference state_dict (mode 1: K.i takes k only, diinn.py:57-64;
    init_q: first_layer = Conv2d(3, 576, 1) and a 576-wide Q.0, diinn.py:48-53). The first_layer entries come LAST so that
    every other tensor keeps its random stream."""
    unfold = in_channels * 9
    shapes = {}
    for i in range(n_layers):
        kin = unfold if i == 0 else (hidden if mode == 1 else hidden + unfold)
        qin = (unfold if init_q else 3) if i == 0 else hidden
        shapes[f"K.{i}.0.weight"] = (hidden, kin, 1, 1)
        shapes[f"K.{i}.0.bias"] = (hidden,)
        shapes[f"Q.{i}.0.weight"] = (hidden, qin, 1, 1)
        shapes[f"Q.{i}.0.bias"] = (hidden,)
    shapes["last_layer.weight"] = (3, hidden, 3, 3) if mode == 4 else (3, hidden, 1, 1)   # diinn.py:89-92
    shapes["last_layer.bias"] = (3,)
    if init_q:
        shapes["first_layer.0.weight"] = (unfold, 3, 1, 1)
        shapes["first_layer.0.bias"] = (unfold,)
    return shapes


def make_weights(seed: int = 0, k_gain: float = 1.0, q_gain: float = 1.0, last_gain: float = 1.0, mode: int = 3,
                 init_q: bool = False, first_gain: float = 1.0):
    """Reference-layout decoder weights as a dict of float32 numpy arrays.

    k_gain/q_gain > 1 give the "stress" set of SURVEY.md §4 item 8 (activations O(1) instead of being
    dominated by last_layer.bias)."""
    out = {}
    shapes = weight_shapes(mode=mode, init_q=init_q)
    for s, (name, shape) in enumerate(shapes.items()):
        wshape = shape if len(shape) == 4 else shapes[name.replace("bias", "weight")]
        fan_in = wshape[1] * wshape[2] * wshape[3]
        bound = 1.0 / np.sqrt(float(fan_in))
        gain = (k_gain if name.startswith("K.") else q_gain if name.startswith("Q.") else
                first_gain if name.startswith("first_layer") else last_gain)
        out[name] = uniform(seed, 100 + s, shape, -bound * gain, bound * gain)
    return out


LIIF_IMNET_SHAPES = {"layers.0.weight": (256, 580), "layers.0.bias": (256,), "layers.2.weight": (256, 256), "layers.2.bias": (256,),
                     "layers.4.weight": (256, 256), "layers.4.bias": (256,), "layers.6.weight": (256, 256), "layers.6.bias": (256,),
                     "layers.8.weight": (3, 256), "layers.8.bias": (3,)}


def make_liif_weights(seed: int = 0, gain: float = 1.0):
    """LIIF's imnet = MLP(580, 3, [256]*4) (liif.py:26, mlp.py:5-15) in state_dict layout, nn.Linear-style bounds
    (uniform(+-1/sqrt(fan_in))); gain > 1 scales the hidden layers for O(1) activations."""
    out = {}
    for s, (name, shape) in enumerate(LIIF_IMNET_SHAPES.items()):
        fan_in = LIIF_IMNET_SHAPES[name.replace("bias", "weight")][1]
        bound = (1.0 if name.startswith("layers.8") else gain) / np.sqrt(float(fan_in))
        out[name] = uniform(seed, 300 + s, shape, -bound, bound)
    return out


def make_feat(seed: int, B: int, H: int, W: int, C: int = IN_CHANNELS, std: float = 0.34) -> np.ndarray:
    """Synthetic encoder output (B,C,H,W); std 0.34 matches the random-init RDN feature std."""
    return normalish(seed, 7, (B, C, H, W), std)


def make_query(seed: int, B: int, Q: int, cell_hw=(2.0 / 96, 2.0 / 96)):
    """Sampled query coords in (-1,1) as (h,w) and constant cells (config c5, sampled form)."""
    coord = uniform(seed, 11, (B, Q, 2), -1.0, 1.0)
    cell = np.empty((B, Q, 2), dtype=np.float32)
    cell[..., 0] = cell_hw[0]
    cell[..., 1] = cell_hw[1]
    return coord, cell


# The five BASELINE.json configs as (B, H, W, H_up, W_up); "510x339" in BASELINE.json is WxH.
CONFIGS = {
    "c1": (1, 48, 48, 192, 192),
    "c2x2": (1, 256, 256, 512, 512),
    "c2x3": (1, 256, 256, 768, 768),
    "c2x4": (1, 256, 256, 1024, 1024),
    "c3": (1, 339, 510, 1356, 2040),
    "c4": (1, 360, 640, 4320, 7680),
    "c5": (16, 48, 48, 48, 48),
}
