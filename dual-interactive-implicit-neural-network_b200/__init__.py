"""diinn_b200 -- B200-native (sm_100a) query decoder for DIINN (arch=diinn, mode=3, init_q=False).

Drop-in for /root/reference/src/models/components/diinn.py:39-173 (``ImplicitDecoder``) behind a C-ABI CUDA
library (include/diinn_b200.h). No CPU fallback: every compute entry needs the built CUDA library and a GPU.
"""
from . import synth  # noqa: F401  (pure numpy; safe without the CUDA library)
from . import _lib  # noqa: F401
from .decoder import FusedImplicitDecoder, SineAct, load_numpy_weights, swap_decoder  # noqa: F401
from .liif import FusedLIIFQuery, load_liif_imnet  # noqa: F401
from .sharding import row_partition, row_align, tile_partition, band_partition, decode_sharded, decode_sharded_fused, broadcast_features  # noqa: F401

__all__ = ["synth", "FusedImplicitDecoder", "FusedLIIFQuery", "load_liif_imnet", "SineAct", "load_numpy_weights", "swap_decoder", "row_partition", "row_align", "tile_partition", "band_partition",
           "decode_sharded", "decode_sharded_fused", "broadcast_features"]
