"""mhla_b200 - B200-native (sm_100a) MHLA forward operator behind the reference's call surface.

Host side: Python/PyTorch shims.  Hot path: hand-written CUDA (TMA + tcgen05) in ``libmhla_b200.so``,
reached through the C ABI declared in ``include/mhla_b200.h``.
"""
from .ops import (  # noqa: F401
    mhla, mhla_blockmix, mhla_blockmix_grid, wan_prep, gated_rmsnorm, gate_add, dwconv3d_tokens, mhla_host, mhla_causal, naive_chunk_simple_mhla_fixed, naive_recurrent_mhla, last_launch_count,
)
from .mixing import BlockDistanceConv, BlockDistanceConv3D, block_distance_matrix  # noqa: F401

__version__ = "0.1.0"
