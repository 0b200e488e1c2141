"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MHLA hot path.

Nothing under ``mhla_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it,
and there only as the checker (or as the timed CPU baseline), never as the product path.
"""
from .mhla_oracle import (  # noqa: F401
    block_distance_matrix,
    blockmix_fwd,
    blockmix_summaries,
    causal_chunk_fwd,
    causal_closed_form,
    recurrent_first_chunk_fwd,
    rope_freqs_wan,
    rope_apply_wan,
    err_ratio,
)
