"""Static sharding of the grid-evaluation work over ranks / devices.

Every grid point and every parameter vector is independent, so there is no exchange step and no
collective on the data path: a single grid is cut into contiguous ROW blocks (C order makes each
block one contiguous slice of the host output), a parameter sweep is cut into contiguous blocks
of parameter vectors.  Coordinates are always computed from GLOBAL row indices, so a shard is
bit-identical to the same rows of an unsharded evaluation.  There is ONE implementation of the
rule: `inflx_shard_of` in csrc/inflx_engine.cpp, which the engine applies across the devices of one
process (inflx_grid_eval) and this module across the ranks of a one-process-per-GPU job.
"""
from __future__ import annotations

import ctypes

from . import _native


def shard(n_rows: int, n_vectors: int, rank: int, world: int) -> tuple[tuple[int, int], tuple[int, int]]:
    """((row_begin, row_end), (vector_begin, vector_end)) owned by `rank` of `world`."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    out = (ctypes.c_uint64 * 4)()
    _native.raise_for_status(_native.lib().inflx_shard_of(n_rows, n_vectors, rank, world, out))
    return (int(out[0]), int(out[1])), (int(out[2]), int(out[3]))
