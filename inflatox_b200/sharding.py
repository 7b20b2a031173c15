"""Static sharding of the grid-evaluation work over ranks / devices.

Every grid point and every parameter vector is independent, so there is no exchange step and no
collective on the data path: a single grid is cut into contiguous ROW blocks (C order makes each
block one contiguous slice of the host output), a parameter sweep is cut into contiguous blocks
of parameter vectors.  Coordinates are always computed from GLOBAL row indices, so a shard is
bit-identical to the same rows of an unsharded evaluation.  The engine applies the same rule
across the devices of one process (csrc/inflx_engine.cpp: inflx_grid_eval).
"""
from __future__ import annotations


def shard(n_rows: int, n_vectors: int, rank: int, world: int) -> tuple[tuple[int, int], tuple[int, int]]:
    """((row_begin, row_end), (vector_begin, vector_end)) owned by `rank` of `world`."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    if n_vectors >= world and world > 1 and n_vectors > 1:
        return (0, n_rows), (n_vectors * rank // world, n_vectors * (rank + 1) // world)
    return (n_rows * rank // world, n_rows * (rank + 1) // world), (0, n_vectors)
