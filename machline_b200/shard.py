"""Row partition of the permuted system over the GPUs of one box.

The AIC is assembled where it is used: rank r builds and keeps rows [row0, row0+nrows) of the permuted system
(contiguous blocks, multiples of 64 rows so that every shard's leading dimension is aligned) and never ships them.
The Krylov solvers exchange one vector per matvec: each rank contributes `shard_pad` entries to an all-gather
(`ml_solve`, csrc/gpu/solve_kernels.cu: Sys::matvec) and the padded shards are compacted into the full vector.
These helpers are the single statement of that layout on the host side (bench.py, tests).

For the direct solver the rows can instead be dealt block-cyclically (`cyclic_rows`, `ml_set_row_shard_cyclic`): with
contiguous blocks and a diagonally dominant matrix the ranks run out of rows below the pivot one after the other, so
the last rank does N^3/P flops instead of 2/3 N^3/P.  The library learns any dealing from the row lists the ranks
exchange (slot tables in ml_solve), so both layouts go through the same kernels."""
from __future__ import annotations

import numpy as np

ALIGN = 64


def rows_per_rank(n_rows: int, world: int, align: int = ALIGN) -> int:
    per = (n_rows + world - 1) // world
    return (per + align - 1) // align * align


def row_shard(n_rows: int, rank: int, world: int, align: int = ALIGN) -> tuple[int, int]:
    """(row0, nrows) of `rank`; trailing ranks may own no rows when n_rows is small."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    per = rows_per_rank(n_rows, world, align)
    row0 = min(n_rows, rank * per)
    return row0, max(0, min(n_rows, row0 + per) - row0)


def all_shards(n_rows: int, world: int, align: int = ALIGN) -> list[tuple[int, int]]:
    return [row_shard(n_rows, r, world, align) for r in range(world)]


def shard_pad(shards, align: int = ALIGN) -> int:
    """Entries every rank contributes to the all-gather of a distributed vector (largest shard, aligned)."""
    mx = max((n for _, n in shards), default=0)
    return (mx + align - 1) // align * align


def compact_gathered(gathered: np.ndarray, shards, pad: int) -> np.ndarray:
    """[world * pad] all-gather buffer -> contiguous full vector (what Sys::matvec does with D2D copies)."""
    n = sum(nr for _, nr in shards)
    out = np.empty(n, dtype=gathered.dtype)
    for r, (r0, nr) in enumerate(shards):
        out[r0:r0 + nr] = gathered[r * pad:r * pad + nr]
    return out


CYCLIC_BLOCK = 128   # the trailing update of the LU works on blocks of 128 local rows


def cyclic_rows(n_rows: int, rank: int, world: int, block: int = CYCLIC_BLOCK) -> np.ndarray:
    """Global rows of `rank` under block-cyclic dealing: blocks b = rank (mod world) of `block` rows, ascending."""
    if not (0 <= rank < world) or block <= 0:
        raise ValueError(f"rank {rank} outside world {world} or block {block} <= 0")
    rows = [np.arange(b0, min(b0 + block, n_rows)) for b0 in range(rank * block, n_rows, world * block)]
    return np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, dtype=np.int32)
