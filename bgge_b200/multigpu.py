"""Host-side plans for the two places the path shards across GPUs:

* independent chains / CV folds -> `shard_units` (no data-path collective), and
* the getK contraction          -> `panel_plan` / `assemble_panels`: 2*world equal tile-row chunks of
  the lower triangle, rank r owns chunks r and 2*world-1-r (equal work), one all-gather of the
  panels, local mirror.  The functions are backend-agnostic (the tests drive them over gloo on
  CPU with a numpy stand-in for the Gram kernel; bench.py drives them over NCCL).
"""
from __future__ import annotations

TILE = 128


def shard_units(n_units: int, world: int, rank: int):
    """Contiguous, balanced assignment of independent units (chains x folds) to ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def panel_plan(L: int, world: int):
    """Returns (chunk_tile_rows, panel_rows, owners) where owners[r] = the two chunk indices of rank r."""
    T = (L + TILE - 1) // TILE
    ch = (T + 2 * world - 1) // (2 * world)
    owners = [(r, 2 * world - 1 - r) for r in range(world)]
    return ch, ch * TILE, owners


def chunk_tile_rows(L: int, world: int, chunk: int):
    """Tile-row range [t0, t1) of one chunk, clamped to the matrix."""
    T = (L + TILE - 1) // TILE
    ch, _, _ = panel_plan(L, world)
    return min(T, chunk * ch), min(T, (chunk + 1) * ch)


def tiles_per_rank(L: int, world: int):
    """Lower-triangle tile count each rank computes (for balance checks / reporting)."""
    out = []
    for r in range(world):
        n = 0
        for c in (r, 2 * world - 1 - r):
            t0, t1 = chunk_tile_rows(L, world, c)
            n += sum(t + 1 for t in range(t0, t1))
        out.append(n)
    return out


def assemble_panels(S, gathered, L: int, world: int):
    """gathered[r][k] is rank r's k-th panel stored [column j][local row]; S is indexed S[j, i]
    (column-major L x L seen as a row-major (L, L) array).  Fills the rows every panel covers."""
    _, R, owners = panel_plan(L, world)
    for r in range(world):
        for k, c in enumerate(owners[r]):
            r0 = c * R
            r1 = min(L, r0 + R)
            if r1 > r0:
                S[:, r0:r1] = gathered[r][k][:, :r1 - r0]
    return S
