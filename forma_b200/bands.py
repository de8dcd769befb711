"""Tile-band partition of a frame over the GPUs of one box.

The sort key's most significant field is tile_y and winding carries only
propagate inside a tile row (cpu/painter/mod.rs:501-567), so contiguous bands
of tile rows are independent: rank r rasterizes the geometry that can reach
its band (lines entirely above/below are culled), sorts and paints only its
band, and one all-gather of the band rows assembles the frame. No halo, no
reduction.
"""
from __future__ import annotations

from dataclasses import dataclass

TILE = 16


@dataclass(frozen=True)
class Band:
    rank: int
    world: int
    rows_per_band: int   # tile rows per rank (last band may be shorter / empty)
    tile_row0: int
    tile_row1: int       # exclusive
    y0: int              # pixel rows [y0, y1) of the frame
    y1: int
    padded_height: int   # world * rows_per_band * 16 (>= height): all-gather buffer rows

    @property
    def empty(self) -> bool:
        return self.y1 <= self.y0


def band_of(height: int, world: int, rank: int) -> Band:
    tiles_y = (height + TILE - 1) // TILE
    per = (tiles_y + world - 1) // world
    r0 = min(rank * per, tiles_y)
    r1 = min((rank + 1) * per, tiles_y)
    return Band(rank, world, per, r0, r1, r0 * TILE, min(r1 * TILE, height), world * per * TILE)


def balanced_rows(costs, world: int):
    """Band boundaries (tile rows) that give every rank about the same share of `costs`, the
    per-tile-row cost of a previous frame (forma_renderer_row_costs): boundary k is the first
    row at which the running cost reaches k / world of the total. Returns world + 1 rows,
    non-decreasing, from 0 to len(costs). Deterministic: every rank computes the same table."""
    n = len(costs)
    total = float(sum(costs))
    if n == 0 or total <= 0.0 or world <= 1:
        per = (n + world - 1) // max(world, 1)
        return [min(k * per, n) for k in range(world)] + [n]
    bounds, run, k = [0], 0.0, 1
    for row in range(n):
        # close band k - 1 before `row` when taking the row would overshoot more than stopping here undershoots
        while k < world and run + 0.5 * float(costs[row]) >= total * k / world:
            bounds.append(row)
            k += 1
        run += float(costs[row])
    while len(bounds) < world:
        bounds.append(n)
    bounds.append(n)
    return bounds


def balanced_band(height: int, world: int, rank: int, costs) -> Band:
    tiles_y = (height + TILE - 1) // TILE
    rows = balanced_rows(list(costs)[:tiles_y] + [0] * max(0, tiles_y - len(costs)), world)
    r0, r1 = rows[rank], rows[rank + 1]
    per = max(b - a for a, b in zip(rows[:-1], rows[1:]))
    return Band(rank, world, per, r0, r1, r0 * TILE, min(r1 * TILE, height), world * per * TILE)


def gather_frame(band_rows, out, dist=None):
    """All-gathers every rank's band (a contiguous (rows_per_band*16, stride)
    uint8 tensor, CPU or CUDA) into `out` ((padded_height, stride)); the first
    `height` rows of `out` are the frame. The only collective of the path."""
    if dist is None:
        import torch.distributed as dist  # noqa: PLC0415
    dist.all_gather_into_tensor(out.view(-1), band_rows.reshape(-1))
    return out
