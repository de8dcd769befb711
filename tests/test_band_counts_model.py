"""NumPy model of the per-row segment counts a band render reports (row_cost_kernel in
forma_b200/csrc/kernels_tables.cu): the sorted 64-bit pixel segments carry tile_y + 1 in their
top 11 bits; row r of the frame owns the keys in [lower((r + 1) << 53), lower((r + 2) << 53)),
the first row also takes the segments above the frame (tile_y clamped to -1, field 0), the last
row those below it. Whatever the band boundaries, the bands' own rows then add up to the frame's
segment count, which is what Timings.n_segments of a multi-device / sliced frame returns."""
import numpy as np


def row_counts(keys: np.ndarray, tiles_y: int) -> np.ndarray:
    n = len(keys)
    lower = lambda k: int(np.searchsorted(keys, np.uint64(k), side="left"))  # noqa: E731
    out = np.zeros(tiles_y, np.int64)
    for row in range(tiles_y):
        s0, s1 = lower((row + 1) << 53), lower((row + 2) << 53)
        out[row] = (n if row + 1 == tiles_y else s1) - (0 if row == 0 else s0)
    return out


def test_rows_partition_the_segments_for_any_band_split():
    rng = np.random.default_rng(5)
    tiles_y = 135
    # tile_y + 1 in [0, tiles_y + 3]: above the frame (0), inside, and below it
    ty = rng.integers(0, tiles_y + 4, size=200_000, dtype=np.uint64)
    low = rng.integers(0, 1 << 53, size=ty.size, dtype=np.uint64)
    keys = np.sort((ty << np.uint64(53)) | low)
    counts = row_counts(keys, tiles_y)
    assert counts.sum() == len(keys)
    inside = np.bincount((keys >> np.uint64(53)).astype(np.int64), minlength=tiles_y + 4)
    assert counts[0] == inside[0] + inside[1]                      # above the frame + row 0
    assert counts[-1] == inside[tiles_y:].sum()                    # last row + everything below
    assert np.array_equal(counts[1:-1], inside[2:tiles_y])
    for n_bands in (1, 2, 3, 8, 16):
        bounds = sorted({0, tiles_y, *rng.integers(1, tiles_y, size=n_bands - 1).tolist()})
        own = [counts[a:b].sum() for a, b in zip(bounds[:-1], bounds[1:])]
        assert sum(own) == len(keys)


def test_empty_and_single_row_frames():
    assert row_counts(np.zeros(0, np.uint64), 4).sum() == 0
    keys = np.sort(np.array([0, 1 << 53, (1 << 53) + 5, 9 << 53], dtype=np.uint64))
    assert row_counts(keys, 1).tolist() == [4]
