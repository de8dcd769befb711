"""The N>1 path on CPU: two gloo ranks each render their tile band (with the
CPU oracle standing in for the device renderer — host-side logic only) and
all-gather the bands; the result must equal the single-rank frame."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 320, 200  # 12.5 tile rows: the last band is partial


def _scene(api):
    import synth
    comp = api.Composition()
    synth.random_mixed(api, comp, 150, W, H, 77)
    return comp


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from forma_b200 import bands
    from forma_b200.binding import RGBA, Color, Rect
    from oracle import oracle
    api = oracle.load()
    comp = _scene(api)
    band = bands.band_of(H, world, rank)
    stride = W * 4
    fb = np.zeros((band.padded_height, stride), np.uint8)
    if not band.empty:
        api.Renderer().render(comp, fb.reshape(-1), W, H, RGBA, Color(1, 1, 1, 0), Rect((0, W), (band.y0, band.y1)), None, stride)
    mine = torch.from_numpy(fb[band.tile_row0 * 16:(band.tile_row0 + band.rows_per_band) * 16].copy())
    if mine.shape[0] < band.rows_per_band * 16:  # ranks past the end of the frame contribute padding
        mine = torch.zeros((band.rows_per_band * 16, stride), dtype=torch.uint8)
    out = torch.zeros((band.padded_height, stride), dtype=torch.uint8)
    bands.gather_frame(mine, out, dist)
    if rank == 0:
        np.save(out_path, out.numpy()[:H])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_band_split_equals_single_frame(tmp_path, world):
    sys.path.insert(0, ROOT)
    from forma_b200.binding import RGBA, Color
    from oracle import oracle
    api = oracle.load()
    ref = np.zeros(W * H * 4, np.uint8)
    api.Renderer().render(_scene(api), ref, W, H, RGBA, Color(1, 1, 1, 0))
    out_path = str(tmp_path / "frame.npy")
    port = 29500 + (os.getpid() % 1000) + world
    mp.spawn(_worker, args=(world, port, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    assert np.array_equal(got.reshape(-1), ref)


def _shared_worker(rank, world, port, out_path):
    """bench.py's multi-GPU end-to-end scheme on CPU: cost-balanced bands (every rank computes
    the same table), every rank renders its band straight into ONE host frame shared by all
    ranks (bench.SharedHostFrame: /dev/shm mapping), no collective on the data path."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from forma_b200 import bands
    from forma_b200.binding import RGBA, Color, Rect
    from oracle import oracle
    api = oracle.load()
    comp = _scene(api)
    tiles_y = (H + 15) // 16
    costs = [1 + (r * 37 % 11) * (5 if r < 4 else 1) for r in range(tiles_y)]  # any deterministic table
    band = bands.balanced_band(H, world, rank, costs)
    table = [bands.balanced_band(H, world, r, costs) for r in range(world)]
    assert table[0].tile_row0 == 0 and table[-1].tile_row1 == tiles_y
    assert all(a.tile_row1 == b.tile_row0 for a, b in zip(table, table[1:]))
    host = bench.SharedHostFrame(torch, dist, H * W * 4, rank, world, "test")
    if not band.empty:
        api.Renderer().render(comp, host.np, W, H, RGBA, Color(1, 1, 1, 0), Rect((0, W), (band.y0, band.y1)), None, W * 4)
    dist.barrier()
    if rank == 0:
        np.save(out_path, np.array(host.np).reshape(H, W * 4))
    dist.barrier()
    host.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_balanced_bands_into_one_shared_host_frame(tmp_path, world):
    sys.path.insert(0, ROOT)
    from forma_b200.binding import RGBA, Color
    from oracle import oracle
    api = oracle.load()
    ref = np.zeros(W * H * 4, np.uint8)
    api.Renderer().render(_scene(api), ref, W, H, RGBA, Color(1, 1, 1, 0))
    out_path = str(tmp_path / "frame.npy")
    port = 29700 + (os.getpid() % 1000) + world
    mp.spawn(_shared_worker, args=(world, port, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    assert np.array_equal(got.reshape(-1), ref)


def test_balanced_rows_properties():
    from forma_b200 import bands
    import random
    rng = random.Random(5)
    for _ in range(200):
        n, world = rng.randint(1, 300), rng.randint(1, 9)
        costs = [rng.random() * (20 if rng.random() < 0.1 else 1) for _ in range(n)]
        rows = bands.balanced_rows(costs, world)
        assert len(rows) == world + 1 and rows[0] == 0 and rows[-1] == n and all(a <= b for a, b in zip(rows, rows[1:]))
        total, shares = sum(costs), [sum(costs[a:b]) for a, b in zip(rows, rows[1:])]
        # no band exceeds its fair share by more than the largest single row (rows are indivisible)
        assert max(shares) <= total / world + max(costs) + 1e-9
    assert bands.balanced_rows([], 4) == [0, 0, 0, 0, 0]
    assert bands.balanced_rows([0, 0, 0, 0], 2) == [0, 2, 4]
