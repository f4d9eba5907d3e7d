"""Host-side sharding logic on CPU: world_size-2 gloo processes (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import harness as H
from canvas_ity_b200 import sharding


@pytest.mark.parametrize("height,world", [(4096, 1), (4096, 2), (4096, 8), (757, 4), (5, 8), (256, 3)])
def test_bands_partition_the_rows(height, world):
    covered = []
    for r in range(world):
        y0, rows = sharding.band(height, r, world)
        covered += list(range(y0, y0 + rows))
    assert covered == list(range(height))
    sizes = [sharding.band(height, r, world)[1] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_round_robin_covers_every_canvas_once():
    for n, world in ((16384, 8), (10, 4), (3, 8)):
        seen = sorted(i for r in range(world) for i in sharding.my_items(n, r, world))
        assert seen == list(range(n))


def _worker(rank, world, port, height, width, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    image = torch.from_numpy(H.golden_rgba8("tiger_512").copy())[:height, :width]
    y0, rows = sharding.band(height, rank, world)
    whole = sharding.gather_bands(image[y0:y0 + rows].contiguous(), height, width)
    ok = torch.equal(whole, image)
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([int(ok)]))
    dist.destroy_process_group()


def test_gather_bands_gloo_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 511, 512, str(tmp_path)), nprocs=2, join=True)     # odd height: uneven bands
    assert all(int(np.load(tmp_path / ("ok_%d.npy" % r))[0]) == 1 for r in range(2))
