"""Multi-GPU host logic on CPU: shard ranges and the gloo world_size-2 gather of report records."""
import os
import socket

import numpy as np
import pytest

from photohive_dsp_b200.shard import gather_records, shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 512, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, rb, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    # each "record" is stamped with its global image index: the gather must restore global order
    local = np.zeros((hi - lo, rb), np.uint8)
    for i in range(lo, hi):
        local[i - lo] = np.frombuffer(np.array([i] * (rb // 8), np.int64).tobytes(), np.uint8)
    got = gather_records(local, n_total, rank, world)
    dist.barrier()
    if rank == 0:
        np.save(out_path, got)
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_gloo_world2_gather_restores_image_order(tmp_path, n_total):
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_total, 64, out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape == (n_total, 64)
    assert np.array_equal(got.view(np.int64)[:, 0], np.arange(n_total))
