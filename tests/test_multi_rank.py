"""N > 1 host logic on CPU: world-size-2 gloo processes shard a batch, run a row-wise function on their shard and gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdnq_b200.parallel import gather_batch, max_over_ranks, shard_batch, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 4, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                     # every rank builds the same global batch and "weights"
    x = torch.randn(n_images, 16, 32)
    w = torch.randn(24, 32)

    def rowwise(t):                                          # stand-in for the Linear stack: rows never interact
        s = t.abs().amax(dim=-1, keepdim=True) / 127
        return (torch.round(t / s) * s) @ w.t()
    y = gather_batch(rowwise(shard_batch(x)), n_images)
    slowest = max_over_ranks(1.0 + rank)
    if rank == 0:
        torch.save({"y": y, "ref": rowwise(x), "slowest": slowest}, os.path.join(out_dir, "out.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5])
def test_two_rank_shard_and_gather_matches_single_process(tmp_path, n_images):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_images, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(tmp_path, "out.pt"))
    assert got["y"].shape == got["ref"].shape
    assert torch.equal(got["y"], got["ref"])                 # sharding must not change a single bit
    assert got["slowest"] == 2.0
