"""CPU: face sharding over ranks (world_size 2, gloo) — host-side logic of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hifidiff_b200.sampler import sample_sharded, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 5, 8, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(4096, 8, 3) == (1536, 2048)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_faces, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def run_local(lo, hi):
        # stand-in for the GPU sampler: a function of the GLOBAL face index only
        idx = torch.arange(lo, hi, dtype=torch.float32)
        return idx[:, None, None, None] * torch.ones(1, 4, 2, 2) + 0.5

    full = sample_sharded(run_local, n_faces)
    local = sample_sharded(run_local, n_faces, gather=False)
    torch.save({"full": full, "local": local}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_faces", [1, 7, 8])  # 1: the second rank owns an EMPTY shard
def test_sample_sharded_world2_gloo(tmp_path, n_faces):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_faces, str(tmp_path)), nprocs=world, join=True)
    want = torch.arange(n_faces, dtype=torch.float32)[:, None, None, None] * torch.ones(1, 4, 2, 2) + 0.5
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(got["full"], want)            # every rank holds all faces in global order
        lo, hi = shard_bounds(n_faces, world, r)
        assert torch.equal(got["local"], want[lo:hi])


def test_sample_sharded_without_process_group():
    out = sample_sharded(lambda lo, hi: torch.arange(lo, hi)[:, None].float(), 5)
    assert out.flatten().tolist() == [0, 1, 2, 3, 4]
