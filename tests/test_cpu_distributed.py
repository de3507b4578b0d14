"""CPU tier: the multi-rank path (world_size 2, gloo): index sharding + chunked all-gather
reassemble the global batch in trajectory order.  The per-chunk "kernels" are faked with
index-stamped tensors; the GPU box runs the same code over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, group, chunks, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from drone_path_planning_python_b200.distributed import ChunkedAllGather, shard_bounds
    lo, hi = shard_bounds(total, world, rank, group)
    count = hi - lo
    n, K, S = 3, 2, 5
    templates = [torch.empty((count, n, K, 8), dtype=torch.float64), torch.empty((count, S), dtype=torch.uint8),
                 torch.empty((count,), dtype=torch.uint8)]
    gather = ChunkedAllGather(count, world, chunks, templates, group)
    calls = []

    def compute(clo, chi):
        calls.append((clo, chi))
        gidx = torch.arange(lo + clo, lo + chi, dtype=torch.float64)
        coef = gidx.view(-1, 1, 1, 1).expand(-1, n, K, 8).contiguous() + 0.25
        hit = (gidx.to(torch.int64) % 7).to(torch.uint8).view(-1, 1).expand(-1, S).contiguous()
        return coef, hit, (gidx.to(torch.int64) % 2).to(torch.uint8)

    coef, hit, any_hit = gather.run(compute)
    assert all((b - a) % group == 0 for a, b in calls) and calls[0][0] == 0 and calls[-1][1] == count
    np.save(os.path.join(out_dir, "coef_%d.npy" % rank), coef.numpy())
    np.save(os.path.join(out_dir, "hit_%d.npy" % rank), hit.numpy())
    np.save(os.path.join(out_dir, "any_%d.npy" % rank), any_hit.numpy())
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    world, total, group, chunks = 2, 40, 5, 3
    mp.spawn(_worker, args=(world, _free_port(), total, group, chunks, str(tmp_path)), nprocs=world, join=True)
    idx = np.arange(total)
    for rank in range(world):
        coef = np.load(tmp_path / ("coef_%d.npy" % rank))
        hit = np.load(tmp_path / ("hit_%d.npy" % rank))
        any_hit = np.load(tmp_path / ("any_%d.npy" % rank))
        assert coef.shape == (total, 3, 2, 8)
        assert np.array_equal(coef[:, 0, 0, 0], idx + 0.25)       # global trajectory order on every rank
        assert np.array_equal(hit[:, 0], idx % 7) and np.array_equal(any_hit, idx % 2)
