"""⑤ multi-GPU host logic on CPU: world_size-2 gloo processes run the sharding + all-gather of global
descriptors (the only collective of the path).  The forward itself is replaced by a deterministic stand-in
because there is no GPU here; the GPU path of the same code runs under bench.py --gpus N."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egonn_b200.parallel import gather_global, shard_clouds


def test_shard_clouds_balanced_and_deterministic():
    sizes = [50, 10, 40, 30, 20, 60, 5]
    parts = shard_clouds(sizes, 3)
    assert sorted(i for p in parts for i in p) == list(range(7))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 20
    assert parts == shard_clouds(sizes, 3)
    assert shard_clouds([3, 2, 1], 8)[3:] == [[] for _ in range(5)]        # more ranks than clouds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = shard_clouds(sizes, world)
    # stand-in for the forward: descriptor of cloud i = i + arange(256)/1000
    local = torch.stack([i + torch.arange(256) / 1000.0 for i in parts[rank]]) if parts[rank] else torch.zeros((0, 256))
    full = gather_global(local, parts)
    if rank == 0:
        torch.save(full, out)
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_global_descriptors_world2(tmp_path):
    sizes = [100, 30, 70, 20, 90]                                          # 5 clouds on 2 ranks: uneven shares
    out = str(tmp_path / "full.pt")
    mp.spawn(_worker, args=(2, _free_port(), sizes, out), nprocs=2, join=True)
    full = torch.load(out)
    expect = torch.stack([i + torch.arange(256) / 1000.0 for i in range(5)])
    assert torch.equal(full, expect)


def test_sharded_batch_partition_and_imbalance():
    """Host logic of the strong-scaling path (bench.py --strong): the rank's share, its batched coordinates and the
    load imbalance, on CPU tensors (no forward)."""
    from egonn_b200.parallel import ShardedBatch
    sizes = [900, 100, 500, 400, 300, 800]
    clouds = [torch.zeros((n, 3), dtype=torch.int32) + i for i, n in enumerate(sizes)]

    def batched(cs):
        return torch.cat([torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32), c], dim=1) for b, c in enumerate(cs)])

    shares = [ShardedBatch(clouds, batched, r, 3) for r in range(3)]
    assert sorted(i for sb in shares for i in sb.mine) == list(range(6))
    assert all(sb.parts == shares[0].parts and sb.loads == shares[0].loads for sb in shares)
    for sb in shares:
        assert sb.coords.shape == (sum(sizes[i] for i in sb.mine), 4)
        assert sb.coords[:, 0].unique().tolist() == list(range(len(sb.mine)))          # batch indices restart at 0 on every rank
        first = sb.coords[sb.coords[:, 0] == 0][0, 1].item()
        assert first == sb.mine[0]                                                      # share keeps ascending original order
    assert 1.0 <= shares[0].imbalance <= 1.1
    empty = ShardedBatch(clouds[:2], batched, 3, 4)                                     # more ranks than clouds
    assert empty.mine == [] and empty.coords is None
