"""⑤ multi-GPU host logic on CPU: world_size-2 gloo processes run the sharding + all-gather of global
descriptors (the only collective of the path).  The forward itself is replaced by a deterministic stand-in
because there is no GPU here; the GPU path of the same code runs under bench.py --gpus N."""
import os
import socket

import pytest
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


def test_ordered_gatherer_issues_in_ticket_order_whatever_the_submission_order():
    """Host logic of the multi-GPU collective path: worker threads submit in arbitrary order, ONE thread issues in ticket
    order (the CUDA plumbing is replaced by no-ops here; NCCL itself runs in tests/test_gpu_multi.py)."""
    import random
    import threading
    import time
    from egonn_b200.parallel import OrderedGatherer

    class CpuGatherer(OrderedGatherer):
        def _make_stream(self):
            return None

        def _record_event(self):
            return None

        def _run_on_stream(self, ev, fn, tensors):
            return fn()

    issued = []
    g = CpuGatherer(torch.device("cpu"), first_ticket=10)
    n, S = 40, 4

    def worker(t):
        rnd = random.Random(t)
        for i in range(10 + t, 10 + n, S):
            time.sleep(rnd.random() * 0.003)
            if i == 23:
                g.submit(i, None)                               # a failed step still consumes its ticket
            else:
                g.submit(i, lambda i=i: issued.append(i) or i * 2)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(S)]
    for th in threads:
        th.start()
    assert g.result(12) == 24                                   # waits for tickets 10, 11, 12
    for th in threads:
        th.join()
    g.drain(10 + n)
    assert issued == [i for i in range(10, 10 + n) if i != 23]
    assert g.result(49) == 98 and g.result(23) is None
    g.forget_results()
    g.submit(50, lambda: (_ for _ in ()).throw(RuntimeError("boom")))
    try:
        g.drain(51)
        raised = False
    except RuntimeError:
        raised = True
    assert raised
    g.close()


def _ordered_worker(rank, world, port, out):
    """Four worker threads per rank finish their steps in rank-dependent random order; every collective is a real (host
    blocking) gloo all-gather issued by the rank's gather thread.  Any order mismatch between the ranks would pair up
    tensors of different steps (caught below) or block for ever (caught by the timeout of the test)."""
    import random
    import threading
    import time
    from egonn_b200.parallel import OrderedGatherer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class CpuGatherer(OrderedGatherer):
        def _make_stream(self):
            return None

        def _record_event(self):
            return None

        def _run_on_stream(self, ev, fn, tensors):
            return fn()

    g = CpuGatherer(torch.device("cpu"))
    n, S = 48, 4

    def gather(i):
        send = torch.tensor([[float(rank), float(i)]])
        recv = torch.empty((world, 2))
        dist.all_gather_into_tensor(recv, send)
        return recv

    def worker(t):
        rnd = random.Random(1000 * rank + t)
        for i in range(t, n, S):
            time.sleep(rnd.random() * 0.004)                    # the ranks' threads drift apart
            g.submit(i, lambda i=i: gather(i))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(S)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    g.drain(n)
    ok = True
    for i in range(n):
        r = g.result(i)
        ok = ok and r[:, 0].tolist() == [float(k) for k in range(world)] and r[:, 1].tolist() == [float(i)] * world
    g.close()
    if rank == 0:
        torch.save(ok, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_ordered_gatherer_pairs_up_collectives_across_two_ranks(tmp_path):
    out = str(tmp_path / "ok.pt")
    mp.spawn(_ordered_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out) is True


def test_ordered_gatherer_default_cuda_plumbing_with_a_fake_torch_cuda(monkeypatch):
    """Every line of the DEFAULT plumbing (stream creation, event hand-over, collective issued under the communication
    stream, record_stream of the inputs) executed against a recording stand-in of ``torch.cuda`` - there is no GPU here; the
    real thing runs under ``bench.py --gpus N`` and tests/test_gpu_multi.py."""
    import contextlib
    import threading
    from egonn_b200 import parallel
    log, state = [], threading.local()

    class FakeStream:
        def __init__(self, name):
            self.name = name

        def wait_event(self, ev):
            log.append(("wait", self.name, ev.recorded_on))

    class FakeEvent:
        recorded_on = None

        def record(self, stream):
            self.recorded_on = stream.name

    class FakeTensor:
        def __init__(self):
            self.streams = []

        def record_stream(self, s):
            self.streams.append(s.name)

    @contextlib.contextmanager
    def fake_stream_ctx(s):
        prev = getattr(state, "cur", None)
        state.cur = s
        try:
            yield
        finally:
            state.cur = prev

    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: FakeStream("comm"))
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: getattr(state, "cur", None) or FakeStream("default"))

    g = parallel.OrderedGatherer(torch.device("cuda", 0))
    tensors = [FakeTensor() for _ in range(6)]

    def worker(t):
        with torch.cuda.stream(FakeStream(f"compute{t}")):
            for i in range(t, 6, 2):
                g.submit(i, lambda i=i: (torch.cuda.current_stream().name, i), tensors[i])

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    g.drain(6)
    assert [g.result(i) for i in range(6)] == [("comm", i) for i in range(6)]          # issued on the communication stream, in order
    assert [e for e in log if e[0] == "wait"] == [("wait", "comm", f"compute{i % 2}") for i in range(6)]   # behind the step's event
    assert all(t.streams == ["comm"] for t in tensors)
    g.close()
