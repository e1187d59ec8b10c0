"""-m gpu, needs >= 2 GPUs (skipped otherwise): the sharded extraction path of SURVEY 8e on real NCCL.

Two processes (one per GPU) run `egonn_b200.parallel.extract_sharded` on the same batch of clouds with the ENGINE's own
communicator (`egn_comm_*` / `egn_allgather_global` of the C ABI): greedy shards by voxel count, forward of the rank's
share, ONE all-gather of the global descriptors, original cloud order restored.  Rank 0 then runs the whole batch on its
GPU alone: the gathered (B, 256) must equal the single-GPU result (<= 5e-5: the per-cloud pooling sums are sliced
differently when the batch composition changes), and each rank's local outputs must be those of its own clouds.
The single-device original of this loop is eval/evaluate.py:454-466."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import numpy as np
    import torch.distributed as dist
    import egonn_b200 as E
    from egonn_b200 import parallel, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sd = torch.load(os.path.join(REPO, "tests", "golden", "egonn_weights.pth"), map_location="cpu", weights_only=True)
    mp_ = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.3)
    model = E.model_factory(mp_)
    model.load_state_dict(sd)
    model = model.eval().to(dev)
    # 7 clouds of very different sizes (uneven shares: 4 + 3), Apollo-shaped scans + small uniform clouds
    clouds = [synth.make_cloud("cfg4", 1), synth.uniform_cloud(3000, 2), synth.make_cloud("cfg4", 3), synth.uniform_cloud(500, 4),
              synth.uniform_cloud(20000, 5), synth.make_cloud("cfg4", 6), synth.uniform_cloud(9000, 7)]
    coords = [mp_.quantizer(torch.from_numpy(pc).to(dev))[0] for pc in clouds]
    comm = parallel.Communicator(dev)
    g_all, local = parallel.extract_sharded(model, coords, E.batched_coordinates, comm=comm)
    torch.cuda.synchronize()
    assert g_all.shape == (len(clouds), 256)
    parts = local["parts"]
    assert sorted(i for p in parts for i in p) == list(range(len(clouds))) and local["cloud_ids"] == parts[rank]
    # single-GPU reference of the whole batch on this rank's device
    bc = E.batched_coordinates(coords)
    full = model.forward_packed({"coords": bc, "features": torch.ones((bc.shape[0], 1), device=dev)})
    torch.cuda.synchronize()
    err = float((g_all - full["global"]).abs().max() / full["global"].abs().max())
    assert err <= 5e-5, f"rank {rank}: gathered global descriptors differ from the single-GPU run: {err:.2e}"
    # this rank's local outputs are those of its own clouds (same rows, same order, per cloud)
    off_l = local["local_offsets"].cpu().numpy()
    off_f = full["local_offsets"].cpu().numpy()
    for j, i in enumerate(local["cloud_ids"]):
        a = local["descriptors"][off_l[j]:off_l[j + 1]]
        b = full["descriptors"][off_f[i]:off_f[i + 1]]
        assert a.shape == b.shape and float((a - b).abs().max()) <= 5e-5, f"cloud {i}"
        ca = local["local_coords"][off_l[j]:off_l[j + 1], 1:]
        cb = full["local_coords"][off_f[i]:off_f[i + 1], 1:]
        assert torch.equal(ca, cb)
    # the collective on a dedicated communication stream (what bench.py does: the compute stream never waits for other ranks)
    side = torch.cuda.Stream(device=dev)
    sb = parallel.ShardedBatch(coords, E.batched_coordinates, rank, world)
    g_side, _ = parallel.run_sharded(model, sb, comm=comm, comm_stream=side)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert float((g_side - g_all).abs().max()) == 0.0
    # a second collective on the same communicator (steady-state use), padded shares
    g2 = parallel.gather_global(full["global"][torch.tensor(parts[rank], device=dev)], parts, comm=comm)
    assert float((g2 - full["global"]).abs().max()) == 0.0
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"err": err, "loads": [sum(int(coords[i].shape[0]) for i in p) for p in parts]}, out_path)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_sharded_extraction_two_gpus_nccl(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    print("\n[2-GPU NCCL] gathered vs single-GPU global descriptors: rel err %.2e; voxels per rank %s" % (res["err"], res["loads"]))
    assert res["err"] <= 5e-5


def _pipeline_worker(rank, world, port, out_path):
    """The pipelined extractor with the collective (parallel.OrderedGatherer: one communicator, one communication stream,
    batch order) on two GPUs: every rank extracts its own batches with two streams / host threads; ``global_all`` of batch i
    must be the concatenation of both ranks' ``global`` of batch i, whatever the order in which the threads finish."""
    import sys
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    import egonn_b200 as E
    from egonn_b200 import parallel, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sd = torch.load(os.path.join(REPO, "tests", "golden", "egonn_weights.pth"), map_location="cpu", weights_only=True)
    mp_ = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.3)
    model = E.model_factory(mp_)
    model.load_state_dict(sd)
    model = model.eval().to(dev)
    comm = parallel.Communicator(dev)
    # 9 batches of 2 clouds; sizes differ between the ranks and between batches, so the ranks' threads drift apart
    batches = [[synth.uniform_cloud(2000 + 900 * ((i + 3 * rank) % 5), 100 * rank + 2 * i),
                synth.uniform_cloud(1500 + 2500 * ((i + rank) % 3), 100 * rank + 2 * i + 1)] for i in range(9)]
    ext = E.Extractor(model, streams=2, topk=32, device=dev, comm=comm)
    mine, gathered = [], []
    for res in ext.extract(iter(batches)):
        mine.append(res["global"])
        gathered.append(res["global_all"])
    ext.close()
    both = [torch.zeros((world * 2, 256), device=dev) for _ in batches]
    for i, g in enumerate(mine):                                     # reference exchange through torch.distributed
        dist.all_gather_into_tensor(both[i], g.to(dev).contiguous())
    torch.cuda.synchronize()
    worst = max(float((a - b.cpu()).abs().max()) for a, b in zip(gathered, both))
    assert worst == 0.0, f"rank {rank}: global_all differs from the per-batch concatenation of the ranks' descriptors ({worst})"
    if rank == 0:
        torch.save({"batches": len(batches)}, out_path)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_pipelined_extractor_with_collective_two_gpus_nccl(tmp_path):
    """Not yet run on hardware when it was written (the round's GPU budget was spent): needs `gpurun --gpus 2`."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_pipeline_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out)["batches"] == 9
