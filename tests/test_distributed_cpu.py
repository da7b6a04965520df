"""CPU, gloo, world size 2: the host-side logic of the batch-sharded evaluation (shard ranges + the detection gather:
one all_gather of counts, then unpadded rows sent once to the destination rank).  The per-shard compute is CUDA-only
and is covered by test_gpu_eval.py; the same gather over NCCL on real GPUs by test_gpu_multi.py and bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modelcompression_b200.eval import DET_COLS, gather_detections, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_detections(lo, hi):
    """Deterministic ragged detections for images [lo, hi): image i has (i % 4) rows (so some images have none)."""
    rows = []
    for i in range(lo, hi):
        for j in range(i % 4):
            rows.append([float(i)] + [float(i * 10 + j + c) for c in range(DET_COLS - 1)])
    return torch.tensor(rows, dtype=torch.float32).view(-1, DET_COLS)


def _worker(rank, world, port, n_images, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_images, rank, world)
        local = _fake_detections(lo, hi)
        want = _fake_detections(0, n_images)
        got = gather_detections(local)  # default: delivered to rank 0 only
        if rank == 0:
            assert got.shape == want.shape, (got.shape, want.shape)
            assert torch.equal(got, want), "rank-major gather must equal the single-process result"
        else:
            assert got.shape == (0, DET_COLS)
        got1 = gather_detections(local, dst=1)
        assert torch.equal(got1, want) if rank == 1 else got1.shape == (0, DET_COLS)
        got = gather_detections(local, dst=None)  # every rank
        assert got.shape == want.shape and torch.equal(got, want)
        torch.save(got, os.path.join(out_dir, 'rank%d.pt' % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 1, 0, 64])
def test_gather_detections_world2(tmp_path, n_images):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_images, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), 'rank0.pt'))
    b = torch.load(os.path.join(str(tmp_path), 'rank1.pt'))
    assert torch.equal(a, b)  # every rank ends with the same, image-ordered table


def test_gather_is_identity_without_process_group():
    x = _fake_detections(0, 5)
    assert gather_detections(x) is x


def _dp_worker(rank, world, port):
    """N4 host logic over gloo: bucketed gradient averaging, parameter broadcast, running-statistics averaging."""
    from modelcompression_b200 import train_dp
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # different initial parameters per rank
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Conv2d(4, 2, 1))
        train_dp.broadcast_parameters(net, src=0)
        torch.manual_seed(100)
        ref = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Conv2d(4, 2, 1))
        for a, b in zip(net.parameters(), ref.parameters()):
            assert torch.equal(a, b)
        for i, p in enumerate(net.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        train_dp.allreduce_gradients(list(net.parameters()))
        for i, p in enumerate(net.parameters()):
            assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)))  # mean of (1, 2) * (i+1)
        net[1].running_mean.fill_(float(rank))
        train_dp.average_buffers(net)
        assert torch.allclose(net[1].running_mean, torch.full((4,), 0.5))
        assert int(net[1].num_batches_tracked) == 0  # integer buffers untouched
    finally:
        dist.destroy_process_group()


def test_data_parallel_helpers_world2():
    mp.spawn(_dp_worker, args=(2, _free_port()), nprocs=2, join=True)
