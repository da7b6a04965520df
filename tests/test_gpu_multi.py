"""GPU, NCCL, world size = min(device_count, 4): the batch-sharded evaluation on REAL GPUs — the detections gathered over
NCCL (all_gather of counts + grouped send/recv of the rows to rank 0) must be torch.equal to the 1-GPU result for the same
seeded images (src/predict.py:116-179 is single-process).  KN-init weights, both threshold modes.
Skipped on a single-GPU box; bench.py --gpus N asserts the same property in its eval-pipeline leg."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg_path, out_dir):
    import torch.distributed as dist
    from conftest import make_darknet
    from modelcompression_b200.eval import evaluate_sharded
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        model = make_darknet(cfg_path, seed=0, kn=True, randbn=True, device=dev)
        n, B = 37, 8
        g = torch.Generator(device=dev).manual_seed(4)  # the same image set on every rank
        images = torch.randint(0, 256, (n, 3, 416, 416), dtype=torch.uint8, device=dev, generator=g)

        def get_batch(lo, hi):
            return images[lo:hi].contiguous()

        for conf, oo, val in ((0.25, 1, False), (0.005, 0, True)):
            got = evaluate_sharded(model, get_batch, n, B, conf, 0.45, oo, rank, world, validation=val)
            everyone = evaluate_sharded(model, get_batch, n, B, conf, 0.45, oo, rank, world, validation=val, gather='all')
            alone = evaluate_sharded(model, get_batch, n, B, conf, 0.45, oo, 0, 1, validation=val, gather=False)
            assert alone.shape[0] > 0 and bool(torch.all(alone[1:, 0] >= alone[:-1, 0]))
            assert torch.equal(everyone, alone), "all-rank gather differs from the 1-GPU result (rank %d)" % rank
            if rank == 0:
                assert torch.equal(got, alone), "gathered detections differ from the 1-GPU result"
            else:
                assert got.shape == (0, 8)
        dist.barrier()
        open(os.path.join(out_dir, 'ok%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_nccl_gathered_detections_equal_single_gpu(cfg_path, tmp_path):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    mp.spawn(_worker, args=(world, _free_port(), cfg_path, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), 'ok%d' % r)) for r in range(world))
