"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path — interleaved tile ownership, QMC sample partition, and the
SUM reduce of per-rank framebuffers — without a GPU.  The device-side equivalent (4 emulated ranks on one GPU summing to the
single-GPU image bit-exactly) is tests/test_path_gpu.py::test_tiles_partition_the_image."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydracore_b200 import multigpu as MG


def test_tile_ownership_partitions_every_pixel_once():
    for (W, H, T, G) in ((128, 96, 32, 4), (1920, 1080, 32, 8), (70, 33, 16, 3), (8, 8, 32, 2)):
        owner = MG.tile_owner_map(W, H, T, G)
        seen = np.zeros(W*H, np.int32)
        for r in range(G):
            px = MG.owned_pixels(W, H, T, r, G)
            assert (owner.reshape(-1)[px] == r).all()
            seen[px] += 1
        assert (seen == 1).all()
    # load balance at 1080p over 8 ranks: interleaving keeps every rank within 2 % of the mean
    cnt = np.bincount(MG.tile_owner_map(1920, 1080, 32, 8).reshape(-1), minlength=8)
    assert cnt.max() - cnt.min() <= 0.02*cnt.mean()
    # QMC: sample slots partition the pass
    allq = np.concatenate([MG.qmc_sample_range(64, 48, r, 3) for r in range(3)])
    assert np.array_equal(np.sort(allq), np.arange(64*48))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, T, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank "renders" a deterministic per-pixel value on the pixels it owns and leaves the others at zero (SUM buffers)
    full = (np.arange(W*H*4, dtype=np.float32).reshape(-1, 4) % 977)*np.float32(0.25)
    part = np.zeros_like(full)
    px = MG.owned_pixels(W, H, T, rank, world)
    part[px] = full[px]
    t = torch.from_numpy(part.reshape(-1))
    MG.reduce_sums(dist, t, dst=0)
    # max-over-ranks timing plumbing used by bench.py
    tm = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = bool(np.array_equal(t.numpy().reshape(-1, 4), full)) and float(tm.item()) == float(world)
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_of_tile_partitioned_sums_over_gloo(tmp_path):
    out = str(tmp_path/"result.txt")
    mp.spawn(_worker, args=(2, _free_port(), 96, 80, 32, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
