"""N > 1 host logic on CPU: two gloo ranks shard a frame list, each computes its frames (with the CPU oracle standing
in for a GPU), rank 0 reassembles in POC order; the result equals the single-process run and the timing reduction is a
MAX over ranks -- the same plumbing bench.py uses under torchrun (no data-path collective)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mipb200 import frames, shard


def test_round_robin_and_merge():
    assert shard.frames_for_rank(7, 0, 2) == [0, 2, 4, 6] and shard.frames_for_rank(7, 1, 2) == [1, 3, 5]
    assert shard.frames_for_rank(3, 3, 8) == []
    merged = shard.merge_in_poc_order([{0: "a", 2: "c"}, {1: "b"}], 3)
    assert merged == ["a", "b", "c"]
    with pytest.raises(ValueError):
        shard.merge_in_poc_order([{0: "a"}, {0: "b"}], 2)
    with pytest.raises(ValueError):
        shard.merge_in_poc_order([{0: "a"}], 2)


def _worker(rank, world, port, n_frames, q):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "vvc-mip-gpu_b200"), os.path.join(root, "oracle")):
        sys.path.insert(0, p)
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.frames_for_rank(n_frames, rank, world)
    sums = {poc: int(O.run_frame(frames.noise_frame(128, 64, 50 + poc), threads=1).astype(np.int64).sum()) for poc in mine}
    t = torch.tensor([float(len(mine))], dtype=torch.float64)   # stand-in for the per-rank device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, sums)
    dist.barrier()
    if rank == 0:
        q.put((shard.merge_in_poc_order(gathered, n_frames), float(t.item())))
    dist.destroy_process_group()


def test_two_ranks_equal_one(oracle):
    n = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [int(oracle.run_frame(frames.noise_frame(128, 64, 50 + poc), threads=1).astype(np.int64).sum()) for poc in range(n)]
    assert merged == want
    assert tmax == 3.0      # rank 0 owns 3 of the 5 frames: MAX over ranks
