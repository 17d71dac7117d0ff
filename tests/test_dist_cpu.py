"""Host-side multi-rank logic on CPU (gloo, world_size 2): channel sharding covers every channel exactly once, the
128-byte communicator id and a spectrum-shaped buffer broadcast from the ingest rank arrive intact, and the max-over-ranks
timing reduction bench.py uses behaves. The data path itself (NCCL broadcast of the device spectrum) needs GPUs and is
exercised by bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ka9q_sdr_b200 import workloads


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = workloads.cfg4(64)
        mine = workloads.shard_channels(plan, rank, world)
        # communicator id travels as a byte tensor (bench.py does the same before ka9q_stream_nccl_init)
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt.copy_(torch.arange(128, dtype=torch.uint8))
        dist.broadcast(idt, 0)
        # spectrum broadcast from the ingest rank: N complex64 per block
        spec = torch.zeros(2 * 4096, dtype=torch.float32)
        if rank == 0:
            spec.copy_(torch.arange(2 * 4096, dtype=torch.float32))
        dist.broadcast(spec, 0)
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counts = torch.tensor([len(mine)], dtype=torch.int64)
        dist.all_reduce(counts)
        q.put((rank, [c.bin for c in mine], idt.numpy().tobytes(), float(spec.sum()), float(t.item()), int(counts.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_broadcast():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plan = workloads.cfg4(64)
    all_bins = sorted(b for r in res for b in r[1])
    assert all_bins == sorted(c.bin for c in plan.channels)          # every channel exactly once
    assert abs(len(res[0][1]) - len(res[1][1])) <= 1                 # balanced
    assert res[0][2] == res[1][2] == bytes(range(128))               # communicator id intact
    assert res[0][3] == res[1][3] == float(np.arange(2 * 4096).sum())  # spectrum intact on every rank
    assert res[0][4] == res[1][4] == 2.0                             # max over ranks
    assert res[0][5] == 64


def test_shard_channels_is_a_partition_for_any_world_size():
    plan = workloads.cfg5(1000)
    for world in (1, 2, 3, 4, 8):
        shards = [workloads.shard_channels(plan, r, world) for r in range(world)]
        flat = sorted(c.bin for s in shards for c in s)
        assert flat == sorted(c.bin for c in plan.channels)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
