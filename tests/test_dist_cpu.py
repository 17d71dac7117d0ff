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


def _arc(bins, N):
    """The arc of the spectrum a set of channels reads (what ka9q_stream_needed_bins returns, restated): every channel
    reads bins [k-1023, k+1024] mod N; the arc is the shortest circular interval covering them, on 16-bin boundaries."""
    signed = sorted((b % N) - N if (b % N) > N // 2 else (b % N) for b in bins)
    lo, hi = signed[0] - 1023, signed[-1] + 1024
    lo -= lo % 16
    hi += (-(hi + 1)) % 16
    return lo, hi - lo + 1


def test_contiguous_sharding_gives_every_rank_one_short_arc():
    """DESIGN.md section 7 relies on it: with frequency-contiguous shards a rank reads 1/G of the spectrum plus a
    2048-bin halo, so that is all the exchange has to deliver (2.5 MB per block and rank at 8 GPUs, not 21 MB)."""
    plan = workloads.cfg5()
    N = plan.N
    allbins = sorted(c.bin for c in plan.channels)
    for world in (1, 2, 4, 8):
        shards = [workloads.shard_contiguous(plan, r, world) for r in range(world)]
        assert sorted(c.bin for s in shards for c in s) == allbins                 # a partition
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1      # balanced
        prev_hi = None
        for s in shards:
            sg = sorted((c.bin % N) - N if (c.bin % N) > N // 2 else (c.bin % N) for c in s)
            if prev_hi is not None:
                assert sg[0] > prev_hi                                             # shards are ordered in frequency
            prev_hi = sg[-1]
            lo, ln = _arc([c.bin for c in s], N)
            assert ln % 16 == 0 and lo % 16 == 0
            spacing = N / len(plan.channels)
            assert ln <= N / world + 2048 + 32 + spacing, (world, ln)              # 1/G of the band + the halo
        if world == 8:
            assert 8 * _arc([c.bin for c in shards[3]], N)[1] * 8 < 2.6e6 * 8 * 1.1   # bytes per block and rank ~2.5 MB


def _xchg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a toy spectrum exchange with the library's protocol shape: blocks sharded by rank, arcs by consumer
        N, B = 4096, 4
        rng = np.random.default_rng(5)
        spec_true = rng.standard_normal((B, N)).astype(np.float32)     # what one GPU would have computed
        plan_bins = [-1500, -900, -300, 200, 700, 1300]                # 6 channels, windows of +-64 bins here
        half = 64
        mine = plan_bins[rank * 3:(rank + 1) * 3]
        lo, hi = min(mine) - half, max(mine) + half
        arcs = [None] * world
        dist.all_gather_object(arcs, (lo, hi))
        cnt = B // world
        have = np.full((B, N), np.nan, dtype=np.float32)
        have[rank * cnt:(rank + 1) * cnt] = spec_true[rank * cnt:(rank + 1) * cnt]      # my FFT blocks
        peer = 1 - rank
        plo, phi = arcs[peer]
        idx_peer = np.arange(plo, phi + 1) % N
        send = torch.from_numpy(np.ascontiguousarray(have[rank * cnt:(rank + 1) * cnt][:, idx_peer]))
        idx_me = np.arange(lo, hi + 1) % N
        recv = torch.empty((cnt, idx_me.size), dtype=torch.float32)
        reqs = [dist.isend(send, peer), dist.irecv(recv, peer)]
        for r_ in reqs:
            r_.wait()
        have[peer * cnt:(peer + 1) * cnt][:, idx_me] = recv.numpy()
        ok = all(np.array_equal(have[:, np.arange(k - half, k + half + 1) % N], spec_true[:, np.arange(k - half, k + half + 1) % N])
                 for k in mine)
        q.put((rank, ok, int(np.isnan(have).sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_arc_exchange_protocol_shape():
    """Blocks sharded by producer, arcs by consumer (ka9q_stream_mgpu_*, restated with gloo send/recv on the CPU): after
    one exchange every rank holds, for every block of the batch, exactly the bins its own channels read."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_xchg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(nan > 0 for _, _, nan in res)      # and nothing like the whole spectrum travelled
