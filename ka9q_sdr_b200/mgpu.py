"""Host plumbing of the channel-sharded multi-GPU channelizer: one process per GPU, `torch.distributed` only carries the
set-up blobs (the NCCL id, every rank's spectrum arc, the CUDA-IPC handles); the data path is inside libka9q_b200.so
(csrc/mgpu.cu: peer-memory stores over NVLink, or grouped ncclSend/ncclRecv)."""
from __future__ import annotations

import torch.distributed as dist

from . import channelizer as ch


def setup_sharded(c: "ch.Channelizer", rank: int, world: int, transport: int = ch.MGPU_P2P) -> None:
    """Collective over the default process group (any backend): after it, c.mgpu_compute() runs sharded batches."""
    arcs = [None] * world
    dist.all_gather_object(arcs, c.needed_bins())
    blobs = None
    if transport == ch.MGPU_NCCL:
        ids = [ch.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        if world > 1:
            c.nccl_init(ids[0], rank, world)
    else:
        parts = [None] * world
        dist.all_gather_object(parts, c.mgpu_export())
        blobs = b"".join(parts)
    c.mgpu_setup(transport, rank, world, arcs, blobs)


def push_batch_share(c: "ch.Channelizer", iq_ptr_base: int, bytes_per_sample: int, first_block: int, nblocks: int,
                     period_samples: int) -> int:
    """Streaming input of a sharded run: copy to the device only the samples this rank's blocks of the batch need. The
    host buffer at iq_ptr_base holds `period_samples` samples of a periodic stimulus (sample n lives at n mod period).
    Returns the number of bytes copied."""
    import ctypes as C
    a, n = c.mgpu_input_range(first_block, nblocks)
    done = 0
    while done < n:
        pos = (a + done) % period_samples
        chunk = min(n - done, period_samples - pos)
        c.push_at(C.c_void_p(iq_ptr_base + pos * bytes_per_sample), a + done, chunk)
        done += chunk
    return n * bytes_per_sample
