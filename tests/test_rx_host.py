"""Receive-side host plumbing (csrc/rx_host.c): the sequence-sorted packet queue of rtp_recv
(main.c:347-361), the ingest step of proc_samples behind it, the block ring, the receive / ingest threads over a real UDP
socket, and the sendmmsg egress. No GPU: the ring is plain memory here (page-locked on a GPU box)."""
import ctypes as C
import socket
import struct
import time

import numpy as np
import pytest

from ka9q_sdr_b200 import _lib, rtp
from test_rtp_glue import iq_datagram


class RxStats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("datagrams_queued", "inserted_out_of_order", "samples", "zero_filled", "ignored",
                                            "rtp_drops", "rtp_dupes")] + [("pinned", C.c_int)]


def _stats(L, rx):
    st = RxStats()
    L.ka9q_rx_get_stats(rx, C.byref(st))
    return st


def _ring_blocks(L, rx, nblocks, block_samples):
    p = L.ka9q_rx_peek_blocks(rx, nblocks, 0)
    assert p
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int16)), (2 * nblocks * block_samples,)).copy()
    assert L.ka9q_rx_consume(rx, nblocks) == 0
    return out


def test_sorted_queue_reorders_what_is_queued_together(ref):
    """Datagrams that sit in the queue at the same time leave it in sequence order (main.c:347-361); the ring then holds
    what the reference's own rtp_process / zero-fill logic produces for that order."""
    L = _lib.lib()
    rng = np.random.default_rng(3)
    n, bs = 240, 960
    seq0, ts0, ssrc = 65530, 1000, 77          # the sequence numbers wrap inside the test
    dg = [iq_datagram(rng, rtp.IQ_PT, seq0 + i, ts0 + i * n, ssrc, n) for i in range(24)]
    del dg[9]                                   # one lost packet: 240 zeros
    rx = L.ka9q_rx_create(rtp.IQ_S16, bs, 8)
    assert rx
    order = []
    for g0 in range(0, len(dg), 4):             # arrive in groups of four, each group shuffled; drained group by group
        grp = list(range(g0, min(g0 + 4, len(dg))))
        perm = list(rng.permutation(grp))
        for i in perm:
            raw = (C.c_ubyte * len(dg[i])).from_buffer_copy(dg[i])
            assert L.ka9q_rx_inject(rx, raw, len(dg[i])) == 0
        # the reference's insert is a plain >= on the 16-bit numbers: across the 65535 -> 0 wrap the "smaller" ones go first
        order += sorted(perm, key=lambda i: (seq0 + i + (1 if i >= 9 else 0)) & 0xFFFF)
        assert L.ka9q_rx_drain(rx) >= 0
    want = []
    g = ref.GlueIngest()
    for i in order:
        r, raw = g.datagram(dg[i])
        if r > 0:
            want.append(np.frombuffer(raw, dtype=np.int16))
    want = np.concatenate(want)
    nblk = want.size // (2 * bs)
    assert L.ka9q_rx_blocks_ready(rx) == nblk >= 5
    got = _ring_blocks(L, rx, nblk, bs)
    assert np.array_equal(got, want[:got.size])
    st = _stats(L, rx)
    assert st.datagrams_queued == len(dg) and st.inserted_out_of_order > 0
    assert st.samples == g.samples.value and st.rtp_drops == g.state.drops and st.rtp_dupes == g.state.dupes
    L.ka9q_rx_destroy(rx)


def test_receive_and_ingest_threads_over_udp():
    L = _lib.lib()
    rng = np.random.default_rng(5)
    n, bs = 480, 3840
    rsock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rsock.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 22)
    rsock.bind(("127.0.0.1", 0))
    ssock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    ssock.connect(rsock.getsockname())
    rx = L.ka9q_rx_create(rtp.IQ_S16, bs, 16)
    assert L.ka9q_rx_start(rx, rsock.fileno()) == 0
    npk = 8 * 10
    payloads = []
    for i in range(npk):
        d = iq_datagram(rng, rtp.IQ_PT, 100 + i, 5000 + i * n, 9, n)
        payloads.append(np.frombuffer(d[12 + 24:], dtype=np.int16))
        ssock.send(d)
        if i % 8 == 7:
            time.sleep(0.002)
    want = np.concatenate(payloads)
    nblk = want.size // (2 * bs)
    got = []
    for _ in range(nblk):
        p = L.ka9q_rx_peek_blocks(rx, 1, 2000)
        assert p, "block did not arrive"
        got.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int16)), (2 * bs,)).copy())
        assert L.ka9q_rx_consume(rx, 1) == 0
    assert np.array_equal(np.concatenate(got), want[:nblk * 2 * bs])
    assert L.ka9q_rx_stop(rx) == 0
    st = _stats(L, rx)
    assert st.datagrams_queued == npk and st.zero_filled == 0 and st.rtp_drops == 0
    L.ka9q_rx_destroy(rx)
    rsock.close()
    ssock.close()


def test_pcm_block_goes_out_with_sendmmsg():
    """Three channels (mono, stereo, one silent) of one 20 ms block: the packets that arrive on the socket are exactly the
    ones ka9q_pcm_packetise emits one by one (audio.c:32-132 semantics incl. silence suppression and the marker bit)."""
    L = _lib.lib()
    rng = np.random.default_rng(8)
    frames = 960
    chans = [1, 2, 1]
    offs = [0, 960, 960 + 1920]
    row = np.zeros(960 + 1920 + 960, dtype=np.int16)
    row[0:960] = rng.integers(-3000, 3000, 960)
    row[960:960 + 1920] = rng.integers(-3000, 3000, 1920)
    row[960 + 480:960 + 960] = 0            # one silent 240-frame stereo chunk in the middle: suppressed, marker after it
    rsock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rsock.bind(("127.0.0.1", 0))
    rsock.settimeout(2.0)
    ssock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    ssock.connect(rsock.getsockname())
    outs = (rtp._PcmOut * 3)()
    ref_outs = [rtp.PcmOut(0x100 + c, 7000, 40 + c) for c in range(3)]
    for c in range(3):
        outs[c].rtp.ssrc, outs[c].rtp.timestamp, outs[c].rtp.seq = 0x100 + c, 7000, 40 + c
    want = []
    for c in range(3):
        want += ref_outs[c].packetise(row[offs[c]:offs[c] + frames * chans[c]], chans[c])
    r = L.ka9q_pcm_send_block(ssock.fileno(), outs, row.ctypes.data_as(C.c_void_p), (C.c_int * 3)(*offs),
                              (C.c_int * 3)(*chans), 3, frames, 4)
    assert r == len(want) == 2 + 3 + 0
    got = [rsock.recv(4096) for _ in range(r)]
    assert got == want
    assert any(p[1] & 0x80 for p in got)     # the marker bit after the suppressed chunk
    for c in range(3):
        assert outs[c].rtp.timestamp == ref_outs[c].st.rtp.timestamp and outs[c].rtp.seq == ref_outs[c].st.rtp.seq
        assert outs[c].silent == ref_outs[c].st.silent
    rsock.close()
    ssock.close()
