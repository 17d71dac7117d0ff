"""Host-side throughput of the wire-format glue (no GPU): datagrams/s through ka9q_ingest_datagram and PCM packets/s
out of ka9q_pcm_packetise, called from C-sized batches through ctypes (the Python loop overhead is included, so these
are lower bounds)."""
import os, sys, time, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from ka9q_sdr_b200 import rtp

rng = np.random.default_rng(0)
n = 240                                   # samples per I/Q datagram (hackrf.c default blocksize 350; funcube 240)
pk = []
for i in range(20000):
    h = struct.pack(">BBHII", 0x80, rtp.IQ_PT, i & 0xFFFF, (i * n) & 0xFFFFFFFF, 7)
    pk.append(h + bytes(24) + rng.integers(-3000, 3000, 2 * n, dtype=np.int16).tobytes())
L = rtp._L()
st = rtp._Ingest(); L.ka9q_ingest_init(C.byref(st), rtp.IQ_S16)
dst = np.zeros(2 * (192000 + 4096), dtype=np.int16); dp = dst.ctypes.data_as(C.c_void_p)
bufs = [(C.c_ubyte * len(p)).from_buffer_copy(p) for p in pk]
t0 = time.perf_counter()
tot = 0
for b in bufs:
    tot += L.ka9q_ingest_datagram(C.byref(st), b, len(b), dp, 192000 + 4096)
dt = time.perf_counter() - t0
print(f"ingest: {len(pk)/dt/1e3:.0f} k datagrams/s = {tot/dt/1e6:.1f} MS/s per host thread (ctypes loop included)")
o = rtp.PcmOut(1)
pcm = rng.integers(-20000, 20000, 960 * 8192, dtype=np.int16)       # one 20 ms block of 8192 mono channels
cnt = [0]
def emit(_u, _p, _n):
    cnt[0] += 1
    return 0
cb = rtp._EMIT(emit)
t0 = time.perf_counter()
for c in range(0, 8192, 8):                                         # every 8th channel: the callback is Python
    row = pcm[c * 960:(c + 1) * 960]
    L.ka9q_pcm_packetise(C.byref(o.st), row.ctypes.data_as(C.c_void_p), 960, 1, cb, None)
dt = time.perf_counter() - t0
print(f"packetise: {cnt[0]/dt/1e3:.0f} k packets/s, {1024/dt:.0f} channel-blocks/s per host thread (Python emit callback included)")
