"""Python driver over the C ABI batch layer (include/ka9q_b200.h section B) — used by tests and bench.py.

It mirrors how the reference is operated: a mode name from the mode table (modes.txt / set_mode, radio.c:322-374)
plus a carrier frequency per channel; the heavy lifting is entirely inside libka9q_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .modes import FM_DEMOD, LINEAR_DEMOD, Mode, get_mode

IQ_S16, IQ_S8 = 1, 2
MGPU_NCCL, MGPU_P2P = 1, 2
FLAG_ISB, FLAG_FLAT, FLAG_PLL, FLAG_SQUARE = 1, 2, 4, 8


def mode_flags(m: Mode) -> int:
    return (FLAG_ISB if m.isb else 0) | (FLAG_FLAT if m.flat else 0) | (FLAG_PLL if m.pll else 0) | \
        (FLAG_SQUARE if m.square else 0)


class PinnedBuffer:
    """Page-locked host buffer (cudaHostAlloc) exposed as a numpy array."""

    def __init__(self, nbytes: int, dtype=np.uint8):
        self._lib = _lib.lib()
        self.ptr = self._lib.ka9q_host_alloc(nbytes)
        if not self.ptr:
            raise RuntimeError("ka9q_host_alloc failed: " + _lib.last_error())
        self.nbytes = nbytes
        buf = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype)

    def free(self):
        if self.ptr:
            self._lib.ka9q_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Channelizer:
    """One I/Q stream on one GPU with K receive channels."""

    def __init__(self, samprate: int, L: int, M: int, decimate: int, *, device: int = 0, iq_format: int = IQ_S16,
                 gain_factor: float = 1.0, max_blocks: int = 1, capture_filter_output: bool = False):
        self.lib = _lib.lib()
        self.cfg = _lib.StreamConfig(device, samprate, L, M, decimate, iq_format, gain_factor, max_blocks,
                                     1 if capture_filter_output else 0)
        h = C.c_void_p()
        _lib.check(self.lib.ka9q_stream_create(C.byref(h), C.byref(self.cfg)), "ka9q_stream_create")
        self.h = h
        self.samprate, self.L, self.M, self.decimate = samprate, L, M, decimate
        self.N = L + M - 1
        self.olen = L // decimate
        self.max_blocks = max_blocks
        self.iq_dtype = np.int16 if iq_format == IQ_S16 else np.int8
        self.channels: list[int] = []     # PCM channel count per channel
        self.committed = False

    # -- channel set-up ------------------------------------------------------------------------------------
    def add_channel(self, mode: str | Mode, bin: int, *, low: float | None = None, high: float | None = None,
                    kaiser_beta: float = 3.0, shift: float | None = None, headroom: float = math.nan,
                    channels: int | None = None) -> int:
        m = get_mode(mode) if isinstance(mode, str) else mode
        p = _lib.ChanParams()
        p.demod_type = m.demod_type
        p.flags = mode_flags(m)
        p.channels = channels if channels is not None else (m.channels if m.demod_type == LINEAR_DEMOD else 1)
        p.bin = int(bin)
        p.low = m.low if low is None else low
        p.high = m.high if high is None else high
        p.kaiser_beta = kaiser_beta
        p.shift = m.shift if shift is None else shift
        p.attack_rate = m.attack
        p.recovery_rate = m.recovery
        p.hangtime = m.hang
        p.headroom = headroom
        idx = _lib.check(self.lib.ka9q_stream_add_channel(self.h, C.byref(p)), "ka9q_stream_add_channel")
        self.channels.append(int(p.channels) if m.demod_type == LINEAR_DEMOD else 1)
        return idx

    def split_carrier(self, carrier_hz: float) -> tuple[int, float]:
        """carrier (Hz from the first LO) -> (nearest grid bin, fraction of a bin left over)."""
        b, f = C.c_longlong(0), C.c_double(0.0)
        _lib.check(self.lib.ka9q_stream_split_carrier(self.h, float(carrier_hz), C.byref(b), C.byref(f)), "split_carrier")
        return int(b.value), float(f.value)

    def add_channel_hz(self, mode: str | Mode, carrier_hz: float, **kw) -> int:
        """add_channel for a carrier anywhere in the band: grid bin + fine LO (ka9q_stream_set_fine_lo)."""
        b, f = self.split_carrier(carrier_hz)
        idx = self.add_channel(mode, b, **kw)
        if f != 0.0:
            _lib.check(self.lib.ka9q_stream_set_fine_lo(self.h, idx, f), "set_fine_lo")
        return idx

    def enable_pl(self, enable: bool = True):
        """PL-tone analyser (fm.c:189-285) for the FM channels; call before commit. plfreq = status['reserved'][..., 1]."""
        _lib.check(self.lib.ka9q_stream_enable_pl(self.h, 1 if enable else 0), "enable_pl")

    def enable_n0(self, enable: bool = True):
        """compute_n0 (radio.c:383-425) for every channel and block; call before commit."""
        _lib.check(self.lib.ka9q_stream_enable_n0(self.h, 1 if enable else 0), "enable_n0")
        self.n0_enabled = enable

    def fetch_n0(self, nblocks: int):
        """(raw, smoothed) noise density of the last computed batch, [nblocks, nchan] float32 each; synchronises."""
        raw = np.empty((nblocks, self.nchan), dtype=np.float32)
        sm = np.empty((nblocks, self.nchan), dtype=np.float32)
        _lib.check(self.lib.ka9q_stream_fetch_n0(self.h, nblocks, raw.ctypes.data_as(C.c_void_p),
                                                 sm.ctypes.data_as(C.c_void_p)), "fetch_n0")
        self.sync()
        return raw, sm

    def commit(self):
        _lib.check(self.lib.ka9q_stream_commit(self.h), "ka9q_stream_commit")
        self.committed = True
        self.nchan = self.lib.ka9q_stream_num_channels(self.h)
        self.pcm_stride = int(self.lib.ka9q_stream_pcm_stride(self.h))
        self.pcm_off = [self.lib.ka9q_stream_pcm_offset(self.h, c) for c in range(self.nchan)]
        self.launches_per_call = self.lib.ka9q_stream_launches_per_call(self.h)

    def set_filter(self, chan: int, low: float, high: float, kaiser_beta: float = 3.0):
        _lib.check(self.lib.ka9q_stream_set_filter(self.h, chan, low, high, kaiser_beta), "ka9q_stream_set_filter")

    # -- processing ----------------------------------------------------------------------------------------
    def process(self, iq: np.ndarray, pcm: np.ndarray | None = None, want_status: bool = True):
        """iq: interleaved I/Q for nblocks*L samples. Returns (pcm[nblocks, pcm_stride], status or None)."""
        iq = np.ascontiguousarray(iq, dtype=self.iq_dtype)
        nblocks = iq.size // (2 * self.L)
        assert nblocks * 2 * self.L == iq.size and 1 <= nblocks <= self.max_blocks
        if pcm is None:
            pcm = np.empty((nblocks, self.pcm_stride), dtype=np.int16)
        status = (_lib.ChanStatus * (nblocks * self.nchan))() if want_status else None
        _lib.check(self.lib.ka9q_stream_process(self.h, iq.ctypes.data_as(C.c_void_p), nblocks,
                                                pcm.ctypes.data_as(C.c_void_p),
                                                C.cast(status, C.c_void_p) if want_status else None),
                   "ka9q_stream_process")
        st = None
        if want_status:
            st = np.frombuffer(status, dtype=np.dtype(_lib.ChanStatus)).reshape(nblocks, self.nchan).copy()
        return pcm, st

    def run(self, iq: np.ndarray, want_status: bool = True):
        """Process an arbitrary number of whole blocks, max_blocks at a time. Returns (pcm[nblocks_total, stride], status)."""
        iq = np.ascontiguousarray(iq, dtype=self.iq_dtype)
        total = iq.size // (2 * self.L)
        pcm = np.empty((total, self.pcm_stride), dtype=np.int16)
        sts = []
        b = 0
        while b < total:
            nb = min(self.max_blocks, total - b)
            _, st = self.process(iq[2 * b * self.L:2 * (b + nb) * self.L], pcm[b:b + nb], want_status)
            if want_status:
                sts.append(st)
            b += nb
        return pcm, (np.concatenate(sts, axis=0) if want_status and sts else None)

    def channel_pcm(self, pcm: np.ndarray, chan: int) -> np.ndarray:
        """Extract one channel's int16 stream (interleaved L/R when stereo) from PCM rows."""
        off = self.pcm_off[chan]
        n = self.olen * self.channels[chan]
        return pcm[:, off:off + n].reshape(-1)

    def push(self, iq_ptr, nblocks):
        _lib.check(self.lib.ka9q_stream_push(self.h, iq_ptr, nblocks), "ka9q_stream_push")

    def compute(self, nblocks):
        _lib.check(self.lib.ka9q_stream_compute(self.h, nblocks), "ka9q_stream_compute")

    def compute_resident(self, nblocks):
        _lib.check(self.lib.ka9q_stream_compute_resident(self.h, nblocks), "ka9q_stream_compute_resident")

    def compute_fft_only(self, nblocks):
        _lib.check(self.lib.ka9q_stream_compute_fft_only(self.h, nblocks), "ka9q_stream_compute_fft_only")

    def compute_channels_only(self, nblocks):
        _lib.check(self.lib.ka9q_stream_compute_channels_only(self.h, nblocks), "ka9q_stream_compute_channels_only")

    def fetch(self, nblocks, pcm_ptr, status_ptr=None):
        _lib.check(self.lib.ka9q_stream_fetch(self.h, nblocks, pcm_ptr, status_ptr), "ka9q_stream_fetch")

    def wait_fetch(self):
        _lib.check(self.lib.ka9q_stream_wait_fetch(self.h), "ka9q_stream_wait_fetch")

    def wait_fetched(self, batches_ago: int = 0):
        """Wait for the fetch issued `batches_ago` fetches ago (0 or 1)."""
        _lib.check(self.lib.ka9q_stream_wait_fetched(self.h, batches_ago), "ka9q_stream_wait_fetched")

    def sync(self):
        _lib.check(self.lib.ka9q_stream_sync(self.h), "ka9q_stream_sync")

    def last_timing(self):
        t, f, c = C.c_float(), C.c_float(), C.c_float()
        _lib.check(self.lib.ka9q_stream_last_timing(self.h, C.byref(t), C.byref(f), C.byref(c)), "last_timing")
        return t.value, f.value, c.value

    def set_overlap(self, enable: bool):
        _lib.check(self.lib.ka9q_stream_set_overlap(self.h, 1 if enable else 0), "set_overlap")

    def timer_start(self, regions: bool = True):
        """regions=False: only the outer event pair (no per-kernel event records inside the timed steps)."""
        if regions:
            _lib.check(self.lib.ka9q_stream_timer_start(self.h), "timer_start")
        else:
            _lib.check(self.lib.ka9q_stream_timer_start_plain(self.h), "timer_start_plain")

    def timer_stop(self):
        """Returns (region_ms, {class: (ms, launches)}) for classes fft, fm, am, linear, bcast."""
        ms = C.c_float()
        cms = (C.c_float * 5)()
        cl = (C.c_int * 5)()
        _lib.check(self.lib.ka9q_stream_timer_stop(self.h, C.byref(ms), cms, cl), "timer_stop")
        names = ("fft", "fm", "am", "linear", "bcast")
        return ms.value, {n: (cms[i], cl[i]) for i, n in enumerate(names)}

    def timer_timeline(self, max_regions: int = 4096):
        """After timer_stop: [(class name, start ms, end ms)] of the bracketed regions, in issue order."""
        cls = (C.c_int * max_regions)()
        t0 = (C.c_float * max_regions)()
        t1 = (C.c_float * max_regions)()
        n = _lib.check(self.lib.ka9q_stream_timer_timeline(self.h, max_regions, cls, t0, t1), "timer_timeline")
        names = ("fft", "fm", "am", "linear", "bcast", "wait")
        return [(names[cls[i]], t0[i], t1[i]) for i in range(n)]

    def nccl_init(self, id128: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(id128, 128)
        _lib.check(self.lib.ka9q_stream_nccl_init(self.h, buf, rank, nranks), "nccl_init")

    def compute_fft_blocks(self, nblocks: int, first: int, count: int):
        _lib.check(self.lib.ka9q_stream_compute_fft_blocks(self.h, nblocks, first, count), "compute_fft_blocks")

    def nccl_allgather_spectrum(self, nblocks: int):
        _lib.check(self.lib.ka9q_stream_nccl_allgather_spectrum(self.h, nblocks), "nccl_allgather_spectrum")

    def nccl_broadcast_spectrum(self, nblocks: int, root: int = 0):
        _lib.check(self.lib.ka9q_stream_nccl_broadcast_spectrum(self.h, nblocks, root), "nccl_broadcast_spectrum")

    # -- channel-sharded multi-GPU (include/ka9q_b200.h: ka9q_stream_mgpu_*) ---------------------------------------
    def needed_bins(self):
        lo, ln = C.c_longlong(), C.c_longlong()
        _lib.check(self.lib.ka9q_stream_needed_bins(self.h, C.byref(lo), C.byref(ln)), "needed_bins")
        return lo.value, ln.value

    def mgpu_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(self.lib.ka9q_stream_mgpu_export(self.h, buf), "mgpu_export")
        return buf.raw

    def mgpu_setup(self, transport: int, rank: int, nranks: int, arcs, blobs: bytes | None = None):
        """arcs: [(lo, len)] of every rank (needed_bins all-gathered); blobs: nranks x 128 bytes (P2P transport)."""
        lo = (C.c_longlong * nranks)(*[a[0] for a in arcs])
        ln = (C.c_longlong * nranks)(*[a[1] for a in arcs])
        b = C.create_string_buffer(blobs, len(blobs)) if blobs else None
        _lib.check(self.lib.ka9q_stream_mgpu_setup(self.h, transport, rank, nranks, lo, ln, b), "mgpu_setup")
        self.mg_rank, self.mg_nranks = rank, nranks

    def mgpu_input_range(self, first_block: int, nblocks: int):
        a, n = C.c_longlong(), C.c_longlong()
        _lib.check(self.lib.ka9q_stream_mgpu_input_range(self.h, first_block, nblocks, C.byref(a), C.byref(n)),
                   "mgpu_input_range")
        return a.value, n.value

    def push_at(self, iq_ptr, first_sample: int, nsamples: int):
        _lib.check(self.lib.ka9q_stream_push_at(self.h, iq_ptr, first_sample, nsamples), "push_at")

    def mgpu_compute(self, nblocks: int, resident: bool = False):
        _lib.check(self.lib.ka9q_stream_mgpu_compute(self.h, nblocks, 1 if resident else 0), "mgpu_compute")

    def blocks_done(self) -> int:
        return int(self.lib.ka9q_stream_blocks_done(self.h))

    def mgpu_error(self) -> int:
        return self.lib.ka9q_stream_mgpu_error(self.h)

    # -- introspection (parity tests) ----------------------------------------------------------------------
    def response(self, chan: int):
        out = np.empty(2048, dtype=np.complex64)
        ng = C.c_float()
        _lib.check(self.lib.ka9q_stream_get_response(self.h, chan, out.ctypes.data_as(C.c_void_p), C.byref(ng)),
                   "get_response")
        return out, ng.value

    def filter_output(self, chan: int, nblocks: int) -> np.ndarray:
        out = np.empty((nblocks, self.olen), dtype=np.complex64)
        _lib.check(self.lib.ka9q_stream_get_filter_output(self.h, chan, nblocks, out.ctypes.data_as(C.c_void_p)),
                   "get_filter_output")
        return out

    def spectrum(self, block: int = 0) -> np.ndarray:
        out = np.empty(self.N, dtype=np.complex64)
        _lib.check(self.lib.ka9q_stream_get_spectrum(self.h, block, out.ctypes.data_as(C.c_void_p)), "get_spectrum")
        return out

    def if_energy(self, nblocks: int) -> np.ndarray:
        out = np.empty(nblocks, dtype=np.float32)
        _lib.check(self.lib.ka9q_stream_get_if_energy(self.h, nblocks, out.ctypes.data_as(C.c_void_p)), "if_energy")
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.ka9q_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().ka9q_nccl_unique_id(buf), "ka9q_nccl_unique_id")
    return buf.raw


def fft_c2c(x: np.ndarray, sign: int = -1, device: int = 0) -> np.ndarray:
    """Generic batched complex FFT through the library (x: [batch, n] or [n])."""
    a = np.ascontiguousarray(x, dtype=np.complex64)
    batch = 1 if a.ndim == 1 else a.shape[0]
    n = a.shape[-1]
    out = np.empty_like(a)
    _lib.check(_lib.lib().ka9q_fft_c2c(device, n, batch, sign, a.ctypes.data_as(C.c_void_p),
                                       out.ctypes.data_as(C.c_void_p)), "ka9q_fft_c2c")
    return out


def fft_plan(n: int):
    sizes = (C.c_int * 4)()
    np_ = _lib.lib().ka9q_fft_plan_describe(n, sizes)
    return None if np_ < 0 else [sizes[i] for i in range(np_)]
