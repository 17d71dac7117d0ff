"""Host-side mirror of the wire-format glue in libka9q_b200.so (include/ka9q_b200.h): I/Q datagram ->
sample stream with the reference's sequence / timestamp repair (main.c:313-344, radio.c:60-100, multicast.c:305-340),
and int16 PCM -> RTP packets (audio.c:32-132). Pure ctypes over the C ABI; no GPU is touched."""
import ctypes as C

import numpy as np

from . import _lib

IQ_PT, IQ_PT8, PCM_MONO_PT, PCM_STEREO_PT = 97, 98, 11, 10
IQ_S16, IQ_S8 = 1, 2


class RtpState(C.Structure):
    """ka9q_rtp_state == struct rtp_state (multicast.h:41-50)"""
    _fields_ = [("ssrc", C.c_uint32), ("init", C.c_int), ("seq", C.c_uint16), ("timestamp", C.c_uint32),
                ("packets", C.c_longlong), ("bytes", C.c_longlong), ("drops", C.c_longlong), ("dupes", C.c_longlong)]


class _Ingest(C.Structure):
    _fields_ = [("rtp", RtpState), ("iq_format", C.c_int), ("samples", C.c_longlong), ("zero_filled", C.c_longlong),
                ("ignored", C.c_longlong)]


class _PcmOut(C.Structure):
    _fields_ = [("rtp", RtpState), ("silent", C.c_int)]


_EMIT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)
_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        L.ka9q_ingest_init.argtypes = [C.c_void_p, C.c_int]
        L.ka9q_ingest_init.restype = None
        L.ka9q_ingest_datagram.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
        L.ka9q_ingest_datagram.restype = C.c_longlong
        L.ka9q_rtp_process.argtypes = [C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint32, C.c_int]
        L.ka9q_pcm_packetise.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _EMIT, C.c_void_p]
        _bound = True
    return L


class Ingest:
    """One I/Q stream's receive state. `datagram()` returns (n, samples): n complex samples appended (lost samples
    zero-filled, then the payload) as an int16 / int8 array of 2n values, or (-1, None) if the datagram was ignored."""

    def __init__(self, iq_format: int = IQ_S16, room: int = 192000 + 65536):
        self.st = _Ingest()
        _L().ka9q_ingest_init(C.byref(self.st), iq_format)
        self.dtype = np.int16 if iq_format == IQ_S16 else np.int8
        self.buf = np.zeros(2 * room, dtype=self.dtype)
        self.room = room

    def datagram(self, data: bytes):
        raw = (C.c_ubyte * len(data)).from_buffer_copy(data)
        n = _L().ka9q_ingest_datagram(C.byref(self.st), raw, len(data), self.buf.ctypes.data_as(C.c_void_p), self.room)
        if n < 0:
            return int(n), None
        return int(n), self.buf[: 2 * n].copy()


class PcmOut:
    """One PCM output stream (demod->output.rtp + the silence flag)."""

    def __init__(self, ssrc: int, timestamp: int = 0, seq: int = 0):
        self.st = _PcmOut()
        self.st.rtp.ssrc = ssrc
        self.st.rtp.timestamp = timestamp
        self.st.rtp.seq = seq

    def packetise(self, pcm: np.ndarray, channels: int):
        """pcm: host-order int16, frames x channels interleaved -> list of RTP packets (bytes)"""
        x = np.ascontiguousarray(pcm, dtype=np.int16)
        out = []

        def emit(_user, pkt, n):
            out.append(C.string_at(pkt, n))
            return 0

        cb = _EMIT(emit)
        r = _L().ka9q_pcm_packetise(C.byref(self.st), x.ctypes.data_as(C.c_void_p), x.size // channels, channels, cb, None)
        if r < 0:
            raise ValueError("ka9q_pcm_packetise: bad arguments")
        assert r == len(out)
        return out
