"""ctypes bindings for oracle/_ref/libka9q_ref.so — the VERBATIM reference sources + harness.

TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; the product package (ka9q_sdr_b200) never does.

The .so is built by `make -C oracle ref` in the build container (where /root/reference exists) and
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libka9q_ref.so")

# enum filtertype (reference filter.h:17-22)
NONE, COMPLEX, CROSS_CONJ, REAL = 0, 1, 2, 3
# RTP payload types (reference multicast.h:19-24)
IQ_PT, IQ_PT8 = 97, 98
# enum demod_type (reference radio.h:20-24)
LINEAR_DEMOD, AM_DEMOD, FM_DEMOD = 0, 1, 2


class RefBlockStatus(C.Structure):
    _fields_ = [
        ("bb_power", C.c_float), ("snr", C.c_float), ("foffset", C.c_float), ("pdeviation", C.c_float),
        ("if_power", C.c_float), ("n0", C.c_float), ("agc_gain", C.c_float), ("cphase", C.c_float),
        ("pll_lock", C.c_int), ("channels", C.c_int), ("plfreq", C.c_float), ("pad", C.c_int),
    ]


class RefChainArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_char_p), ("samprate", C.c_int), ("L", C.c_int), ("M", C.c_int), ("decimate", C.c_int),
        ("carrier_hz", C.c_double), ("lo_cycles", C.c_double), ("low", C.c_float), ("high", C.c_float),
        ("shift", C.c_double), ("kaiser_beta", C.c_float), ("gain_factor", C.c_float), ("headroom", C.c_float),
        ("pkt_samples", C.c_int), ("pkt_type", C.c_int), ("channels", C.c_int),
    ]


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def mkl_lib_path() -> str | None:
    try:
        import torch  # noqa: F401  (only to locate libtorch_cpu.so, which exports MKL DFTI)
        p = os.path.join(os.path.dirname(torch.__file__), "lib", "libtorch_cpu.so")
        return p if os.path.exists(p) else None
    except Exception:
        return None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
    if "KA9Q_ORACLE_MKL_LIB" not in os.environ:
        p = mkl_lib_path()
        if p:
            os.environ["KA9Q_ORACLE_MKL_LIB"] = p
    L = C.CDLL(LIB_PATH)
    L.ka9q_oracle_set_fft_backend.argtypes = [C.c_char_p]
    L.ka9q_oracle_set_fft_backend.restype = C.c_int
    L.ka9q_oracle_fft_backend.restype = C.c_char_p
    L.ref_build_info.restype = C.c_char_p
    L.ref_modes_clear.restype = C.c_int
    L.ref_modes_load.argtypes = [C.c_char_p]
    L.ref_modes_count.restype = C.c_int
    L.ref_modes_add.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_int, C.c_int]
    L.ref_modes_get.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)]
    L.ref_chain_run.argtypes = [C.POINTER(RefChainArgs), C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long,
                                C.POINTER(C.c_long), C.c_void_p, C.c_long, C.c_void_p, C.c_int]
    L.ref_chain_run.restype = C.c_int
    L.ref_filter_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float),
                                 C.c_void_p]
    L.ref_filter_run.restype = C.c_int
    L.ref_osc_run.argtypes = [C.c_double, C.c_double, C.c_long, C.c_void_p]
    L.ref_glue_ingest.restype = C.c_longlong
    L.ref_glue_ingest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ref_glue_send.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.ref_glue_status.argtypes = [C.c_int, C.c_int] + [C.c_float] * 7 + [C.c_int, C.c_void_p]
    L.ref_hb15.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.ref_hb3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.ref_make_kaiser.argtypes = [C.c_void_p, C.c_uint, C.c_float]
    L.ref_window_filter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float]
    L.ref_window_rfilter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float]
    # raw FFT entry points of the shim (for validating the stand-in FFT itself)
    L.fftwf_plan_dft_1d.restype = C.c_void_p
    L.fftwf_plan_dft_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    L.fftwf_plan_dft_r2c_1d.restype = C.c_void_p
    L.fftwf_plan_dft_r2c_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
    L.fftwf_plan_dft_c2r_1d.restype = C.c_void_p
    L.fftwf_plan_dft_c2r_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
    L.fftwf_execute.argtypes = [C.c_void_p]
    L.fftwf_destroy_plan.argtypes = [C.c_void_p]
    L.ref_frontend_create.restype = C.c_void_p
    L.ref_frontend_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
    L.ref_frontend_destroy.argtypes = [C.c_void_p]
    L.ref_frontend_set_estimates.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
    L.ref_frontend_status.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_frontend_process.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    _lib = L
    return L


def set_fft_backend(name: str) -> bool:
    return lib().ka9q_oracle_set_fft_backend(name.encode()) == 0


def fft_backend() -> str:
    return lib().ka9q_oracle_fft_backend().decode()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def load_modes(table) -> None:
    """table: iterable of ka9q_sdr_b200.modes.Mode-like objects (name, demod_type, low, high, shift, attack,
    recovery, hang, channels, isb, flat, pll, square)."""
    L = lib()
    L.ref_modes_clear()
    for m in table:
        flags = (1 if m.isb else 0) | (2 if m.flat else 0) | (4 if m.pll else 0) | (8 if m.square else 0)
        L.ref_modes_add(m.name.encode(), int(m.demod_type), m.low, m.high, m.shift, m.attack, m.recovery, m.hang,
                        m.channels, flags)


def load_modes_file(directory: str) -> int:
    return lib().ref_modes_load(directory.encode())


def get_modes():
    L = lib()
    out = []
    for i in range(L.ref_modes_count()):
        name = C.create_string_buffer(32)
        dt = C.c_int()
        vals = (C.c_float * 6)()
        ch = C.c_int()
        fl = C.c_int()
        L.ref_modes_get(i, name, C.byref(dt), vals, C.byref(ch), C.byref(fl))
        out.append((name.value.decode(), dt.value, tuple(float(v) for v in vals), ch.value, fl.value))
    return out


@dataclass
class ChainResult:
    pcm: np.ndarray          # int16, interleaved when stereo
    filt: np.ndarray | None  # complex64 [nblocks, olen] (FM/AM only: linear modifies output in place)
    status: np.ndarray       # structured array of RefBlockStatus
    nblocks: int
    channels: int


def chain_run(mode: str, samprate: int, L: int, M: int, decimate: int, iq: np.ndarray, *, carrier_hz: float = 0.0,
              lo_cycles: float = math.nan, low: float = math.nan, high: float = math.nan, shift: float = math.nan,
              kaiser_beta: float = 3.0, gain_factor: float = 1.0, headroom: float = math.nan, pkt_samples: int = 1024,
              pkt_type: int = IQ_PT, drop: np.ndarray | None = None, want_filt: bool = False,
              channels: int = 0) -> ChainResult:
    """Run the reference receive chain (proc_samples -> filter -> demod) on interleaved int16/int8 I/Q."""
    lb = lib()
    iq = np.ascontiguousarray(iq)
    assert iq.dtype in (np.int16, np.int8)
    nsamples = iq.size // 2
    nblocks_max = nsamples // L + 2
    olen = L // decimate
    pcm = np.zeros(nblocks_max * olen * 2, dtype=np.int16)
    pcm_len = C.c_long(0)
    filt = np.zeros((nblocks_max, olen), dtype=np.complex64) if want_filt else None
    st = (RefBlockStatus * nblocks_max)()
    args = RefChainArgs(mode.encode(), samprate, L, M, decimate, carrier_hz, lo_cycles, low, high, shift, kaiser_beta,
                        gain_factor, headroom, pkt_samples, pkt_type, channels)
    dropa = np.ascontiguousarray(drop, dtype=np.uint8) if drop is not None else None
    nb = lb.ref_chain_run(C.byref(args), _ptr(iq), nsamples, _ptr(dropa), _ptr(pcm), pcm.size, C.byref(pcm_len),
                          _ptr(filt), nblocks_max if want_filt else 0, C.cast(st, C.c_void_p), nblocks_max)
    if nb < 0:
        raise RuntimeError(f"ref_chain_run failed: {nb}")
    status = np.frombuffer(st, dtype=np.dtype(RefBlockStatus), count=nblocks_max)[:nb].copy()
    ch = int(status["channels"][0]) if nb else 1
    return ChainResult(pcm[:pcm_len.value].copy(), filt[:nb].copy() if want_filt else None, status, nb, ch)


def filter_run(L: int, M: int, decimate: int, in_type: int, out_type: int, x: np.ndarray, *, low: float = 0.0,
               high: float = 0.0, beta: float = 3.0, response: np.ndarray | None = None, want_fdomain: bool = False):
    """Run create_filter_input/output + execute_* over nblocks of L samples. Returns dict."""
    lb = lib()
    N = L + M - 1
    N_dec = N // decimate
    olen = L // decimate
    if in_type == REAL:
        x = np.ascontiguousarray(x, dtype=np.float32)
        nblocks = x.size // L
    else:
        x = np.ascontiguousarray(x, dtype=np.complex64)
        nblocks = x.size // L
    rbins = N_dec // 2 + 1 if out_type == REAL else N_dec
    out = np.zeros(nblocks * olen, dtype=np.float32 if out_type == REAL else np.complex64)
    resp_out = np.zeros(rbins if response is not None else N_dec, dtype=np.complex64)
    ng = C.c_float(0)
    fd = np.zeros(N if in_type != REAL else N // 2 + 1, dtype=np.complex64) if want_fdomain else None
    resp_in = np.ascontiguousarray(response, dtype=np.complex64) if response is not None else None
    if resp_in is not None:
        assert resp_in.size >= rbins
    r = lb.ref_filter_run(L, M, decimate, in_type, out_type, low, high, beta, _ptr(resp_in), _ptr(x), nblocks, _ptr(out),
                          _ptr(resp_out), C.byref(ng), _ptr(fd))
    if r < 0:
        raise RuntimeError(f"ref_filter_run failed: {r}")
    return {"out": out.reshape(nblocks, olen), "response": resp_out, "noise_gain": ng.value, "fdomain": fd}


def osc_run(freq: float, rate: float, nsteps: int) -> np.ndarray:
    out = np.zeros(2 * nsteps, dtype=np.float64)
    lib().ref_osc_run(freq, rate, nsteps, _ptr(out))
    return out[0::2] + 1j * out[1::2]


def make_kaiser(M: int, beta: float) -> np.ndarray:
    w = np.zeros(M, dtype=np.float32)
    lib().ref_make_kaiser(_ptr(w), M, beta)
    return w


def window_filter(L: int, M: int, response: np.ndarray, beta: float) -> np.ndarray:
    r = np.ascontiguousarray(response, dtype=np.complex64).copy()
    assert r.size == L + M - 1
    lib().ref_window_filter(L, M, _ptr(r), beta)
    return r


def window_rfilter(L: int, M: int, response: np.ndarray, beta: float) -> np.ndarray:
    r = np.ascontiguousarray(response, dtype=np.complex64).copy()
    assert r.size == (L + M - 1) // 2 + 1
    lib().ref_window_rfilter(L, M, _ptr(r), beta)
    return r


def hb15(state16: np.ndarray, x: np.ndarray) -> np.ndarray:
    """state16: float32[16] = coeffs[4], even[4], odd[4], old_odd[4]; updated in place. x: 2*cnt floats."""
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    cnt = x.size // 2
    out = np.zeros(cnt, dtype=np.float32)
    lib().ref_hb15(_ptr(state16), _ptr(out), _ptr(x), cnt)
    return out


def hb3(state1: np.ndarray, x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    cnt = x.size // 2
    out = np.zeros(cnt, dtype=np.float32)
    lib().ref_hb3(_ptr(state1), _ptr(out), _ptr(x), cnt)
    return out


def raw_fft(x: np.ndarray, sign: int) -> np.ndarray:
    """c2c through the shim's fftwf_* entry points (validates the FFT backend itself)."""
    lb = lib()
    a = np.ascontiguousarray(x, dtype=np.complex64).copy()
    b = np.zeros_like(a)
    p = lb.fftwf_plan_dft_1d(a.size, _ptr(a), _ptr(b), sign, 1 << 6)
    lb.fftwf_execute(p)
    lb.fftwf_destroy_plan(p)
    return b


def raw_rfft(x: np.ndarray) -> np.ndarray:
    lb = lib()
    a = np.ascontiguousarray(x, dtype=np.float32).copy()
    b = np.zeros(a.size // 2 + 1, dtype=np.complex64)
    p = lb.fftwf_plan_dft_r2c_1d(a.size, _ptr(a), _ptr(b), 1 << 6)
    lb.fftwf_execute(p)
    lb.fftwf_destroy_plan(p)
    return b


def raw_irfft(X: np.ndarray, n: int) -> np.ndarray:
    lb = lib()
    a = np.ascontiguousarray(X, dtype=np.complex64).copy()
    b = np.zeros(n, dtype=np.float32)
    p = lb.fftwf_plan_dft_c2r_1d(n, _ptr(a), _ptr(b), 1 << 6)
    lb.fftwf_execute(p)
    lb.fftwf_destroy_plan(p)
    return b


# ---- wire-format glue ----
class RtpState(C.Structure):
    """struct rtp_state (multicast.h:41-50)"""
    _fields_ = [("ssrc", C.c_uint32), ("init", C.c_int), ("seq", C.c_uint16), ("timestamp", C.c_uint32),
                ("packets", C.c_longlong), ("bytes", C.c_longlong), ("drops", C.c_longlong), ("dupes", C.c_longlong)]


class GlueIngest:
    """rtp_recv parsing + proc_samples packet head, on the reference's own ntoh_rtp / rtp_process."""

    def __init__(self):
        self.state = RtpState()
        self.samples = C.c_longlong(0)

    def datagram(self, data: bytes):
        """-> (ret, raw bytes appended): ret = complex samples appended, or -1 if ignored / dropped"""
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        out = np.zeros((192000 + 4096) * 4, dtype=np.uint8)
        r = lib().ref_glue_ingest(C.byref(self.state), C.byref(self.samples), buf, len(data), _ptr(out))
        if r < 0:
            return r, b""
        ptype = data[1] & 0x7F
        return r, out[: r * (4 if ptype == 97 else 2)].tobytes()


def glue_send(channels: int, state: dict, samples: np.ndarray):
    """The real send_mono_output / send_stereo_output (audio.c:32-132) on float samples; returns the list of packets and
    updates state {ssrc, timestamp, seq, silent, packets}. Call with a few packets' worth at a time."""
    x = np.ascontiguousarray(samples, dtype=np.float32)
    frames = x.size // channels
    st = np.array([state["ssrc"], state["timestamp"], state["seq"], state["silent"], 0], dtype=np.int64)
    out = np.zeros(64 * 2048, dtype=np.uint8)
    lens = np.zeros(64, dtype=np.int32)
    n = lib().ref_glue_send(channels, _ptr(st), _ptr(x), frames, _ptr(out), out.size, _ptr(lens), 64)
    assert n >= 0
    state.update(timestamp=int(st[1]) & 0xFFFFFFFF, seq=int(st[2]) & 0xFFFF, silent=int(st[3]),
                 packets=state.get("packets", 0) + int(st[4]))
    pk, off = [], 0
    for i in range(n):
        pk.append(out[off:off + lens[i]].tobytes())
        off += int(lens[i])
    return pk


def glue_status(demod_type: int, isb: int, noise_bw: float, if_power: float, bb_power: float, gain: float, pdev: float,
                foffset: float, snr: float, channels: int) -> bytes:
    """The reference's status.c encoders applied in radio_status.c's order to the fields the product computes."""
    out = np.zeros(256, dtype=np.uint8)
    n = lib().ref_glue_status(demod_type, isb, noise_bw, if_power, bb_power, gain, pdev, foffset, snr, channels, _ptr(out))
    return out[:n].tobytes()


# ---- front-end decimator service: oracle/frontend_ref.c, a restatement of hackrf.c's sample path around the
# verbatim decimate.c ----
class Frontend:
    def __init__(self, out_samprate: int, decimate: int, offset: int, callback_samples: int, dc_alpha: float = 1e-7,
                 power_alpha: float = 1.0):
        self.h = lib().ref_frontend_create(out_samprate, decimate, offset, callback_samples, dc_alpha, power_alpha)
        self.decimate = decimate

    def set_estimates(self, dc_i, dc_q, imbalance, sinphi):
        lib().ref_frontend_set_estimates(self.h, dc_i, dc_q, imbalance, sinphi)

    def process(self, iq8: np.ndarray) -> np.ndarray:
        iq8 = np.ascontiguousarray(iq8, dtype=np.int8)
        n = iq8.size // 2
        out = np.zeros(2 * (n // self.decimate), dtype=np.int16)
        r = lib().ref_frontend_process(self.h, _ptr(iq8), n, _ptr(out))
        if r < 0:
            raise RuntimeError("ref_frontend_process failed")
        return out

    def status(self):
        o = np.zeros(7, dtype=np.float32)
        lib().ref_frontend_status(self.h, _ptr(o))
        return dict(dc_i=float(o[0]), dc_q=float(o[1]), imbalance=float(o[2]), sinphi=float(o[3]), in_power=float(o[4]),
                    clips=int(o[5]))

    def close(self):
        if self.h:
            lib().ref_frontend_destroy(self.h)
            self.h = None
