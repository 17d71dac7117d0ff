"""Python driver of the front-end decimator service (include/ka9q_b200.h: ka9q_frontend_*): the sample path of the
reference's `hackrf` daemon (hackrf.c:129-345) on the device. Used by tests and bench.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class Frontend:
    def __init__(self, out_samprate: int, decimate: int = 64, offset: int = 1, callback_samples: int = 131072, *,
                 device: int = 0, dc_alpha: float = 1e-7, power_alpha: float = 1.0):
        self.lib = _lib.lib()
        self.cfg = _lib.FrontendConfig(device, out_samprate, decimate, offset, callback_samples, dc_alpha, power_alpha)
        h = C.c_void_p()
        _lib.check(self.lib.ka9q_frontend_create(C.byref(h), C.byref(self.cfg)), "ka9q_frontend_create")
        self.h = h
        self.decimate = decimate

    def set_estimates(self, dc_i: float, dc_q: float, imbalance: float, sinphi: float):
        _lib.check(self.lib.ka9q_frontend_set_estimates(self.h, dc_i, dc_q, imbalance, sinphi), "set_estimates")

    def process(self, iq8: np.ndarray) -> np.ndarray:
        iq8 = np.ascontiguousarray(iq8, dtype=np.int8)
        n = iq8.size // 2
        out = np.empty(2 * (n // self.decimate), dtype=np.int16)
        _lib.check(self.lib.ka9q_frontend_process(self.h, iq8.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)),
                   "ka9q_frontend_process")
        return out

    def process_to_stream(self, iq8: np.ndarray, channelizer) -> int:
        iq8 = np.ascontiguousarray(iq8, dtype=np.int8)
        n = iq8.size // 2
        _lib.check(self.lib.ka9q_frontend_process_to_stream(self.h, iq8.ctypes.data_as(C.c_void_p), n, channelizer.h),
                   "ka9q_frontend_process_to_stream")
        return n // self.decimate

    def rerun_resident(self, nsamples: int) -> float:
        ms = C.c_float()
        _lib.check(self.lib.ka9q_frontend_rerun_resident(self.h, nsamples, C.byref(ms)), "rerun_resident")
        return ms.value

    def status(self) -> dict:
        st = _lib.FrontendStatus()
        _lib.check(self.lib.ka9q_frontend_get_status(self.h, C.byref(st)), "frontend_get_status")
        return {f: getattr(st, f) for f, _ in st._fields_}

    def close(self):
        if getattr(self, "h", None):
            self.lib.ka9q_frontend_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
