"""TEST / BENCH INFRASTRUCTURE ONLY — a CPU restatement of the *shared-FFT channelizer* (NOT of the reference).

The reference runs one `radio` process per channel: per-sample LO + its own N-point forward FFT + N/D-point inverse +
demodulator (oracle/_ref and oracle/port.py restate that). The product restructures the path: ONE forward FFT per
stream block, then per channel a rotated 2048-bin window x response -> inverse FFT -> demodulator.
This file does that restructured algorithm on the host with numpy/scipy, vectorised over channels, for NBFM, so that
bench.py can report the algorithmic speed-up (shared FFT) and the hardware speed-up (B200 vs host cores) separately
(SURVEY 8d). Only bench.py's CPU legs and tests/ may import it.

Per block and channel, with the reference file:line each step restates:
  ingest int16 * SCALE16 * gain               radio.c:113-122
  window = last M-1 samples + L new, FFT      filter.c:146-172
  Y[p] = H[p] * X[(bin + s(p)) mod N]         filter.c:206-227 with the bin rotation of Appendix C
  y = IFFT_2048(Y), keep the last olen        filter.c:250, :131
  squelch statistics, discriminator, audio    fm.c:86-173 (blanking: samples below 0.55*avg repeat the last good one)
  filter (REAL overlap-save), gain, scaleclip fm.c:162-171, audio.c:22-28
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft as sfft

from . import port

F = np.float32


class FmChannelizer:
    """K NBFM channels of one I/Q stream; process(iq_block) -> int16 PCM [K, olen] per block."""

    def __init__(self, samprate: int, L: int, M: int, decimate: int, bins, low: float, high: float, *,
                 kaiser_beta: float = 3.0, headroom: float | None = None, gain_factor: float = 1.0, workers: int = -1):
        self.fs, self.L, self.M, self.D = samprate, L, M, decimate
        self.N = L + M - 1
        self.Nd = self.N // decimate
        self.olen = L // decimate
        self.workers = workers
        self.gain_factor = F(gain_factor)
        self.bins = np.asarray(bins, dtype=np.int64)
        K = self.bins.size
        if headroom is None:
            headroom = float(F(math.pow(10.0, -15.0 / 20)))                                      # main.c:117
        dsr = F(F(samprate) / F(decimate))
        self.dsr = dsr
        lo, hi = (low, high) if low <= high else (high, low)
        self.H, _ = port.set_filter_response(L, M, decimate, port.COMPLEX, F(F(lo) / dsr), F(F(hi) / dsr), kaiser_beta)
        # bin offsets of the N_dec window: DC..+Nyquist, then the negative frequencies (filter.c:206-227)
        p = np.arange(self.Nd)
        s = np.where(p <= self.Nd // 2, p, p - self.Nd)
        self.idx = (self.bins[:, None] + s[None, :]) % self.N                                    # [K, N_dec]
        self.hist = np.zeros(self.M - 1, dtype=np.complex64)                                     # filter.c:77
        self.block = 0
        # FM state per channel (fm.c:26-31)
        self.state = np.ones(K, dtype=np.complex64)
        self.lastaudio = np.zeros(K, dtype=F)
        self.below = np.zeros(K, dtype=np.int32)
        # post-detection audio filter (fm.c:39-66): REAL overlap-save, AL = olen, AM = N_dec - olen + 1
        AL, AM, AN = self.olen, self.Nd - self.olen + 1, self.Nd
        ar = np.zeros(AN // 2 + 1, dtype=np.complex64)
        j = np.arange(AN // 2 + 1)
        f = (j.astype(F) * dsr / F(AN)).astype(F)
        sel = (f >= 300) & (f <= 6000)
        ar[sel] = (F(10.0 / AN) * F(300.0) / f[sel]).astype(F)
        self.AR = port.window_rfilter(AL, AM, ar, kaiser_beta)
        self.audio_hist = np.zeros((K, AM - 1), dtype=F)
        self.gain = F((headroom * (1 / math.pi) * float(dsr)) / abs(float(F(lo) - F(hi))))       # fm.c:86

    def process(self, iq_block: np.ndarray) -> np.ndarray:
        L, N, Nd, olen = self.L, self.N, self.Nd, self.olen
        v = np.asarray(iq_block).reshape(-1, 2)
        assert v.shape[0] == L
        x = ((v[:, 0].astype(F) * port.SCALE16) * self.gain_factor + 1j * ((v[:, 1].astype(F) * port.SCALE16) * self.gain_factor)).astype(np.complex64)
        win = np.concatenate((self.hist, x))
        X = sfft.fft(win, workers=self.workers)                                                  # ONE forward FFT per block
        self.hist = win[L:].copy()
        Y = X[self.idx] * self.H[None, :]
        y = (sfft.ifft(Y, axis=1, workers=self.workers) * F(Nd)).astype(np.complex64)[:, Nd - olen:]
        # per-block LO phase exp(j*2*pi*((-k*s_m) mod N)/N), s_m = m*L - (M-1)  (Appendix C)
        s_m = self.block * L - (self.M - 1)
        e = (-(self.bins * s_m)) % N
        y = y * np.exp(2j * np.pi * e / N).astype(np.complex64)[:, None]
        self.block += 1
        # ---- demod_fm, vectorised over channels
        t = (y.real ** 2 + y.imag ** 2).astype(F)
        bb_power = t.sum(axis=1, dtype=F) / F(2 * olen)
        avg_amp = np.sqrt(t).sum(axis=1, dtype=F) / F(math.sqrt(2) * olen)
        var = bb_power - avg_amp * avg_amp
        with np.errstate(divide="ignore", invalid="ignore"):
            snr = avg_amp * avg_amp / (F(2) * var) - F(1)
        snr = np.where(np.isnan(snr), snr, np.maximum(F(0), snr))
        self.below = np.where(snr > 2, 0, np.minimum(self.below + 1, 1000)).astype(np.int32)
        is_open = self.below < 2
        min_ampl = (F(0.55 * 0.55) * avg_amp * avg_amp).astype(F)
        good = t > min_ampl[:, None]
        n = np.arange(olen)[None, :]
        last_good = np.maximum.accumulate(np.where(good, n, -1), axis=1)                         # last good index <= n
        prev_good = np.concatenate((np.full((y.shape[0], 1), -1), last_good[:, :-1]), axis=1)    # last good index < n
        rows = np.arange(y.shape[0])[:, None]
        st = np.where(prev_good >= 0, np.conj(y[rows, np.maximum(prev_good, 0)]), self.state[:, None]).astype(np.complex64)
        prod = y * st
        ang = np.arctan2(prod.imag, prod.real).astype(F)
        audio = np.where(last_good >= 0, ang[rows, np.maximum(last_good, 0)], self.lastaudio[:, None]).astype(F)
        audio = np.where(is_open[:, None], audio, F(0))
        any_good = last_good[:, -1] >= 0
        new_state = np.where(any_good, np.conj(y[rows[:, 0], np.maximum(last_good[:, -1], 0)]), self.state)
        self.state = np.where(is_open, new_state, np.complex64(0)).astype(np.complex64)           # fm.c:156
        self.lastaudio = np.where(is_open, audio[:, -1], F(0)).astype(F)
        # ---- audio filter: REAL overlap-save of length N_dec
        a_in = np.concatenate((self.audio_hist, audio), axis=1)
        self.audio_hist = a_in[:, olen:].copy()
        A = sfft.rfft(a_in, axis=1, workers=self.workers)
        out = sfft.irfft(A * self.AR[None, :], n=Nd, axis=1, workers=self.workers)[:, Nd - olen:] * F(Nd)
        return port.scaleclip((out.astype(F) * self.gain).astype(F))
