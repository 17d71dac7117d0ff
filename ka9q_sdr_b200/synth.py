"""Synthetic int16 I/Q stimulus for the parity tests and the bench.

All stimuli are what an SDR front end would put on the wire as RTP payload type IQ_PT: interleaved
int16 I,Q in host byte order, full scale 32767 (reference radio.c:38,113-114). Every stimulus carries
AWGN: on noise-free input the reference's FM squelch estimator is a rounding-noise coin flip
(fm.c:101-115).
"""
from __future__ import annotations

import numpy as np


def geometry(samprate: int, out_rate: int = 48000, block_ms: int = 20):
    """Return (D, L, M, N) with N = 2048*D at 48 kHz out — the reference default geometry
    (main.c:113-114: L=3840, M=4353 at 192 kHz) scaled to the input rate."""
    if samprate % out_rate:
        raise ValueError("input rate must be an integer multiple of the output rate (radio_status.c:266)")
    D = samprate // out_rate
    olen = out_rate * block_ms // 1000
    L = olen * D
    ndec = 1 << int(np.ceil(np.log2(2 * olen)))
    N = ndec * D
    M = N - L + 1
    return D, L, M, N


def _quantize(x: np.ndarray) -> np.ndarray:
    iq = np.empty(2 * x.size, dtype=np.int16)
    iq[0::2] = np.clip(np.rint(x.real * 32767.0), -32767, 32767).astype(np.int16)
    iq[1::2] = np.clip(np.rint(x.imag * 32767.0), -32767, 32767).astype(np.int16)
    return iq


def awgn(rng: np.random.Generator, n: int, sigma: float) -> np.ndarray:
    return sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


def fm_carrier(n: int, samprate: float, f_c: float, tone_hz: float, deviation_hz: float, amplitude: float,
               phase0: float = 0.0) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / samprate
    beta = deviation_hz / tone_hz
    return amplitude * np.exp(1j * (2 * np.pi * f_c * t + beta * np.sin(2 * np.pi * tone_hz * t) + phase0))


def am_carrier(n: int, samprate: float, f_c: float, tone_hz: float, depth: float, amplitude: float) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / samprate
    return amplitude * (1 + depth * np.sin(2 * np.pi * tone_hz * t)) * np.exp(2j * np.pi * f_c * t)


def ssb_two_tone(n: int, samprate: float, f_c: float, tones_hz, amplitudes, ramp_hz: float = 0.0,
                 ramp_db: float = 0.0) -> np.ndarray:
    """Upper-sideband audio tones at f_c + tone; optional slow amplitude ramp to exercise AGC."""
    t = np.arange(n, dtype=np.float64) / samprate
    x = np.zeros(n, dtype=np.complex128)
    for f, a in zip(tones_hz, amplitudes):
        x += a * np.exp(2j * np.pi * (f_c + f) * t)
    if ramp_hz > 0:
        env = 10.0 ** ((ramp_db / 20.0) * 0.5 * (1 + np.sin(2 * np.pi * ramp_hz * t)))
        x *= env
    return x


def cfg1_fm(nblocks: int, seed: int = 1):
    """cfg1: 192 kS/s, one NBFM carrier at bin +2048 (+48 kHz), 1 kHz tone, 3 kHz deviation."""
    fs = 192000
    D, L, M, N = geometry(fs)
    n = nblocks * L
    rng = np.random.default_rng(seed)
    k = 2048
    x = fm_carrier(n, fs, k * fs / N, 1000.0, 3000.0, 0.25) + awgn(rng, n, 0.02)
    return dict(samprate=fs, D=D, L=L, M=M, N=N, iq=_quantize(x), bins=[k], modes=["FM"])


def cfg2_usb(nblocks: int, seed: int = 2, samprate: int = 1920000):
    """cfg2: USB two-tone (700+1900 Hz) at bin +4096 with a slow 6 dB ramp."""
    fs = samprate
    D, L, M, N = geometry(fs)
    n = nblocks * L
    rng = np.random.default_rng(seed)
    k = 4096
    x = ssb_two_tone(n, fs, k * fs / N, (700.0, 1900.0), (0.05, 0.05), ramp_hz=0.5, ramp_db=6.0) + awgn(rng, n, 0.005)
    return dict(samprate=fs, D=D, L=L, M=M, N=N, iq=_quantize(x), bins=[k], modes=["USB"])


def multi_channel(samprate: int, nblocks: int, bins, modes, seed: int, amplitude: float, sigma: float,
                  tone0: float = 300.0, tone_step: float = 37.0, deviation: float = 2500.0):
    """Equal-power multiplex: channel j at bin bins[j] with mode modes[j] (FM / AM / USB / LSB ...).
    Generated blockwise in the time domain; cost O(channels * samples)."""
    fs = samprate
    D, L, M, N = geometry(fs)
    n = nblocks * L
    rng = np.random.default_rng(seed)
    x = awgn(rng, n, sigma)
    for j, (k, m) in enumerate(zip(bins, modes)):
        f_c = k * fs / N
        tone = tone0 + tone_step * (j % 64)
        mu = m.upper()
        if mu.startswith("FM"):
            x += fm_carrier(n, fs, f_c, tone, deviation, amplitude, phase0=0.37 * j)
        elif mu == "AM":
            x += am_carrier(n, fs, f_c, 1000.0, 0.5, amplitude)
        elif mu in ("LSB", "CWL"):
            x += ssb_two_tone(n, fs, f_c, (-tone, -tone - 600.0), (amplitude / 2, amplitude / 2))
        else:
            x += ssb_two_tone(n, fs, f_c, (tone, tone + 600.0), (amplitude / 2, amplitude / 2))
    return dict(samprate=fs, D=D, L=L, M=M, N=N, iq=_quantize(x), bins=list(bins), modes=list(modes))


def comb_spectrum_iq(samprate: int, nblocks: int, bins, seed: int, amplitude: float, sigma: float,
                     tone0: float = 300.0, tone_step: float = 50.0, deviation: float = 1000.0):
    """Cheap wide-band stimulus for the throughput configs (thousands of channels, tens of MS/s).

    Every channel carries a continuous, phase-coherent NBFM signal: the modulating tones are multiples of the block
    rate (50 Hz) and the carriers sit on the L-point grid (multiples of Fs/L = 50 Hz, i.e. within 25 Hz of the channel
    centre bin*Fs/N), so the whole multiplex is exactly periodic in one block. It is therefore synthesised once, in
    the frequency domain (one 960-point FFT per channel placed around its carrier, one L-point inverse FFT: cost
    O(samples log samples), not O(channels * samples)), and repeated for every block with fresh AWGN on top.
    Parity cases use multi_channel() (time-domain synthesis on the exact bin grid) instead."""
    fs = samprate
    D, L, M, N = geometry(fs)
    rng = np.random.default_rng(seed)
    prng = np.random.default_rng(seed + 1000)
    olen = L // D
    t48 = np.arange(olen, dtype=np.float64) / (fs / D)
    spec = np.zeros(L, dtype=np.complex128)
    h = olen // 2
    for j, k in enumerate(bins):
        tone = tone0 + tone_step * (j % 16)
        # random carrier and modulation phases keep the crest factor of the multiplex Gaussian-like (no clipping)
        ph = (deviation / tone) * np.sin(2 * np.pi * tone * t48 + prng.uniform(0, 2 * np.pi)) + prng.uniform(0, 2 * np.pi)
        S = np.fft.fft(np.exp(1j * ph))
        kc = int(round(k * L / N))
        idx = (kc + np.arange(-h, h)) % L
        spec[idx] += np.concatenate([S[-h:], S[:h]]) * (amplitude * D)
    clean = np.fft.ifft(spec).astype(np.complex64)
    out = np.empty(2 * nblocks * L, dtype=np.int16)
    for b in range(nblocks):
        x = clean + (sigma * (rng.standard_normal(L) + 1j * rng.standard_normal(L))).astype(np.complex64)
        out[2 * b * L:2 * (b + 1) * L] = _quantize(x)
    return dict(samprate=fs, D=D, L=L, M=M, N=N, iq=out, bins=list(bins))
