"""The five BASELINE.json configs as concrete channel plans.

A plan is: input rate, geometry (L, M, N, D with N = 2048*D), and a list of channels (mode name, carrier bin,
optional per-channel filter edges). Carriers sit on the FFT bin grid f_c = bin*Fs/N, where the shared-FFT channelizer
is provably equal to the reference's mix-then-FFT.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from . import synth


@dataclass
class ChannelSpec:
    mode: str
    bin: int
    low: float | None = None
    high: float | None = None


@dataclass
class Plan:
    name: str
    samprate: int
    D: int
    L: int
    M: int
    N: int
    channels: list[ChannelSpec] = field(default_factory=list)
    amplitude: float = 0.02
    sigma: float = 0.004
    seed: int = 1
    deviation: float = 2500.0


def _plan(name, fs, seed, amplitude, sigma, deviation=2500.0):
    D, L, M, N = synth.geometry(fs)
    return Plan(name, fs, D, L, M, N, [], amplitude, sigma, seed, deviation)


def cfg1() -> Plan:
    p = _plan("cfg1: 1x NBFM @192 kS/s", 192000, 1, 0.25, 0.02, 3000.0)
    p.channels = [ChannelSpec("FM", 2048)]
    return p


def cfg2() -> Plan:
    p = _plan("cfg2: 1x USB+AGC @1.92 MS/s", 1920000, 2, 0.05, 0.005)
    p.channels = [ChannelSpec("USB", 4096)]
    return p


def cfg3() -> Plan:
    p = _plan("cfg3: 64x NBFM @1.92 MS/s", 1920000, 3, 0.02, 0.004)
    p.channels = [ChannelSpec("FM", 1024 * (j - 32) + 512) for j in range(64)]
    return p


def cfg4(nchan: int = 1024) -> Plan:
    p = _plan("cfg4: 1024x FM/FM/AM/USB @19.2 MS/s", 19200000, 4, 0.005, 0.002)
    pattern = ("FM", "FM", "AM", "USB")
    p.channels = [ChannelSpec(pattern[j % 4], 800 * (j - 512)) for j in range(nchan)]
    return p


def cfg5(nchan: int = 8192, offset_bins: int = 0) -> Plan:
    """8192 NBFM channels on 61.44 MS/s, 7.03125 kHz raster (300 bins), edges narrowed to +-3 kHz so neighbours do
    not overlap. offset_bins shifts the whole raster (used to give every GPU rank distinct carriers)."""
    p = _plan("cfg5: 8192x NBFM @61.44 MS/s", 61440000, 5, 0.002, 0.003, 1000.0)
    p.channels = [ChannelSpec("FM", 300 * (j - nchan // 2) + offset_bins, -3000.0, 3000.0) for j in range(nchan)]
    return p


def shard_channels(plan: Plan, rank: int, world: int) -> list[ChannelSpec]:
    """Channel sharding for multi-GPU runs: channel c -> rank c mod world. Channels are independent after
    the forward FFT (in the reference they are independent processes), so no channel state ever moves between GPUs."""
    return [c for i, c in enumerate(plan.channels) if i % world == rank]


def shard_contiguous(plan: Plan, rank: int, world: int) -> list[ChannelSpec]:
    """Frequency-contiguous sharding: channels sorted by signed carrier bin, rank r takes the r-th chunk. A
    rank's channels then read one arc of the spectrum (1/world of it plus a 2048-bin halo), which is all the multi-GPU
    exchange has to deliver to it (ka9q_stream_mgpu_*)."""
    N = plan.N

    def signed(c):
        b = c.bin % N
        return b - N if b > N // 2 else b
    order = sorted(plan.channels, key=signed)
    n = len(order)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return order[lo:hi]


CONFIGS = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}

# Algorithmic bytes per channel-block (SURVEY §8d / BASELINE.md §2): spectrum window + own response + PCM + carried state
NDEC = 2048


def channel_block_bytes(mode_demod: int, pcm_channels: int, olen: int, flat: bool = False) -> int:
    b = 8 * NDEC + 8 * NDEC + 2 * olen * pcm_channels
    if mode_demod == 2 and not flat:          # FM: audio history read + write, 2*4*(M_dec-1)
        b += 2 * 4 * (NDEC - olen)
    else:
        b += 64
    return b


def stream_block_bytes(L: int, N: int, bytes_per_sample: int = 4) -> int:
    """int16 I/Q as on the wire + the N-point spectrum written once."""
    return bytes_per_sample * L + 8 * N
