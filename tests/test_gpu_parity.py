"""GPU parity tests (run with -m gpu on the B200 box). Everything goes through the C ABI (libka9q_b200.so); the checker
is the oracle: committed golden vectors (produced by the verbatim reference), the verbatim reference library itself
(oracle/_ref, which travels with the snapshot) and the numpy port. Tolerances are the ones BASELINE.json states:
filter output <= 1e-5 relative RMS (fp32), PCM within +-1 LSB after 16-bit quantisation; integer/bit-exact where the
comparison is GPU-vs-GPU (batching, sharding)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from ka9q_sdr_b200 import _lib, channelizer as ch, modes, synth, workloads
from oracle import port

sys.path.insert(0, os.path.join(ROOT, "scripts"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu

FILT_TOL = 1e-5   # relative RMS, north_star
PCM_TOL = 1       # LSB, north_star


def rel_rms(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def run_gpu(cfg, chans, nblocks, max_blocks=4, capture=True, **kw):
    """chans: list of (mode, bin, extra kwargs). Returns (channelizer, pcm rows, status)."""
    c = ch.Channelizer(cfg["samprate"], cfg["L"], cfg["M"], cfg["D"], max_blocks=max_blocks,
                       capture_filter_output=capture, **kw)
    for mode, k, extra in chans:
        c.add_channel(mode, k, **extra)
    c.commit()
    iq = cfg["iq"][:2 * nblocks * cfg["L"]]
    filt = [[] for _ in chans]
    pcm = np.empty((nblocks, c.pcm_stride), dtype=np.int16)
    sts = []
    b = 0
    while b < nblocks:
        nb = min(max_blocks, nblocks - b)
        _, st = c.process(iq[2 * b * cfg["L"]:2 * (b + nb) * cfg["L"]], pcm[b:b + nb])
        sts.append(st)
        if capture:
            for i in range(len(chans)):
                filt[i].append(c.filter_output(i, nb))
        b += nb
    filt = [np.concatenate(f) for f in filt] if capture else None
    return c, pcm, np.concatenate(sts), filt


def pcm_channels(mode):
    m = modes.get_mode(mode)
    return m.channels if m.demod_type == modes.LINEAR_DEMOD else 1


def check_pcm(mode, got, want, olen, label=""):
    n = min(got.size, want.size)
    d = np.abs(got[:n].astype(np.int32) - want[:n].astype(np.int32))
    # AM / linear: the AGC normalises the near-zero start-up ramp of block 0
    skip = 0 if modes.get_mode(mode).demod_type == modes.FM_DEMOD else olen * pcm_channels(mode)
    assert d[skip:].max() <= PCM_TOL, f"{label}{mode}: PCM differs by {d[skip:].max()} LSB"
    assert (d[skip:] == 0).mean() > 0.97, f"{label}{mode}: only {(d[skip:] == 0).mean():.4f} of samples bit-equal"


# ------------------------------------------------------------------------------------------ K1+K2: forward FFT

@pytest.mark.parametrize("n", [64, 1600, 2048, 8192, 49152, 81920, 819200, 2621440])
def test_fft_sizes_against_float64(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for sign in (-1, +1):
        y = ch.fft_c2c(x, sign)
        want = np.fft.fft(x.astype(np.complex128)) if sign < 0 else np.fft.ifft(x.astype(np.complex128)) * n
        assert rel_rms(y, want) < 5e-7


def test_fft_batched_and_linear():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((5, 8192)) + 1j * rng.standard_normal((5, 8192))).astype(np.complex64)
    y = ch.fft_c2c(x, -1)
    for i in range(5):
        assert rel_rms(y[i], np.fft.fft(x[i].astype(np.complex128))) < 5e-7
    # round trip: IFFT(FFT(x)) = N x
    back = ch.fft_c2c(y, +1)
    assert rel_rms(back, x * 8192) < 1e-6


def test_ingest_and_spectrum_match_reference_forward_fft(ref):
    """int16 ingest (radio.c:113-122) + overlap-save window (filter.c:159-170) + forward FFT (filter.c:151), k = 0."""
    cfg = synth.cfg1_fm(3)
    L, M, N = cfg["L"], cfg["M"], cfg["N"]
    c, _, _, _ = run_gpu(cfg, [("FM", 0, {})], 3, max_blocks=3, gain_factor=0.5)
    x = (cfg["iq"][0::2].astype(np.float32) * np.float32(1 / 32767) * np.float32(0.5)) + \
        1j * (cfg["iq"][1::2].astype(np.float32) * np.float32(1 / 32767) * np.float32(0.5))
    r = ref.filter_run(L, M, 4, ref.COMPLEX, ref.COMPLEX, x.astype(np.complex64), low=-0.1, high=0.1, want_fdomain=True)
    assert rel_rms(c.spectrum(2), r["fdomain"]) < 1e-6
    # if_power numerator (radio.c:123): sum |x|^2 over the L new samples of each block
    e = c.if_energy(3)
    want = [(np.abs(x[b * L:(b + 1) * L]) ** 2).sum() for b in range(3)]
    np.testing.assert_allclose(e, want, rtol=1e-4)


# ------------------------------------------------------------------------------------------ K4: filter design

def test_filter_design_matches_golden():
    g = np.load(os.path.join(GOLDEN, "design.npz"))
    cfg = synth.cfg1_fm(1)
    c = ch.Channelizer(cfg["samprate"], cfg["L"], cfg["M"], cfg["D"])
    c.add_channel("FM", 100)
    c.add_channel("USB", 200)   # the 3000 Hz edge falls exactly on a bin: the inclusive float compare decides
    c.add_channel("ISB", 300)
    c.commit()
    for i, name in enumerate(("fm", "usb", "isb")):
        H, ng = c.response(i)
        assert rel_rms(H, g[name]) < 2e-6, name
        assert ng == pytest.approx(float(g[name + "_ng"][0]), rel=1e-4)
    # retune after commit (display.c:163-177 path)
    c.set_filter(0, -4000.0, 4000.0, 3.0)
    H2, ng2 = c.response(0)
    want, wng = port.set_filter_response(cfg["L"], cfg["M"], cfg["D"], port.COMPLEX, np.float32(-4000) / np.float32(48000),
                                         np.float32(4000) / np.float32(48000), 3.0)
    assert rel_rms(H2, want) < 2e-6 and ng2 == pytest.approx(wng, rel=1e-4)


# ------------------------------------------------------------------------------------------ K3: golden chain parity

GOLDEN_CASES = [("fm", "FM"), ("fm", "FMF"), ("am", "AM"), ("usb", "USB"), ("usb", "LSB"), ("usb", "IQ"), ("usb", "ISB"),
                ("usb", "CWU")]


@pytest.mark.parametrize("stim,mode", GOLDEN_CASES)
def test_chain_matches_golden(stim, mode):
    g = np.load(os.path.join(GOLDEN, f"chain_{mode.lower()}.npz"))
    cfg = make_golden.stimulus(stim)
    assert int(cfg["iq"].astype(np.int64).sum()) == int(g["iq_crc"][0])
    fs, L, M, D, N, k, nb = (int(v) for v in g["meta"])
    c, pcm, st, filt = run_gpu(cfg, [(mode, k, {})], nb, max_blocks=3)
    check_pcm(mode, c.channel_pcm(pcm, 0), g["pcm"], L // D)
    if g["filt"].size:
        assert rel_rms(filt[0], g["filt"]) < FILT_TOL
    if mode == "FM":
        np.testing.assert_allclose(st["bb_power"][:, 0], g["bb_power"], rtol=1e-4)
        np.testing.assert_allclose(st["snr"][1:, 0], g["snr"][1:], rtol=3e-2)
        np.testing.assert_allclose(st["pdeviation"][2:, 0], g["pdeviation"][2:], rtol=2e-3)
        np.testing.assert_allclose(st["foffset"][2:, 0], g["foffset"][2:], atol=0.05)
        assert st["squelch_open"][:, 0].all()


@pytest.mark.parametrize("mode", ["FM", "AM", "USB"])
def test_long_run_against_reference(ref, mode):
    """150 blocks = 3 s: AGC attack / hang / recovery and squelch steady state."""
    nb = 150
    if mode == "FM":
        cfg = synth.cfg1_fm(nb)
    elif mode == "USB":
        cfg = synth.cfg2_usb(nb, samprate=192000)
    else:
        D, L, M, N = synth.geometry(192000)
        rng = np.random.default_rng(3)
        n = nb * L
        x = synth.am_carrier(n, 192000, 1024 * 192000 / N, 1000.0, 0.5, 0.1) + synth.awgn(rng, n, 0.01)
        cfg = dict(samprate=192000, D=D, L=L, M=M, N=N, iq=synth._quantize(x), bins=[1024])
    fs, L, M, D, N = cfg["samprate"], cfg["L"], cfg["M"], cfg["D"], cfg["N"]
    k = cfg["bins"][0]
    c, pcm, st, filt = run_gpu(cfg, [(mode, k, {})], nb, max_blocks=8)
    r = ref.chain_run(mode, fs, L, M, D, cfg["iq"], carrier_hz=k * fs / N, lo_cycles=-k / N, want_filt=mode != "USB")
    check_pcm(mode, c.channel_pcm(pcm, 0), r.pcm, L // D)
    if r.filt is not None:
        assert rel_rms(filt[0], r.filt[:nb]) < FILT_TOL
    # AGC gain trajectory (reference captures it at send time, i.e. after the block)
    if mode != "FM":
        np.testing.assert_allclose(st["agc_gain"][5:, 0], r.status["agc_gain"][5:nb], rtol=2e-4)


@pytest.mark.parametrize("L,M", [(4096, 4097), (3072, 5121)])
def test_other_overlap_geometries_against_reference(ref, L, M):
    """N/decimate stays 2048 but olen = L/4 is not the default 960 (1024 and 768): the run-time-olen kernels
    (fm_kernel<0>, agc_kernel<..., 0>), another audio-filter length (AM = 2049 - olen, fm.c:39-43) and, for FM, one
    paired and one unpaired channel."""
    fs, D, N, nb = 192000, 4, 8192, 12
    rng = np.random.default_rng(L)
    n = nb * L
    kf, kf2, ka, ku = 2048, -2304, -1024, 512
    x = (synth.fm_carrier(n, fs, kf * fs / N, 1000.0, 3000.0, 0.2)
         + synth.fm_carrier(n, fs, kf2 * fs / N, 700.0, 2500.0, 0.2)
         + synth.am_carrier(n, fs, ka * fs / N, 1000.0, 0.5, 0.1)
         + synth.ssb_two_tone(n, fs, ku * fs / N, [700.0, 1900.0], [0.05, 0.05])
         + synth.awgn(rng, n, 0.01))
    cfg = dict(samprate=fs, D=D, L=L, M=M, N=N, iq=synth._quantize(x))
    chans = [("FM", kf, {}), ("AM", ka, {}), ("USB", ku, {}), ("FM", kf2, {}), ("FM", kf, {})]   # 3 FM: a pair + one
    c, pcm, st, filt = run_gpu(cfg, chans, nb, max_blocks=5)
    assert c.olen == L // D
    for i, (mode, k, _) in enumerate(chans):
        r = ref.chain_run(mode, fs, L, M, D, cfg["iq"], carrier_hz=k * fs / N, lo_cycles=-k / N, want_filt=mode != "USB")
        check_pcm(mode, c.channel_pcm(pcm, i), r.pcm, L // D, label=f"{mode}@{k}")
        if r.filt is not None:
            assert rel_rms(filt[i], r.filt[:nb]) < FILT_TOL


def test_split_and_fused_agc_kernels_are_bit_identical(monkeypatch):
    """AM / linear run as front + per-lane recurrence + output kernels by default; the fused single-kernel form
    (KA9Q_B200_AGC_FUSED=1) must give the same PCM and status bit for bit, also across batch boundaries, with more
    than 32 channels per class (several recurrence warps, a partial last warp) and an olen that is not a multiple of 32."""
    for L, M, nb in ((3840, 4353, 11), (4000, 4193, 7)):      # olen 960 and 1000
        fs, D, N = 192000, 4, 8192
        rng = np.random.default_rng(L)
        n = nb * L
        x = synth.awgn(rng, n, 0.02)
        chans = []
        for i in range(37):
            k = -3000 + 160 * i
            x = x + synth.am_carrier(n, fs, k * fs / N, 400.0 + 20 * i, 0.5, 0.01 + 0.001 * i)
            chans.append((["AM", "USB", "IQ", "CWU", "ISB"][i % 5], k, {}))
        t = np.arange(n) / fs
        x = x * (1.0 + 0.8 * np.sin(2 * np.pi * 2.0 * t))       # slow fading: attack, hang and recovery all happen
        cfg = dict(samprate=fs, D=D, L=L, M=M, N=N, iq=synth._quantize(x))
        res = {}
        for fused in ("0", "1"):
            monkeypatch.setenv("KA9Q_B200_AGC_FUSED", fused)
            c, pcm, st, filt = run_gpu(cfg, chans, nb, max_blocks=4, capture=True)
            res[fused] = (pcm.copy(), st.copy(), filt)
        assert np.array_equal(res["0"][0], res["1"][0])
        for i, (mode, _, _) in enumerate(chans):   # the debug capture of the filter output takes a separate path in each form
            if mode != "ISB":                      # (no capture for CROSS_CONJ outputs: two sidebands, not one signal)
                assert np.array_equal(res["0"][2][i], res["1"][2][i]), i
        for f in ("bb_power", "agc_gain", "reserved"):
            assert np.array_equal(res["0"][1][f], res["1"][1][f], equal_nan=True), f
        assert np.abs(res["0"][0]).max() > 1000


def test_fm_block_split_form_is_bit_identical(monkeypatch):
    """With few FM pairs per GPU the FM kernel splits every pair's blocks over a cluster of 2 or 4 CTAs (launch_fm,
    fm_split_wait: state and ring order handed from block to block through global memory). PCM, status and the state
    carried into later batches must equal the one-CTA-per-pair form bit for bit: carriers that drop in and out (squelch
    state machine, the carried discriminator state), a weak carrier (threshold-extension blanking), an odd channel count
    (a pair without partner), FLAT channels (no audio filter), batches of 7, 5 and 1 blocks."""
    fs, D, L, M, N = 192000, 4, 3840, 4353, 8192
    nb = 27
    n = nb * L
    rng = np.random.default_rng(77)
    t = np.arange(n) / fs
    x = synth.awgn(rng, n, 0.01)
    chans = []
    for i in range(9):
        k = -3200 + 800 * i
        gate = ((t * (3 + i)) % 1.0) < (0.55 + 0.04 * i) if i % 3 == 0 else 1.0     # every third carrier keys on and off
        amp = 0.012 if i == 4 else 0.08                                              # one near the FM threshold
        x = x + gate * synth.fm_carrier(n, fs, k * fs / N, 300.0 + 100 * i, 2500.0, amp)
        chans.append(("FMF" if i in (2, 7) else "FM", k, {}))
    cfg = dict(samprate=fs, D=D, L=L, M=M, N=N, iq=synth._quantize(x))
    res = {}
    for split in ("1", "2", "4"):
        monkeypatch.setenv("KA9Q_B200_FM_SPLIT", split)
        out = []
        for mb in (7, 5, 1):
            c, pcm, st, filt = run_gpu(cfg, chans, nb, max_blocks=mb, capture=False)
            out.append((pcm.copy(), st.copy()))
            c.close()
        res[split] = out
    monkeypatch.delenv("KA9Q_B200_FM_SPLIT")
    c, pcm, st, filt = run_gpu(cfg, chans, nb, max_blocks=7, capture=False)          # the launcher's own choice
    res["auto"] = [(pcm.copy(), st.copy())]
    c.close()
    ref_pcm, ref_st = res["1"][0]
    assert np.abs(ref_pcm).max() > 1000
    opened = ref_st["squelch_open"]
    assert opened[:, 0].min() == 0 and opened[:, 0].max() == 1                       # the squelch really moved
    assert (ref_st["reserved"][:, 4, 0] == 0).any()                                  # the blanking path really ran
    for key, outs in res.items():
        for j, (p, s_) in enumerate(outs):
            assert np.array_equal(p, ref_pcm), (key, j)
            for f in ("bb_power", "snr", "foffset", "pdeviation", "squelch_open", "reserved"):
                assert np.array_equal(s_[f], ref_st[f], equal_nan=True), (key, j, f)


def test_fm_squelch_closes_on_noise_like_the_reference(ref):
    cfg = synth.cfg1_fm(12)
    rng = np.random.default_rng(9)
    iq = cfg["iq"].copy()
    L = cfg["L"]
    noise = synth._quantize(synth.awgn(rng, 6 * L, 0.02))
    iq[2 * 4 * L:2 * 10 * L] = noise  # carrier disappears for blocks 4..9
    cfg = dict(cfg, iq=iq)
    fs, M, D, N = cfg["samprate"], cfg["M"], cfg["D"], cfg["N"]
    k = cfg["bins"][0]
    c, pcm, st, _ = run_gpu(cfg, [("FM", k, {})], 12)
    r = ref.chain_run("FM", fs, L, M, D, iq, carrier_hz=k * fs / N, lo_cycles=-k / N)
    got = c.channel_pcm(pcm, 0)
    assert (st["squelch_open"][6:9, 0] == 0).all() and st["squelch_open"][11, 0] == 1
    d = np.abs(got.astype(np.int32) - r.pcm.astype(np.int32)).reshape(12, -1)
    # identical while the carrier is present or the squelch is shut; the blocks where noise is being demodulated and the
    # block where the squelch re-opens on a zeroed state (fm.c:156: cargf(x*0), a sign-of-zero coin flip) are excluded
    for b in (0, 1, 2, 3, 7, 8, 9):
        assert d[b].max() <= PCM_TOL, f"block {b}: {d[b].max()}"
    assert np.abs(got.reshape(12, -1)[8]).max() == 0  # shut squelch sends zeros (fm.c:155-160)


# ------------------------------------------------------------------------------------------ multi-channel configs

def test_cfg3_64_fm_channels_share_one_forward_fft(ref):
    plan = workloads.cfg3()
    nb = 6
    cfg = synth.multi_channel(plan.samprate, nb, [s.bin for s in plan.channels], [s.mode for s in plan.channels],
                              plan.seed, plan.amplitude, plan.sigma, deviation=plan.deviation)
    chans = [(s.mode, s.bin, {}) for s in plan.channels]
    c, pcm, st, filt = run_gpu(cfg, chans, nb, max_blocks=3)
    assert c.launches_per_call == 2 + 1  # two FFT passes for N=81920 + one FM launch for all 64 channels
    assert st["squelch_open"].all()
    for j in (0, 1, 31, 32, 63):  # both band edges and around DC
        k = plan.channels[j].bin
        r = ref.chain_run("FM", plan.samprate, plan.L, plan.M, plan.D, cfg["iq"], carrier_hz=k * plan.samprate / plan.N,
                          lo_cycles=-k / plan.N, want_filt=True, pkt_samples=4096)
        assert rel_rms(filt[j], r.filt[:nb]) < FILT_TOL, f"channel {j}"
        check_pcm("FM", c.channel_pcm(pcm, j), r.pcm, plan.L // plan.D, label=f"ch{j} ")


def test_mixed_modes_wraparound_and_negative_bins(ref):
    """cfg4 pattern (FM, FM, AM, USB) at 1.92 MS/s incl. a channel at -Fs/2 whose window wraps around the spectrum."""
    fs = 1920000
    D, L, M, N = synth.geometry(fs)
    nb = 5
    bins = [-N // 2, -N // 2 + 800, -20000, 12345, N // 2 - 1024, 3000]
    mds = ["FM", "AM", "USB", "FM", "LSB", "IQ"]
    cfg = synth.multi_channel(fs, nb, bins, mds, 4, 0.02, 0.004)
    c, pcm, st, filt = run_gpu(cfg, [(m, k, {}) for m, k in zip(mds, bins)], nb, max_blocks=5)
    for j, (m, k) in enumerate(zip(mds, bins)):
        r = ref.chain_run(m, fs, L, M, D, cfg["iq"], carrier_hz=k * fs / N, lo_cycles=-k / N,
                          want_filt=m in ("FM", "AM"), pkt_samples=4096)
        if r.filt is not None:
            assert rel_rms(filt[j], r.filt[:nb]) < FILT_TOL, (m, k)
        check_pcm(m, c.channel_pcm(pcm, j), r.pcm, L // D, label=f"bin {k} ")


def test_int8_input_and_gain_factor(ref):
    cfg = synth.cfg1_fm(4)
    iq8 = (cfg["iq"].astype(np.int32) // 258).astype(np.int8)
    fs, L, M, D, N = cfg["samprate"], cfg["L"], cfg["M"], cfg["D"], cfg["N"]
    k = cfg["bins"][0]
    c = ch.Channelizer(fs, L, M, D, max_blocks=4, iq_format=ch.IQ_S8, gain_factor=0.7, capture_filter_output=True)
    c.add_channel("FM", k)
    c.commit()
    pcm, _ = c.process(iq8)
    r = ref.chain_run("FM", fs, L, M, D, iq8, carrier_hz=k * fs / N, lo_cycles=-k / N, pkt_type=ref.IQ_PT8,
                      gain_factor=0.7, want_filt=True)
    assert rel_rms(c.filter_output(0, 4), r.filt[:4]) < FILT_TOL
    check_pcm("FM", c.channel_pcm(pcm, 0), r.pcm, L // D)


def test_lost_packet_zero_fill_matches_reference(ref):
    """radio.c:81-100: lost samples become zeros and the LO phase keeps counting; in the channelizer the host writes the
    zeros into the ring and the bin-rotation phase is a function of the sample index, so nothing else is needed."""
    cfg = synth.cfg1_fm(5)
    fs, L, M, D, N = cfg["samprate"], cfg["L"], cfg["M"], cfg["D"], cfg["N"]
    k = cfg["bins"][0]
    pkt = 960
    drop = np.zeros(cfg["iq"].size // 2 // pkt, dtype=np.uint8)
    drop[[3, 9]] = 1
    r = ref.chain_run("FM", fs, L, M, D, cfg["iq"], carrier_hz=k * fs / N, lo_cycles=-k / N, pkt_samples=pkt, drop=drop)
    iq = cfg["iq"].copy()
    for p in (3, 9):
        iq[2 * p * pkt:2 * (p + 1) * pkt] = 0
    c, pcm, _, _ = run_gpu(dict(cfg, iq=iq), [("FM", k, {})], 5, capture=False)
    check_pcm("FM", c.channel_pcm(pcm, 0), r.pcm, L // D)


# ------------------------------------------------------------------------------------------ GPU-vs-GPU invariances

def test_batching_and_channel_set_invariance_bit_exact():
    """Same samples, different launch shapes: blocks per call (1 vs 4) and channel population (alone vs among others,
    i.e. what sharding across GPUs changes). Blocks-per-call never changes a bit. Re-sharding never changes a bit of an
    AM / linear channel; a de-emphasised FM channel shares one complex audio transform with a partner channel (two real
    signals as re/im), so a different partner perturbs it at fp32 rounding level: <= 1 LSB, almost always 0."""
    plan = workloads.cfg3()
    nb = 8
    cfg = synth.multi_channel(plan.samprate, nb, [s.bin for s in plan.channels][:16], ["FM"] * 16, 3, 0.02, 0.004)
    sel = [(s.mode, s.bin, {}) for s in plan.channels[:16]]
    sel[5] = ("AM", sel[5][1], {})
    sel[6] = ("USB", sel[6][1], {})
    c1, pcm1, _, _ = run_gpu(cfg, sel, nb, max_blocks=1, capture=False)
    c4, pcm4, _, _ = run_gpu(cfg, sel, nb, max_blocks=4, capture=False)
    assert np.array_equal(pcm1, pcm4)
    # shard: even / odd channels on two independent channelizers (two "ranks")
    for shard in (0, 1):
        sub = sel[shard::2]
        cs, pcms, _, _ = run_gpu(cfg, sub, nb, max_blocks=4, capture=False)
        for i, j in enumerate(range(shard, 16, 2)):
            a, b = cs.channel_pcm(pcms, i), c4.channel_pcm(pcm4, j)
            if sel[j][0] == "FM":
                d = np.abs(a.astype(np.int32) - b.astype(np.int32))
                assert d.max() <= 1 and (d == 0).mean() > 0.995, f"channel {j}"
            else:
                assert np.array_equal(a, b), f"channel {j}"


def test_cfg5_full_size_properties():
    """Full BASELINE size (8192 NBFM channels, N = 2 621 440): properties that do not need the oracle at that size —
    every squelch opens, the demodulated deviation is the stimulus's, and sampled channels equal the same channels run
    alone in a small channelizer (the sharding invariance at full size)."""
    plan = workloads.cfg5()
    nb = 3
    iq = synth.comb_spectrum_iq(plan.samprate, nb, [s.bin for s in plan.channels], plan.seed, plan.amplitude, plan.sigma,
                                deviation=plan.deviation)["iq"]
    cfg = dict(samprate=plan.samprate, L=plan.L, M=plan.M, D=plan.D, N=plan.N, iq=iq)
    chans = [(s.mode, s.bin, dict(low=s.low, high=s.high)) for s in plan.channels]
    c, pcm, st, _ = run_gpu(cfg, chans, nb, max_blocks=3, capture=False)
    assert st["squelch_open"][1:].mean() > 0.999
    pdev = st["pdeviation"][2]
    assert abs(np.median(pdev) - plan.deviation) < 0.1 * plan.deviation
    sample = [0, 1, 4095, 4096, 8190, 8191]
    cs, pcms, _, _ = run_gpu(cfg, [chans[j] for j in sample], nb, max_blocks=3, capture=False)
    for i, j in enumerate(sample):
        d = np.abs(cs.channel_pcm(pcms, i).astype(np.int32) - c.channel_pcm(pcm, j).astype(np.int32))
        assert d.max() <= 1 and (d == 0).mean() > 0.995, f"channel {j}"  # FM pair partner differs: rounding level only
    c.close()


def test_cfg5_sampled_channels_against_reference(ref):
    """SURVEY §8d-5 parity leg: a handful of cfg5 channels, verbatim reference at N = 2 621 440 (MKL-free, so slow:
    kept to 3 channels x 3 blocks)."""
    plan = workloads.cfg5(256)
    nb = 3
    sel = [0, 128, 255]
    chans = [plan.channels[j] for j in sel]
    # sigma chosen for ~30 dB in-channel SNR: on a cleaner signal the reference's squelch estimator is a rounding-noise
    # coin flip (fm.c:101-115)
    cfgd = synth.multi_channel(plan.samprate, nb, [s.bin for s in chans], ["FM"] * len(chans), 5, 0.05, 0.1,
                               tone0=400.0, tone_step=50.0, deviation=1000.0)
    c, pcm, st, filt = run_gpu(cfgd, [(s.mode, s.bin, dict(low=s.low, high=s.high)) for s in chans], nb, max_blocks=3)
    ref.set_fft_backend("mkl")
    try:
        for i, s in enumerate(chans):
            r = ref.chain_run("FM", plan.samprate, plan.L, plan.M, plan.D, cfgd["iq"],
                              carrier_hz=s.bin * plan.samprate / plan.N, lo_cycles=-s.bin / plan.N, low=s.low, high=s.high,
                              want_filt=True, pkt_samples=4096)
            assert rel_rms(filt[i], r.filt[:nb]) < FILT_TOL, f"channel {sel[i]}"
            check_pcm("FM", c.channel_pcm(pcm, i), r.pcm, plan.L // plan.D, label=f"ch{sel[i]} ")
    finally:
        ref.set_fft_backend("standin")


# ------------------------------------------------------------------------------------------ drop-in layer (filter.h)

def _dropin_run(L, M, D, in_type, out_type, x, low, high, beta):
    lib = _lib.lib()
    m = lib.create_filter_input(L, M, in_type)
    assert m
    s = lib.create_filter_output(m, None, D, out_type)
    assert s
    assert lib.set_filter(s, low, high, beta) == 0
    fin = _FilterIn.from_address(m)
    fout = _FilterOut.from_address(s)
    N = L + M - 1
    Nd = N // D
    olen = L // D
    nb = x.size // L
    outs = []
    cplx_in = in_type != 3
    for b in range(nb):
        src = np.ascontiguousarray(x[b * L:(b + 1) * L])
        C.memmove(fin.input, src.ctypes.data, src.nbytes)
        assert lib.execute_filter_input(m) == 0
        assert lib.execute_filter_output(s) == 0
        if out_type == 3:
            outs.append(np.ctypeslib.as_array(C.cast(fout.output, C.POINTER(C.c_float)), (olen,)).copy())
        else:
            outs.append(np.ctypeslib.as_array(C.cast(fout.output, C.POINTER(C.c_float)), (2 * olen,)).copy().view(np.complex64))
    nbins = N if cplx_in else N // 2 + 1
    fd = np.ctypeslib.as_array(C.cast(fin.fdomain, C.POINTER(C.c_float)), (2 * nbins,)).copy().view(np.complex64)
    rb = Nd
    resp = np.ctypeslib.as_array(C.cast(fout.response, C.POINTER(C.c_float)), (2 * rb,)).copy().view(np.complex64)
    ng = fout.noise_gain
    assert fin.blocknum == nb and fout.blocknum == nb and fout.olen == olen and fin.ilen == L
    assert lib.delete_filter_output(s) == 0 and lib.delete_filter_input(m) == 0
    return np.array(outs), fd, resp, ng


class _FilterIn(C.Structure):  # struct filter_in (include/ka9q_b200.h; reference filter.h:54-66)
    _fields_ = [("in_type", C.c_int), ("ilen", C.c_uint), ("impulse_length", C.c_uint), ("fdomain", C.c_void_p),
                ("input_buffer", C.c_void_p), ("input", C.c_void_p), ("fwd_plan", C.c_void_p), ("blocknum", C.c_uint),
                ("filter_mutex", C.c_byte * 40), ("filter_cond", C.c_byte * 48)]


class _FilterOut(C.Structure):  # struct filter_out (reference filter.h:67-80)
    _fields_ = [("master", C.c_void_p), ("out_type", C.c_int), ("response", C.c_void_p), ("response_mutex", C.c_byte * 40),
                ("f_fdomain", C.c_void_p), ("noise_gain", C.c_float), ("output_buffer", C.c_void_p), ("output", C.c_void_p),
                ("rev_plan", C.c_void_p), ("decimate", C.c_uint), ("olen", C.c_uint), ("blocknum", C.c_uint)]


@pytest.mark.parametrize("in_type,out_type", [(1, 1), (1, 2), (1, 3), (3, 3), (3, 1)])
def test_dropin_filter_api_matches_reference(ref, in_type, out_type):
    """create/execute/set_filter/delete with the reference's names, struct fields and every in/out type combination
    (filter.c:206-249), decimation 4; host mirrors (fdomain, output, response, noise_gain) are what callers read."""
    rng = np.random.default_rng(in_type * 10 + out_type)
    L, M, D = 960, 1089, 4
    nb = 3
    x = rng.standard_normal(nb * L).astype(np.float32)
    if in_type == 1:
        x = (x + 1j * rng.standard_normal(nb * L)).astype(np.complex64)
    r = ref.filter_run(L, M, D, in_type, out_type, x, low=-0.2, high=0.3, beta=3.0, want_fdomain=True)
    out, fd, resp, ng = _dropin_run(L, M, D, in_type, out_type, x, -0.2, 0.3, 3.0)
    assert rel_rms(resp, r["response"]) < 2e-6
    assert ng == pytest.approx(r["noise_gain"], rel=1e-4)
    assert rel_rms(fd, r["fdomain"]) < 1e-6
    assert rel_rms(out, r["out"]) < FILT_TOL


def test_dropin_fm_audio_filter_path(ref):
    """The REAL->REAL de-emphasis filter of demod_fm (fm.c:39-66): window_rfilter + create/execute with decimate 1."""
    lib = _lib.lib()
    AL, AM = 960, 1089
    AN = AL + AM - 1
    ar = np.zeros(AN // 2 + 1, dtype=np.complex64)
    for j in range(AN // 2 + 1):
        f = np.float32(j) * np.float32(48000) / np.float32(AN)
        if 300 <= f <= 6000:
            ar[j] = np.float32(10.0 / AN * 300.0 / f)
    want = ref.window_rfilter(AL, AM, ar, 3.0)
    got = ar.copy()
    assert lib.window_rfilter(AL, AM, got.ctypes.data_as(C.c_void_p), 3.0) == 0
    assert rel_rms(got, want) < 2e-6
    w = ref.window_filter(AL, AM, np.where(np.arange(AN) < 200, 1 / AN, 0).astype(np.complex64), 3.0)
    g2 = np.where(np.arange(AN) < 200, 1 / AN, 0).astype(np.complex64)
    assert lib.window_filter(AL, AM, g2.ctypes.data_as(C.c_void_p), 3.0) == 0
    assert rel_rms(g2, w) < 2e-6


# ------------------------------------------------------------------------------------------ K5: half-band decimators

def test_halfband_decimators_match_golden_and_reference(ref):
    lib = _lib.lib()
    g = np.load(os.path.join(GOLDEN, "osc_hb.npz"))
    x = g["hb_x"].copy()
    st = _lib.Hb15State()
    coeffs = np.array([-6, 33, -116, 490], dtype=np.float32) / np.float32(802)  # hackrf.c:229-237
    for i in range(4):
        st.coeffs[i] = coeffs[i]
    y = np.zeros(2048, dtype=np.float32)
    a, b = x[:2048].copy(), x[2048:].copy()
    lib.hb15_block(C.byref(st), y.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), 1024)
    lib.hb15_block(C.byref(st), y[1024:].ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 1024)
    np.testing.assert_allclose(y, g["hb15_y"], rtol=0, atol=2e-6)   # state carried across calls, drop-in struct
    np.testing.assert_allclose(np.ctypeslib.as_array(st.old_odd_samples), g["hb15_state"][12:16], rtol=0, atol=0)
    s3 = np.zeros(1, dtype=np.float32)
    z = np.zeros(2048, dtype=np.float32)
    xin = x.copy()
    lib.hb3_block(s3.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), xin.ctypes.data_as(C.c_void_p), 2048)
    np.testing.assert_allclose(z, g["hb3_y"], rtol=0, atol=2e-6)
    # /64 cascade as hackrf.c:297-318 (stage 5 first ... stage 0 last), int16 after rounding (hackrf.c:309)
    rng = np.random.default_rng(2)
    n = 64 * 700
    sig = (0.3 * np.sin(2 * np.pi * 0.0007 * np.arange(n)) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    states = (_lib.Hb15State * 6)()
    rstates = [np.zeros(16, dtype=np.float32) for _ in range(6)]
    for j in range(6):
        for i in range(4):
            states[j].coeffs[i] = coeffs[i]
        rstates[j][:4] = coeffs
    out = np.zeros(n // 64, dtype=np.float32)
    assert lib.ka9q_hb15_cascade(0, 6, states, sig.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)) == 0
    buf = sig.copy()
    for j in range(5, -1, -1):
        buf = ref.hb15(rstates[j], buf)
    atten = np.float32(0.5) ** 6
    got16 = np.round(32767 * out * atten).astype(np.int32)
    want16 = np.round(32767 * buf * atten).astype(np.int32)
    assert np.abs(got16 - want16).max() <= 1
    # second chunk: the six carried states; and the fused one-launch cascade against the stage-by-stage form, bit for bit
    sig2 = (0.3 * np.sin(2 * np.pi * 0.0007 * (n + np.arange(n))) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    out2 = np.zeros(n // 64, dtype=np.float32)
    assert lib.ka9q_hb15_cascade(0, 6, states, sig2.ctypes.data_as(C.c_void_p), n, out2.ctypes.data_as(C.c_void_p)) == 0
    buf = sig2.copy()
    for j in range(5, -1, -1):
        buf = ref.hb15(rstates[j], buf)
    assert np.abs(np.round(32767 * out2 * atten) - np.round(32767 * buf * atten)).max() <= 1
    os.environ["KA9Q_B200_HB15_FUSED"] = "0"
    try:
        st2 = (_lib.Hb15State * 6)()
        for j in range(6):
            for i in range(4):
                st2[j].coeffs[i] = coeffs[i]
        o1, o2 = np.zeros(n // 64, dtype=np.float32), np.zeros(n // 64, dtype=np.float32)
        assert lib.ka9q_hb15_cascade(0, 6, st2, sig.ctypes.data_as(C.c_void_p), n, o1.ctypes.data_as(C.c_void_p)) == 0
        assert lib.ka9q_hb15_cascade(0, 6, st2, sig2.ctypes.data_as(C.c_void_p), n, o2.ctypes.data_as(C.c_void_p)) == 0
    finally:
        del os.environ["KA9Q_B200_HB15_FUSED"]
    assert np.array_equal(o1, out) and np.array_equal(o2, out2)
    assert bytes(st2) == bytes(states)
