"""CPU tests of the oracle itself (no GPU): the FFT stand-in, the verbatim reference build (oracle/_ref) against the
committed golden vectors, and the numpy port (oracle/port.py) against both. The golden vectors were produced by running
the reference's own C (scripts/make_golden.py); the reference ships no fixtures of its own."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from ka9q_sdr_b200 import modes, synth
from oracle import port

sys.path.insert(0, os.path.join(ROOT, "scripts"))
import make_golden  # noqa: E402  (stimulus() is shared with the generator)

CASES = [("fm", "FM"), ("fm", "FMF"), ("am", "AM"), ("usb", "USB"), ("usb", "LSB"), ("usb", "IQ"), ("usb", "ISB"),
         ("usb", "CWU")]


def _golden(mode):
    return np.load(os.path.join(GOLDEN, f"chain_{mode.lower()}.npz"))


def _pcm_channels(mode):
    m = modes.get_mode(mode)
    return m.channels if m.demod_type == modes.LINEAR_DEMOD else 1


# ------------------------------------------------------------------------------------------ FFT shim

@pytest.mark.parametrize("n", [8, 60, 2048, 8192, 81920])
def test_standin_fft_matches_numpy(ref, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for sign in (-1, +1):
        y = ref.raw_fft(x, sign)
        want = np.fft.fft(x.astype(np.complex128)) if sign < 0 else np.fft.ifft(x.astype(np.complex128)) * n
        assert np.linalg.norm(y - want) / np.linalg.norm(want) < 1e-7
    xr = rng.standard_normal(n).astype(np.float32)
    X = ref.raw_rfft(xr)
    want = np.fft.rfft(xr.astype(np.float64))
    assert np.linalg.norm(X - want) / np.linalg.norm(want) < 1e-7
    back = ref.raw_irfft(want.astype(np.complex64), n)
    assert np.linalg.norm(back - xr * n) / np.linalg.norm(xr * n) < 1e-6


def test_mkl_backend_agrees_with_standin_when_available(ref):
    if not ref.set_fft_backend("mkl"):
        pytest.skip("MKL DFTI not loadable")
    try:
        assert "mkl" in ref.fft_backend()
        rng = np.random.default_rng(5)
        for n in (2048, 8192):
            x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
            a = ref.raw_fft(x, -1)
            want = np.fft.fft(x.astype(np.complex128))
            assert np.linalg.norm(a - want) / np.linalg.norm(want) < 1e-6
    finally:
        ref.set_fft_backend("standin")


# ------------------------------------------------------------------------------------------ mode table

def test_mode_table_matches_readmodes_on_the_reference_file(ref):
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present")
    assert ref.load_modes_file("/root/reference") == 0
    theirs = ref.get_modes()
    ref.load_modes(modes.MODES.values())
    assert ref.get_modes() == theirs


def test_mode_line_parser_rules():
    m = modes.parse_mode_line("XYZ LINEAR +3000 -100 700 50 -6 -1.1 conj mono # comment")
    assert (m.low, m.high) == (-100.0, 3000.0) and m.attack == -50 and m.recovery == 6 and m.hang == pytest.approx(1.1)
    assert m.isb and m.channels == 1 and not m.pll
    assert modes.parse_mode_line("# only a comment") is None
    assert modes.parse_mode_line("FOO BOGUS 1 2") is None
    assert modes.parse_mode_line("S LINEAR -1 1 0 0 0 0 square").pll
    assert modes.get_mode("usb").name == "USB"
    with pytest.raises(KeyError):
        modes.get_mode("nope")


# ------------------------------------------------------------------------------------------ reference == golden

@pytest.mark.parametrize("stim,mode", CASES)
def test_reference_reproduces_golden_exactly(ref, stim, mode):
    """The verbatim reference is deterministic: this pins the _ref build itself (compiler flags, shim)."""
    g = _golden(mode)
    c = make_golden.stimulus(stim)
    assert int(c["iq"].astype(np.int64).sum()) == int(g["iq_crc"][0]), "stimulus generator changed"
    fs, L, M, D, N, k, nb = (int(v) for v in g["meta"])
    r = ref.chain_run(mode, fs, L, M, D, c["iq"], carrier_hz=k * fs / N, lo_cycles=-k / N,
                      want_filt=mode in ("FM", "FMF", "AM"))
    assert r.nblocks == nb
    d = np.abs(r.pcm.astype(np.int32) - g["pcm"].astype(np.int32))
    assert d.max() <= 1 and (d == 0).mean() > 0.999  # libm / FMA differences between hosts can flip a truncation
    if g["filt"].size:
        assert np.linalg.norm(r.filt - g["filt"]) / np.linalg.norm(g["filt"]) < 1e-6


# ------------------------------------------------------------------------------------------ port == golden

@pytest.mark.parametrize("stim,mode", CASES)
def test_port_matches_golden(stim, mode):
    g = _golden(mode)
    c = make_golden.stimulus(stim)
    fs, L, M, D, N, k, nb = (int(v) for v in g["meta"])
    p = port.run_channel(mode, fs, L, M, D, c["iq"], k)
    ch = _pcm_channels(mode)
    olen = L // D
    d = np.abs(p["pcm"].astype(np.int32) - g["pcm"].astype(np.int32))
    skip = 0 if modes.get_mode(mode).demod_type == modes.FM_DEMOD else olen * ch  # AGC start-up, SURVEY Appendix D-5
    assert d[skip:].max() <= 1, f"{mode}: {d[skip:].max()} LSB"
    assert (d[skip:] == 0).mean() > 0.97
    if g["filt"].size:
        assert np.linalg.norm(p["filt"] - g["filt"]) / np.linalg.norm(g["filt"]) < 1e-5
    if mode == "FM":
        st = p["status"]
        np.testing.assert_allclose([s["bb_power"] for s in st], g["bb_power"], rtol=1e-4)
        np.testing.assert_allclose([s["snr"] for s in st][1:], g["snr"][1:], rtol=2e-2)
        np.testing.assert_allclose([s["pdeviation"] for s in st][2:], g["pdeviation"][2:], rtol=1e-3)


def test_port_filter_design_matches_golden():
    g = np.load(os.path.join(GOLDEN, "design.npz"))
    D, L, M, N = synth.geometry(192000)
    F = np.float32
    st = F(F(4) / F(192000))
    cases = {"fm": (port.COMPLEX, F(-8000) / F(48000), F(8000) / F(48000)),
             "usb": (port.COMPLEX, st * F(100), st * F(3000)),   # the 3000 Hz edge lands exactly on a bin (Appendix A)
             "isb": (port.CROSS_CONJ, F(-5000) / F(48000), F(5000) / F(48000))}
    for name, (ot, lo, hi) in cases.items():
        resp, ng = port.set_filter_response(L, M, D, ot, float(lo), float(hi), 3.0)
        assert np.linalg.norm(resp - g[name]) / np.linalg.norm(g[name]) < 2e-6, name
        assert ng == pytest.approx(float(g[name + "_ng"][0]), rel=1e-5)
    np.testing.assert_allclose(port.make_kaiser(1089, 3.0), g["kaiser_1089_3"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(port.make_kaiser(64, 2.0), g["kaiser_64_2"], rtol=2e-6, atol=1e-7)


def test_port_oscillator_and_halfband_match_golden():
    g = np.load(os.path.join(GOLDEN, "osc_hb.npz"))
    o = port.Osc()
    o.set(-2048 / 8192)
    z = o.run(40000)
    np.testing.assert_allclose(z[g["osc_idx"]], g["osc"], rtol=0, atol=1e-11)  # closed-form powers vs repeated multiply
    x = g["hb_x"]
    st = dict(coeffs=(np.array([-6, 33, -116, 490], dtype=np.float32) / np.float32(802)), even=np.zeros(4, np.float32),
              odd=np.zeros(4, np.float32), old_odd=np.zeros(4, np.float32))
    y = np.concatenate([port.hb15(st, x[:2048]), port.hb15(st, x[2048:])])  # state carried across calls
    np.testing.assert_allclose(y, g["hb15_y"], rtol=0, atol=2e-6)
    s3 = np.zeros(1, np.float32)
    np.testing.assert_allclose(port.hb3(s3, x), g["hb3_y"], rtol=0, atol=2e-6)


def test_port_filter_variants_match_reference(ref):
    """REAL/COMPLEX in x REAL/COMPLEX/CROSS_CONJ out, with decimation (filter.c:206-249)."""
    rng = np.random.default_rng(11)
    L, M, D = 960, 1089, 4
    N = L + M - 1
    nb = 3
    for in_type, out_type in ((port.COMPLEX, port.COMPLEX), (port.COMPLEX, port.CROSS_CONJ), (port.COMPLEX, port.REAL),
                              (port.REAL, port.REAL), (port.REAL, port.COMPLEX)):
        x = rng.standard_normal(nb * L).astype(np.float32)
        if in_type == port.COMPLEX:
            x = (x + 1j * rng.standard_normal(nb * L)).astype(np.complex64)
        r = ref.filter_run(L, M, D, in_type, out_type, x, low=-0.2, high=0.3, beta=3.0)
        m = port.FilterIn(L, M, in_type)
        resp = r["response"]  # all N_dec bins as designed by set_filter (filter.c:523)
        assert resp.size == N // D
        f = port.FilterOut(m, resp, D, out_type)
        out = []
        for b in range(nb):
            m.execute(x[b * L:(b + 1) * L])
            out.append(f.execute())
        out = np.array(out)
        assert np.linalg.norm(out - r["out"]) / np.linalg.norm(r["out"]) < 2e-6, (in_type, out_type)


def test_reference_zero_fills_lost_packets_and_keeps_lo_phase(ref):
    """radio.c:81-100: a lost packet becomes zeros while the LOs keep stepping. Dropping a packet must equal feeding
    explicit zeros in its place."""
    c = synth.cfg1_fm(4)
    fs, L, M, D, N = c["samprate"], c["L"], c["M"], c["D"], c["N"]
    k = c["bins"][0]
    pkt = 960
    npk = c["iq"].size // 2 // pkt
    drop = np.zeros(npk, dtype=np.uint8)
    drop[5] = 1
    a = ref.chain_run("FM", fs, L, M, D, c["iq"], carrier_hz=k * fs / N, lo_cycles=-k / N, pkt_samples=pkt, drop=drop)
    iq2 = c["iq"].copy()
    iq2[2 * 5 * pkt:2 * 6 * pkt] = 0
    b = ref.chain_run("FM", fs, L, M, D, iq2, carrier_hz=k * fs / N, lo_cycles=-k / N, pkt_samples=pkt)
    assert np.array_equal(a.pcm, b.pcm)


def test_cpu_channelizer_restatement_matches_the_per_channel_port():
    """oracle/channelizer_port.py (shared forward FFT + bin rotation, vectorised over channels: the algorithm the GPU path
    runs, used by bench.py as the second CPU figure) against oracle/port.py (the reference's mix-then-FFT order, one
    channel at a time) on three FM channels of one stream: same PCM within 1 LSB."""
    from oracle import channelizer_port as cp
    from ka9q_sdr_b200 import synth
    fs, D, L, M, N, nb = 192000, 4, 3840, 4353, 8192, 6
    rng = np.random.default_rng(11)
    n = nb * L
    bins = [2048, -1536, 300]
    x = synth.awgn(rng, n, 0.01)
    for i, k in enumerate(bins):
        x = x + synth.fm_carrier(n, fs, k * fs / N, 700.0 + 200 * i, 2500.0, 0.2)
    iq = synth._quantize(x)
    ch = cp.FmChannelizer(fs, L, M, D, bins, -8000.0, 8000.0, workers=1)
    got = np.concatenate([ch.process(iq[2 * b * L:2 * (b + 1) * L]) for b in range(nb)], axis=1)
    for i, k in enumerate(bins):
        want = port.run_channel("FM", fs, L, M, D, iq, k)["pcm"]
        d = np.abs(got[i].astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1, (k, d.max())
        assert np.abs(want).max() > 1000


@pytest.mark.parametrize("eps", [0.37, -0.5, 0.0])
def test_off_grid_lo_identity_against_the_reference(ref, eps):
    """DESIGN.md section 1.2, pinned on the CPU: the reference mixing with an arbitrary-double second LO ahead of its own
    FFT (radio.c:217,299) equals — to its own fp32 rounding — the shared-FFT form: bin rotation by the nearest grid bin k,
    the channel's impulse response times exp(+j 2 pi eps m / 2048), the kept samples times exp(-j 2 pi eps n' / 2048).
    float64 numpy on one side, the verbatim reference (oracle/_ref) on the other."""
    fs = 192000
    D, L, M, N = synth.geometry(fs)
    nb, k = 6, 1024
    n = nb * L
    rng = np.random.default_rng(5)
    fc = (k + eps) * fs / N
    x = (synth.fm_carrier(n, fs, fc, 1000.0, 3000.0, 0.2) + synth.fm_carrier(n, fs, fc + 9000, 700.0, 2500.0, 0.2)
         + synth.awgn(rng, n, 0.01))
    iq = synth._quantize(x)
    want = ref.chain_run("FM", fs, L, M, D, iq, carrier_hz=fc, want_filt=True).filt[:nb]
    xs = (iq[0::2].astype(np.float64) + 1j * iq[1::2]) / 32767.0
    m = modes.get_mode("FM")
    nd, md, olen = N // D, (M - 1) // D + 1, L // D
    f = np.fft.fftfreq(nd)
    H = ((f >= m.low / (fs / D)) & (f <= m.high / (fs / D))).astype(np.complex128) / N      # filter.c:518-535
    h = np.fft.ifft(H) * nd                                                                 # unnormalised backward
    w = np.asarray(port.make_kaiser(md, 3.0), dtype=np.float64)
    hh = np.zeros(nd, complex)
    idx = np.arange(md)
    hh[:md] = h[(idx - md // 2) % nd] * w / nd                                              # filter.c:386-392
    hh[:md] *= np.exp(2j * np.pi * eps * idx / nd)                                          # the phase ramp
    R = np.fft.fft(hh)
    buf = np.concatenate([np.zeros(M - 1, complex), xs])
    s = np.arange(nd)
    s = np.where(s <= nd // 2, s, s - nd)
    for b in range(nb):
        X = np.fft.fft(buf[b * L:b * L + N])
        y = np.fft.ifft(X[(k + s) % N] * R) * nd
        y = y * np.exp(2j * np.pi * ((-k * (b * L - (M - 1))) % N) / N)
        yk = y[nd - olen:] * np.exp(-2j * np.pi * eps * (b * olen + np.arange(olen)) / nd)  # rotation at the output rate
        err = np.linalg.norm(yk - want[b]) / np.linalg.norm(want[b])
        assert err < 5e-7, (eps, b, err)
