"""Generate tests/golden/*.npz by running the VERBATIM reference (oracle/_ref) on seeded synthetic I/Q.

Run in the build container (where /root/reference exists and `make -C oracle ref` has been run):
    python scripts/make_golden.py
The fixtures pin (a) the oracle port (oracle/port.py) and (b) the CUDA path to the reference's own output; the reference
ships no golden vectors of its own. Inputs are regenerated from the seed at test time (synth.py is
deterministic), so only the reference OUTPUTS are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ka9q_sdr_b200 import modes, synth  # noqa: E402
from oracle import refbind as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NB = 8  # blocks per case


def stimulus(case: str):
    """Seeded stimuli shared by make_golden.py and the tests."""
    if case == "fm":
        return synth.cfg1_fm(NB)
    fs = 192000
    D, L, M, N = synth.geometry(fs)
    n = NB * L
    if case == "am":
        rng = np.random.default_rng(3)
        x = synth.am_carrier(n, fs, 1024 * fs / N, 1000.0, 0.5, 0.1) + synth.awgn(rng, n, 0.01)
        return dict(samprate=fs, D=D, L=L, M=M, N=N, iq=synth._quantize(x), bins=[1024])
    return synth.cfg2_usb(NB, samprate=fs)  # two-tone SSB stimulus for every linear mode


CASES = [("fm", "FM"), ("fm", "FMF"), ("am", "AM"), ("usb", "USB"), ("usb", "LSB"), ("usb", "IQ"), ("usb", "ISB"),
         ("usb", "CWU")]


def main():
    R.lib()
    R.set_fft_backend("standin")
    R.load_modes(modes.MODES.values())
    os.makedirs(OUT, exist_ok=True)
    for stim, mode in CASES:
        c = stimulus(stim)
        fs, L, M, D, N = c["samprate"], c["L"], c["M"], c["D"], c["N"]
        k = c["bins"][0]
        r = R.chain_run(mode, fs, L, M, D, c["iq"], carrier_hz=k * fs / N, lo_cycles=-k / N,
                        want_filt=mode in ("FM", "FMF", "AM"))
        st = r.status
        np.savez_compressed(
            os.path.join(OUT, f"chain_{mode.lower()}.npz"), pcm=r.pcm,
            filt=(r.filt.astype(np.complex64) if r.filt is not None else np.zeros(0, np.complex64)),
            bb_power=st["bb_power"], snr=st["snr"], foffset=st["foffset"], pdeviation=st["pdeviation"],
            agc_gain=st["agc_gain"], meta=np.array([fs, L, M, D, N, k, NB]), iq_crc=np.array([int(c["iq"].astype(np.int64).sum())]))
        print(mode, r.pcm.shape, "pcm rms %.1f" % r.pcm.astype(float).std())
    # filter design fixtures: set_filter responses + noise gains (filter.c:500-546)
    D, L, M, N = synth.geometry(192000)
    resp = {}
    for name, (ot, lo, hi) in {"fm": (R.COMPLEX, -8000 / 48000, 8000 / 48000),
                               "usb": (R.COMPLEX, np.float32(np.float32(4) / np.float32(192000)) * np.float32(100),
                                       np.float32(np.float32(4) / np.float32(192000)) * np.float32(3000)),
                               "isb": (R.CROSS_CONJ, -5000 / 48000, 5000 / 48000)}.items():
        x = np.zeros(L, dtype=np.complex64)
        o = R.filter_run(L, M, D, R.COMPLEX, ot, x, low=float(lo), high=float(hi), beta=3.0)
        resp[name] = o["response"]
        resp[name + "_ng"] = np.array([o["noise_gain"]], dtype=np.float32)
    resp["kaiser_1089_3"] = R.make_kaiser(1089, 3.0)
    resp["kaiser_64_2"] = R.make_kaiser(64, 2.0)
    np.savez_compressed(os.path.join(OUT, "design.npz"), **resp)
    # oscillator + half-band fixtures (osc.c, decimate.c)
    osc = R.osc_run(-2048 / 8192, 0.0, 40000)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(4096).astype(np.float32)
    st = np.zeros(16, dtype=np.float32)
    st[:4] = np.array([-6, 33, -116, 490], dtype=np.float32) / np.float32(802)  # hackrf.c:229-237
    y1 = R.hb15(st, x[:2048])
    y2 = R.hb15(st, x[2048:])
    s3 = np.zeros(1, dtype=np.float32)
    z = R.hb3(s3, x)
    np.savez_compressed(os.path.join(OUT, "osc_hb.npz"), osc_idx=np.array([0, 1, 2, 16383, 16384, 16385, 39999]),
                        osc=osc[[0, 1, 2, 16383, 16384, 16385, 39999]], hb_x=x, hb15_y=np.concatenate([y1, y2]),
                        hb15_state=st, hb3_y=z)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
