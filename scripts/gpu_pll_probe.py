"""Diagnostics: per-block loop state of a coherent channel, GPU vs the verbatim reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ka9q_sdr_b200 import channelizer as ch, modes, synth
from oracle import refbind as R
R.lib(); R.set_fft_backend("standin"); R.load_modes(modes.MODES.values())
fs = 192000
D, L, M, N = synth.geometry(fs)
nb = 120
rng = np.random.default_rng(5)
n = nb * L
k, off = 1024, 37.3
t = np.arange(n) / fs
x = 0.1 * (1 + 0.5 * np.sin(2 * np.pi * 1000 * t)) * np.exp(2j * np.pi * (k * fs / N + off) * t) + synth.awgn(rng, n, 0.005)
iq = synth._quantize(x)
for mode in sys.argv[1:] or ["CAM"]:
    c = ch.Channelizer(fs, L, M, D, max_blocks=8)
    c.add_channel(mode, k)
    c.commit()
    pcm, st = c.run(iq)
    r = R.chain_run(mode, fs, L, M, D, iq, carrier_hz=k * fs / N, lo_cycles=-k / N)
    rs = r.status[:nb]
    got = c.channel_pcm(pcm, 0).astype(np.int32); want = r.pcm.astype(np.int32)
    per = got.size // nb
    d = np.abs(got - want[:got.size]).reshape(nb, per)
    print("==", mode)
    for b in list(range(0, 6)) + list(range(30, 44)) + list(range(100, 120, 3)):
        print(b, "lock", st["squelch_open"][b, 0], rs["pll_lock"][b], "cphase %.5f %.5f" % (st["reserved"][b, 0, 0], rs["cphase"][b]),
              "snr %.3f %.3f" % (st["snr"][b, 0], rs["snr"][b]), "foff %.5f %.5f" % (st["foffset"][b, 0], rs["foffset"][b]),
              "gain %.4f %.4f" % (st["agc_gain"][b, 0], rs["agc_gain"][b]), "pcm maxdiff", d[b].max(), "frac>1 %.3f" % (d[b] > 1).mean())
    c.close()
