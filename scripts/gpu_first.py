"""First GPU bring-up: FFT correctness + cfg1 FM/AM/USB parity vs the verbatim reference .so."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ka9q_sdr_b200 import channelizer as ch, synth, modes
from oracle import refbind as R

print("devices", ch._lib.lib().ka9q_device_count(), ch._lib.lib().ka9q_version())
rng = np.random.default_rng(0)
for n in (64, 2048, 8192, 81920, 819200, 2621440):
    print("plan", n, ch.fft_plan(n))
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for sign in (-1, 1):
        y = ch.fft_c2c(x, sign)
        ref = np.fft.fft(x.astype(np.complex128)) if sign < 0 else np.fft.ifft(x.astype(np.complex128)) * n
        print("  fft", n, sign, "relerr %.3e" % (np.linalg.norm(y - ref) / np.linalg.norm(ref)))

R.load_modes(modes.MODES.values())

def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

def run_case(mode, cfg, nblocks, **kw):
    fs, L, M, D, N = cfg["samprate"], cfg["L"], cfg["M"], cfg["D"], cfg["N"]
    k = cfg["bins"][0]
    c = ch.Channelizer(fs, L, M, D, max_blocks=4, capture_filter_output=True)
    c.add_channel(mode, k, **kw)
    c.commit()
    iq = cfg["iq"]
    pcm_all = []; filt_all = []; st_all = []
    for b in range(0, nblocks, 4):
        nb = min(4, nblocks - b)
        pcm, st = c.process(iq[2 * b * L:2 * (b + nb) * L])
        pcm_all.append(c.channel_pcm(pcm, 0).copy()); st_all.append(st)
        filt_all.append(c.filter_output(0, nb))
    pcm = np.concatenate(pcm_all); filt = np.concatenate(filt_all); st = np.concatenate(st_all)
    ref = R.chain_run(mode, fs, L, M, D, iq, carrier_hz=k * fs / N, lo_cycles=-k / N, want_filt=True, **kw)
    H, ng = c.response(0)
    out = {"mode": mode}
    if mode in ("FM", "AM"):
        out["filt_rel_rms"] = rel(filt[:nblocks], ref.filt[:nblocks])
    d = pcm.astype(np.int32) - ref.pcm[:pcm.size].astype(np.int32)
    out["pcm_maxdiff"] = int(np.abs(d).max()); out["pcm_maxdiff_after_blk1"] = int(np.abs(d[960 * c.channels[0]:]).max())
    out["pcm_frac_exact"] = float((d == 0).mean()); out["pcm_rms"] = float(ref.pcm.astype(float).std())
    out["n_gt1"] = int((np.abs(d) > 1).sum())
    print(out)
    print("   status gpu", st[min(5, nblocks - 1), 0], "\n   status ref", ref.status[min(5, nblocks - 1)])
    c.close()
    return out

nb = 24
run_case("FM", synth.cfg1_fm(nb), nb)
D, L, M, N = synth.geometry(192000)
n = nb * L; rg = np.random.default_rng(3)
x = synth.am_carrier(n, 192000, 1024 * 192000 / N, 1000, 0.5, 0.1) + synth.awgn(rg, n, 0.01)
run_case("AM", dict(samprate=192000, D=D, L=L, M=M, N=N, iq=synth._quantize(x), bins=[1024]), nb)
cu = synth.cfg2_usb(nb, samprate=192000)
run_case("USB", cu, nb)
run_case("IQ", cu, nb)
run_case("ISB", cu, nb)
run_case("CWU", cu, nb)
run_case("LSB", cu, nb)
