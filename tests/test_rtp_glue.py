"""Wire-format glue (host C in libka9q_b200.so) against the reference's own code: ntoh_rtp / rtp_process
(multicast.c) behind a restatement of their callers, and the real send_mono_output / send_stereo_output (audio.c)
writing into a socket pair. No GPU."""
import struct

import numpy as np
import pytest

from ka9q_sdr_b200 import rtp


def rtp_header(pt, seq, ts, ssrc, marker=0, cc=0, ext_words=None, pad=0):
    b0 = (2 << 6) | (pad << 5) | ((1 if ext_words is not None else 0) << 4) | cc
    h = struct.pack(">BBHII", b0, (marker << 7) | pt, seq & 0xFFFF, ts & 0xFFFFFFFF, ssrc)
    h += b"".join(struct.pack(">I", 0x1000 + i) for i in range(cc))
    if ext_words is not None:
        # the reference skips 4 + get16(length) bytes after the 4-byte extension header (multicast.c:271-274)
        h += struct.pack(">HH", 0xBEDE, ext_words) + bytes(4 + ext_words)
    return h


def iq_datagram(rng, pt, seq, ts, ssrc, nsamp, **kw):
    pad = kw.pop("padbytes", 0)
    if pt == rtp.IQ_PT:
        payload = rng.integers(-32768, 32767, 2 * nsamp, dtype=np.int16).tobytes()
    else:
        payload = rng.integers(-128, 127, 2 * nsamp, dtype=np.int8).tobytes()
    d = rtp_header(pt, seq, ts, ssrc, pad=1 if pad else 0, **kw) + bytes(range(24)) + payload
    if pad:
        d += bytes(pad - 1) + bytes([pad])
    return d


def make_stream(rng, pt, npk=400):
    """A packet sequence with losses, duplicates, old packets, an SSRC restart, CSRCs, extensions, padding, foreign
    payload types, runts and one absurd timestamp jump."""
    out, seq, ts, ssrc = [], int(rng.integers(0, 65536)), int(rng.integers(0, 2**32)), 0x12345678
    last = None
    for i in range(npk):
        n = int(rng.choice([240, 350, 1024, 1]))
        ev = rng.random()
        kw = {}
        if ev < 0.06:          # lost packets: sequence and timestamp both jump
            k = int(rng.integers(1, 4))
            seq += k
            ts += k * n
        elif ev < 0.10 and last is not None:
            out.append(last)   # duplicate
            continue
        elif ev < 0.13:
            out.append(iq_datagram(rng, pt, seq - 5, ts - 5 * n, ssrc, n))   # old packet
            continue
        elif ev < 0.15:
            ssrc += 1          # sender restart: new SSRC, unrelated numbering
            seq, ts = int(rng.integers(0, 65536)), int(rng.integers(0, 2**32))
        elif ev < 0.18:
            kw["cc"] = int(rng.integers(1, 4))
        elif ev < 0.21:
            kw["ext_words"] = int(rng.integers(0, 3)) * 4
        elif ev < 0.24:
            kw["padbytes"] = int(rng.integers(1, 5)) * (4 if pt == rtp.IQ_PT else 2)
        elif ev < 0.26:
            out.append(rtp_header(rtp.PCM_MONO_PT, seq, ts, ssrc) + bytes(200))   # not I/Q
            continue
        elif ev < 0.28:
            out.append(bytes(int(rng.integers(0, 12))))                             # runt
            continue
        elif ev < 0.29:
            ts += 500000       # a jump of more than 192000 samples is dropped, not filled
        elif ev < 0.31:
            ts += int(rng.integers(1, 3000))   # timestamp gap without sequence gap (sender-side silence)
        d = iq_datagram(rng, pt, seq, ts, ssrc, n, **kw)
        out.append(d)
        last = d
        seq += 1
        ts += n
    return out


@pytest.mark.parametrize("pt", [rtp.IQ_PT, rtp.IQ_PT8])
def test_ingest_matches_reference_rtp_handling(ref, pt):
    rng = np.random.default_rng(7 + pt)
    stream = make_stream(rng, pt)
    ours = rtp.Ingest(rtp.IQ_S16 if pt == rtp.IQ_PT else rtp.IQ_S8)
    theirs = ref.GlueIngest()
    appended = ignored = filled = 0
    for d in stream:
        n, x = ours.datagram(d)
        rn, rraw = theirs.datagram(d)
        assert n == rn
        if n >= 0:
            assert x.tobytes() == rraw
            appended += n
        else:
            ignored += 1
        for f in ("ssrc", "init", "seq", "timestamp", "packets", "drops", "dupes"):
            assert getattr(ours.st.rtp, f) == getattr(theirs.state, f), f
        assert ours.st.samples == theirs.samples.value
    filled = ours.st.zero_filled
    assert appended > 50000 and ignored > 10 and filled > 1000 and ours.st.rtp.dupes > 0 and ours.st.rtp.drops > 0
    assert ours.st.ignored == ignored


def test_ingest_edge_cases():
    g = rtp.Ingest(rtp.IQ_S16, room=1000)
    rng = np.random.default_rng(1)
    assert g.datagram(b"")[0] == -1 and g.datagram(bytes(11))[0] == -1
    d0 = iq_datagram(rng, rtp.IQ_PT, 10, 1000, 99, 100)
    n, x = g.datagram(d0)
    assert n == 100 and x.tobytes() == d0[36:]
    # a datagram of the stream's other sample format is ignored
    assert g.datagram(iq_datagram(rng, rtp.IQ_PT8, 11, 1100, 99, 100))[0] == -1
    # no room: -2 and the RTP state is untouched, so the same datagram can be retried
    before = bytes(g.st.rtp)
    big = iq_datagram(rng, rtp.IQ_PT, 11, 1100 + 950, 99, 100)     # 950 lost + 100 > room
    assert g.datagram(big)[0] == -2 and bytes(g.st.rtp) == before
    g.room = 2000
    g.buf = np.zeros(4000, dtype=np.int16)
    n, x = g.datagram(big)
    assert n == 1050 and not x[:1900].any() and x[1900:].tobytes() == big[36:]
    # header only (no samples) is accepted with zero samples, like the reference (sampcount = 0)
    assert g.datagram(rtp_header(rtp.IQ_PT, 12, 2150, 99) + bytes(24))[0] == 0
    # extension bit set but the datagram ends before the extension header does
    h = rtp_header(rtp.IQ_PT, 13, 2150, 99)
    assert g.datagram(bytes([h[0] | 0x10]) + h[1:])[0] == -1


def scaleclip(x):
    """audio.c:22-28"""
    x = np.asarray(x, dtype=np.float32)
    v = (np.float32(32767.0) * x).astype(np.float32)
    out = np.trunc(np.clip(v, -40000, 40000)).astype(np.int32)
    out = np.where(x >= 1.0, 32767, np.where(x <= -1.0, -32768, out))
    return out.astype(np.int16)


@pytest.mark.parametrize("channels", [1, 2])
def test_pcm_packetiser_matches_reference_send_output(ref, channels):
    rng = np.random.default_rng(11 + channels)
    state = dict(ssrc=0xCAFEF00D, timestamp=0xFFFFFF00, seq=0xFFFE, silent=0)      # both counters wrap in the run
    ours = rtp.PcmOut(state["ssrc"], state["timestamp"], state["seq"])
    total = 0
    for blk in range(60):
        frames = int(rng.choice([960, 960, 480, 1, 241, 1500]))
        x = (0.5 * rng.standard_normal(frames * channels)).astype(np.float32)
        kind = rng.random()
        if kind < 0.25:
            x[:] = 0                                  # squelched block: nothing is sent, timestamps advance
        elif kind < 0.45:
            lo = int(rng.integers(0, frames)) * channels
            x[lo:] = 0                                # silence starts inside the block
        elif kind < 0.55:
            x *= 4                                    # clipping
        elif kind < 0.60:
            x[:] = 1e-6                               # quantises to zero: silent after scaleclip
        want = ref.glue_send(channels, state, x)
        got = ours.packetise(scaleclip(x), channels)
        assert got == want, f"block {blk}"
        total += len(got)
        assert ours.st.rtp.timestamp == state["timestamp"] and ours.st.rtp.seq == state["seq"]
        assert ours.st.silent == state["silent"]
    assert total > 40 and ours.st.rtp.packets == state["packets"]
    # header spot check on the last packet that went out: version 2, payload type by channel count
    assert got == [] or (got[-1][0] == 0x80 and (got[-1][1] & 0x7F) == (rtp.PCM_STEREO_PT if channels == 2 else rtp.PCM_MONO_PT))


def test_packetiser_rejects_bad_arguments():
    o = rtp.PcmOut(1)
    with pytest.raises(ValueError):
        o.packetise(np.zeros(10, dtype=np.int16), 3)


def test_status_tlv_matches_reference_encoders(ref):
    """ka9q_status_encode_signals against status.c's encode_float / encode_byte / encode_int32 / encode_eol called in
    radio_status.c's order (leading-zero suppression included: 0.0 encodes with length 0, small ints with one byte)."""
    import ctypes as C
    from ka9q_sdr_b200 import _lib
    L = _lib.lib()
    L.ka9q_status_encode_signals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int]
    rng = np.random.default_rng(5)
    buf = (C.c_ubyte * 128)()
    for demod_type in (0, 1, 2):
        for trial in range(20):
            st = _lib.ChanStatus()
            vals = rng.standard_normal(6).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 6)
            if trial == 0:
                vals[:] = 0                      # every float suppressed to zero length
            if trial == 1:
                vals[:] = np.float32(np.nan)
            st.bb_power, st.snr, st.foffset, st.pdeviation, st.agc_gain = (float(v) for v in vals[:5])
            isb, nch = int(trial % 2), 1 + int(trial % 2)
            ifp, nbw = float(vals[5]), float(np.float32(trial * 1234.5))
            n = L.ka9q_status_encode_signals(C.byref(st), demod_type, isb, ifp, nbw, nch, buf, 128)
            want = ref.glue_status(demod_type, isb, nbw, ifp, st.bb_power, st.agc_gain, st.pdeviation, st.foffset, st.snr, nch)
            assert n == len(want) and bytes(buf[:n]) == want, (demod_type, trial)
    assert L.ka9q_status_encode_signals(C.byref(st), 2, 0, 1.0, 1.0, 1, buf, 3) == -1      # no room


def test_status_type_numbers_match_the_reference_header():
    import os
    import re
    hdr = "/root/reference/status.h"
    if not os.path.exists(hdr):
        pytest.skip("reference tree not present")
    txt = open(hdr).read()
    body = re.sub(r"//.*", "", txt[txt.index("enum status_type {"):txt.index("};", txt.index("enum status_type {"))])
    names = [n.strip().split("=")[0].strip() for n in body.split("{")[1].split(",") if n.strip()]
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ka9q_sdr_b200", "csrc", "rtp_glue.c")).read()
    for name in ("NOISE_BANDWIDTH", "IF_POWER", "BASEBAND_POWER", "DEMOD_MODE", "INDEPENDENT_SIDEBAND", "DEMOD_SNR",
                 "DEMOD_GAIN", "FREQ_OFFSET", "PEAK_DEVIATION", "OUTPUT_CHANNELS", "EOL"):
        m = re.search(rf"ST_{name} = (\d+)", src)
        assert m and int(m.group(1)) == names.index(name), name
