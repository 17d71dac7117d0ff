"""C-ABI library: loads without a GPU, exports every symbol include/ka9q_b200.h declares, keeps the reference's struct
layouts, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT
from ka9q_sdr_b200 import _lib, channelizer as ch

HEADER = os.path.join(ROOT, "include", "ka9q_b200.h")
REF = "/root/reference"


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"typedef\s+[^;{]*\(\s*\*[^;]*;", "", txt)   # function-pointer typedefs declare no symbol
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", txt))
    names |= set(re.findall(r"extern\s+float\s+([A-Za-z_][A-Za-z0-9_]*)\s*;", txt))
    return {n for n in names if n not in ("defined",)}


def test_library_loads_and_exports_everything_declared():
    L = _lib.lib()
    declared = declared_symbols()
    assert len(declared) > 50
    missing = []
    for name in sorted(declared):
        try:
            getattr(L, name)
        except AttributeError:
            missing.append(name)
    assert not missing, f"declared in the header but not exported: {missing}"
    # and the binding table agrees with the header
    assert set(_lib.EXPORTED) == declared, (set(_lib.EXPORTED) ^ declared)


def test_version_and_device_count():
    L = _lib.lib()
    assert b"sm_100a" in L.ka9q_version()
    assert L.ka9q_device_count() >= 0


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure path")
def test_no_cpu_fallback_without_gpu():
    L = _lib.lib()
    assert L.ka9q_device_count() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ch.Channelizer(192000, 3840, 4353, 4)
    assert not L.create_filter_input(3840, 4353, 1)
    x = np.zeros(64, dtype=np.complex64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ch.fft_c2c(x)
    r = np.zeros(2048, dtype=np.complex64)
    assert L.window_filter(960, 1089, r.ctypes.data_as(C.c_void_p), 3.0) != 0
    st = _lib.Hb15State()
    xin = np.zeros(64, dtype=np.float32)
    out = np.zeros(32, dtype=np.float32)
    assert L.ka9q_hb15_cascade(0, 1, C.byref(st), xin.ctypes.data_as(C.c_void_p), 64, out.ctypes.data_as(C.c_void_p)) != 0


def test_argument_validation_matches_reference_error_behaviour():
    L = _lib.lib()
    # NULL arguments return -1 / 0 like filter.c:147-149,176-178,254-256,501-505
    assert L.execute_filter_input(None) == -1
    assert L.execute_filter_output(None) == -1
    assert L.delete_filter_input(None) == 0
    assert L.delete_filter_output(None) == 0
    assert L.set_filter(None, 0.1, 0.2, 3.0) == -1
    assert L.make_kaiser(None, 16, 3.0) == -1
    assert L.window_filter(10, 7, None, 3.0) == -1
    assert not L.create_filter_output(None, None, 1, 1)
    assert np.isnan(L.noise_gain(None))


def test_fft_planner():
    assert ch.fft_plan(8192) == [64, 128]
    assert ch.fft_plan(81920) == [256, 320]
    assert ch.fft_plan(819200) == [80, 80, 128]
    assert ch.fft_plan(2621440) == [128, 128, 160]
    assert ch.fft_plan(2048) == [32, 64]
    assert ch.fft_plan(7 * 1024) is None  # radix 7 unsupported: must be rejected, not mis-computed
    for n in (64, 1600, 4096, 61440, 3 * 2 ** 14):
        p = ch.fft_plan(n)
        assert p is not None and int(np.prod(p)) == n


C_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include <complex.h>
%s
int main(void){
  printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\n", sizeof(struct filter_in), offsetof(struct filter_in, fdomain),
     offsetof(struct filter_in, input), offsetof(struct filter_in, fwd_plan), offsetof(struct filter_in, blocknum),
     offsetof(struct filter_in, filter_cond), offsetof(struct filter_in, ilen), offsetof(struct filter_in, impulse_length),
     offsetof(struct filter_in, input_buffer));
  printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\n", sizeof(struct filter_out), offsetof(struct filter_out, out_type),
     offsetof(struct filter_out, response), offsetof(struct filter_out, response_mutex), offsetof(struct filter_out, f_fdomain),
     offsetof(struct filter_out, noise_gain), offsetof(struct filter_out, output), offsetof(struct filter_out, decimate),
     offsetof(struct filter_out, olen), offsetof(struct filter_out, blocknum));
  printf("%%zu %%zu %%zu %%zu\n", sizeof(struct osc), offsetof(struct osc, phasor), offsetof(struct osc, mutex), offsetof(struct osc, steps));
  printf("%%zu %%zu\n", sizeof(struct hb15_state), offsetof(struct hb15_state, old_odd_samples));
  printf("%%d %%d %%d %%d\n", (int)NONE, (int)COMPLEX, (int)CROSS_CONJ, (int)REAL);
  return 0;
}
"""


def _probe(includes, cflags):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.c")
        exe = os.path.join(d, "p")
        open(src, "w").write(C_PROBE % includes)
        subprocess.run(["gcc", "-std=gnu11", "-o", exe, src] + cflags, check=True, capture_output=True)
        return subprocess.run([exe], check=True, capture_output=True, text=True).stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_struct_layouts_match_the_reference_headers():
    ours = _probe('#include "ka9q_b200.h"', ["-I", os.path.join(ROOT, "include")])
    theirs = _probe('#include <pthread.h>\n#include "filter.h"\n#include "osc.h"\n#include "decimate.h"',
                    ["-I", os.path.join(ROOT, "oracle", "shim"), "-I", REF])
    assert ours == theirs


def test_oscillator_matches_reference_golden():
    """set_osc/step_osc/renorm_osc (osc.c:22-59) incl. the renormalisation at step 16384."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "osc_hb.npz"))
    out = np.zeros(2 * 40000)
    assert _lib.lib().ka9q_osc_run(-2048 / 8192, 0.0, 40000, out.ctypes.data_as(C.c_void_p)) == 0
    z = out[0::2] + 1j * out[1::2]
    np.testing.assert_allclose(z[g["osc_idx"]], g["osc"], rtol=0, atol=1e-13)
    assert abs(abs(z[-1]) - 1) < 1e-12


def test_make_kaiser_matches_reference_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "design.npz"))
    for M, beta, key in ((1089, 3.0, "kaiser_1089_3"), (64, 2.0, "kaiser_64_2")):
        w = np.zeros(M, dtype=np.float32)
        assert _lib.lib().make_kaiser(w.ctypes.data_as(C.c_void_p), M, beta) == 0
        np.testing.assert_allclose(w, g[key], rtol=2e-6, atol=1e-7)
        if M & 1:
            assert w[(M - 1) // 2] == 1.0


def test_oscillator_chirp_matches_reference(ref):
    """The Doppler form of the oscillator: a non-zero sweep rate makes phasor_step itself step (osc.c:31-34,44-46) and
    renorm_osc normalise both (osc.c:57-58). Checked against the verbatim reference across two renormalisations, for the
    rates set_doppler produces (radio.c:180-184: -rate / samprate^2) and for a zero start frequency."""
    for f, r in ((-1500.0 / 192000, -40.0 / 192000 ** 2), (0.0, 3e-9), (0.123, -2.5e-8)):
        n = 40000
        out = np.zeros(2 * n)
        assert _lib.lib().ka9q_osc_run(f, r, n, out.ctypes.data_as(C.c_void_p)) == 0
        got = out[0::2] + 1j * out[1::2]
        want = ref.osc_run(f, r, n)
        # a one-ulp difference in cos / sin of the tiny sweep angle (libm vs gcc's -funsafe-math code in the reference) is a
        # step-of-the-step error and grows with n^2 / 2 = 8e8: 1.5e-8 after 40 000 steps is that, not an algorithmic difference
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-8)
        np.testing.assert_allclose(got[:2000], want[:2000], rtol=0, atol=5e-10)      # (n^2 / 2 = 2e6 there)
        if f != 0.0:
            assert abs(np.angle(got[-1] * np.conj(got[-2])) / (2 * np.pi) - (f + r * (n - 2))) < 1e-9   # it really sweeps
