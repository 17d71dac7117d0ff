"""ctypes loader for libka9q_b200.so (the C ABI declared in include/ka9q_b200.h).

The product path has no CPU fallback: if the shared library is missing it fails loudly here, and every compute
entry point of the library itself fails when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libka9q_b200.so")


class ChanParams(C.Structure):
    """struct ka9q_chan_params"""
    _fields_ = [
        ("demod_type", C.c_int), ("flags", C.c_int), ("channels", C.c_int), ("reserved", C.c_int),
        ("bin", C.c_longlong), ("low", C.c_float), ("high", C.c_float), ("kaiser_beta", C.c_float),
        ("shift", C.c_float), ("attack_rate", C.c_float), ("recovery_rate", C.c_float), ("hangtime", C.c_float),
        ("headroom", C.c_float),
    ]


class ChanStatus(C.Structure):
    """struct ka9q_chan_status"""
    _fields_ = [
        ("bb_power", C.c_float), ("snr", C.c_float), ("foffset", C.c_float), ("pdeviation", C.c_float),
        ("agc_gain", C.c_float), ("squelch_open", C.c_int), ("reserved", C.c_float * 2),
    ]


class StreamConfig(C.Structure):
    """struct ka9q_stream_config"""
    _fields_ = [
        ("device", C.c_int), ("samprate", C.c_int), ("L", C.c_int), ("M", C.c_int), ("decimate", C.c_int),
        ("iq_format", C.c_int), ("gain_factor", C.c_float), ("max_blocks", C.c_int),
        ("capture_filter_output", C.c_int),
    ]


class FrontendConfig(C.Structure):
    """struct ka9q_frontend_config"""
    _fields_ = [("device", C.c_int), ("out_samprate", C.c_int), ("decimate", C.c_int), ("offset", C.c_int),
                ("callback_samples", C.c_int), ("dc_alpha", C.c_float), ("power_alpha", C.c_float)]


class FrontendStatus(C.Structure):
    """struct ka9q_frontend_status"""
    _fields_ = [("dc_i", C.c_float), ("dc_q", C.c_float), ("imbalance", C.c_float), ("sinphi", C.c_float),
                ("in_power", C.c_float), ("clips", C.c_longlong), ("samples", C.c_longlong)]


class Hb15State(C.Structure):
    """struct hb15_state (reference decimate.h:4-9)"""
    _fields_ = [("coeffs", C.c_float * 4), ("even_samples", C.c_float * 4), ("odd_samples", C.c_float * 4),
                ("old_odd_samples", C.c_float * 4)]


_lib = None

# every symbol include/ka9q_b200.h declares (checked by tests/test_abi.py)
EXPORTED = [
    "window_filter", "window_rfilter", "create_filter_input", "create_filter_output", "execute_filter_input",
    "execute_filter_output", "delete_filter_input", "delete_filter_output", "make_kaiser", "set_filter", "noise_gain",
    "Kaiser_beta", "ka9q_alloc", "ka9q_free", "set_osc", "step_osc", "renorm_osc", "is_phasor_init", "hb15_block",
    "hb3_block", "ka9q_last_error", "ka9q_version", "ka9q_device_count", "ka9q_stream_create", "ka9q_stream_destroy",
    "ka9q_stream_add_channel", "ka9q_stream_commit", "ka9q_stream_set_filter", "ka9q_stream_num_channels",
    "ka9q_stream_pcm_stride", "ka9q_stream_pcm_offset", "ka9q_stream_olen", "ka9q_stream_fft_size",
    "ka9q_stream_launches_per_call", "ka9q_stream_process", "ka9q_stream_push", "ka9q_stream_compute",
    "ka9q_stream_compute_resident", "ka9q_stream_fetch", "ka9q_stream_sync", "ka9q_stream_last_timing",
    "ka9q_stream_spectrum_ptr", "ka9q_stream_compute_fft_only", "ka9q_stream_compute_channels_only",
    "ka9q_nccl_unique_id", "ka9q_stream_nccl_init", "ka9q_stream_nccl_broadcast_spectrum", "ka9q_stream_get_response",
    "ka9q_stream_get_filter_output", "ka9q_stream_get_spectrum", "ka9q_stream_get_if_energy", "ka9q_fft_c2c",
    "ka9q_fft_plan_describe", "ka9q_hb15_cascade", "ka9q_host_alloc", "ka9q_host_free", "ka9q_stream_timer_start",
    "ka9q_stream_timer_stop", "ka9q_osc_run", "ka9q_stream_wait_fetch", "ka9q_stream_compute_fft_blocks",
    "ka9q_stream_nccl_allgather_spectrum", "ka9q_stream_set_overlap", "ka9q_ingest_init", "ka9q_ingest_datagram",
    "ka9q_rtp_process", "ka9q_pcm_packetise", "ka9q_stream_wait_fetched", "ka9q_status_encode_signals",
    "ka9q_stream_needed_bins", "ka9q_stream_mgpu_export", "ka9q_stream_mgpu_setup", "ka9q_stream_mgpu_input_range",
    "ka9q_stream_push_at", "ka9q_stream_mgpu_compute", "ka9q_stream_mgpu_error", "ka9q_stream_blocks_done",
    "ka9q_stream_enable_n0", "ka9q_stream_fetch_n0", "ka9q_stream_enable_pl", "ka9q_stream_timer_timeline", "ka9q_stream_timer_start_plain", "ka9q_stream_set_fine_lo", "ka9q_stream_split_carrier", "ka9q_frontend_create", "ka9q_frontend_destroy",
    "ka9q_frontend_process", "ka9q_frontend_process_to_stream", "ka9q_frontend_rerun_resident",
    "ka9q_frontend_set_estimates", "ka9q_frontend_get_status", "ka9q_stream_push_device",
    "ka9q_stream_sync_input", "ka9q_rx_create", "ka9q_rx_destroy", "ka9q_rx_inject", "ka9q_rx_drain", "ka9q_rx_start", "ka9q_rx_stop",
    "ka9q_rx_peek_blocks", "ka9q_rx_consume", "ka9q_rx_blocks_ready", "ka9q_rx_get_stats", "ka9q_pcm_send_block",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -m ka9q_sdr_b200.build` "
                           "(there is no CPU fallback for the product path)")
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, cll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    L.ka9q_last_error.restype = C.c_char_p
    L.ka9q_version.restype = C.c_char_p
    L.ka9q_device_count.restype = ci
    L.ka9q_stream_create.argtypes = [C.POINTER(vp), C.POINTER(StreamConfig)]
    L.ka9q_stream_destroy.argtypes = [vp]
    L.ka9q_stream_add_channel.argtypes = [vp, C.POINTER(ChanParams)]
    L.ka9q_stream_commit.argtypes = [vp]
    L.ka9q_stream_set_filter.argtypes = [vp, ci, cf, cf, cf]
    L.ka9q_stream_num_channels.argtypes = [vp]
    L.ka9q_stream_pcm_stride.argtypes = [vp]
    L.ka9q_stream_pcm_stride.restype = cll
    L.ka9q_stream_pcm_offset.argtypes = [vp, ci]
    L.ka9q_stream_olen.argtypes = [vp]
    L.ka9q_stream_fft_size.argtypes = [vp]
    L.ka9q_stream_launches_per_call.argtypes = [vp]
    L.ka9q_stream_process.argtypes = [vp, vp, ci, vp, vp]
    L.ka9q_stream_push.argtypes = [vp, vp, ci]
    L.ka9q_stream_compute.argtypes = [vp, ci]
    L.ka9q_stream_compute_resident.argtypes = [vp, ci]
    L.ka9q_stream_compute_fft_only.argtypes = [vp, ci]
    L.ka9q_stream_compute_channels_only.argtypes = [vp, ci]
    L.ka9q_stream_fetch.argtypes = [vp, ci, vp, vp]
    L.ka9q_stream_sync.argtypes = [vp]
    L.ka9q_stream_set_overlap.argtypes = [vp, ci]
    L.ka9q_stream_compute_fft_blocks.argtypes = [vp, ci, ci, ci]
    L.ka9q_stream_nccl_allgather_spectrum.argtypes = [vp, ci]
    L.ka9q_stream_wait_fetch.argtypes = [vp]
    L.ka9q_stream_wait_fetched.argtypes = [vp, ci]
    L.ka9q_stream_last_timing.argtypes = [vp, C.POINTER(cf), C.POINTER(cf), C.POINTER(cf)]
    L.ka9q_stream_spectrum_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(cll)]
    L.ka9q_nccl_unique_id.argtypes = [vp]
    L.ka9q_stream_nccl_init.argtypes = [vp, vp, ci, ci]
    L.ka9q_stream_nccl_broadcast_spectrum.argtypes = [vp, ci, ci]
    L.ka9q_stream_get_response.argtypes = [vp, ci, vp, C.POINTER(cf)]
    L.ka9q_stream_get_filter_output.argtypes = [vp, ci, ci, vp]
    L.ka9q_stream_get_spectrum.argtypes = [vp, ci, vp]
    L.ka9q_stream_get_if_energy.argtypes = [vp, ci, vp]
    L.ka9q_fft_c2c.argtypes = [ci, ci, ci, ci, vp, vp]
    L.ka9q_fft_plan_describe.argtypes = [ci, C.POINTER(ci)]
    L.ka9q_hb15_cascade.argtypes = [ci, ci, vp, vp, ci, vp]
    L.ka9q_stream_timer_start.argtypes = [vp]
    L.ka9q_stream_timer_stop.argtypes = [vp, C.POINTER(cf), C.POINTER(cf), C.POINTER(ci)]
    L.ka9q_osc_run.argtypes = [C.c_double, C.c_double, C.c_long, vp]
    L.ka9q_stream_needed_bins.argtypes = [vp, C.POINTER(cll), C.POINTER(cll)]
    L.ka9q_stream_mgpu_export.argtypes = [vp, vp]
    L.ka9q_stream_mgpu_setup.argtypes = [vp, ci, ci, ci, C.POINTER(cll), C.POINTER(cll), vp]
    L.ka9q_stream_mgpu_input_range.argtypes = [vp, cll, ci, C.POINTER(cll), C.POINTER(cll)]
    L.ka9q_stream_push_at.argtypes = [vp, vp, cll, cll]
    L.ka9q_stream_mgpu_compute.argtypes = [vp, ci, ci]
    L.ka9q_stream_mgpu_error.argtypes = [vp]
    L.ka9q_frontend_create.argtypes = [C.POINTER(vp), C.POINTER(FrontendConfig)]
    L.ka9q_frontend_destroy.argtypes = [vp]
    L.ka9q_frontend_process.argtypes = [vp, vp, cll, vp]
    L.ka9q_frontend_process_to_stream.argtypes = [vp, vp, cll, vp]
    L.ka9q_frontend_rerun_resident.argtypes = [vp, cll, C.POINTER(cf)]
    L.ka9q_frontend_set_estimates.argtypes = [vp, cf, cf, cf, cf]
    L.ka9q_frontend_get_status.argtypes = [vp, C.POINTER(FrontendStatus)]
    L.ka9q_stream_push_device.argtypes = [vp, vp, cll]
    L.ka9q_stream_sync_input.argtypes = [vp]
    L.ka9q_rx_create.argtypes = [ci, cll, ci]
    L.ka9q_rx_create.restype = vp
    L.ka9q_rx_destroy.argtypes = [vp]
    L.ka9q_rx_destroy.restype = None
    L.ka9q_rx_inject.argtypes = [vp, vp, ci]
    L.ka9q_rx_drain.argtypes = [vp]
    L.ka9q_rx_drain.restype = cll
    L.ka9q_rx_start.argtypes = [vp, ci]
    L.ka9q_rx_stop.argtypes = [vp]
    L.ka9q_rx_peek_blocks.argtypes = [vp, ci, ci]
    L.ka9q_rx_peek_blocks.restype = vp
    L.ka9q_rx_consume.argtypes = [vp, ci]
    L.ka9q_rx_blocks_ready.argtypes = [vp]
    L.ka9q_rx_blocks_ready.restype = cll
    L.ka9q_rx_get_stats.argtypes = [vp, vp]
    L.ka9q_rx_get_stats.restype = None
    L.ka9q_pcm_send_block.argtypes = [ci, vp, vp, vp, vp, ci, ci, ci]
    L.ka9q_stream_enable_pl.argtypes = [vp, ci]
    L.ka9q_stream_timer_timeline.argtypes = [vp, ci, vp, vp, vp]
    L.ka9q_stream_timer_start_plain.argtypes = [vp]
    L.ka9q_stream_set_fine_lo.argtypes = [vp, ci, C.c_double]
    L.ka9q_stream_split_carrier.argtypes = [vp, C.c_double, vp, vp]
    L.ka9q_stream_enable_n0.argtypes = [vp, ci]
    L.ka9q_stream_fetch_n0.argtypes = [vp, ci, vp, vp]
    L.ka9q_stream_blocks_done.argtypes = [vp]
    L.ka9q_stream_blocks_done.restype = cll
    L.ka9q_host_alloc.argtypes = [C.c_size_t]
    L.ka9q_host_alloc.restype = vp
    L.ka9q_host_free.argtypes = [vp]
    # drop-in layer
    L.create_filter_input.argtypes = [C.c_uint, C.c_uint, ci]
    L.create_filter_input.restype = vp
    L.create_filter_output.argtypes = [vp, vp, C.c_uint, ci]
    L.create_filter_output.restype = vp
    L.execute_filter_input.argtypes = [vp]
    L.execute_filter_output.argtypes = [vp]
    L.delete_filter_input.argtypes = [vp]
    L.delete_filter_output.argtypes = [vp]
    L.set_filter.argtypes = [vp, cf, cf, cf]
    L.noise_gain.argtypes = [vp]
    L.noise_gain.restype = cf
    L.make_kaiser.argtypes = [vp, C.c_uint, cf]
    L.window_filter.argtypes = [ci, ci, vp, cf]
    L.window_rfilter.argtypes = [ci, ci, vp, cf]
    L.ka9q_alloc.argtypes = [C.c_size_t]
    L.ka9q_alloc.restype = vp
    L.ka9q_free.argtypes = [vp]
    L.hb15_block.argtypes = [vp, vp, vp, ci]
    L.hb15_block.restype = None
    L.hb3_block.argtypes = [vp, vp, vp, ci]
    L.hb3_block.restype = None
    _lib = L
    return L


def last_error() -> str:
    return lib().ka9q_last_error().decode(errors="replace")


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise RuntimeError(f"{what} failed: {last_error()}")
    return rc
