#!/usr/bin/env python
"""bench.py — headline benchmark: channel·MS/s demodulated on the cfg5 workload (8192 NBFM channels per GPU on a
61.44 MS/s complex int16 stream), device-resident (`value`) and end to end through the C ABI with host buffers (`e2e`),
plus the roofline of the dominant kernel and the reference's CPU path timed on this box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg5] [--blocks B]

One "step" = one pass of the hot path over one batch of B consecutive 20 ms blocks of synthetic I/Q
(forward FFT once per block + every channel: bin rotation, response multiply, inverse FFT, overlap discard, FM
demodulation, de-emphasis filter, int16 PCM).
Multi-GPU (torchrun, one rank per GPU), default `--mgpu sharded`: STRONG scaling of the same workload — cfg5's 8192
channels are sharded frequency-contiguously over the ranks, the forward FFT is sharded by block, and every producer stores
each peer's arc of the spectrum into that peer's HBM over NVLink (csrc/mgpu.cu; `--transport nccl` uses grouped
ncclSend/ncclRecv instead). A WEAK line (every rank its own 8192 distinct channels, no inter-rank coupling at all) is
measured in the same run and reported under "weak".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the library runs its forward FFT / multi-GPU exchange / channel kernels / copies on separate streams: give every one
# of them its own hardware queue (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "channel·MS/s demodulated"
UNIT = "channel·MS/s"


def make_plan(name: str, nchan: int | None):
    from ka9q_sdr_b200 import workloads
    if name == "cfg5":
        return workloads.cfg5(nchan or 8192)
    if name == "cfg4":
        return workloads.cfg4(nchan or 1024)
    return workloads.CONFIGS[name]()


def make_input(plan, nblocks: int) -> np.ndarray:
    from ka9q_sdr_b200 import synth
    if len(plan.channels) > 128:
        fm_bins = [c.bin for c in plan.channels]
        return synth.comb_spectrum_iq(plan.samprate, nblocks, fm_bins, plan.seed, plan.amplitude, plan.sigma,
                                      deviation=plan.deviation)["iq"]
    return synth.multi_channel(plan.samprate, nblocks, [c.bin for c in plan.channels], [c.mode for c in plan.channels],
                               plan.seed, plan.amplitude, plan.sigma, deviation=plan.deviation)["iq"]


def algorithmic_bytes(plan, nblocks: int):
    """(per-step bytes of the channel kernels, per-step bytes of ingest + forward FFT)"""
    from ka9q_sdr_b200 import modes, workloads
    olen = plan.L // plan.D
    chan = 0
    for c in plan.channels:
        m = modes.get_mode(c.mode)
        pcm_ch = m.channels if m.demod_type == modes.LINEAR_DEMOD else 1
        chan += workloads.channel_block_bytes(m.demod_type, pcm_ch, olen, m.flat)
    return chan * nblocks, workloads.stream_block_bytes(plan.L, plan.N) * nblocks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        try:
            time.sleep(0.15)
            self.proc.terminate()
            self.proc.wait(timeout=5)
            sm, smax, reasons = [], [], set()
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                       "samples": len(sm)}
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(gpu_index: int):
    """Multi-GPU end-to-end leg: every rank moves 83 MB per step over its own PCIe link; run the rank (and so allocate
    its pinned buffers) on the CPUs NVML reports as local to the GPU. Best effort: returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_cpu_channelizer(plan, iq_blocks: np.ndarray, nsample: int = 1024):
    """Second CPU figure, clearly NOT the reference: the shared-FFT channelizer restated in numpy/scipy
    (oracle/channelizer_port.py: one forward FFT per block, rotated windows, vectorised over channels) on the host cores,
    so that the algorithmic gain (shared FFT) and the hardware gain (B200 vs CPU) can be told apart. Bounded sample: the
    forward FFT of every block is done in full, the channel part on `nsample` of the plan's channels and scaled."""
    from oracle import channelizer_port as cp
    fm = [c for c in plan.channels if c.mode in ("FM", "NBFM")]
    if len(fm) != len(plan.channels) or not fm:
        return None
    cores = os.cpu_count() or 1
    sub = fm[:: max(1, len(fm) // nsample)][:nsample]
    ch = cp.FmChannelizer(plan.samprate, plan.L, plan.M, plan.D, [c.bin for c in sub], sub[0].low, sub[0].high, workers=cores)
    nb = iq_blocks.size // (2 * plan.L)
    import scipy.fft as sfft
    t_fft = t_all = 0.0
    for b in range(nb):
        blk = iq_blocks[2 * b * plan.L:2 * (b + 1) * plan.L]
        t0 = time.perf_counter()
        ch.process(blk)
        dt = time.perf_counter() - t0
        if b:                           # block 0 warms the FFT plans up
            t_all += dt
            x = np.zeros(plan.N, dtype=np.complex64)
            t0 = time.perf_counter()
            sfft.fft(x, workers=cores)
            t_fft += time.perf_counter() - t0
    nt = nb - 1
    per_block = t_fft / nt + (t_all - t_fft) / nt * (len(fm) / len(sub))       # full plan: one FFT + all channels
    value = len(fm) * (plan.L / 1e6) / per_block
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "what": "shared-FFT channelizer restated in numpy/scipy (oracle/channelizer_port.py) - NOT the reference",
            "sample": f"{nt} blocks: full forward FFT + {len(sub)} of {len(fm)} channels, channel part scaled linearly",
            "fft_ms_per_block": 1e3 * t_fft / nt, "channels_ms_per_block_full_plan": 1e3 * (t_all - t_fft) / nt * len(fm) / len(sub)}


# ------------------------------------------------------------------------------------------------ reference arm

def run_reference(args, plan, emit=True):
    """The reference's own CPU implementation (oracle/_ref: verbatim reference C, one `radio`-equivalent chain per
    channel = per-sample LO + own N-point forward FFT + inverse FFT + demodulator), one channel per host thread."""
    from concurrent.futures import ThreadPoolExecutor
    from ka9q_sdr_b200 import modes
    from oracle import refbind as R
    cores = os.cpu_count() or 1
    if not R.available():
        line = {"impl": "reference", "unavailable": "oracle/_ref/libka9q_ref.so not built (needs /root/reference at build time)"}
        if emit:
            print(json.dumps(line))
        return line
    R.lib()
    backend_ok = R.set_fft_backend("auto")
    backend = R.fft_backend()
    R.load_modes(modes.MODES.values())
    nthreads = min(cores, len(plan.channels))
    # bounded sample: `nthreads` channels x nb blocks per step
    per_chan_block_cost = plan.N / 2621440 * 0.06 + 0.0007      # rough seconds, only used to size the sample
    nb = int(max(2, min(16, round(1.0 / per_chan_block_cost))))
    if args.ref_blocks:
        nb = args.ref_blocks
    iq = make_input(plan, nb)
    chans = [plan.channels[int(i * len(plan.channels) / nthreads)] for i in range(nthreads)]

    def one(c):
        kw = {}
        if c.low is not None:
            kw["low"], kw["high"] = c.low, c.high
        r = R.chain_run(c.mode, plan.samprate, plan.L, plan.M, plan.D, iq, carrier_hz=c.bin * plan.samprate / plan.N,
                        lo_cycles=-c.bin / plan.N, pkt_samples=4096 if plan.L >= 4096 else 240, **kw)
        return r.nblocks

    def step():
        with ThreadPoolExecutor(nthreads) as ex:
            return list(ex.map(one, chans))

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ms_per_step = dt / args.steps * 1e3
    value = nthreads * (nb * plan.L / 1e6) / (dt / args.steps)
    sample = (f"{nthreads} channels x {nb} blocks of {plan.name} per step, one channel per thread, FFT backend {backend}"
              f"{'' if backend_ok else ' (auto-select failed)'}")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": plan.name, "samprate": plan.samprate, "L": plan.L, "M": plan.M, "N": plan.N,
                   "channels_in_plan": len(plan.channels), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "realtime_channels": value / (plan.samprate / 1e6),
    }
    if emit:
        print(json.dumps(line))
    return line


def _resident_class_times(c, in_ptr, B, steps=20):
    """prime the ring, then time `steps` resident steps with the FFT/channel overlap off: ms per step by kernel class"""
    for _ in range(2):
        c.push(in_ptr, B)
        c.compute(B)
        c.sync()
    for _ in range(10):
        c.compute_resident(B)
    c.sync()
    c.set_overlap(False)
    for _ in range(3):
        c.compute_resident(B)
    c.sync()
    c.timer_start()
    for _ in range(steps):
        c.compute_resident(B)
    ms, classes = c.timer_stop()
    return ms / steps, {k: v[0] / steps for k, v in classes.items()}


def run_other_workloads(args, plan5, pin_in, B):
    import ctypes as C
    from ka9q_sdr_b200 import channelizer as ch, frontend, modes, workloads
    peak, _ = measured_peak_hbm()
    out = []
    olen = plan5.L // plan5.D
    in_ptr = C.c_void_p(pin_in.ptr)
    # 8192 channels of one AM / linear mode on the cfg5 stream
    for mode, cls in (("AM", "am"), ("USB", "linear")):
        c = ch.Channelizer(plan5.samprate, plan5.L, plan5.M, plan5.D, max_blocks=B)
        for spec in plan5.channels:
            c.add_channel(mode, spec.bin)
        c.commit()
        ms, cm = _resident_class_times(c, in_ptr, B)
        m = modes.get_mode(mode)
        nbytes = len(plan5.channels) * B * workloads.channel_block_bytes(m.demod_type, 1, olen)
        out.append({"workload": f"{len(plan5.channels)}x {mode} on the cfg5 stream", "blocks_per_step": B, "ms_per_step": ms,
                    "value": len(plan5.channels) * (B * plan5.L / 1e6) / (ms / 1e3), "unit": UNIT,
                    "kernel_ms": cm[cls], "kernels": "agc_front_kernel + agc_serial_kernel",
                    "roofline": {"bound": "hbm", "achieved": nbytes / (cm[cls] / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": nbytes / (cm[cls] / 1e3) / 1e9 / peak, "algorithmic_bytes_per_step": nbytes}})
        c.close()
    # cfg4: 1024 mixed channels at 19.2 MS/s (N = 819 200)
    p4 = workloads.cfg4()
    iq4 = ch.PinnedBuffer(B * p4.L * 4, np.int16)
    iq4.array[:] = make_input(p4, B)
    c = ch.Channelizer(p4.samprate, p4.L, p4.M, p4.D, max_blocks=B)
    for spec in p4.channels:
        c.add_channel(spec.mode, spec.bin)
    c.commit()
    ms, cm = _resident_class_times(c, C.c_void_p(iq4.ptr), B)
    nb4 = sum(workloads.channel_block_bytes(modes.get_mode(s_.mode).demod_type, 1, olen) for s_ in p4.channels) * B
    out.append({"workload": p4.name, "blocks_per_step": B, "ms_per_step": ms,
                "value": len(p4.channels) * (B * p4.L / 1e6) / (ms / 1e3), "unit": UNIT, "class_ms_per_step": cm,
                "speedup_over_realtime": B * 20.0 / ms,
                "note": "FM, AM and linear kernels run concurrently on three streams; ms_per_step is the serialised sum "
                        "of forward FFT and channel kernels",
                "roofline": {"bound": "hbm", "achieved": nb4 / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": nb4 / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_step": nb4,
                             "what": "all channel kernels + forward FFT of the step"}})
    c.close()
    iq4.free()
    # K5: front-end decimator service, /64 (12.288 MS/s int8 -> 192 kS/s int16 in the reference's hackrf set-up)
    cb, nblk = 131072, 64
    n = cb * nblk
    rng = np.random.default_rng(11)
    iq8 = rng.integers(-40, 40, 2 * n, dtype=np.int8)
    fe = frontend.Frontend(192000, 64, 1, cb)
    fe.set_estimates(0.0, 0.0, 1.0, 0.0)
    fe.process(iq8)
    for _ in range(3):
        fe.rerun_resident(n)
    t = [fe.rerun_resident(n) for _ in range(10)]
    ms = float(np.median(t))
    fbytes = n * 2 + (n // 64) * 4
    out.append({"workload": "front-end decimator service: int8 I/Q, DC / gain / phase correction, Fs/4 rotation, /64 half-band "
                            "cascade on both planes, int16 out (hackrf.c:129-345)", "input_samples_per_step": n,
                "ms_per_step": ms, "value": n / 1e6 / (ms / 1e3), "unit": "MS/s (input rate)",
                "launches_per_step": nblk + 1, "reference_published": "14.8 MS/s per Atom core for the /64 cascade (dcc2018.pdf p.9)",
                "roofline": {"bound": "hbm", "achieved": fbytes / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": fbytes / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_step": fbytes,
                             "what": "2 B in + 4/64 B out per complex input sample; the estimate chain (one launch per "
                                     "131072-sample callback block) is launch-bound, not bandwidth-bound"}})
    fe.close()
    return out


# ------------------------------------------------------------------------------------------------ our arm

def run_ours(args, plan):
    import torch
    import torch.distributed as dist
    from ka9q_sdr_b200 import channelizer as ch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if multi else None   # before any pinned allocation (first touch decides the node)
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = args.blocks if args.blocks else 4
    sharded = multi and args.mgpu == "sharded"
    if multi and args.mgpu in ("allgather", "sharded"):
        B = max(B, world)
        B += (-B) % world          # the block-sharded forward FFT needs blocks_per_step % n_gpus == 0
    if sharded and not args.blocks:
        # Strong scaling shrinks the per-GPU share of a step; keep it larger than the 126 MB L2 (timing rule: no step may
        # run out of a warm L2) by batching more blocks per step: 4 / 4 / 8 / 16 blocks at 1 / 2 / 4 / 8 GPUs on cfg5.
        def per_gpu_bytes(nb):
            kr = len(plan.channels) / world
            olen = plan.L // plan.D
            return (kr * 2048 * 4 + (plan.N / world + 2048) * 8 * nb + (nb // world) * (plan.L * 4 + plan.N * 8 * 3)
                    + nb * kr * olen * 2)
        while per_gpu_bytes(B) < 1.05 * 126e6 and B < 8 * world:
            B += world
    from ka9q_sdr_b200 import workloads
    my_channels = workloads.shard_contiguous(plan, rank, world) if sharded else plan.channels
    c = ch.Channelizer(plan.samprate, plan.L, plan.M, plan.D, device=local, max_blocks=B)
    for spec in my_channels:
        c.add_channel(spec.mode, spec.bin, low=spec.low, high=spec.high)
    c.commit()
    K = c.nchan
    K_total = len(plan.channels) if sharded else world * K     # channels demodulated by the whole job

    # synthetic input (rank 0 ingests the stream); pinned host buffers for the end-to-end leg
    nbytes_in = B * plan.L * 4
    pin_in = ch.PinnedBuffer(nbytes_in, np.int16)
    if rank == 0 or args.mgpu in ("allgather", "replicate", "sharded"):
        pin_in.array[:] = make_input(plan, B)   # block-sharded FFT: every rank ingests (its part of) the int16 stream
    pin_pcm = [ch.PinnedBuffer(B * c.pcm_stride * 2, np.int16) for _ in range(2)]
    in_ptr = C.c_void_p(pin_in.ptr)
    pcm_ptrs = [C.c_void_p(p.ptr) for p in pin_pcm]

    replicate = multi and args.mgpu == "replicate"
    arc = None
    if sharded:
        from ka9q_sdr_b200 import mgpu
        mgpu.setup_sharded(c, rank, world, ch.MGPU_NCCL if args.transport == "nccl" else ch.MGPU_P2P)
        arc = c.needed_bins()
    if multi and not replicate and not sharded:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(ch.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        c.nccl_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        c.sync()

    def spectrum_step():
        """Multi-GPU: produce the batch's spectra on every rank. `allgather` (default): rank r transforms blocks
        [r*B/G, (r+1)*B/G) and the spectra are all-gathered (forward FFT sharded by block, every rank ingests the int16
        stream); `broadcast`: rank 0 transforms everything and broadcasts. Both run on the library's FFT stream and
        overlap the channel kernels of the previous batch."""
        if args.mgpu == "allgather":
            c.compute_fft_blocks(B, rank * (B // world), B // world)
            c.nccl_allgather_spectrum(B)
        else:
            if rank == 0:
                c.compute_fft_only(B)
            c.nccl_broadcast_spectrum(B, 0)

    def step_resident():
        if not multi or replicate:
            c.compute_resident(B)
        elif sharded:
            c.mgpu_compute(B, resident=True)
        else:
            spectrum_step()
            c.compute_channels_only(B)

    e2e_count = [0]
    e2e_block0 = [0]      # stream position (in blocks) where the end-to-end leg starts
    e2e_h2d = [nbytes_in]

    def step_e2e():
        """One batch through the public streaming calls: H2D of the batch's I/Q from pinned host memory, forward FFT +
        all channels, D2H of its PCM into pinned host memory. The calls are asynchronous and the device PCM buffer is
        double-buffered, so the copy-out of batch k overlaps the compute of batch k+1. The copy-out of batch k is queued
        behind that of batch k-1 before the host waits for k-1 (wait_fetched(1)), so the D2H link — the bottleneck of this
        leg — never idles on a host round trip."""
        i = e2e_count[0]
        e2e_count[0] += 1
        if sharded:
            # every rank uploads only the samples its own blocks of the batch need (its blocks + M-1 samples of overlap)
            fb = e2e_block0[0] + i * B
            if i == 0:
                e2e_h2d[0] = mgpu.push_batch_share(c, pin_in.ptr, 4, fb, B, B * plan.L)
            mgpu.push_batch_share(c, pin_in.ptr, 4, fb + B, B, B * plan.L)
            c.mgpu_compute(B, resident=False)
        elif not multi or rank == 0 or args.mgpu in ("allgather", "replicate"):
            if i == 0:
                c.push(in_ptr, B)      # prime: batch 0
            c.push(in_ptr, B)          # batch i+1 goes up while batch i is computed (the ring holds two batches)
        if sharded:
            pass
        elif not multi or replicate:
            c.compute(B)
        else:
            # every rank returns its own PCM rows to the host
            spectrum_step()
            c.compute_channels_only(B)
        c.fetch(B, pcm_ptrs[i & 1])    # batch i: queued behind its compute and behind the copy-out of batch i-1
        if i > 0:
            c.wait_fetched(1)          # batch i-1 has landed in host memory (its buffer is re-used by batch i+1)

    # make the ring resident (all ranks keep a ring; only rank 0's is meaningful in multi-GPU runs)
    # two batches, so that the history in front of the resident batch (the M-1 samples of overlap, the FM audio rings)
    # is the tail of an identical batch — the stimulus is periodic over one batch — and not the start-up zeros, which
    # would put a transient (and the discriminator's blanking path) into block 0 of every resident step
    for _ in range(2):
        c.push(in_ptr, B)
        c.compute(B)
        c.sync()

    # ---- device-resident leg: W warm-up steps, then exactly K timed steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # nvidia-smi needs a few hundred ms to come up, so ~1.2 s of untimed steps run first. The multi-GPU step contains a
    # collective, so the count must be the same on every rank: it is derived from a step time agreed by all-reduce.
    barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        step_resident()
    barrier()
    est = torch.tensor([(time.perf_counter() - t0) / 20], dtype=torch.float64, device="cuda")
    if multi:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)     # same count on every rank
    n_extra = max(300, int(1.2 / max(float(est.item()), 1e-5)))   # ~1.2 s under load before the timed steps
    for _ in range(args.warmup + n_extra):
        step_resident()
    barrier()
    # the headline interval carries no per-kernel event records (they cost microseconds per step); the per-class times
    # come from the serialised pass below
    c.timer_start(regions=bool(args.timeline))
    for _ in range(args.steps):
        step_resident()
    ms_total, _ = c.timer_stop()
    if args.timeline:
        # where the step's time goes when everything overlaps: the bracketed regions of the last steps of the timed
        # interval, per rank (start = the region's stream reached it, end = its kernels finished)
        tl = c.timer_timeline()
        last = [r for r in tl if r[1] >= ms_total - 8 * ms_total / args.steps]
        with open(f"{args.timeline}.rank{rank}.json", "w") as f:
            json.dump({"rank": rank, "ms_per_step": ms_total / args.steps, "t_end_ms": ms_total,
                       "regions": [[n, round(a, 4), round(b, 4)] for n, a, b in last]}, f)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel durations for the roofline: same steps with the FFT/channel overlap switched off, so every event pair
    # brackets one kernel running alone (the overlapped region above is what `value` reports)
    c.set_overlap(False)
    for _ in range(3):
        step_resident()
    barrier()
    c.timer_start()
    nser = max(5, min(args.steps, 20))
    for _ in range(nser):
        step_resident()
    ms_serial, classes = c.timer_stop()
    c.set_overlap(True)
    barrier()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if multi:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    ms_per_step = ms_total_max / args.steps
    ms_blocks_per_s = K_total * (B * plan.L / 1e6) / (ms_per_step / 1e3)   # channel·MS/s of the whole job

    # ---- 1 block per step: the latency-oriented operating point (a 20 ms block as the reference processes it); the
    # responses are then fetched from DRAM every block instead of being re-read from L2 for blocks 2..B of a batch
    b1 = None
    if not multi:
        for _ in range(5):
            c.compute_resident(1)
        barrier()
        c.timer_start()
        n1 = max(20, args.steps)
        for _ in range(n1):
            c.compute_resident(1)
        ms1, cl1 = c.timer_stop()
        barrier()
        b1 = {"blocks_per_step": 1, "ms_per_step": ms1 / n1, "value": K * (plan.L / 1e6) / (ms1 / n1 / 1e3), "unit": UNIT,
              "speedup_over_realtime": 20.0 / (ms1 / n1)}

    # ---- other workloads of the path, device-resident, each with its own HBM-roofline fraction (N = 1 only): 8192 AM and
    # 8192 USB channels on the same stream (the AM / linear kernels), cfg4 (1024 mixed FM/FM/AM/USB channels at 19.2 MS/s),
    # and the front-end decimator service (K5: int8 in, /64, int16 out)
    others = None
    if not multi and args.extras:
        others = []
        try:
            others = run_other_workloads(args, plan, pin_in, B)
        except Exception as e:   # the headline numbers stand on their own
            others = [{"error": str(e)}]

    # ---- end-to-end leg (host pinned buffers, H2D + D2H inside the timed region)
    c.sync()
    if sharded:
        # streaming continues where the resident leg left the stream: the next batch boundary
        e2e_block0[0] = c.blocks_done()
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step_e2e()
    c.wait_fetch()                     # the last batch's PCM is in host memory
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if multi:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = K_total * (B * plan.L / 1e6) / e2e_s
    launches_pc, pcm_stride = c.launches_per_call, c.pcm_stride
    if sharded:
        t = torch.tensor([float(e2e_h2d[0]), float(B * c.pcm_stride * 2)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)                      # bytes over all ranks
        h2d, d2h = int(t[0].item()), int(t[1].item())
    else:
        h2d = nbytes_in * (world if (multi and args.mgpu in ("allgather", "replicate")) else 1)
        d2h = world * B * c.pcm_stride * 2

    # ---- weak line (multi-GPU): every rank demodulates its own 8192 DISTINCT channels (carrier raster shifted by 16 bins
    # per rank) from the stream; every rank ingests the stream and runs its own forward FFT, so the ranks are not coupled
    # at all (the I/Q stream is a multicast any number of receivers can join). Device-resident, same timing rules.
    weak = None
    if multi and args.weak:
        c.close()
        wplan = make_plan(args.config, args.channels) if args.config != "cfg5" else workloads.cfg5(args.channels or 8192, 16 * rank)
        Bw = args.blocks if args.blocks else 4
        cw = ch.Channelizer(wplan.samprate, wplan.L, wplan.M, wplan.D, device=local, max_blocks=Bw)
        for spec in wplan.channels:
            cw.add_channel(spec.mode, spec.bin, low=spec.low, high=spec.high)
        cw.commit()
        for _ in range(2):
            cw.push(in_ptr, Bw)
            cw.compute(Bw)
            cw.sync()
        for _ in range(200):
            cw.compute_resident(Bw)
        cw.sync()
        if multi:
            dist.barrier()
        cw.timer_start(regions=False)
        for _ in range(args.steps):
            cw.compute_resident(Bw)
        msw, _ = cw.timer_stop()
        t = torch.tensor([msw], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        msw = float(t.item()) / args.steps
        weak = {"scaling": "weak", "channels_per_gpu": cw.nchan, "channels_total": world * cw.nchan, "blocks_per_step": Bw,
                "ms_per_step": msw, "value": world * cw.nchan * (Bw * wplan.L / 1e6) / (msw / 1e3), "unit": UNIT,
                "coupling": "none: every rank ingests the stream and runs its own forward FFT on distinct carriers "
                            "(raster shifted by 16 bins per rank)"}
        cw.close()

    if rank != 0:
        if multi:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the FM channel kernel), measured live with CUDA events
    import copy
    rplan = copy.copy(plan)
    rplan.channels = my_channels                 # the dominant kernel's launch covers this rank's channels
    chan_bytes, stream_bytes = algorithmic_bytes(rplan, B)
    peak, peak_src = measured_peak_hbm()
    dom = max(("fm", "am", "linear"), key=lambda k: classes[k][0])
    dom_ms, dom_n = classes[dom]
    roofline = None
    if dom_n:
        avg_ms = dom_ms / dom_n
        ach = chan_bytes / (avg_ms / 1e3) / 1e9 if len({m.mode for m in rplan.channels}) == 1 else None
        if ach is None:
            # mixed plans: only the dominant class's channels count for its kernel
            from ka9q_sdr_b200 import modes
            olen = plan.L // plan.D
            dt = {"fm": 2, "am": 1, "linear": 0}[dom]
            cb = sum(workloads.channel_block_bytes(dt, modes.get_mode(s.mode).channels if dt == 0 else 1, olen,
                                                   modes.get_mode(s.mode).flat)
                     for s in rplan.channels if modes.get_mode(s.mode).demod_type == dt) * B
            ach = cb / (avg_ms / 1e3) / 1e9
        # DRAM traffic per launch comes from an `ncu --set full` capture of this same workload (profiles/traffic.json
        # names the report); it only applies to the single-GPU launch shape it was captured on
        traffic = traffic_src = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and not multi and B == 4 and args.config == "cfg5" and not args.channels:
            try:
                tj = json.load(open(tp))
                traffic, traffic_src = tj.get(f"{dom}_kernel_bytes_per_launch"), tj.get("source")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "avg_launch_ms": avg_ms, "launches": dom_n,
                    "algorithmic_bytes_per_launch": chan_bytes,
                    "share_of_step": dom_ms / ms_serial if ms_serial else None,
                    "timing": "kernel timed alone: FFT/channel overlap switched off for this leg"}

    # ---- CPU baseline on this box's host cores (bounded sample), N=1 only
    cpu = None
    if not multi and not args.no_cpu_baseline:
        try:
            ra = argparse.Namespace(**vars(args))
            ra.steps, ra.warmup = 2, 1
            cpu = run_reference(ra, plan, emit=False).get("cpu_baseline")
        except Exception as e:  # the GPU numbers stand on their own
            cpu = {"error": str(e)}
    cpu_chan = None
    if not multi and not args.no_cpu_baseline:
        try:
            cpu_chan = run_cpu_channelizer(plan, np.asarray(pin_in.array[:2 * min(B, 3) * plan.L]))
        except Exception as e:
            cpu_chan = {"error": str(e)}

    launches_per_step = launches_pc + ((4 if args.transport == "p2p" else 1) if sharded else (1 if multi and not replicate else 0))
    if sharded and multi:
        # per GPU: audio rings, the arc of the B spectra its channels read, its own blocks' transform (input, two scratch
        # buffers, output), its PCM rows
        bown = B // world
        work_mb = (K * 2048 * 4 + arc[1] * 8 * B + bown * (plan.L * 4 + plan.N * 8 * 3) + B * pcm_stride * 2) / 1e6
    else:
        work_mb = (K * 2048 * 8 + K * 2048 * 4 + B * plan.N * 8 * 2 + B * pcm_stride * 2) / 1e6
    if work_mb > 126:
        l2_note = f"per-step working set {work_mb:.0f} MB > 126 MB L2 (responses+state+spectra+PCM); no flush needed"
    else:
        l2_note = (f"per-GPU working set of a step {work_mb:.0f} MB, of the order of the 126 MB L2 and rewritten every step "
                   "(the peers store fresh spectrum arcs over NVLink, the PCM rows are overwritten); no flush: the channel "
                   "kernels are bound by the L1 data pipe with DRAM at 7 % (DESIGN.md section 3), so L2 residency does not "
                   "move this number")
    line = {
        "metric": METRIC, "value": ms_blocks_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if sharded else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": plan.name, "channels_per_gpu": K, "channels_total": K_total, "samprate": plan.samprate,
                   "L": plan.L, "M": plan.M,
                   "N": plan.N, "decimate": plan.D, "blocks_per_step": B, "block_ms": 20,
                   "parallelism": "1 GPU" if not multi else {
                       "sharded": f"{K_total} channels sharded frequency-contiguously over {world} GPUs (strong scaling); "
                                  f"forward FFT sharded by block ({B // world} of {B} blocks per rank and step); every "
                                  "producer sends each peer only the arc of the spectrum its channels read: "
                                  + ("peer-memory stores over NVLink (copy engines by default, this library's copy kernel with "
                                     "KA9Q_B200_MGPU_CE=0) + device-memory sequence flags raised by this library's kernels "
                                     "(no collective library on the data path)" if args.transport == "p2p"
                                     else "grouped ncclSend/ncclRecv"),
                       "replicate": f"channels x{world} (weak); every rank ingests the int16 stream and runs its own "
                                    "forward FFT, no data-path collective",
                       "allgather": f"channels x{world} (weak); forward FFT sharded by block + NCCL all-gather of spectra",
                       "broadcast": f"channels x{world} (weak); NCCL spectrum broadcast from rank 0"}[args.mgpu],
                   **({"cpus_bound_per_rank": numa} if numa else {}),
                   "l2": l2_note},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "cpu_channelizer": cpu_chan,
        "realtime_channels": ms_blocks_per_s / (plan.samprate / 1e6),
        "speedup_over_realtime": (B * 20.0) / ms_per_step,
        "class_ms_per_step": {k: v[0] / nser for k, v in classes.items()},
        "serialised_ms_per_step": ms_serial / nser,
    }
    if b1:
        line["operating_points"] = [b1]
    if others:
        line["other_workloads"] = others
    if weak:
        line["weak"] = weak
    if sharded:
        arc_lo, arc_len = arc
        if args.transport != "p2p":
            form = "grouped ncclSend / ncclRecv of the arcs"
        elif os.environ.get("KA9Q_B200_MGPU_FUSED", "0") not in ("", "0"):
            form = "arcs stored into the peers' HBM by the forward FFT's last pass (fused)"
        elif os.environ.get("KA9Q_B200_MGPU_PULL", "0") not in ("", "0"):
            form = "consumers load their arcs out of the producers' HBM (pull kernel)"
        elif os.environ.get("KA9Q_B200_MGPU_CE", "1") in ("", "0"):
            form = "arcs stored into the peers' HBM by this library's copy kernel; flags at system scope"
        else:
            form = "peer-to-peer copy-engine transfers of the arcs into the peers' HBM + this library's flag kernels (default)"
        line["exchange"] = {"transport": args.transport, "form": form, "arc_bins_this_rank": arc_len,
                            "pipeline": "FFT and exchange of later batches run under the channel kernels of the current one "
                                        "(3 spectrum buffers, high-priority FFT / exchange streams; DESIGN.md section 7)",
                            "nvlink_bytes_in_per_step_this_rank": int(arc_len * 8 * B * (world - 1) / world),
                            "broadcast_would_move_bytes_per_step": int(plan.N * 8 * B * (world - 1) / world)}
    print(json.dumps(line))
    if multi:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--channels", type=int, default=None)
    ap.add_argument("--blocks", type=int, default=None,
                    help="20 ms blocks per step (default 4; sharded multi-GPU runs grow it until a rank's share of a step "
                         "exceeds L2)")
    ap.add_argument("--e2e-steps", type=int, default=100,
                    help="timed steps of the end-to-end leg (the last batch's copy-out drains inside the timed region)")
    ap.add_argument("--ref-blocks", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mgpu", default="sharded", choices=["sharded", "replicate", "allgather", "broadcast"],
                    help="multi-GPU form: sharded (default) = channels sharded over the ranks, FFT sharded by block, "
                         "per-peer spectrum arcs over NVLink (strong scaling); replicate = every rank runs the full "
                         "plan on its own FFT (weak, no coupling); allgather / broadcast = full-plan ranks with the "
                         "spectrum all-gathered / broadcast by NCCL")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="sharded mode: exchange transport")
    ap.add_argument("--timeline", default=None, help="write the per-region timeline of the last timed steps to FILE.rankR.json")
    ap.add_argument("--no-weak", dest="weak", action="store_false", help="skip the weak-scaling line of multi-GPU runs")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the other workloads (AM / USB at scale, cfg4, front-end service) of the N = 1 run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    plan = make_plan(args.config, args.channels)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        run_reference(args, plan)
        return 0
    run_ours(args, plan)
    return 0


if __name__ == "__main__":
    sys.exit(main())
