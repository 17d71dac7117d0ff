"""FM kernel at the per-GPU shapes of strong scaling (cfg5 stream, a contiguous share of the channels, B blocks per step):
ms per launch with one CTA per pair and with the block-split form (KA9Q_B200_FM_SPLIT = 1 / 2 / 4)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import bench
from ka9q_sdr_b200 import channelizer as ch

plan = bench.make_plan("cfg5", None)
shapes = [(1024, 8), (2048, 4), (1024, 4), (512, 8), (64, 8)]
for nch, B in shapes:
    iq = bench.make_input(plan, B)
    pin = ch.PinnedBuffer(iq.nbytes, np.int16)
    pin.array[:] = iq
    for split in ("1", "2", "4", ""):
        if split:
            os.environ["KA9Q_B200_FM_SPLIT"] = split
        else:
            os.environ.pop("KA9Q_B200_FM_SPLIT", None)
        c = ch.Channelizer(plan.samprate, plan.L, plan.M, plan.D, max_blocks=B)
        for s in plan.channels[:nch]:
            c.add_channel("FM", s.bin)
        c.commit()
        ms, cm = bench._resident_class_times(c, C.c_void_p(pin.ptr), B, steps=30)
        print(f"channels {nch} blocks {B} split {split or 'auto'}: step {ms:.4f} ms  fm {cm.get('fm', 0):.4f}  fft {cm.get('fft', 0):.4f}", flush=True)
        c.close()
    pin.free()
