"""8192 channels of one mode (default USB) on the cfg5 stream, a few resident steps: for per-kernel launch lists under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import bench
from ka9q_sdr_b200 import channelizer as ch
mode = sys.argv[1] if len(sys.argv) > 1 else "USB"
plan = bench.make_plan("cfg5", None)
B = 4
iq = bench.make_input(plan, B)
c = ch.Channelizer(plan.samprate, plan.L, plan.M, plan.D, max_blocks=B)
for s in plan.channels:
    c.add_channel(mode, s.bin)
c.commit()
pin = ch.PinnedBuffer(iq.nbytes, np.int16); pin.array[:] = iq
for _ in range(2):
    c.push(C.c_void_p(pin.ptr), B); c.compute(B); c.sync()
for _ in range(6):
    c.compute_resident(B)
c.sync()
