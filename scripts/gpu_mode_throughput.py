"""Channel-kernel throughput by mode at scale: K channels of one mode on the cfg5 stream geometry, device-resident."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from ka9q_sdr_b200 import channelizer as ch

plan = bench.make_plan("cfg5", None)
B = 4
iq = bench.make_input(plan, B)
for mode, K in (("FM", 8192), ("AM", 8192), ("USB", 8192), ("IQ", 8192), ("AM", 1024), ("USB", 1024)):
    c = ch.Channelizer(plan.samprate, plan.L, plan.M, plan.D, max_blocks=B)
    for s in plan.channels[:K]:
        c.add_channel(mode, s.bin, low=(s.low if mode == "FM" else None), high=(s.high if mode == "FM" else None))
    c.commit()
    pin = ch.PinnedBuffer(iq.nbytes, np.int16); pin.array[:] = iq
    import ctypes as C
    for _ in range(2):
        c.push(C.c_void_p(pin.ptr), B); c.compute(B); c.sync()
    for _ in range(20): c.compute_resident(B)
    c.sync()
    c.set_overlap(False)
    for _ in range(3): c.compute_resident(B)
    c.sync(); c.timer_start()
    for _ in range(20): c.compute_resident(B)
    ms, classes = c.timer_stop()
    print(mode, K, "ms/step %.4f" % (ms / 20), {k: round(v[0] / 20, 4) for k, v in classes.items() if v[1]}, flush=True)
    del c
