import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from ka9q_sdr_b200 import channelizer as ch
plan = bench.make_plan("cfg5", 512)
iq = bench.make_input(plan, 4)
c = ch.Channelizer(plan.samprate, plan.L, plan.M, plan.D, max_blocks=4, capture_filter_output=True)
for s in plan.channels: c.add_channel(s.mode, s.bin, low=s.low, high=s.high)
c.commit()
pcm, st = c.process(iq)
print("squelch open frac", st["squelch_open"].mean(), "snr median", np.median(st["snr"]), "pdev median", np.median(st["pdeviation"]))
y = c.filter_output(100, 4)
a = np.abs(y[2])
print("chan100 |y| min/mean", a.min(), a.mean(), "frac below 0.55*avg:", (a*a < 0.3025*(a.mean()/1.0)**2).mean())
bad = 0
for k in range(0, 512, 37):
    y = c.filter_output(k, 4)[2]; a = np.abs(y); bad += int((a*a < 0.3025*a.mean()**2).any())
print("channels with any blanked sample (of 14 probed):", bad)
print("pcm rms", pcm.astype(float).std())
plan = bench.make_plan("cfg5", None)
iq = bench.make_input(plan, 4)
c = ch.Channelizer(plan.samprate, plan.L, plan.M, plan.D, max_blocks=4)
for s in plan.channels: c.add_channel(s.mode, s.bin, low=s.low, high=s.high)
c.commit()
for it in range(3):
    pcm, st = c.process(iq)
    r = st["reserved"][:, :, 0]
    print("iter", it, "squelch open", st["squelch_open"].mean(), "all_good frac per block", (r == 1).mean(axis=1), "snr median", np.median(st["snr"]))
import ctypes as C
st_buf = (ch._lib.ChanStatus * (4 * c.nchan))()
pcm2 = np.empty((4, c.pcm_stride), dtype=np.int16)
for it in range(3):
    c.compute_resident(4); c.fetch(4, pcm2.ctypes.data_as(C.c_void_p), C.cast(st_buf, C.c_void_p)); c.sync()
    st = np.frombuffer(st_buf, dtype=np.dtype(ch._lib.ChanStatus)).reshape(4, c.nchan)
    r = st["reserved"][:, :, 0]
    print("resident iter", it, "open", st["squelch_open"].mean(), "all_good per block", (r == 1).mean(axis=1))
