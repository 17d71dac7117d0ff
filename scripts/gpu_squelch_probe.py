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
