import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ka9q_sdr_b200 import frontend
from oracle import refbind as R
import test_gpu_frontend as T
R.lib()
decimate, cb, nblk = 64, 65536, 6
n = cb * nblk
iq = T._adc_stream(n, 7 + decimate, decimate, clip=True, scale=0.2)
fe = frontend.Frontend(192000, decimate, 1, cb); fr = R.Frontend(192000, decimate, 1, cb)
got, want = [], []
for part in (iq[:2 * cb * 2], iq[2 * cb * 2:]):
    got.append(fe.process(part)); want.append(fr.process(part))
    print("status gpu", fe.status()); print("status ref", fr.status())
got, want = np.concatenate(got).astype(np.int32), np.concatenate(want).astype(np.int32)
d = np.abs(got - want)
idx = np.nonzero(d > 1)[0]
print("n>1:", idx.size, "of", d.size, "first idx", idx[:20], "plane", idx[:20] % 2, "out sample", idx[:20] // 2, "cb block", (idx[:20] // 2) * decimate // cb)
for i in idx[:10]: print(i, got[i], want[i])
print("frac exact", (d == 0).mean(), "per-block max", [int(d[2 * b * cb // decimate:2 * (b + 1) * cb // decimate].max()) for b in range(nblk)])
