"""Print the per-rank region timelines bench.py --timeline wrote: one line per region, times in microseconds relative
to the first region shown."""
import json, sys, glob
for f in sorted(glob.glob(sys.argv[1] + ".rank*.json")):
    d = json.load(open(f))
    regs = d["regions"]
    if not regs:
        continue
    t0 = regs[0][1]
    print(f"== rank {d['rank']}  {d['ms_per_step'] * 1e3:.1f} us/step")
    for n, a, b in regs:
        print(f"  {n:6s} {1e3 * (a - t0):8.1f} -> {1e3 * (b - t0):8.1f}   ({1e3 * (b - a):6.1f})")
