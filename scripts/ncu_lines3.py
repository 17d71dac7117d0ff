"""Per-source-line table for one kernel of an ncu report: warp instructions, shared wavefronts, global tag requests and
sampled stalls, joined with nvdisasm line info from the matching .so.
usage: ncu_lines3.py <report.ncu-rep> <lib.so> <cubin-substring> <kernel-substring> [top]"""
import csv, re, sys, collections, subprocess, os, tempfile, glob
rep, lib, cubin_sub, kern = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in glob.glob(tmp + "/*.cubin") if cubin_sub in f][0]
dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
addr2line = {}; fn = None; cur = None
for line in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); cur = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s", line)
    if m and fn and kern in fn and cur: addr2line[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
col = {n: h.index(n) for n in ("Instructions Executed", "# Samples", "Source", "L1 Wavefronts Shared", "L1 Tag Requests Global",
                               "L1 Wavefronts Shared Ideal", "L2 Theoretical Sectors Global")}
agg = collections.defaultdict(lambda: collections.Counter()); base = None; tot = collections.Counter()
def num(x):
    try: return int(float(x))
    except Exception: return 0
for r in rows[hi + 1:]:
    if len(r) <= col["Source"] or not r[col["Instructions Executed"]].isdigit():
        if r and r[0] == "Address": break
        continue
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    for n, c in col.items():
        if n == "Source": continue
        v = num(r[c]); agg[key][n] += v; tot[n] += v
print("totals:", dict(tot))
files = {}
def text(f, l):
    if f not in files:
        cand = [p for p in glob.glob("ka9q_sdr_b200/csrc/*") if p.endswith("/" + f)]
        files[f] = open(cand[0]).read().split("\n") if cand else None
    return files[f][l - 1].strip()[:70] if files[f] and 0 < l <= len(files[f]) else ""
T = tot["Instructions Executed"]; S = tot["# Samples"]; W = max(1, tot["L1 Wavefronts Shared"]); G = max(1, tot["L1 Tag Requests Global"])
print(f"{'file':14s} line  inst%  samp%  shWF%  glRq%  source")
for (f, l), c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print(f"{f[:14]:14s} {l:4d} {100*c['Instructions Executed']/T:6.1f} {100*c['# Samples']/S:6.1f} "
          f"{100*c['L1 Wavefronts Shared']/W:6.1f} {100*c['L1 Tag Requests Global']/G:6.1f}  {text(f,l)}")
