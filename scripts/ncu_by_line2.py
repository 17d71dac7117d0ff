"""Per-source-line dynamic instruction mix for one kernel of an ncu report (needs the matching .so for line info).
usage: ncu_by_line2.py <report.ncu-rep> <lib.so> <cubin-substring> <kernel-substring> [top]"""
import csv, re, sys, collections, subprocess, os, tempfile, glob
rep, lib, cubin_sub, kern = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 45
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
h = rows[hi]; iE = h.index("Instructions Executed"); iS = h.index("# Samples"); iSrc = h.index("Source")
def cls(op):
    if op.startswith(("FADD", "FMUL", "FFMA", "FSET", "FMNMX", "FSEL", "MUFU", "F2I", "I2F", "F2F", "HFMA", "FCHK")): return "fp"
    if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "ATOM", "RED", "LDGSTS")): return "mem"
    if op.startswith(("BRA", "BSSY", "BSYNC", "BREAK", "BAR", "CALL", "RET", "EXIT", "WARPSYNC", "NOP")): return "ctl"
    return "int"
ex = collections.defaultdict(collections.Counter); sm = collections.Counter(); tot = 0; tots = 0; base = None
for r in rows[hi + 1:]:
    if len(r) <= iSrc or not r[iE].isdigit(): 
        if r and r[0] == "Address": break
        continue
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    if base is None: base = a
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iSrc])
    key = addr2line.get(a - base, ("?", 0))
    n = int(r[iE]); ex[key][cls(m.group(2)) if m else "int"] += n; sm[key] += int(r[iS]); tot += n; tots += int(r[iS])
print("total warp-instr", tot, "samples", tots)
byf = collections.defaultdict(collections.Counter)
for (f, l), c in ex.items():
    for k, v in c.items(): byf[f][k] += v
for f, c in sorted(byf.items(), key=lambda kv: -sum(kv[1].values())):
    print(f"  {f:24s} {100*sum(c.values())/tot:5.1f}%  " + " ".join(f"{k} {100*v/tot:4.1f}" for k, v in c.most_common()))
files = {}
for (f, l), c in sorted(ex.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    if f not in files:
        cand = [p for p in glob.glob("ka9q_sdr_b200/csrc/*") if p.endswith("/" + f)]
        files[f] = open(cand[0]).read().split("\n") if cand else None
    text = files[f][l - 1].strip()[:80] if files[f] and l - 1 < len(files[f]) else ""
    print(f"  {f[:14]:14s} L{l:4d} {100*sum(c.values())/tot:5.1f}% ({' '.join(f'{k}{100*v/tot:.1f}' for k, v in c.most_common())}) samp {100*sm[(f,l)]/tots:4.1f}%  {text}")
