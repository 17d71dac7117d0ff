"""Join an ncu --page source CSV with nvdisasm -g line info: dynamic warp-instructions and stall samples per source line."""
import csv, re, sys, collections
lines_txt, src_csv, kernel, srcfile = sys.argv[1:5]
addr2line = {}
cur = None; fn = None
for line in open(lines_txt):
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s', line)
    if m and fn and kernel in fn and cur: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
base = None
ex = collections.Counter(); sm = collections.Counter(); tot = 0; tots = 0
for r in rows[hi + 1:]:
    try: a = int(r[0], 16) if r[0].startswith('0x') else int(r[0]); n = int(r[iE]); s = int(r[iS])
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base, ('?', 0))
    ex[key] += n; sm[key] += s; tot += n; tots += s
src = open(srcfile).read().split('\n')
byfile = collections.Counter(); sfile = collections.Counter()
for (f, l), n in ex.items(): byfile[f] += n; sfile[f] += sm[(f, l)]
print('total warp-instr', tot, 'samples', tots)
for f, n in byfile.most_common(): print(f'  {f:28s} {100*n/tot:5.1f}% instr  {100*sfile[f]/tots:5.1f}% samples')
name = srcfile.split('/')[-1]
print('top lines of', name)
for (f, l), n in sorted(((k, v) for k, v in ex.items() if k[0] == name), key=lambda kv: -sm[kv[0]])[:45]:
    print(f'  L{l:4d} {100*n/tot:5.1f}% instr {100*sm[(f,l)]/tots:5.1f}% samp  {src[l-1].strip()[:100]}')
