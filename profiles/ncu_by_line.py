"""Join an ncu SASS source page (csv) with nvdisasm --print-line-info to get per-CUDA-line
instruction counts and stall samples.  usage: ncu_by_line.py src.csv dis.txt kernel_symbol"""
import csv, re, sys, collections
src_csv, dis, sym = sys.argv[1:4]
# nvdisasm: offset -> (file line, inlined chain)
off2line = {}
cur = None; active = False
for l in open(dis):
    if l.startswith('.text.'):
        active = (sym in l)
        continue
    if not active: continue
    m = re.search(r'//## File ".*?", line (\d+)(.*)', l)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp, ith = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
base = None
by = collections.defaultdict(lambda: [0, 0, 0])
ops = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ii or not r[ia]: continue
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    if base is None: base = a
    off = a - base
    n = int(float(r[ii] or 0)); s = int(float(r[isamp] or 0)); t = int(float(r[ith] or 0))
    line, sass = off2line.get(off, (None, ''))
    by[line][0] += n; by[line][1] += s; by[line][2] += t
    ops[sass.split()[0].split('.')[0] if sass else '?'] += n
    tot += n
print("total warp instructions", tot)
srcl = open('/root/repo/frenetix_motion_planner_b200/csrc/frx_kernels.cu').read().split('\n')
for line, (n, s, t) in sorted(by.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    txt = srcl[line - 1].strip()[:100] if line else '?'
    print(f"{n:>10} {100*n/tot:5.1f}%  samples={s:>6}  thr/inst={t/max(n,1):4.1f}  L{line}: {txt}")
print("top opcodes:", [(k, f"{100*v/tot:.1f}%") for k, v in ops.most_common(25)])
