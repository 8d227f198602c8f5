"""Join an ncu SASS source page (csv) with `nvdisasm --print-line-info-inline` output: per-CUDA-line instruction
counts, issue share and stall samples of one kernel, attributed to the line of the kernel BODY (helpers such as
ddivf / poly_pos / libm are charged to their call site).

usage: ncu_by_line.py src.csv dis.txt kernel_symbol [top_n] [body_first_line body_last_line]
  src.csv : ncu -i prof.ncu-rep --page source --csv
  dis.txt : cuobjdump -xelf all lib.so; nvdisasm --print-line-info-inline frx_kernels.sm_100a.cubin
"""
import collections
import csv
import os
import re
import sys

src_csv, dis, sym = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
body_lo = int(sys.argv[5]) if len(sys.argv) > 5 else 0
body_hi = int(sys.argv[6]) if len(sys.argv) > 6 else 10 ** 9
KERNEL_FILES = ("frx_eval_tile.cuh", "frx_obstacle.cuh", "frx_kernels.cu", "frx_reference.cuh")   # attribution preference: bodies first, helpers second

off2 = {}
chain = []
active = False
fresh = True
for l in open(dis):
    if l.startswith('.text.'):
        active = (sym in l)
        continue
    if not active:
        continue
    if '//## File' in l:
        if fresh:
            chain = []
            fresh = False
        for m in re.finditer(r'(?:File |inlined at )"([^"]*)", line (\d+)', l):
            e = (os.path.basename(m.group(1)), int(m.group(2)))
            if not chain or chain[-1] != e:
                chain.append(e)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        fresh = True
        body = None
        for kf in KERNEL_FILES:
            for f, ln in chain:
                if f == kf and (kf not in KERNEL_FILES[:2] or body_lo <= ln <= body_hi):
                    body = (kf, ln)
                    break
            if body is not None:
                break
        leaf = chain[0] if chain else (None, None)
        off2[int(m.group(1), 16)] = (body, leaf, m.group(2).strip())

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: k for k, h in enumerate(hdr)}
ia, ii, isamp, ith = col['Address'], col['Instructions Executed'], col['# Samples'], col['Thread Instructions Executed']
stall_cols = [(h, k) for h, k in col.items() if h.startswith('stall_') and '(Not Issued)' not in h]
base = None
by = collections.defaultdict(lambda: [0, 0, 0])
ops = collections.Counter()
ops_s = collections.Counter()
stall_tot = collections.Counter()
stall_by_line = collections.defaultdict(collections.Counter)
tot = tot_s = 0
for r in rows[2:]:
    if len(r) <= ii or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    if base is None:
        base = a
    off = a - base
    n = int(float(r[ii] or 0)); s = int(float(r[isamp] or 0)); t = int(float(r[ith] or 0))
    line, leaf, sass = off2.get(off, (None, (None, None), ''))
    by[line][0] += n; by[line][1] += s; by[line][2] += t
    op = sass.split()[0] if sass else '?'
    if op.startswith('@'):
        op = sass.split()[1]
    op = op.split('.')[0]
    ops[op] += n; ops_s[op] += s
    for h, k in stall_cols:
        v = int(float(r[k] or 0))
        stall_tot[h] += v
        stall_by_line[line][h] += v
    tot += n; tot_s += s
print("total warp instructions", tot, " samples", tot_s)
print("stalls:", ", ".join(f"{h[6:]}={100*v/max(tot_s,1):.1f}%" for h, v in stall_tot.most_common(9)))
print("opcodes (dynamic):", ", ".join(f"{o}={100*n/tot:.1f}%/{100*ops_s[o]/max(tot_s,1):.1f}%s" for o, n in ops.most_common(24)))
srcl = {kf: open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'frenetix_motion_planner_b200', 'csrc',
                             kf)).read().split('\n') for kf in KERNEL_FILES}
for line, (n, s, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top_n]:
    top = ",".join(f"{h[6:]}:{v}" for h, v in stall_by_line[line].most_common(3))
    txt = srcl[line[0]][line[1] - 1].strip()[:90] if line else '?'
    line = f"{line[0][4:8]}:{line[1]}" if line else None
    print(f"{n:>9} {100*n/tot:5.1f}%i {100*s/max(tot_s,1):5.1f}%s thr={t/max(n,1):4.1f} [{top}] L{line}: {txt}")
