"""Per-source-line instruction counts and active lanes of one kernel: joins `ncu --page source --csv` (SASS view) with
the line table of the cubin (`cuobjdump -xelf <name> lib.so; nvdisasm --print-line-info <cubin>`; `ncu -i rep --page source --csv --print-source sass`).  usage: ncu_lines.py <src.csv> <disasm> <kernel-substring> [raw.csv]"""
import csv, re, sys
src, dis, kern = sys.argv[1:4]
amap, cur, inkern = {}, None, False
for ln in open(dis):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)          # nvdisasm --print-line-info of a cubin
    if m or ln.startswith('.text.'):
        inkern = kern in ln
        continue
    if ln.lstrip().startswith('.section'):
        inkern = False
        continue
    if not inkern:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
h = rows[1]
ia, ii, it, isamp = h.index('Address'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
base, agg, tot, tots, tott = None, {}, 0, 0, 0
for r in rows[2:]:
    if len(r) <= it:
        continue
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    if base is None:
        base = a
    ie, te, sm = float(r[ii] or 0), float(r[it] or 0), float(r[isamp] or 0)
    d = agg.setdefault(amap.get(a - base), [0, 0, 0])
    d[0] += ie; d[1] += te; d[2] += sm
    tot += ie; tots += sm; tott += te
print(f'total warp inst {tot:.4g}  thread inst {tott:.4g}  avg active {tott / tot:.2f}  samples {tots:.0f}')
for k, (ie, te, sm) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    print(f"{sm / tots * 100:5.1f}% samp {ie / tot * 100:5.1f}% inst  act={te / max(ie, 1):5.1f}  {k}")
