#!/usr/bin/env python
"""Dynamic instruction / stall-sample distribution per CUDA source line:
   python scripts/ncu_by_line.py <ncu source-page csv> <nvdisasm -g -c listing> <mangled kernel substring>"""
import csv, re, collections, sys
srccsv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
L = open(dis).read().split('\n')
start = next(i for i, l in enumerate(L) if l.startswith('.text.') and kname in l and l.rstrip().endswith(':'))
lines = []
cur = None
for ln in L[start + 1:]:
    if ln.startswith('//-----'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', ln): lines.append(cur)
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
ia, ie, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
body = rows[2:]
assert len(body) == len(lines), (len(body), len(lines))
inst = collections.Counter(); samp = collections.Counter()
for r, ln in zip(body, lines):
    inst[ln] += float(r[ie]); samp[ln] += float(r[isamp])
ti, ts = sum(inst.values()), sum(samp.values())
src = {}
print(f"total warp-instructions {ti:.3e}, samples {ts:.0f}")
for (f, l), v in sorted(inst.items(), key=lambda x: -x[1])[:45]:
    if f not in src:
        try: src[f] = open('/root/repo/midaspom_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = None
    text = src[f][l - 1].strip()[:90] if src[f] else ''
    print(f"{100*v/ti:5.1f}% inst {100*samp[(f,l)]/ts:5.1f}% samp  {f}:{l}  {text}")

if len(sys.argv) > 4:      # regions: name:file:lo-hi,...
    print()
    regs = []
    for spec in sys.argv[4].split(','):
        name, f, rng = spec.split(':'); lo, hi = map(int, rng.split('-')); regs.append((name, f, lo, hi))
    agg = collections.Counter(); aggs = collections.Counter()
    for (k, v) in inst.items():
        if k is None: continue
        f, l = k
        for name, rf, lo, hi in regs:
            if f == rf and lo <= l <= hi: agg[name] += v; aggs[name] += samp[k]; break
        else: agg['other'] += v; aggs['other'] += samp[k]
    for name, v in agg.most_common(): print(f"{100*v/ti:5.1f}% inst {100*aggs[name]/ts:5.1f}% samp  {name}")
