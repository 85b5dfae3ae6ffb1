"""Attribute executed warp-instructions of one kernel to call-site lines of its body (outermost inlining frame).
usage: ncu_attr2.py <ncu source-page csv (sass)> <nvdisasm -gi output> <kernel> <body-file> <first-body-line>"""
import csv, re, collections, sys
rep_csv, dis, fn, body_file, body_first = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
rows = list(csv.reader(open(rep_csv)))
hdr = rows[1]; isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
sass = [(r[isrc].strip(), int(r[iex])) for r in rows[2:] if len(r) > iex]
on = False; chain = []; out = []
for line in open(dis):
    st = line.strip()
    if st.startswith('.section'):
        on = ('.text.' + fn) in st
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        f, ln, rest = m.group(1).split('/')[-1], int(m.group(2)), m.group(3)
        if 'inlined at' in rest:
            chain.append((f, ln))
        else:
            chain = [(f, ln)]
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        # outermost frame inside the kernel body
        site = None
        for f, ln in reversed(chain):
            if f == body_file and ln >= body_first:
                site = ln; break
        out.append((m.group(2).strip(), site, chain[0] if chain else None))
assert len(out) == len(sass), (len(out), len(sass))
tot = sum(c for _, c in sass)
by = collections.Counter(); byop = collections.defaultdict(collections.Counter)
for (s, c), (d, site, inner) in zip(sass, out):
    by[site] += c
    toks = s.split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    byop[site][op.split('.')[0]] += c
print('total warp instr', tot)
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:25]:
    ops = ', '.join(f'{o} {n/v*100:.0f}%' for o, n in byop[k].most_common(6))
    print(f'line {k}: {v/tot*100:5.1f}%   [{ops}]')
