"""Per call-site stall samples of one kernel: ncu_stalls.py <sass source csv> <nvdisasm -gi> <kernel> <body-file> <first-body-line>"""
import csv, re, collections, sys
rep_csv, dis, fn, body_file, body_first = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
rows = list(csv.reader(open(rep_csv)))
hdr = rows[1]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); iall = hdr.index('Warp Stall Sampling (All Samples)')
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
data = [r for r in rows[2:] if len(r) > iex]
on = False; chain = []; sites = []
for line in open(dis):
    st = line.strip()
    if st.startswith('.section'): on = ('.text.' + fn) in st
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        loc = (m.group(1).split('/')[-1], int(m.group(2)))
        if 'inlined at' in m.group(3): chain.append(loc)
        else: chain = [loc]
        continue
    if re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line):
        site = None
        for f, ln in reversed(chain):
            if f == body_file and ln >= body_first: site = ln; break
        sites.append(site)
assert len(sites) == len(data), (len(sites), len(data))
tot = sum(int(r[iall]) for r in data)
by = collections.Counter(); reasons = collections.defaultdict(collections.Counter); inst = collections.Counter()
for r, site in zip(data, sites):
    by[site] += int(r[iall]); inst[site] += int(r[iex])
    for h, i in stall_cols.items():
        v = int(r[i] or 0)
        if v: reasons[site][h.replace('stall_', '')] += v
print('total samples', tot)
for k, v in by.most_common(16):
    rs = ', '.join(f'{n} {c/v*100:.0f}%' for n, c in reasons[k].most_common(5))
    print(f'line {k}: samples {v/tot*100:5.1f}%  instr {inst[k]/sum(inst.values())*100:5.1f}%  [{rs}]')
