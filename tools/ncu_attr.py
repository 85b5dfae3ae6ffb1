import csv, re, collections, sys
rep_csv, dis_all, fn = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(rep_csv)))
hdr = rows[1]; isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
sass = [(r[isrc].strip(), int(r[iex])) for r in rows[2:] if len(r) > iex]
cur = None; dis = []; on = False
for line in open(dis_all):
    if line.strip().startswith('.section'):
        on = ('.text.' + fn) in line
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m: dis.append((m.group(2).strip(), cur))
print(len(sass), len(dis))
tot = sum(c for _, c in sass)
byline = collections.Counter(); byop = collections.Counter()
for (s, c), (d, loc) in zip(sass, dis):
    byline[loc] += c
    toks = s.split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    byop[op.split('.')[0]] += c
print('total warp instr', tot)
print('--- by opcode')
for k, v in byop.most_common(30): print(f'{k:12s} {v/tot*100:5.1f}%')
print('--- by source line')
for k, v in byline.most_common(45): print(k, f'{v/tot*100:5.1f}%')
