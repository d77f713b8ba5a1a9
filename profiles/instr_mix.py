import csv, re, collections, subprocess, sys
rep = sys.argv[1]; nlev = float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
seen = set(); c = collections.Counter(); st = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < 10: continue
    a = r[ix['Address']]
    try: int(a, 16)
    except: continue
    if a in seen: continue
    seen.add(a)
    t = re.sub(r'^\s*@!?U?P\d+\s+', '', r[ix['Source']].strip())
    op = t.split()[0]; key = op.split('.')[0]
    if op.startswith('IMAD.MOV'): key = 'IMAD.MOV'
    elif op.startswith('IMAD.WIDE'): key = 'IMAD.WIDE'
    try: c[key] += float(r[ix['Instructions Executed']]); st[key] += float(r[ix['# Samples']])
    except: pass
tot = sum(c.values()); ts = sum(st.values())
print("instr per warp-level %.1f" % (tot / nlev))
for k, v in c.most_common(28): print("%7.1f  %-10s samples %5.1f%%" % (v / nlev, k, 100 * st[k] / ts))
