import csv, re, collections, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
seen = set(); seq = []
for r in rows[hi + 1:]:
    if len(r) < 10: continue
    a = r[ix['Address']]
    try: int(a, 16)
    except: continue
    if a in seen: continue
    seen.add(a); seq.append(r)
def f(r, k):
    try: return float(r[ix[k]])
    except: return 0.0
tot = sum(f(r, '# Samples') for r in seq)
print("instrs", len(seq), "samples", tot, "inst executed %.4g" % sum(f(r, 'Instructions Executed') for r in seq))
for k in ['stall_long_sb', 'stall_wait', 'stall_math', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_lg', 'stall_mio', 'stall_dispatch', 'stall_branch_resolving', 'stall_no_inst', 'stall_barrier', 'stall_membar', 'stall_sleep', 'stall_drain', 'stall_tex', 'stall_misc']:
    if k in ix: print("%-24s %5.1f%%" % (k, 100 * sum(f(r, k) for r in seq) / tot))
print("top instructions:")
for r in sorted(seq, key=lambda r: -f(r, '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    best = max(((k, f(r, k)) for k in ix if k.startswith('stall_') and 'Not Issued' not in k), key=lambda kv: kv[1])
    print(r[ix['Address']][-5:], "%5.1f%%" % (100 * f(r, '# Samples') / tot), best[0], "exec %.3g" % f(r, 'Instructions Executed'), r[ix['Source']].strip()[:80])
