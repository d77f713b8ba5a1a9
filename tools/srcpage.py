"""Summarise an `ncu --page source --csv --print-source sass` dump: instructions executed and stall samples per code
region (split at the given SASS addresses or by executed-count plateaus)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, name):
    try: return float(r[ix[name]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# segment by "Instructions Executed" level: contiguous runs with similar execution count
segs = []
cur = None
for n, r in enumerate(data):
    ex = f(r, "Instructions Executed")
    if cur is None or not (0.7 * cur["ex"] <= ex <= 1.4 * cur["ex"]) :
        cur = {"start": n, "ex": ex, "n": 0, "inst": 0.0, "samples": 0.0, "st": collections.Counter(), "ops": collections.Counter()}
        segs.append(cur)
    cur["n"] += 1; cur["inst"] += ex; cur["samples"] += f(r, "# Samples")
    cur["ex"] = (cur["ex"] * (cur["n"] - 1) + ex) / cur["n"]
    for s_ in stalls: cur["st"][s_] += f(r, s_)
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    cur["ops"][op[0].split(".")[0] if op else "?"] += 1
tot_i = sum(s["inst"] for s in segs); tot_s = sum(s["samples"] for s in segs)
print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
for s in segs:
    if s["inst"] < 0.004 * tot_i and s["samples"] < 0.004 * tot_s: continue
    top = ", ".join(f"{k[6:]} {v / max(s['samples'], 1):.2f}" for k, v in s["st"].most_common(4))
    ops = ", ".join(f"{k} {v}" for k, v in s["ops"].most_common(8))
    print(f"sass[{s['start']:5d}..{s['start'] + s['n'] - 1:5d}] n={s['n']:4d} exec/instr {s['ex']:.3e} inst {100 * s['inst'] / tot_i:5.1f}% samples {100 * s['samples'] / tot_s:5.1f}% | {top} | {ops}")
