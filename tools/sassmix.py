"""Static SASS opcode mix of one kernel (or of an address range inside it).
usage: sassmix.py lib.so mangled-substring [lo hi]"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name: continue
    c = collections.Counter(); n = 0; bra = []
    for line in b.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m: continue
        a = int(m.group(1), 16)
        ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
        op = ins.split()[0]
        if op.startswith("BRA") :
            t = re.search(r"0x([0-9a-f]+)", ins)
            if t and int(t.group(1), 16) < a: bra.append((hex(int(t.group(1),16)), hex(a)))
        if a < lo or a > hi: continue
        key = op.split(".")[0]
        if op.startswith("IMAD.MOV"): key = "IMAD.MOV"
        elif op.startswith("IMAD.WIDE"): key = "IMAD.WIDE"
        c[key] += 1; n += 1
    print(name); print("backward branches (target, at):", bra); print("instructions in range:", n)
    for k, v in c.most_common(40): print("%5d %s" % (v, k))
