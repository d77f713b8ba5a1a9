"""Where do the LDGs of the hot loop sit?  Prints, for the main loop of a kernel, the index of every LDG / first
consumer-looking stall candidate relative to the loop length. usage: loadpos.py lib.so pattern"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for b in txt.split("Function : ")[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name: continue
    ins = []
    for line in b.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    back = [(int(re.search(r"0x([0-9a-f]+)", t).group(1), 16), a) for a, t in ins if "BRA" in t and re.search(r"0x([0-9a-f]+)", t) and int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a]
    lo, hi = max(back, key=lambda x: x[1] - x[0])
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    print(name, "loop", hex(lo), hex(hi), "instructions", len(body))
    print("LDG positions:", [i for i, (a, t) in enumerate(body) if "LDG" in t and "CONSTANT" in t])
    print("LDL/STL:", sum(1 for a, t in body if "LDL" in t or "STL" in t))
