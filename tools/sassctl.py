"""Decode the scheduling control fields (stall, yield, write / read barrier, wait mask) of a cuobjdump -sass listing.
usage: sassctl.py file.sass [first_addr_hex last_addr_hex]"""
import re, sys
lines = open(sys.argv[1]).read().split("\n")
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
i = 0
pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
pat2 = re.compile(r"/\* (0x[0-9a-f]{16}) \*/")
while i < len(lines):
    m = pat.search(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.search(lines[i + 1])
        if m2:
            addr = int(m.group(1), 16)
            hiw = int(m2.group(1), 16)
            ctrl = (hiw >> 41) & 0x1fffff          # bits 105.. of the 128-bit word
            stall, yld, wb, rb, wait = ctrl & 0xf, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3f
            if lo <= addr <= hi:
                w = "".join(str(b) if wait >> b & 1 else "-" for b in range(6))
                print(f"{addr:05x} st{stall:2d} {'Y' if yld else ' '} W{wb if wb != 7 else '-'} R{rb if rb != 7 else '-'} wait[{w}] {m.group(2).strip()[:90]}")
            i += 2
            continue
    i += 1
