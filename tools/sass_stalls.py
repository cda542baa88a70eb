"""Static issue-time estimate of a SASS address range: sum of the per-instruction stall counts of the control words
(cuobjdump -sass prints the two 64-bit words of every instruction; stall = bits 41..44 of the second).  Scoreboard waits
(MUFU / LDS / LDC) come on top.   python tools/sass_stalls.py <sass file> <function substring> <lo hex> <hi hex>"""
import re, sys
path, fn, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
lines = open(path).read().splitlines()
infn = False
ins = []
i = 0
while i < len(lines):
    l = lines[i]
    if "Function :" in l:
        infn = fn in l
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if infn and m:
        addr = int(m.group(1), 16)
        m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
        hiw = int(m2.group(1), 16)
        ins.append((addr, m.group(2).strip(), (hiw >> 41) & 0xf, (hiw >> 45) & 1, (hiw >> 46) & 7, (hiw >> 49) & 7, (hiw >> 52) & 0x3f))
        i += 2
        continue
    i += 1
sel = [x for x in ins if lo <= x[0] < hi]
tot = sum(x[2] for x in sel)
print(f"{len(sel)} instructions, stall sum {tot} cycles, {tot/ max(1,len(sel)):.2f} per instruction; waits on scoreboards: {sum(1 for x in sel if x[6])}")
if len(sys.argv) > 5:
    for x in sel:
        print(f"{x[0]:05x} st={x[2]:2d} y={x[3]} wr={x[4]} rd={x[5]} wait={x[6]:06b}  {x[1]}")
