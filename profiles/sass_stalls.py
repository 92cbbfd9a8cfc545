"""Decode the per-instruction control words (stall count, yield, barriers) of a cuobjdump -sass listing and
sum the static stall cycles between two markers.  usage: python profiles/sass_stalls.py file.sass [regex_from regex_to]"""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
ins = []
i = 0
while i < len(lines):
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 52) & 0x3f))
            i += 2
            continue
    i += 1
frm = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
to = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
on = frm is None
n = st = 0
import collections
by = collections.Counter(); cnt = collections.Counter()
for a, txt, stall, y, w in ins:
    if not on and frm.search(txt):
        on = True
    if on:
        n += 1; st += max(stall, 1)
        op = re.sub(r'^@!?U?P\d+\s+', '', txt).split()[0].split('.')[0]
        by[op] += max(stall, 1); cnt[op] += 1
        if to is not None and n > 1 and to.search(txt):
            break
print(f"{n} instructions, {st} static stall cycles ({st / max(n, 1):.2f} per instruction)")
for op, c in by.most_common(14):
    print(f"  {op:10s} n={cnt[op]:4d} stall={c:5d}")
