"""List registers / spills per kernel from a verbose build (python build.py --force -v | python profiles/ptxas_report.py)."""
import re, subprocess, sys
txt = sys.stdin.read()
name = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = m.groups()
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        print(f"{m.group(1):>4} regs  stack {spill[0]:>4}  spill st/ld {spill[1]:>4}/{spill[2]:>4}  {name}")
        name = None
