"""Summarise `nvcc -Xptxas -v` output read from stdin: one line per kernel (registers, stack, spills)."""
import re
import subprocess
import sys

cur, rows = None, []
for ln in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        cur = {'fn': m.group(1)}
        rows.append(cur)
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', ln)
    if m and cur is not None and 'stack' not in cur:
        cur['stack'], cur['st'], cur['ld'] = map(int, m.groups())
    m = re.search(r'Used (\d+) registers', ln)
    if m and cur is not None:
        cur['regs'] = int(m.group(1))
names = subprocess.run(['c++filt'] + [r['fn'] for r in rows], capture_output=True, text=True).stdout.splitlines() if rows else []
for r, n in zip(rows, names):
    n = re.sub(r'\(.*', '', n)
    print(f"{r.get('regs', '?'):>4} regs  stack {r.get('stack', 0):>4}  spill st/ld {r.get('st', 0):>4}/{r.get('ld', 0):<4}  {n}")
