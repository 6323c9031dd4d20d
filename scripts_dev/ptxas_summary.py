"""condense `nvcc -Xptxas -v` output: one line per kernel of interest (registers, spills, shared memory)"""
import re, sys, subprocess
txt = sys.stdin.read()
cur = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1); continue
    if 'error' in line.lower() and 'ptxas info' not in line: print(line)
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: spill = m.groups(); continue
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        name = subprocess.run(['c++filt', cur], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", name)[:110]
        if any(k in short for k in ('k_admm_cols_tma', 'k_gd_cols_tma', 'k_rows_ifft', 'k_rows_spatial', 'k_gd_rows')):
            print(f"{short:110s} regs {m.group(1):>3s} spill {spill[1]}/{spill[2]}")
