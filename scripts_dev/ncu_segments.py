"""Aggregate an `ncu --page source --csv` export (SASS view) into code segments delimited by
barriers / mbarrier ops / bulk copies: share of stall samples, top stall reasons, top opcodes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def new(): return {'n': 0, 'samples': 0, 'exec': 0, 'stalls': {}, 'ops': {}, 'wf': 0, 'wfi': 0}
seg = []; cur = new(); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    if r[0] == 'Address':
        break  # second launch
    src = r[isrc]; smp = int(r[ismp] or 0); ex = int(r[iex] or 0)
    cur['n'] += 1; cur['samples'] += smp; cur['exec'] += ex; tot += smp
    cur['wf'] += int(r[iw] or 0); cur['wfi'] += int(r[iwi] or 0)
    t = src.split()
    op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')).split('.')[0]
    cur['ops'][op] = cur['ops'].get(op, 0) + smp
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: cur['stalls'][hdr[i]] = cur['stalls'].get(hdr[i], 0) + v
    if 'BAR.' in src or 'SYNCS' in src or 'UBLKCP' in src:
        cur['end'] = src[:60]; seg.append(cur); cur = new()
seg.append(cur)
print("total samples", tot)
for s in seg:
    if s['samples'] < tot * 0.004: continue
    st = sorted(s['stalls'].items(), key=lambda x: -x[1])[:4]
    ops = sorted(s['ops'].items(), key=lambda x: -x[1])[:5]
    print(f"{s['n']:5d} instr {100*s['samples']/tot:5.1f}% samples exec {s['exec']:9d} smem wf {s['wf']}/{s['wfi']} end={s.get('end','')}\n      stalls {st}\n      ops {ops}")
