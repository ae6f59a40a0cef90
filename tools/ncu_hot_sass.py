#!/usr/bin/env python
"""Hot SASS of one kernel from an .ncu-rep (source page): share of executed instructions and of stall samples.
Usage: python tools/ncu_hot_sass.py REPORT KERNEL_REGEX [min_pct]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix['Instructions Executed']].isdigit()]
tot = sum(int(r[ix['Instructions Executed']]) for r in data)
tots = sum(int(r[ix['# Samples']]) for r in data)
print('total warp inst', tot, 'samples', tots, 'sass lines', len(data))
for k, r in enumerate(data):
    n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    if n > thr / 100 * tot or s > thr * 1.5 / 100 * tots:
        print(f"{k:4d} {n / tot * 100:5.2f}% s{s / tots * 100:5.2f}% thr{r[ix['Avg. Predicated-On Threads Executed']]:>5} {r[ix['Source']].strip()}")
