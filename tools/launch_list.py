#!/usr/bin/env python
"""Developer tool: print the kernels of one steady-state bench step from an ncu launch-list CSV
(ncu --metrics gpu__time_duration.sum --csv ...)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
idx = [i for i, r in enumerate(rows) if 'preprocess_fwd' in r[4]]
s, e = idx[3], idx[4]
tot = 0.0
for r in rows[s:e]:
    name = r[4].split('(')[0]
    if 'FillFunctor<unsigned char>' in name:
        continue
    tot += float(r[14])
    print(f"{float(r[14]) / 1e3:9.2f} us  {r[7]:>14} {r[8]:>16}  {name[-70:]}")
print(f"{tot / 1e3:9.2f} us  total")
