"""Developer script: hot code footprint of a kernel from `ncu --page source --print-source sass --csv`:
how many bytes of SASS cover 50/80/90/99/100 % of the executed warp instructions, per 128-byte line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
lines = {}
tot = 0
n_ins = 0
for r in rows[2:]:
    if len(r) <= ie:
        continue
    try:
        a, e = int(r[ia], 16), int(r[ie])
    except ValueError:
        continue
    n_ins += 1
    lines[a // 128] = lines.get(a // 128, 0) + e
    tot += e
hot = sorted(lines.values(), reverse=True)
print(f"{n_ins} instructions ({n_ins * 16 / 1024:.0f} KB), {len(hot)} lines of 128 B, {sum(1 for h in hot if h > 0)} lines executed "
      f"({sum(1 for h in hot if h > 0) * 128 / 1024:.0f} KB), {tot} warp instructions")
acc, k = 0, 0
for frac in (0.5, 0.8, 0.9, 0.95, 0.99, 0.999):
    while acc < frac * tot:
        acc += hot[k]
        k += 1
    print(f"  {100 * frac:5.1f} % of the executed instructions come from {k * 128 / 1024:6.1f} KB of code")
