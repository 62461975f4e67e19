"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump: samples and stall reasons per source line.

    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_hot_lines.py src.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: collections.Counter())
text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "":  # a source line row (aggregated over its SASS)
        key = (cur_file, int(r[0]))
        text[key] = r[1].strip()[:90]
        for name, v in zip(hdr[4:], r[4:]):
            if name.startswith("stall_") and "Not Issued" not in name or name in ("# Samples", "Instructions Executed"):
                try:
                    agg[key][name] += float(v)
                except ValueError:
                    pass
total = sum(a["# Samples"] for a in agg.values())
print(f"total samples {total:.0f}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    stalls = sorted(((k, v) for k, v in a.items() if k.startswith("stall_") and v > 0), key=lambda kv: -kv[1])[:3]
    st = " ".join(f"{k[6:]}={v:.0f}" for k, v in stalls)
    print(f"{a['# Samples']:8.0f} {100 * a['# Samples'] / total:5.1f}%  inst={a['Instructions Executed']:10.0f}  {key[0]}:{key[1]:<4d} {st:42s} | {text[key]}")
