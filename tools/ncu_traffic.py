"""Developer script: DRAM traffic and instruction counts per launch of the two planner kernels from an .ncu-rep captured
with `ncu --set full` over `tools/profile_target.py <frames> <iters> stage` -> profiles/ncu_traffic.json (read by bench.py).

    python tools/ncu_traffic.py gpurun_out/r1_t.ncu-rep 10240
"""
import csv
import io
import json
import subprocess
import sys

rep, frames = sys.argv[1], int(sys.argv[2])
out = {"source": rep.split("/")[-1], "frames_per_launch": frames,
       "note": "per launch over the whole batch (stage entry points), ncu --set full --clock-control none"}
for name, rx in (("sort_match_kernel", "sort_match_kernel"), ("path_kernel", "path_kernel")):
    raw = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + rx, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[-1]
    units = rows[1]
    g = lambda k: float(vals[hdr.index(k)].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = g("dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = g("dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
    out[name] = {"read_bytes": rd, "write_bytes": wr, "warp_inst": g("inst_executed"),
                 "warp_inst_per_frame": g("inst_executed") / frames,
                 "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "duration_ms_under_ncu": g("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1)}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
