"""Developer script: DRAM traffic, instruction counts and pipe utilisation per launch of the planner kernels from an
.ncu-rep captured with `ncu --set full` over `tools/profile_target.py <frames> <iters> stage` -> profiles/ncu_traffic.json
(read by bench.py for roofline.traffic / roofline.issue / roofline.cost_matrix).

    python tools/ncu_traffic.py gpurun_out/r2h.ncu-rep 10240
"""
import csv
import io
import json
import subprocess
import sys

rep, frames = sys.argv[1], int(sys.argv[2])
out = {"source": rep.split("/")[-1], "frames_per_launch": frames,
       "note": "per launch over the whole batch (stage entry points), ncu --set full --clock-control none; the first "
               "captured launch of each kernel"}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for name in ("sort_kernel", "match_kernel", "path_kernel", "knn_kernel"):
    hit = [r for r in rows[2:] if ("::" + name + "<") in r[hdr.index("Kernel Name")]]
    if not hit:
        continue
    vals = hit[0]
    g = lambda k: float(vals[hdr.index(k)].replace(",", ""))
    u = lambda k: units[hdr.index(k)]
    out[name] = {"read_bytes": g("dram__bytes_read.sum") * scale[u("dram__bytes_read.sum")],
                 "write_bytes": g("dram__bytes_write.sum") * scale[u("dram__bytes_write.sum")],
                 "warp_inst": g("inst_executed"), "warp_inst_per_frame": g("inst_executed") / frames,
                 "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "fp64_pipe_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                 "lanes_per_warp_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
                 "icc_hit_pct": g("sm__icc_request_hit_rate.pct"),
                 "local_loads": g("sass__inst_executed_local_loads"), "local_stores": g("sass__inst_executed_local_stores"),
                 "duration_ms_under_ncu": g("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u("gpu__time_duration.sum"), 1)}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
