"""Summaries of an .ncu-rep for profiles/: key raw metrics and instruction count / samples per function."""
import collections, csv, io, re, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
kfilter = ["--kernel-name", "regex:" + sys.argv[4]] if len(sys.argv) > 4 else []  # reports holding several kernels
raw = subprocess.run(["ncu", "-i", rep, *kfilter, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
n_launches = max(len(rows) - 2, 1)  # the source page below aggregates every captured launch of the kernel
keep = ["gpu__time_duration.sum", "launch__", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__pcsamp_warps_issue_stalled",
        "sm__pipe_fp64_cycles_active", "sm__inst_executed.sum", "sm__icc_request", "gcc__cache_requests_type_instruction", "gcc__average_cache_request_hit_rate",
        "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate.pct", "sass__inst_executed_local"]
with open(f"profiles/{tag}_ncu_raw.txt", "w") as f:
    for h, v in zip(hdr, vals):
        if any(w in h for w in keep) and "not_issued" not in h:
            f.write(f"{h} = {v}\n")
src = subprocess.run(["ncu", "-i", rep, *kfilter, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
funcs = {}
import glob, os
for path in glob.glob("ft_fsd_path_planning_b200/csrc/*.cu*"):
    cur, m = None, {}
    for i, l in enumerate(open(path), 1):
        mm = re.match(r"^(?:FSD_DEVFN|FSD_DEV|__global__|__device__)[^(]*?\b([A-Za-z_0-9]+)\(", l)
        if mm:
            cur = mm.group(1)
        m[i] = cur
    funcs[os.path.basename(path)] = m
inst, samp = collections.Counter(), collections.Counter()
cur_file, h2 = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        h2 = r
    elif r[0] not in ("Function Name", "") and h2 is not None:
        try:
            ln, n, s = int(r[0]), float(r[7]), float(r[6])
        except ValueError:
            continue
        fn = funcs.get(cur_file, {}).get(ln, cur_file)
        inst[(cur_file, fn)] += n
        samp[(cur_file, fn)] += s
tot, ts = sum(inst.values()), sum(samp.values())
frames = (int(sys.argv[3]) if len(sys.argv) > 3 else 10240) * n_launches
with open(f"profiles/{tag}_by_function.txt", "w") as f:
    f.write(f"warp instructions per frame: {tot / frames:.0f}\n")
    for k, v in inst.most_common(24):
        f.write(f"{v / frames:9.0f} inst/frame {100 * v / tot:5.1f}%  samples {100 * samp[k] / ts:5.1f}%  {k[0]}:{k[1]}\n")
print(open(f"profiles/{tag}_by_function.txt").read())
