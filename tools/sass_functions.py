"""Developer script: per-device-function code size and instruction mix of one kernel in a built .so.
usage: python tools/sass_functions.py <lib.so> <kernel substring, e.g. path_kernelIf> [function substring to dump]"""
import collections
import re
import subprocess
import sys

lib, kern = sys.argv[1], sys.argv[2]
dump = sys.argv[3] if len(sys.argv) > 3 else None
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
funcs = []
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+)", line)
    if m and kern in m.group(3) and "merc" not in line:
        name = m.group(3).split("$")[-1]
        funcs.append((int(m.group(1), 16), int(m.group(2), 16), name))
funcs = sorted(set(funcs))
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = False
ins = {}
for line in sass.splitlines():
    if "Function :" in line:
        cur = kern in line
        continue
    if not cur:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins[int(m.group(1), 16)] = m.group(2).strip()
print(f"{kern}: {len(ins)} instructions, {len(ins) * 16} bytes")
first = funcs[0][0] if funcs else 0
funcs = [(0, first, "<kernel body>")] + funcs
for off, size, name in funcs:
    body = [v for a, v in ins.items() if off <= a < off + size]
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", b).split(".")[0].split()[0] for b in body)
    top = " ".join(f"{k}:{v}" for k, v in mix.most_common(8))
    print(f"{size:7d} B {len(body):5d} ins  {name[:60]:60s} {top}")
    if dump and dump in name:
        for a in sorted(a for a in ins if off <= a < off + size):
            print(f"    {a:06x}  {ins[a]}")
