"""Developer script: SASS evidence for profiles/ -- per kernel of the built libfsdplan.so: code bytes, opcode histogram,
the TMA / mbarrier instructions (UBLKCP, SYNCS), fp64 / shuffle / shared-memory / local-memory instruction counts and
the largest device functions.  Runs without a GPU (cuobjdump on the cross-compiled library).

    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ft_fsd_path_planning_b200", "csrc", "libfsdplan.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        kernels[cur].append(re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip()))
def demangle(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    d = d.replace("(anonymous namespace)::", "").replace("void ", "")
    m = re.match(r"([\w:]+(?:<[^>]*>)?)", d)
    return m.group(1) if m else d
funcs = collections.defaultdict(list)  # kernel -> [(size, name)]
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+)", line)
    if m and "merc" not in line:
        kern, _, fn = m.group(3).partition("$")
        funcs[kern].append((int(m.group(2), 16), fn))
print(f"SASS summary of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); one block per kernel\n")
for k, ins in kernels.items():
    ops = collections.Counter(i.split()[0].split(".")[0] for i in ins)
    full = collections.Counter(i.split()[0] for i in ins)
    name = demangle(k)
    print(f"== {name}\n   {len(ins)} instructions, {16 * len(ins)} bytes of code")
    print("   top opcodes: " + " ".join(f"{o}:{c}" for o, c in ops.most_common(14)))
    tma = {o: c for o, c in full.items() if o.startswith(("UBLKCP", "SYNCS", "UTMA", "CCTL", "ATOMG", "RED", "ATOM"))}
    if tma:
        print("   TMA / mbarrier / atomics / cache control: " + " ".join(f"{o}:{c}" for o, c in sorted(tma.items())))
    grp = lambda *p: sum(c for o, c in ops.items() if o in p)
    print(f"   fp64 (DADD DMUL DFMA DSETP MUFU): {grp('DADD', 'DMUL', 'DFMA', 'DSETP', 'MUFU')}   shuffles / votes (SHFL VOTE REDUX MATCH): "
          f"{grp('SHFL', 'VOTE', 'REDUX', 'MATCH')}   shared (LDS STS): {grp('LDS', 'STS')}   local (LDL STL): {grp('LDL', 'STL')}   "
          f"global (LDG STG): {grp('LDG', 'STG')}   barriers (BAR WARPSYNC): {grp('BAR', 'WARPSYNC')}")
    fl = sorted(funcs.get(k, []), reverse=True)[:8]
    if fl:
        print("   largest device functions: " + ", ".join(f"{demangle(fn).split('::')[-1]} {sz} B" for sz, fn in fl))
    print()
