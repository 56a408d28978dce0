"""Static evidence that needs no GPU: per kernel of csrc/liboduck_cuda.so the registers / stack / shared memory (cuobjdump
--dump-resource-usage), the SASS instruction count and the counts of the mnemonics that identify the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier, CREDUX/REDUX =
redux.sync, SHFL, LDS/STS, LDL/STL = local memory, BAR).

    python tools/sass_summary.py [lib.so] > profiles/r01_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "CREDUX", "REDUX", "SHFL", "LDS", "STS", "LDL", "STL", "LDG", "STG", "BAR", "MUFU", "FFMA", "BSSY"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "open_duck_playground_b200", "csrc", "liboduck_cuda.so")
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
    usage, name = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            usage[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            name = None
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            total[cur] += 1
            op = m.group(1)
            if op in MNEMONICS:
                counts[cur][op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(total), capture_output=True, text=True).stdout.splitlines()
    print(f"# Static SASS summary of `{os.path.relpath(lib, ROOT)}` (sm_100a; `python tools/sass_summary.py`)\n")
    print("Counts are static (one per instruction in the listing, warp-collective fallback stubs included), not executed instructions.\n")
    print("| kernel | regs | stack B | static smem B | SASS instr | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|---|---|---|" + "---|" * len(MNEMONICS))
    for mangled, pretty in sorted(zip(total, demangle), key=lambda x: -total[x[0]]):
        u = usage.get(mangled, {})
        pretty = re.sub(r"\(.*", "", pretty).replace("void ", "")
        print(f"| `{pretty}` | {u.get('REG', '')} | {u.get('STACK', '')} | {u.get('SHARED', '')} | {total[mangled]} | " + " | ".join(str(counts[mangled][k] or "") for k in MNEMONICS) + " |")


if __name__ == "__main__":
    main()
