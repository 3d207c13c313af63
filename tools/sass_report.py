"""Static evidence from the built library (no GPU needed): per-kernel registers / stack / shared memory from
`cuobjdump --dump-resource-usage`, and per-kernel counts of the SASS mnemonics that show how a kernel moves data and
issues its arithmetic (bulk-TMA copies UBLKCP, mbarrier SYNCS, packed FP32 FFMA2, 3-input min FMNMX3, FP64 DFMA/DADD,
atomics, local-memory spills STL/LDL).
usage: python tools/sass_report.py [lib/libicpb200.so] > profiles/r02_sass_resource_usage.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "icp-slam-prototype_b200", "lib", "libicpb200.so")
MNEMONICS = ["UBLKCP", "SYNCS", "FFMA2", "FFMA", "FMNMX3", "FMNMX", "DFMA", "DADD", "DMUL", "MUFU", "ATOM", "ATOMG", "RED",
             "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "MATCH", "VOTE", "BAR", "ACQBULK", "UTMALDG", "UTCHMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    names = demangle(sorted(usage))
    print(f"# {os.path.relpath(LIB, ROOT)}: sm_100a only; produced by tools/sass_report.py (cuobjdump {len(usage)} kernels)")
    print("# columns: registers, stack bytes, static shared bytes, SASS instructions, then the non-zero mnemonic counts")
    tot = collections.Counter()
    for k in sorted(usage, key=lambda k: names[k]):
        reg, stack, shared, local = usage[k]
        c = counts.get(k, {})
        short = re.sub(r"\(.*\)$", "", names[k]).replace("icpb::", "")
        short = re.sub(r"^void ", "", short)
        ops = " ".join(f"{m}={c[m]}" for m in MNEMONICS if c.get(m))
        for m in MNEMONICS:
            tot[m] += c.get(m, 0)
        print(f"{short:60s} reg={reg:3d} stack={stack:4d} smem={shared:6d} sass={c.get('_total', 0):6d}  {ops}")
    print("# library totals: " + " ".join(f"{m}={tot[m]}" for m in MNEMONICS if tot[m]))
    print("# UBLKCP = 1-D bulk TMA copy (cp.async.bulk); no UTMALDG / UTC*MMA: nothing on this path is a tiled tensor or a "
          "GEMM (K = 3 contraction, north star)")


if __name__ == "__main__":
    main()
