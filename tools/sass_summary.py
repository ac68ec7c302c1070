"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell-native mnemonics (tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP, fp64 mma.sync -> DMMA) in every kernel of librange_b200.so.
    python tools/sass_summary.py > profiles/sass_summary.txt          (runs anywhere: cuobjdump, no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "range_b200", "librange_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "HMMA", "MUFU.EX2", "SYNCS",
       "ST.E", "STG.E"]


def pretty(mangled):
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = name.replace("(anonymous namespace)::", "").replace("rangeb200::", "")
    name = re.sub(r"^void ", "", name)
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):              # cut the argument list: the first "(" outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return name[:cut]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = pretty(m.group(1))
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        for p in PAT:
            if re.search(r"\b" + re.escape(p), line):
                counts[fn][p] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    print("# SASS summary of range_b200/librange_b200.so (cuobjdump -sass; sm_100a).  Columns: instruction counts per kernel.")
    print("# tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP,")
    print("# mma.sync fp64 -> DMMA, mbarrier -> SYNCS; STG/ST.E on the routed epilogue are the peer (NVLink) stores.")
    print("| kernel | " + " | ".join(PAT) + " |")
    print("|---|" + "---|" * len(PAT))
    for fn, c in sorted(counts.items()):
        print(f"| {fn} | " + " | ".join(str(c.get(p, 0)) if c.get(p, 0) else "" for p in PAT) + " |")
    print("\n# resource usage (cuobjdump -res-usage)")
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = pretty(m.group(1))
        elif "REG:" in line and name:
            print(f"{name}: {line.strip()}")
            name = None


if __name__ == "__main__":
    main()
