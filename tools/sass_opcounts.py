"""Per-kernel SASS opcode evidence of the built library: counts of the tcgen05 / TMA / TMEM mnemonics (UTCHMMA = tcgen05.mma,
UTMALDG = TMA load, UTMAPF = TMA prefetch, UTMASTG = TMA store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = bulk copy) and of
the generic ones that would betray a fallback (HMMA = mma.sync).   usage: python tools/sass_opcounts.py > profiles/r2_sass_opcounts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200", "libb2seg.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "REDG", "RED", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            if op in OPS:
                counts[cur][op] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels; opcode counts per kernel (0 columns omitted)")
    tot = collections.Counter()
    for (name, c), d in zip(counts.items(), dem):
        d = re.sub(r"\(.*", "", d)
        cols = " ".join(f"{op}={c[op]}" for op in OPS if c[op])
        print(f"{d:<70} SASS={c['_total']:<6} {cols}")
        tot.update({k: v for k, v in c.items() if k != "_total"})
    print("# totals: " + " ".join(f"{op}={tot[op]}" for op in OPS if tot[op]))


if __name__ == "__main__":
    sys.exit(main())
