"""Per-kernel counts of the Blackwell-native SASS opcodes in the shipped library (evidence that the contractions run on
tcgen05 / TMEM / TMA, not on recompiled mma.sync code):  python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = TMA tile
load (cp.async.bulk.tensor), UTMAPF = TMA prefetch, SYNCS = mbarrier ops, HMMA / HGMMA = legacy warp-level MMA (must be 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "laff_b200", "_lib", "liblaff_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "MUFU", "HMMA", "HGMMA", "ATOM", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    names = list(counts)
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except Exception:
        pass
    print("SASS opcode counts per kernel of laff_b200/_lib/liblaff_b200.so (sm_100a), cuobjdump -sass")
    print("%-92s %7s " % ("kernel", "instrs") + " ".join("%7s" % o for o in OPS))
    tot = collections.Counter()
    for k, c in counts.items():
        name = re.sub(r"\(.*", "", demangle.get(k, k))[:92]
        print("%-92s %7d " % (name, c["_total"]) + " ".join("%7d" % c[o] for o in OPS))
        tot.update(c)
    print("%-92s %7d " % ("TOTAL", tot["_total"]) + " ".join("%7d" % tot[o] for o in OPS))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy MMA opcodes found"
    assert tot["UTCHMMA"] > 0 and tot["UTMALDG"] > 0 and tot["LDTM"] > 0


if __name__ == "__main__":
    main()
