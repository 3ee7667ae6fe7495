#!/usr/bin/env python3
"""Per-kernel SASS opcode counts of librpc_b200.so (cuobjdump -sass): the Blackwell-specific instructions the DESIGN claims.

    python tools/sass_opcodes.py > profiles/sass_opcodes_r02.txt

DMMA = FP64 tensor-core mma.sync (tcgen05 has no FP64 kind); UBLKCP = TMA bulk copy (cp.async.bulk, both directions);
UTMALDG = tensor-map TMA load (cp.async.bulk.tensor.2d); SYNCS = mbarrier operations; USETMAXREG = setmaxnreg.
Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "randomly_pivoted_cholesky_b200", "librpc_b200.so")
OPS = ["DMMA", "DFMA", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "USETMAXREG", "SHFL", "BAR", "BRX", "LDS", "STG", "LDG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    names = demangle(list(counts))
    print("# cuobjdump -sass %s : %d kernels (sm_100a)" % (os.path.relpath(LIB, ROOT), len(counts)))
    print("# " + " ".join("%9s" % o for o in ["instrs"] + OPS) + "  kernel")
    tot = collections.Counter()
    for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["DMMA"] * 1000 - kv[1]["_total"]):
        tot.update(c)
        short = re.sub(r"\(.*", "", names.get(k, k))
        short = short.replace("(int)", "").replace("(bool)", "")
        print("  " + " ".join("%9d" % c[o] for o in ["_total"] + OPS) + "  " + short[:150])
    print("# totals:")
    print("  " + " ".join("%9d" % tot[o] for o in ["_total"] + OPS))
    for absent in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HGMMA"):
        print("# %s: %d (tcgen05 / TMEM / wgmma have no FP64 kind: none expected)" % (absent, sum(c[absent] for c in counts.values())))


if __name__ == "__main__":
    sys.exit(main())
