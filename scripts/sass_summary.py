"""Per-kernel SASS evidence for libctr_b200.so: counts of the Blackwell-specific mnemonics
(tcgen05 MMA = UTCHMMA/UTCQMMA..., TMEM loads = LDTM, TMA tensor loads = UTMALDG, 1-D bulk copy =
UBLKCP, vector reductions = REDG / RED, match-any = MATCH, atomics = ATOMG).
Usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt   (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "recsys_b200", "libctr_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS",
       "REDG", "RED.", "ATOMG", "ATOMS", "MATCH", "HMMA", "MEMBAR", "ERRBAR", "CCTL", "ACQBULK",
       "UCGABAR", "VOTE"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_insts"] += 1
        for p in PAT:
            if op.startswith(p):
                counts[cur][p] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("SASS summary of recsys_b200/libctr_b200.so (cuobjdump -sass, sm_100a); per kernel: "
          "instruction count and the counts of the mnemonics that matter")
    print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM), UTMALDG = TMA tensor load, UBLKCP = "
          "cp.async.bulk (1-D TMA), REDG = red.global, MATCH = match.any, HMMA = mma.sync, "
          "UCGABAR = barrier.cluster, VOTE = ballot\n")
    for (k, c), name in zip(counts.items(), demangle):
        name = re.sub(r"\(.*", "", name)[:90]
        tags = "  ".join("%s=%d" % (p, c[p]) for p in PAT if c[p])
        print("%-92s insts=%-6d %s" % (name, c["_insts"], tags))


if __name__ == "__main__":
    main()
