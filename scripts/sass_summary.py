"""SASS instruction summary per object (cuobjdump -sass): the mnemonics that prove tcgen05 / TMEM / TMA usage.

    python scripts/sass_summary.py [out.txt]        # needs bubbleformer_b200/build/*.o (python bubbleformer_b200/build.py)
UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce,
UBLKCP = cp.async.bulk, LDGSTS = cp.async, HMMA = mma.sync, SYNCS = mbarrier, UTCBAR = tcgen05.commit.
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "LDGSTS", "HMMA",
        "LDSM", "SYNCS", "MUFU.TANH", "MUFU.EX2", "REDG", "RED.", "ATOMS", "ATOMG", "ACQBULK", "ELECT"]


def main():
    out = []
    for o in sorted(glob.glob(os.path.join(ROOT, "bubbleformer_b200", "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        cnt = collections.Counter()
        total = 0
        kernels = 0
        for line in sass.splitlines():
            if "Function :" in line:
                kernels += 1
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            total += 1
            op = m.group(2)
            for k in KEYS:
                if op.startswith(k):
                    cnt[k] += 1
        row = f"{os.path.basename(o):22s} kernels {kernels:3d}  instructions {total:7d}  " + \
            "  ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k])
        out.append(row)
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write("# cuobjdump -sass instruction counts per object (scripts/sass_summary.py), sm_100a\n" + text)
    print(text)


if __name__ == "__main__":
    main()
