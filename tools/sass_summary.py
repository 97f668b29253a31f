#!/usr/bin/env python
"""Static evidence that the contraction / convolution kernels run on the 5th-generation tensor
cores with TMA-staged operands: counts of the sm_100a SASS mnemonics that only tcgen05 / TMEM /
TMA code contains (UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA
tensor load / store, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM allocation) per kernel of
libdusty_b200.so.  No GPU needed.

    python tools/sass_summary.py > profiles/r01_sass_tcgen05_tma.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "HMMA", "SYNCS")


def main():
    lib = os.path.join(ROOT, "dusty_gan_v2_b200", "libdusty_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                           capture_output=True, text=True).stdout.splitlines()
    table, cur, it = OrderedDict(), None, iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it, m.group(1))
            cur = cur.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("dusty::", "")
            cur = re.sub(r"\((?:int|bool|long|unsigned int)\)", "", cur)       # template-argument casts
            cur = re.sub(r"\(.*$", "", cur)                                  # parameter list
            table[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            table[cur]["_all"] += 1
            for k in KEYS:
                if m.group(1).startswith(k):
                    table[cur][k] += 1
    print("# SASS mnemonic counts per kernel (static instructions), sm_100a build of libdusty_b200.so")
    print(f"{'kernel':78s} " + " ".join(f"{k:>10s}" for k in KEYS) + f" {'all':>7s}")
    for name, c in table.items():
        if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
            print(f"{name[:78]:78s} " + " ".join(f"{c[k]:10d}" for k in KEYS) + f" {c['_all']:7d}")
    rest = [n for n, c in table.items() if not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"])]
    print(f"# {len(rest)} further kernels (memory-bound stencils, elementwise, reductions) use none of these")
    legacy = [n for n, c in table.items() if c["HMMA"] and not c["UTCHMMA"]]
    print(f"# kernels using legacy mma.sync (HMMA) without tcgen05: {len(legacy)} {legacy[:5]}")


if __name__ == "__main__":
    sys.exit(main())
