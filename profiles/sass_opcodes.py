#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove (or disprove) a Blackwell-native kernel, from the shipped library:

    python profiles/sass_opcodes.py [path/to/libb200rank.so] > profiles/rNN_sass_opcodes.txt

tcgen05.mma -> UTCHMMA (.2CTA for cta_group::2), tcgen05.ld / .st -> LDTM / STTM, tcgen05.commit -> UTCBAR, TMA loads / stores /
reductions -> UTMALDG / UTMASTG / UTMAREDG, cp.async.bulk -> UBLKCP, mma.sync -> HMMA, MUFU.EX2 = the softmax exponentials
(B200_PROFILING.md "What proves a Blackwell-native kernel"). Needs only cuobjdump (no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "MUFU.EX2", "SYNCS"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "llm-rankers_b200", "libb200rank.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    counts, order, cur, it = {}, [], None, iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it, m.group(1))
            cur = cur.replace("void ", "").replace("b200::", "")
            cur = cur[:cur.find(">(") + 1] if ">(" in cur else cur.split("(")[0]
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            c = counts[cur]
            c["_total"] += 1
            if op.startswith("UTCHMMA"):
                c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
            elif op.startswith("MUFU.EX2"):
                c["MUFU.EX2"] += 1
            else:
                for o in ("UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "SYNCS"):
                    if op.startswith(o):
                        c[o] += 1
                        break
    print(f"# {os.path.relpath(so, ROOT)}: {len(order)} kernels; counts are static SASS instructions per kernel")
    print("kernel".ljust(64) + "".join(o.rjust(13) for o in OPS) + "total".rjust(9))
    tot = collections.Counter()
    for k in sorted(order):
        c = counts[k]
        tot.update(c)
        print(k[:63].ljust(64) + "".join(str(c.get(o, 0) or "-").rjust(13) for o in OPS) + str(c["_total"]).rjust(9))
    print("ALL".ljust(64) + "".join(str(tot.get(o, 0)).rjust(13) for o in OPS) + str(tot["_total"]).rjust(9))


if __name__ == "__main__":
    main()
