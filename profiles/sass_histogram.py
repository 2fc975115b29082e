#!/usr/bin/env python
"""SASS mnemonic histogram of the shipped objects (customnerf_b200/_obj/*.o): the tcgen05 / TMEM / TMA / async-copy / atomic
instructions per kernel -- evidence that the field network runs on the 5th-generation tensor cores (UTCHMMA = tcgen05.mma
kind::f16, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMASTG = cp.async.bulk.tensor store, LDGSTS = cp.async) and of
what the scatter kernels issue (RED.E.ADD.F32x2).  Run here (no GPU needed):  python profiles/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "customnerf_b200", "_obj")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMASTG", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "REDG", "RED", "ATOMG", "ATOMS",
         "SHFL", "REDUX", "LDG", "STG", "LDS", "STS", "MUFU", "F2I", "I2F", "FRND", "HMMA", "BAR", "LDGMC", "ELECT", "MEMBAR"]


def main():
    files = sorted(f for f in os.listdir(OBJ) if f.endswith(".o"))
    only = sys.argv[1:] or None
    for f in files:
        if only and not any(o in f for o in only):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], stdout=subprocess.PIPE, text=True).stdout
        kern, hist = None, {}
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
                kern = re.sub(r"\(anonymous namespace\)::", "", kern)
                kern = kern.split("(")[0][:90]
                hist[kern] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and kern:
                op = m.group(1)
                hist[kern]["_total"] += 1
                for w in WATCH:
                    if op == w or op.startswith(w + "."):
                        hist[kern][w] += 1
        print("== %s" % f)
        for k, h in hist.items():
            if h["_total"] < 40:
                continue
            parts = ["%s %d" % (w, h[w]) for w in WATCH if h[w]]
            print("  %-92s %5d instr | %s" % (k, h["_total"], ", ".join(parts)))


if __name__ == "__main__":
    main()
