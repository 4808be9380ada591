"""Per-kernel SASS evidence of the Blackwell-native paths (B200_PROFILING.md: the PTX names never appear in SASS):

    python tools/sass_counts.py [lib.so] > profiles/r2_sass_counts.txt

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA loads / stores, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync (only the layers with one channel on one side), LDGSTS = cp.async, ELECT / R2UR = what the single-issuer
loops pay when ptxas does not keep them on the uniform datapath.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gan-inpainting_b200", "libb200inpaint.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "LDGSTS", "ELECT", "R2UR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


funcs = re.split(r"\n\s*Function : ", sass)[1:]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l)]
    c = collections.Counter()
    for l in ins:
        m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        op = m.group(1)
        for k in MN:
            if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA.2CTA")):
                c[k] += 1
    rows.append((name, len(ins), c))
dm = demangle([r[0] for r in rows])
print("%-118s %6s " % ("kernel", "instrs") + " ".join("%12s" % k for k in MN))
tot = collections.Counter()
for name, n, c in sorted(rows, key=lambda r: dm[r[0]]):
    short = dm[name].replace("b200::", "").replace("(bool)", "").replace("(int)", "")
    short = re.sub(r"\(CUtensorMap_st.*", "", short).replace("void ", "")
    print("%-118s %6d " % (short[:118], n) + " ".join("%12d" % c[k] for k in MN))
    tot.update(c)
print("%-118s %6s " % ("TOTAL", "") + " ".join("%12d" % tot[k] for k in MN))
