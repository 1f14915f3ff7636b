"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libdeepim_b200.so (cuobjdump -sass): proof that the hot
kernels issue tcgen05 MMAs (UTCHMMA / UTCQMMA...), TMA loads (UTMALDG), tensor-memory loads / stores (LDTM / STTM) and
mbarrier traffic (SYNCS), and how many of each sit in the compiled loops.

    python tools/sass_counts.py [path/to/lib.so] > profiles/r02_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mx-deepim_b200", "libdeepim_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS",
         "ELECT", "HMMA", "LDGSTS", "LDSM", "REDUX", "FFMA", "LDG", "STG", "LDS", "STS", "ATOMG", "RED")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            op = m.group(1)
            counts[kern]["_all"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[kern][w] += 1
    print("# %s: SASS mnemonic counts per kernel (static instructions)" % os.path.relpath(LIB, ROOT))
    print("# UTCHMMA = tcgen05.mma kind::f16, UTMALDG = cp.async.bulk.tensor (TMA) load, LDTM / STTM = tcgen05.ld / st, "
          "UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = mma.sync (fc6 only)")
    cols = [w for w in WATCH if any(c[w] for c in counts.values())]
    print("%-78s %6s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
    for k, c in counts.items():
        if any(c[w] for w in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA", "UTCBAR")) or "raster" in k or "zoom_fused" in k:
            print("%-78s %6d " % (k[:78], c["_all"]) + " ".join("%7d" % c[w] for w in cols))
            for w in cols:
                total[w] += c[w]
    print("%-78s %6s " % ("total (listed kernels)", "") + " ".join("%7d" % total[w] for w in cols))


if __name__ == "__main__":
    main()
