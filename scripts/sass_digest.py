"""Per-kernel instruction digest of the built library (cuobjdump -sass): the mnemonics that prove the tensor-core / TMEM /
bulk-copy / cluster paths are real, counted per kernel.  CPU-only.

    python scripts/sass_digest.py [libnws_b200.so] > profiles/r2_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "HMMA", "FFMA2", "MUFU", "SHFL", "LDGSTS", "UCGABAR", "ACQBULK",
         "SYNCS", "ELECT", "LDS", "STS", "ATOM", "RED", "BAR", "CCTL", "MAPA", "ERRBAR"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "neural_waveshaping_synthesis_b200", "libnws_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            op = m.group(1)
            for wname in WATCH:
                if op == wname or op.startswith(wname):
                    cur[wname] += 1
                    break
    print("# cuobjdump -sass digest of %s (sm_100a); per kernel: total SASS instructions and the watched mnemonics" % os.path.basename(lib))
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,")
    print("# HMMA = mma.sync (warp-level tensor core), LDGSTS = cp.async, UCGABAR = barrier.cluster, FFMA2 = fma.rn.f32x2\n")
    tot = collections.Counter()
    for name, c in kernels.items():
        nice = demangle(name).replace("(bool)", "").replace("(int)", "").replace("<unnamed>::", "")
        nice = re.sub(r">\(.*$", ">", nice) if ">(" in nice else re.sub(r"\(.*$", "", nice)
        marks = "  ".join("%s %d" % (k, c[k]) for k in WATCH if c[k])
        print("%-52s %6d  %s" % (nice[:52], c["_total"], marks))
        tot.update(c)
    print("\nTOTAL %d kernels, %d instructions:  %s" % (len(kernels), tot["_total"], "  ".join("%s %d" % (k, tot[k]) for k in WATCH if tot[k])))


if __name__ == "__main__":
    main()
