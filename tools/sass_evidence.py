"""Blackwell-native SASS instructions per kernel of the built library -> profiles/r2_sass_blackwell_instructions.txt
(cuobjdump -sass; runs in the build container, no GPU needed).  usage: python tools/sass_evidence.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nanograd_b200", "lib", "libnanograd_b200.so")
OUT = os.path.join(ROOT, "profiles", "r2_sass_blackwell_instructions.txt")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|UTMALDG|UTMAPF|UTCBAR|LDTM|STTM|UTCATOMSWS|UTMACCTL|UTCCP|SYNCS|"
                 r"UCGABAR_ARV|UCGABAR_WAIT|ELECT)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, samples = None, collections.OrderedDict(), {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        mm = PAT.search(line) if cur else None
        if mm:
            counts.setdefault(cur, collections.Counter())[mm.group(1)] += 1
            samples.setdefault((cur, mm.group(1)), line.strip())
    out = ["# SASS evidence: Blackwell-native instructions per kernel of nanograd_b200/lib/libnanograd_b200.so",
           "# (cuobjdump -sass, sm_100a; UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM/STTM = tcgen05.ld/st (TMEM),",
           "#  UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS = TMEM alloc/dealloc, SYNCS = mbarrier ops, UCGABAR = cluster barrier)",
           ""]
    for k, c in counts.items():
        if not any(x in c for x in ("UTCHMMA", "UTMALDG", "LDTM", "STTM")):
            continue
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        out.append(name)
        out.append("    " + "  ".join(f"{n} x{v}" for n, v in sorted(c.items())))
        for n in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR"):
            if (k, n) in samples:
                out.append("      e.g. " + samples[(k, n)][:150])
        out.append("")
    with open(OUT, "w") as f:
        f.write("\n".join(out))
    print(f"{OUT}: {sum(1 for l in out if l and not l.startswith((' ', '#')))} kernels")


if __name__ == "__main__":
    main()
