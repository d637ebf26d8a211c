"""Counts of the SASS mnemonics that prove which hardware paths each object of the library uses (the table in
/opt/skills/guides/B200_PROFILING.md: tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG, mbarrier = SYNCS;
fp64 tensor = DMMA).  Runs without a GPU on the objects build() leaves in build/.

    python scripts/sass_evidence.py > profiles/r1_sass_mnemonics.md"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build")
WHAT = [("UTC[A-Z]*MMA", "tcgen05.mma"), ("UTCBAR", "tcgen05.commit"), ("LDTM", "tcgen05.ld"), ("UTMALDG", "TMA load"),
        ("SYNCS", "mbarrier"), ("DMMA", "fp64 mma.sync"), ("DFMA", "fp64 fma"), ("SHFL", "warp shuffle"),
        ("LDGSTS", "cp.async"), ("MUFU", "SFU")]


def main():
    print("# SASS mnemonics per object (`cuobjdump -sass build/*.o`, sm_100a)\n")
    print("| object | " + " | ".join("%s (`%s`)" % (w, p) for p, w in WHAT) + " |")
    print("|---|" + "---|" * len(WHAT))
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith(".o") or f.startswith("_"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", txt, flags=re.M)
        print("| `%s` | " % f + " | ".join(str(sum(1 for o in ops if re.fullmatch(p, o))) for p, _ in WHAT) + " |")


if __name__ == "__main__":
    main()
