"""Per-kernel resource usage of the built library (registers, local-memory stack = spills, static shared memory),
from `cuobjdump --dump-resource-usage` -- runs without a GPU.  Writes a markdown table.

    python scripts/resource_usage.py > profiles/r1_resource_usage.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "riemann_b200", "libriemann_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::|\(.*$", "", n) for n in out]


def main():
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    arch = re.findall(r"arch = (sm_\w+)", txt)
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt)
    names = demangle([r[0] for r in rows])
    print("# Resource usage of every kernel in `libriemann_b200.so`\n")
    print("`cuobjdump --dump-resource-usage` (arch %s), written by `scripts/resource_usage.py`.  STACK > 0 means local-memory"
          % ", ".join(sorted(set(arch))))
    print("frames (spills or local arrays); dynamic shared memory is not listed (see the launch sites).\n")
    print("| kernel | registers | stack (B) | static smem (B) |")
    print("|---|---|---|---|")
    for n, r in sorted(zip(names, rows)):
        print("| `%s` | %s | %s | %s |" % (n, r[1], r[2], r[3]))
    spilled = [n for n, r in zip(names, rows) if int(r[2]) > 0]
    print("\n%d kernels; %d with a local-memory frame%s" % (len(rows), len(spilled), (": " + ", ".join("`%s`" % s for s in sorted(set(spilled)))) if spilled else ""))


if __name__ == "__main__":
    sys.exit(main())
