"""Instruction budget of the headline kernel from a saved `ncu --set full --import-source on` report (no GPU needed):
executed warp-instructions per warp-step, by how often each instruction runs (1.0 = every step, 0.82 = the
trans-dimensional block, 0.25 = the thinned diagnostics ...), by opcode, and by source file:line region.

    python scripts/ncu_instruction_budget.py gpurun_out/r22_changepoint_full.ncu-rep 8192 1000 > profiles/r1_changepoint_instruction_budget.md

(8192 = warps of the launch, 1000 = MH iterations of the launch: one warp-step = 8 chain-steps at 4 lanes per chain.)"""
import collections
import csv
import io
import re
import subprocess
import sys

REGIONS = [   # (file suffix, first line, last line, label) -- riemann_b200/csrc at the profiled commit
    ("common.cuh", 40, 75, "Philox4x32-10 rounds"),
    ("common.cuh", 76, 100, "Box-Muller normals"),
    ("changepoint.cuh", 40, 95, "binary search of the x table / uniforms"),
    ("changepoint.cu", 63, 82, "group ballot / butterfly sums"),
    ("changepoint.cu", 83, 119, "shift_prev / shift_next (neighbour element)"),
    ("changepoint.cu", 120, 141, "shift_by (insert / delete gather)"),
    ("changepoint.cu", 142, 165, "elem_at / count_below"),
    ("changepoint.cu", 166, 233, "log-posterior of the proposal"),
    ("changepoint.cu", 240, 294, "prologue (state load, Philox prefetch)"),
    ("changepoint.cu", 295, 349, "per-step randoms and move selection"),
    ("changepoint.cu", 350, 366, "fixed-dimension proposal by selection"),
    ("changepoint.cu", 367, 410, "birth / death arithmetic and rebuild"),
    ("changepoint.cu", 411, 437, "padding + run-boundary search"),
    ("changepoint.cu", 438, 460, "Philox prefetch, accept, state update"),
    ("changepoint.cu", 461, 478, "thinned diagnostics"),
    ("changepoint.cu", 479, 540, "trace stores + epilogue"),
]


def main():
    rep, warps, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    base = float(warps * iters)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    path, line = "", 0
    seen = {}                                # address -> [count, sass, [(file, line), ...]]: inlined code is listed once
    for r in rows:                           # per frame of its inline stack, so every address is counted ONCE
        if len(r) == 2 and r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name"):
            continue
        if r[0] != "":                       # a source line header; its SASS rows follow
            line = int(r[0]) if r[0].isdigit() else 0
            continue
        if not r[2].startswith("0x"):
            continue
        try:
            n = int(r[7])
        except ValueError:
            continue
        seen.setdefault(r[2], [n, r[3], []])[2].append((path, line))
    by_freq, by_op, by_region = (collections.Counter() for _ in range(3))
    static_freq = collections.Counter()
    total = 0
    rank = {"changepoint.cu": 3, "changepoint.cuh": 2, "common.cuh": 1}
    for n, sass_txt, frames in seen.values():
        total += n
        f = round(n / base, 2)
        by_freq[f] += n
        static_freq[f] += 1
        sass = sass_txt.split()
        op = sass[1] if sass and sass[0].startswith("@") else (sass[0] if sass else "?")
        by_op[op.split(".")[0]] += n
        # attribute to the outermost frame ncu lists inside our sources (helpers are defined above their callers)
        fpath, fline = max(frames, key=lambda fr: (rank.get(fr[0], 0) if fr[1] > 0 else -1, fr[1]))
        label = "other (inlined math library: log, division, sqrt slow paths; intrinsics headers)"
        for suffix, lo, hi, lab in REGIONS:
            if fpath == suffix and lo <= fline <= hi:
                label = lab
                break
        by_region[label] += n
    print("# Instruction budget of `changepoint_kernel<0,2,4>` (from `%s`)\n" % rep.split("/")[-1])
    print("%.3e executed warp-instructions = **%.0f per warp-step = %.0f per chain-step** (8 chains per warp).\n"
          % (total, total / base, total / base / 8))
    print("## By how often an instruction runs\n\n| runs on this fraction of warp-steps | static instructions | executed per warp-step | share |\n|---|---|---|---|")
    for f, n in sorted(by_freq.items(), key=lambda kv: -kv[1])[:10]:
        print("| %.2f | %d | %.0f | %.1f %% |" % (f, static_freq[f], n / base, 100.0 * n / total))
    print("\n## By source region\n\n| region | executed per warp-step | share |\n|---|---|---|")
    for lab, n in sorted(by_region.items(), key=lambda kv: -kv[1]):
        print("| %s | %.0f | %.1f %% |" % (lab, n / base, 100.0 * n / total))
    print("\n## By opcode\n\n| opcode | executed per warp-step | share |\n|---|---|---|")
    for op, n in by_op.most_common(16):
        print("| `%s` | %.0f | %.1f %% |" % (op, n / base, 100.0 * n / total))


if __name__ == "__main__":
    main()
