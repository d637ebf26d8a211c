#!/usr/bin/env python
"""
Summarise an ncu report for profiles/: key counters (raw page) + the source lines that
execute the most instructions / collect the most stall samples (source page, needs -lineinfo).

    python scripts/ncu_summary.py gpurun_out/X_full.ncu-rep [launches.csv] > profiles/X.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
]
STALLS = "smsp__average_warps_issue_stalled_"


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = raw[0]
    units = raw[1]
    vals = raw[2]
    name = vals[hdr.index("Kernel Name")]
    print("# ncu summary: `%s`\n" % rep.split("/")[-1])
    print("kernel: `%s`\n" % name[:160])
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("| %s | %s | %s |" % (k, vals[i], units[i]))
    st = [(float(vals[i].replace(",", "")), h) for i, h in enumerate(hdr)
          if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and vals[i] not in ("", "n/a")]
    if st:
        print("\nwarp stall reasons (avg warps stalled per issue-active cycle, top 8):\n")
        for v, h in sorted(st, reverse=True)[:8]:
            print("* %s: %.3f" % (h[len(STALLS):-len("_per_issue_active.ratio")], v))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    rows, tot_i, tot_s = [], 0.0, 0.0
    fname = ""
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if len(r) > 7 and r[0].isdigit():
            try:
                n, s = float(r[7]), float(r[6])
            except ValueError:
                continue
            rows.append((n, s, fname, r[0], r[1].strip()[:100]))
            tot_i += n
            tot_s += s
    if rows:
        print("\ntotal warp instructions executed: %.4g; stall samples: %d\n" % (tot_i, tot_s))
        print("| % inst | % samples | file:line | source |\n|---|---|---|---|")
        for n, s, f, ln, code in sorted(rows, reverse=True)[:25]:
            print("| %.1f | %.1f | %s:%s | `%s` |" % (100 * n / tot_i, 100 * s / max(tot_s, 1), f, ln,
                                                   code.replace("|", "\\|")))
    if len(sys.argv) > 2:
        agg = {}
        total = 0.0
        for r in csv.reader(open(sys.argv[2])):
            if len(r) > 5 and r[-1].replace(".", "").isdigit() and "gpu__time_duration" in r[-3]:
                k = r[4].split("(")[0][-70:]
                v = float(r[-1])
                if r[-2] in ("us", "usecond"):
                    v *= 1e3
                elif r[-2] in ("ms", "msecond"):
                    v *= 1e6
                agg.setdefault(k, [0, 0.0])
                agg[k][0] += 1
                agg[k][1] += v
                total += v
        print("\nlaunch list (`%s`): share of summed device time per kernel\n" % sys.argv[2].split("/")[-1])
        print("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, v / 1e6, 100 * v / total))


if __name__ == "__main__":
    main()
