"""Read an .ncu-rep (no GPU needed) and print the per-kernel numbers DESIGN.md / bench.py quote.

    python profiles/summarise.py gpurun_out/prof.ncu-rep [--lines N] [--kernel regex]
--lines N adds the N most-executed source lines (needs --import-source at capture time).
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    n_lines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
    kern = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else None
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        print("==", name[:90])
        print("  " + "  ".join("%s=%s%s" % (short, r[ix[k]], units[ix[k]] if short in ("time", "dram_rd", "dram_wr") else "")
                               for k, short in KEYS if k in ix))
        st = []
        for s in ("long_scoreboard", "short_scoreboard", "barrier", "wait", "math_pipe_throttle", "mio_throttle",
                  "not_selected", "branch_resolving", "no_instruction", "lg_throttle", "dispatch_stall", "sleeping"):
            k = STALLS % s
            if k in ix and r[ix[k]]:
                st.append("%s=%.2f" % (s, float(r[ix[k]])))
        print("  stalls/issue: " + " ".join(st))
    if n_lines:
        args = ["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
        if kern:
            args += ["-k", "regex:" + kern]
        rows = list(csv.reader(io.StringIO(ncu(args + ["--launch-count", "1"]))))
        cur, agg, byfile = None, [], collections.Counter()
        for r in rows:
            if len(r) == 2 and r[0] == "File Path":
                cur = r[1].split("/")[-1]
                continue
            if len(r) < 8 or r[0] in ("Line No", ""):
                continue
            try:
                n, s = int(r[7]), int(r[6])
            except ValueError:
                continue
            agg.append((n, s, cur, r[0], r[1][:100]))
            byfile[cur] += n
        print("warp instructions by file:", dict(byfile))
        agg.sort(reverse=True)
        for n, s, f, l, src in agg[:n_lines]:
            print("%10d %6d %s:%s %s" % (n, s, f, l, src))


if __name__ == "__main__":
    main()
