#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof/x.ncu-rep profiles/r01_x.txt [--hot 0.01]

Raw page: duration, registers, occupancy limits, issue/pipe utilisation, DRAM bytes, SIMD efficiency,
stall reasons per issue. Source page: opcode mix and the hottest SASS lines (needs -lineinfo).
"""
import csv
import collections
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    hot = float(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 0.01
    lines = [f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on)"]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        lines.append(f"\n## launch: {row[hdr.index('Kernel Name')]}  grid {row[hdr.index('Grid Size')]} block {row[hdr.index('Block Size')]}")
        for h, u, v in zip(hdr, units, row):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                lines.append(f"{h:90s} {v:>18s} {u}")
    src = page(rep, "source")
    # the source page is one block per kernel: "Kernel Name" row, header row, rows
    i = 0
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            kname = src[i][1]
            h = src[i + 1]
            ia, isrc, iex, ith, isamp = (h.index(k) for k in ("Address", "Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
            j = i + 2
            data = []
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                r = src[j]
                if len(r) > iex and r[iex].isdigit():
                    data.append((r[ia], r[isrc], int(r[iex]), float(r[ith] or 0), int(r[isamp] or 0)))
                j += 1
            tot = sum(d[2] for d in data) or 1
            ts = sum(d[4] for d in data) or 1
            lines.append(f"\n## SASS of {kname}: {len(data)} instructions, {tot} warp-instructions executed, "
                         f"avg active threads {sum(d[2] * d[3] for d in data) / tot:.2f}")
            op, ops = collections.Counter(), collections.Counter()
            for a, s, e, t, sm in data:
                parts = s.split()
                k = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
                op[k] += e
                ops[k] += sm
            lines.append("opcode                     %inst   %stall-samples")
            for k, v in op.most_common(24):
                lines.append(f"{k:24s} {v / tot * 100:7.2f} {ops[k] / ts * 100:9.2f}")
            lines.append(f"\nhottest lines (>= {hot * 100:.1f}% of executed warp-instructions):")
            for a, s, e, t, sm in data:
                if e >= tot * hot:
                    lines.append(f"{a[-5:]} {e / tot * 100:5.2f}% thr={t:4.1f} samp={sm / ts * 100:5.2f}%  {s[:110]}")
            i = j
        else:
            i += 1
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst, len(lines), "lines")


if __name__ == "__main__":
    main()
