#!/usr/bin/env python
"""Fill profiles/issue.json and profiles/traffic.json (the `roofline.ncu` / `roofline.traffic` keys of the bench line)
from ncu --set full captures of ONE launch each.

    python tools/ncu_json.py <tag> c2=gpurun_out/x/mega_c2.ncu-rep c3=... c4=...
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
WANT = {"issue_slots_busy_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "active_lanes_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
        "fma_pipe_cycles_active_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "alu_pipe_cycles_active_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "registers_per_thread": "launch__registers_per_thread",
        "warp_instructions": "smsp__inst_executed.sum",
        "duration_ms_under_ncu": "gpu__time_duration.sum"}


def num(v, unit=""):
    x = float(v.replace(",", ""))
    return x


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, row = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, row)}


def scaled(v, u, to):
    x = num(v)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
    return x * mult


def main():
    tag = sys.argv[1]
    issue = {"_note": f"from the {tag} ncu --set full captures of ONE launch (profiles/{tag}_ncu_full_megakernel_*.txt): share of issue slots "
                      "used (an FFMA2 counts as one issue but holds the port two cycles), average active lanes per executed warp-instruction, "
                      "FMA/ALU pipe cycles active. Reported next to the FP32 fraction because config 2 (3 spheres) is bounded by instruction "
                      "issue, not by the sphere loop."}
    traffic = {"_note": f"dram__bytes_read.sum + dram__bytes_write.sum of ONE megakernel launch, from the {tag} ncu --set full captures. c2: 1920x1080 "
                        "x 256 spp launch of megakernel_wq, accumulation written once (33.2 MB algorithmic) and still resident in the 126 MB L2 when "
                        "the kernel ends, plus since r02k the 96 B per pixel of launch constants the claiming lanes read (199 MB at 1080p, written by "
                        "pixel_prologue_kernel just before: the price of running the per-pixel prologue 32 lanes wide instead of 3); c3/c4: 3840x2160 "
                        "launches of the packed forms (132.7 MB algorithmic write), partly evicted to HBM."}
    for spec in sys.argv[2:]:
        name, rep = spec.split("=")
        r = raw(rep)
        d = {}
        for k, m in WANT.items():
            v, u = r[m]
            d[k] = scaled(v, u, None) if k == "duration_ms_under_ncu" else num(v)
        d["kernel"] = r["Kernel Name"][0]
        issue[name] = d
        traffic[name] = int(scaled(*r["dram__bytes_read.sum"], None) + scaled(*r["dram__bytes_write.sum"], None))
    (ROOT / "profiles" / "issue.json").write_text(json.dumps(issue, indent=2) + "\n")
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(traffic, indent=2) + "\n")
    print(json.dumps(issue, indent=1))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
