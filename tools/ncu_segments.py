#!/usr/bin/env python
"""Split the SASS of a profiled kernel into runs of equal execution count (loop bodies) and print each
run's share of executed warp-instructions, stall samples and average active threads.
    python tools/ncu_segments.py x.ncu-rep [min_share]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.003
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ith, isamp = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
data = [(int(r[ia], 16), r[isrc], int(r[iex]), float(r[ith] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot = sum(d[2] for d in data); ts = sum(d[4] for d in data); base = data[0][0]
seg, cur = [], None
for a, s, e, t, sm in data:
    off = a - base
    if cur and abs(cur["e"] - e) <= 0.15 * max(cur["e"], e) + 1:
        cur["n"] += 1; cur["tot"] += e; cur["samp"] += sm; cur["thr"] += t * e; cur["end"] = off
    else:
        if cur: seg.append(cur)
        cur = {"start": off, "end": off, "e": e, "n": 1, "tot": e, "samp": sm, "thr": t * e, "first": s}
seg.append(cur)
print(f"{tot} warp-instructions, {ts} samples")
for c in seg:
    if c["tot"] / tot > thr:
        print(f"{c['start']:6x}-{c['end']:6x} n={c['n']:4d} inst={c['tot'] / tot * 100:6.2f}% samp={c['samp'] / ts * 100:6.2f}% "
              f"thr={c['thr'] / max(c['tot'], 1):5.1f} exec/instr={c['e'] / 1e6:9.1f}M  {c['first'][:60]}")
