#!/usr/bin/env python3
"""Split an `ncu --page source --print-source sass --csv` dump into regions delimited by BAR.SYNC and
print, per region, the stall-sample share, executed warp instructions and the dominant stall reasons.
Usage: python tools/ncu_regions.py dump.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
regions, cur = [], dict(start=None, n=0, samples=0, inst=0, st={s: 0 for s in stalls}, first="", ops={})
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    smp = int(r[col["# Samples"]] or 0)
    ins = int(r[col["Instructions Executed"]] or 0)
    if cur["start"] is None:
        cur["start"], cur["first"] = r[col["Address"]], src
    cur["n"] += 1; cur["samples"] += smp; cur["inst"] += ins
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + ins
    for s in stalls:
        cur["st"][s] += int(r[col[s]] or 0)
    total += smp
    if "BAR.SYNC" in src or src.startswith("EXIT"):
        cur["end"] = src
        regions.append(cur)
        cur = dict(start=None, n=0, samples=0, inst=0, st={s: 0 for s in stalls}, first="", ops={})
if cur["n"]:
    regions.append(cur)
print("total samples", total)
for k, g in enumerate(regions):
    if g["samples"] < total * 0.002:
        continue
    top = sorted(g["st"].items(), key=lambda kv: -kv[1])[:5]
    ops = sorted(g["ops"].items(), key=lambda kv: -kv[1])[:8]
    print("region %2d  @%s  %4d instr  samples %5.1f%%  warp-inst %6.2f M  | %s" % (
        k, g["start"], g["n"], 100.0 * g["samples"] / total, g["inst"] / 1e6,
        ", ".join("%s %.0f%%" % (a.replace("stall_", ""), 100.0 * b / max(1, g["samples"])) for a, b in top)))
    print("           ops: " + ", ".join("%s %.1fM" % (a, b / 1e6) for a, b in ops))
