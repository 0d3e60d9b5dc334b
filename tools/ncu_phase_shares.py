"""Share of the warp-stall samples, executed instructions and shared-memory wavefronts per PHASE of a kernel, where the phases are
the stretches of SASS between two BAR.SYNC instructions (source page of an `ncu --set full --import-source on` capture).

    python tools/ncu_phase_shares.py gpurun_out/prof_asm5.ncu-rep [kernel-name-substring]
"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
want = sys.argv[2] if len(sys.argv) > 2 else ""
for b, i0 in enumerate(start):
    name = rows[i0][1]
    if want not in name:
        continue
    hdr = rows[i0 + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    end = start[b + 1] if b + 1 < len(start) else len(rows)
    segs, cur = [], dict(samples=0, wave=0, ideal=0, inst=0, lines=0)
    for r in rows[i0 + 2:end]:
        if len(r) < len(hdr):
            continue
        cur["samples"] += int(r[ix["# Samples"]] or 0)
        cur["wave"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
        cur["ideal"] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        cur["inst"] += int(r[ix["Instructions Executed"]] or 0)
        cur["lines"] += 1
        if "BAR.SYNC" in r[ix["Source"]] or r[ix["Source"]].strip().endswith("EXIT"):
            segs.append(cur)
            cur = dict(samples=0, wave=0, ideal=0, inst=0, lines=0)
    segs.append(cur)
    tot = sum(s["samples"] for s in segs) or 1
    print(name)
    print(f"{'phase (ends at the next BAR.SYNC / EXIT)':42s} {'SASS lines':>10s} {'samples':>9s} {'share':>7s} {'warp instr.':>12s} {'smem wavefronts':>16s} {'ideal':>12s}")
    for k, s in enumerate(segs):
        if s["samples"] + s["inst"] == 0:
            continue
        print(f"{'phase ' + str(k):42s} {s['lines']:10d} {s['samples']:9d} {100 * s['samples'] / tot:6.1f}% {s['inst']:12d} {s['wave']:16d} {s['ideal']:12d}")
