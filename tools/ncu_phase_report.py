"""Per-phase reading of an `ncu --set full --import-source on` capture of one kernel, made on the CPU box:

    python tools/ncu_phase_report.py <report.ncu-rep> <object or cubin with the kernel> <kernel name regex> <units (e.g. tetrahedra)> [top=25]

The SASS page of the report (instructions executed, stall samples, shared-memory wavefronts per instruction) is matched line by line
with `nvdisasm -g` of the same build (source line of every instruction), cut into phases at the kernel's BAR.SYNC instructions, and the
headline counters of the raw page are printed beside it.  Used for profiles/r02_*.txt."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kre, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=False).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, vals = raw[0], raw[2]
R = dict(zip(hdr, vals))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "smsp__inst_executed_pipe_lsu.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"]
print("== raw counters")
for k in want:
    if k in R:
        print("  %-80s %s" % (k, R[k]))
for k, v in R.items():
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(v or 0) > 0.3:
        print("  %-80s %s" % (k, v))
print("  instructions per unit: %.1f" % (float(R["smsp__inst_executed.sum"]) / units))
src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
sh, data = src[1], src[2:]
iex, ism, iw, iwi, isrc = sh.index("Instructions Executed"), sh.index("# Samples"), sh.index("L1 Wavefronts Shared"), sh.index("L1 Wavefronts Shared Ideal"), sh.index("Source")
tmp = tempfile.mkdtemp()
if obj.endswith(".cubin"):
    cub = obj
else:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = run(["nvdisasm", "-g", "-c", cub]).splitlines()
insts, cur, on = [], None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        on = re.search(kre, ln) is not None and not insts
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        insts.append((m.group(2).strip(), cur))
if len(insts) != len(data):
    print("!! SASS of the report (%d instructions) and of %s (%d) differ: rebuild mismatch" % (len(data), obj, len(insts)))
    sys.exit(1)
bars = [i for i, (t, c) in enumerate(insts) if t.split()[-2 if t.startswith("@") else 0].startswith("BAR") or " BAR." in " " + t]
print("== phases (cut at BAR.SYNC, source line of the barrier)")
prev = 0
tot_s = sum(int(d[ism]) for d in data)
for b in bars + [len(insts)]:
    ex = sum(int(d[iex]) for d in data[prev:b]); sm = sum(int(d[ism]) for d in data[prev:b]); w = sum(int(d[iw] or 0) for d in data[prev:b]); wi = sum(int(d[iwi] or 0) for d in data[prev:b])
    where = insts[b][1] if b < len(insts) else None
    print("  SASS %5d..%5d  inst/unit %6.1f  samples %5.1f %%  smem wavefronts/unit %5.1f (ideal %5.1f)  ends at %s" % (prev, b, ex / units, 100.0 * sm / max(tot_s, 1), w / units, wi / units, where))
    prev = b
print("== hottest instructions (stall samples)")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ism]))[:top]
for i in sorted(order):
    print("  %5d  inst/unit %5.2f  samples %5d  line %-28s %s" % (i, int(data[i][iex]) / units, int(data[i][ism]), insts[i][1], insts[i][0][:90]))
print("== opcodes per unit")
hist = collections.Counter()
for (t, c), d in zip(insts, data):
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    hist[op.split(".")[0]] += int(d[iex])
print("  " + ", ".join("%s %.1f" % (k, v / units) for k, v in hist.most_common(24)))
