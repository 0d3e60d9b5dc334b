"""Per-kernel histogram of the SASS mnemonics that identify the Blackwell paths of libwae_b200.so (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): DMMA (FP64 tensor op, mma.sync.m8n8k4.f64 -- tcgen05 has no f64 kind), UBLKCP / UTMALDG (TMA bulk copies),
SYNCS (mbarrier), UCGABAR (thread-block-cluster barrier), LDGSTS (cp.async), UTC*MMA / LDTM / STTM (tcgen05 + TMEM: none expected for an FP64 path), HMMA (legacy tensor path:
none expected), RED / ATOM (atomics).

    python tools/sass_ops.py [path to libwae_b200.so] > profiles/r02_sass_ops.txt
"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "wavesandeigenvalues.jl_b200", "libwae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WATCH = [("DMMA", r"\bDMMA\b"), ("UBLKCP", r"\bUBLKCP\b"), ("UBLKPF", r"\bUBLKPF\b"), ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"),
         ("SYNCS", r"\bSYNCS\b"), ("UCGABAR", r"\bUCGABAR_\w+\b"), ("LDGSTS", r"\bLDGSTS\b"), ("UTCxMMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
         ("HMMA", r"\bHMMA\b"), ("RED", r"\bREDG?\b"), ("ATOM", r"\bATOM[SG]?\b"), ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("SHFL", r"\bSHFL\b"),
         ("BAR", r"\bBAR\b"), ("DFMA", r"\bDFMA\b"), ("DADD", r"\bDADD\b"), ("DMUL", r"\bDMUL\b")]
kern, rows, arch = None, {}, None
for ln in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", ln)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        rows[kern] = collections.Counter()
        rows[kern]["arch"] = arch
        continue
    if kern is None or "/*" not in ln:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if not m:
        continue
    ins = m.group(1)
    rows[kern]["instructions"] += 1
    for name, pat in WATCH:
        if re.search(pat, ins):
            rows[kern][name] += 1


def demangle(names):
    try:
        r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, r))
    except Exception:  # noqa: BLE001
        return {n: n for n in names}


dm = demangle(list(rows))
print(f"# SASS mnemonics per kernel of {os.path.basename(so)} (cuobjdump -sass; regenerate: python tools/sass_ops.py)")
print("# columns: instructions | " + " ".join(n for n, _ in WATCH))
tot = collections.Counter()
for k in sorted(rows, key=lambda k: -rows[k]["instructions"]):
    c = rows[k]
    name = re.sub(r"\(.*", "", dm[k].replace("(anonymous namespace)::", ""))
    name = re.sub(r"^void ", "", name)
    print(f"{name[:64]:64s} {c['arch']} {c['instructions']:6d} | " + " ".join(f"{n}={c[n]}" for n, _ in WATCH if c[n]))
    for n, _ in WATCH:
        tot[n] += c[n]
print("# totals: " + " ".join(f"{n}={tot[n]}" for n, _ in WATCH))
print("# expected for this workload: DMMA > 0 (FP64 tensor pipe in lu_gemm_kernel), UBLKCP + SYNCS > 0 (1-D bulk TMA + mbarrier in the assembly "
      "kernels), LDGSTS > 0 (cp.async ring of the GEMM), UTC*MMA = LDTM = STTM = HMMA = 0 (no f64 kind in tcgen05; no legacy half tensor path)")
