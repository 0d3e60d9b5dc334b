#!/bin/bash
# Profiling call of round 2 (one B200): launch list of the bench command and one `ncu --set full` capture each of the assembly kernel and the
# LU GEMM, every ncu run directly after the same command has exited 0 without ncu (B200_PROFILING.md).  Usage:
#   gpurun --timeout 1500 -- 'bash tools/r2_profile_call.sh'          (set WAE_* switches in front to profile a non-default configuration)
# Outputs: gpurun_out/r2p_*; summarise with tools/launch_summary.py / tools/ncu_summary.py / tools/ncu_phase_shares.py into profiles/r02_*.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --skip-extras"
A="python tools/prof_star.py 64 quad 3"
L="python tools/bench_lu.py 20 20 300 quad 1"
$B > gpurun_out/r2p_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r2p_launches_bench.csv $B > gpurun_out/r2p_bench_ncu.log 2>&1
echo "launch list exit $?"
$A > gpurun_out/r2p_asm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assemble_tet_stars -s 2 -c 1 -f -o gpurun_out/r2p_asm $A > gpurun_out/r2p_asm_ncu.log 2>&1
echo "assembly capture exit $?"
WAE_NFACTOR=1 WAE_PROBE_ONLY=1 $L > gpurun_out/r2p_lu_plain.log 2>&1 &&
WAE_LU_GROUPS=1 WAE_LU_PROFILE_DEPTH=4 WAE_NFACTOR=1 WAE_PROBE_ONLY=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lu_gemm_kernel -c 1 -f -o gpurun_out/r2p_gemm $L > gpurun_out/r2p_lu_ncu.log 2>&1
echo "gemm capture (Schur complement of depth 4, one launch over all 16 fronts) exit $?"
# three solve kernels of the first forward / backward sweep after the factorisation (one launch each: ncu saves and restores the 23 GB
# factor around every replay, ~7 s per captured launch -- a capture of all 200 launches of a sweep pair took 23 minutes and 220 MB):
# the fused deep-level kernel on the 7844-front level, the forward update of depth 4 (window 0) and the backward update of depth 4 (window 0)
cap() {  # name, kernel regex, launches to skip
  WAE_NFACTOR=1 WAE_PROBE_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/r2p_$1 $L > gpurun_out/r2p_$1_ncu.log 2>&1
  echo "$1 capture exit $?"
}
cap solve_fused lu_fwd_fused_kernel 1
cap solve_fwd_update lu_fwd_update_kernel 9
cap solve_bwd_update lu_bwd_update2_kernel 34
# gpurun copies gpurun_out/ back only if it stays under 64 MiB: drop anything large rather than lose everything
find gpurun_out -size +30M -print -delete
du -sh gpurun_out
ls -la gpurun_out/r2p_* | head -20
