#!/bin/bash
# First GPU call of round 2 (one B200, no ncu in this call): everything that was written after round 1's GPU budget was spent runs here for the
# first time.  Usage:  gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# Outputs (gpurun_out/r2a_*): the GPU test log, the full assembly sweep, the shape-sensitivity / LU-knob tools and one bench line.
# Nothing stops on a failure: every step writes its own log and exit code, so one call answers all open questions.
set -u
mkdir -p gpurun_out
run() {  # run <tag> <timeout s> <cmd...>
  local tag=$1 to=$2; shift 2
  timeout "$to" "$@" > "gpurun_out/r2a_${tag}.log" 2>&1
  echo "$tag exit $?" | tee -a gpurun_out/r2a_summary.txt
}
: > gpurun_out/r2a_summary.txt
python __graft_entry__.py > gpurun_out/r2a_build.log 2>&1; echo "build exit $?" | tee -a gpurun_out/r2a_summary.txt
# the verified tests first (-x), then every never-run file on its own so that one failure does not hide the others
run tests_verified 900 python -m pytest tests -m gpu -x -q --deselect tests/test_zw_tutorial_01_gpu.py --deselect tests/test_zx_forcing.py \
    --deselect tests/test_zy_shape_sensitivity_gpu.py --deselect tests/test_zz_assembly_variants_gpu.py --deselect tests/test_zzz_release_gpu.py
for f in zw_tutorial_01_gpu zx_forcing zy_shape_sensitivity_gpu zz_assembly_variants_gpu zzz_release_gpu; do
  run "test_$f" 600 python -m pytest "tests/test_$f.py" -m gpu -q
done
run sweep_assembly 400 python tools/sweep_assembly.py 64 quad 5
run lu_knobs 300 python tools/bench_lu_knobs.py 20 20 300 quad 2
run shape_sens 300 python tools/bench_shape_sens.py 10 10 150 5
run bench 900 python bench.py
tail -n 3 gpurun_out/r2a_tests_verified.log gpurun_out/r2a_test_*.log
grep -h '^{' gpurun_out/r2a_sweep_assembly.log | tail -n 1 | cut -c1-1500
grep -h '^{' gpurun_out/r2a_lu_knobs.log | tail -n 1 | cut -c1-2500
cat gpurun_out/r2a_summary.txt
