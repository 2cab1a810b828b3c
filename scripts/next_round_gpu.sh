#!/bin/bash
# Runbook for the first GPU call of the next round (everything written after this round's GPU budget was spent).
#   gpurun --timeout 1500 -- 'bash scripts/next_round_gpu.sh'
# Order: cheapest / most informative first; every step has its own timeout so a hang cannot eat the call.
set -x
mkdir -p gpurun_out
# 0. the validated suite must still be green after the graph / ABI changes (frame reader header, selector scopes, ABI 3)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/nr_gpu_suite.txt
# 1. the LLD sink test that was committed after the last GPU run of this round (part of the suite above; kept for the log)
# 2. the opt-in tests: selector grouping, formant kernel vs its host build, shipped GeMAPS / eGeMAPS end to end
export OSM_B200_RUN_UNVERIFIED=1
timeout 900 python -m pytest tests/test_zz_lld_sinks_gpu.py tests/test_zz_select_gpu.py tests/test_zzz_formant_gpu.py tests/test_zzz_gemaps_gpu.py -q 2>&1 | tail -30 | tee gpurun_out/nr_unverified.txt
# 2b. per-column parity table (eGeMAPS / GeMAPS vs the reference's rows)
timeout 300 python scripts/parity_report.py gpurun_out/nr_parity_report.md 2>&1 | tail -30
# 3. memcheck over the new kernels (small inputs)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_zzz_formant_gpu.py tests/test_zzz_gemaps_gpu.py -x -q \
  > gpurun_out/nr_memcheck.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/nr_memcheck.txt
tail -8 gpurun_out/nr_memcheck.txt
# 4. racecheck over the pitch chain + new kernels (owed since the previous round)
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_pitch_gpu.py -x -q -k "variant or batch" \
  > gpurun_out/nr_racecheck.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/nr_racecheck.txt
tail -8 gpurun_out/nr_racecheck.txt
# 5. timing: eGeMAPS workload (BASELINE configs[2]) + per-kernel launch list
timeout 600 python bench.py --workload egemaps --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/nr_egemaps_bench.json
OSM_BENCH_N_UTT=2000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/nr_egemaps_launches.csv \
  python bench.py --workload egemaps --steps 2 --warmup 3 > gpurun_out/nr_egemaps_ncu.log 2>&1
# 6. the default bench line, unchanged path
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/nr_bench.json
# 7. one ncu --set full capture each of the two new kernels (small batch), summaries -> profiles/ by scripts/ncu_summary.py
for K in formant_kernel harmonics_kernel; do
  OSM_BENCH_N_UTT=500 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/nr_$K \
    python bench.py --workload egemaps --steps 1 --warmup 3 > gpurun_out/nr_${K}_ncu.log 2>&1
done
ls -la gpurun_out | tail -20
