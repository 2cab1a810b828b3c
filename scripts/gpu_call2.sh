#!/bin/bash
# round 2, GPU call 2: strict GeMAPS-family parity through the reference-order FFT, plugin through the reference's SMILExtract
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/c2_gpu_suite.txt
timeout 300 python scripts/parity_report.py gpurun_out/c2_parity_report.md 2>&1 | tail -40
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_zzz_formant_gpu.py tests/test_plugin_gpu.py -x -q > gpurun_out/c2_memcheck.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/c2_memcheck.txt
tail -5 gpurun_out/c2_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_zzz_formant_gpu.py -x -q > gpurun_out/c2_racecheck.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/c2_racecheck.txt
tail -5 gpurun_out/c2_racecheck.txt
timeout 600 python bench.py --workload egemaps --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c2_egemaps_bench.json
OSM_BENCH_N_UTT=2000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lld_kernel|shs_|viterbi|jitter|seq_post|post_kernel|formant|harmonics|spectral|energy|mzcr|rasta|plp_tail|intensity|acf_pitch|pitch_smooth' -c 200 --csv --log-file gpurun_out/c2_egemaps_launches.csv \
  python bench.py --workload egemaps --steps 2 --warmup 3 > gpurun_out/c2_egemaps_ncu.log 2>&1
OSM_BENCH_N_UTT=500 timeout 900 ncu --set full --clock-control none --import-source on -k regex:formant_kernel -c 1 -o gpurun_out/c2_formant_kernel \
    python bench.py --workload egemaps --steps 1 --warmup 3 > gpurun_out/c2_formant_kernel_ncu.log 2>&1
ls -la gpurun_out | tail -12
