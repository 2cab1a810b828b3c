#!/bin/bash
# One GPU session for the ComParE_2016 full-LLD workload (BASELINE configs[3]): GPU test suite, bench line,
# per-kernel launch list and ncu --set full captures of the pitch-chain kernels.  Writes to gpurun_out/.
#   usage (under gpurun): bash scripts/gpu_compare16.sh [n_utt]
N=${1:-10000}
C=oracle/_ref/config/compare16/ComParE_2016.conf
export OSM_BENCH_OPTS=lldcsvoutput=x.csv
export OSM_BENCH_N_UTT=$N
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench.py --workload compare16"; timeout 400 python bench.py --workload compare16 --steps 5 --warmup 3 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; tail -c 1500 gpurun_out/c16_bench.json; tail -3 gpurun_out/c16_bench.err
echo "== kernel times"; timeout 400 bash scripts/kernel_times.sh $C $N 48000 > gpurun_out/c16_kernel_times.txt 2>&1; cat gpurun_out/c16_kernel_times.txt
for k in ${KERNELS-jitter_kernel shs_kernel viterbi_kernel seq_post_kernel}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/c16_$k python scripts/bench_general.py $C $N 48000 > /dev/null 2>&1
  python scripts/ncu_summary.py gpurun_out/c16_$k.ncu-rep > gpurun_out/c16_${k}_summary.txt 2>&1
  echo "== $k"; head -40 gpurun_out/c16_${k}_summary.txt
done
