#!/bin/bash
# Runs on the GPU box (under gpurun): bench (both arms), launch list, full ncu captures of K1 and K2.
# Outputs land in gpurun_out/ and are summarised into profiles/ by tools/summarise_profiles.py back here.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_ref_${TAG}.json
# launch list of the bench command (short run): per-launch device time, compare SHARES
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
# full capture of the dominant kernel (whole-batch launch, second occurrence) and of K2
ncu --set full --clock-control none --import-source on -k regex:hamming_knn2 -s 1 -c 1 -o gpurun_out/prof_k1_${TAG} -f \
    python bench.py --kernels-only > gpurun_out/ncu_k1_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:track_solve -s 1 -c 1 -o gpurun_out/prof_k2_${TAG} -f \
    python bench.py --kernels-only > gpurun_out/ncu_k2_${TAG}.log 2>&1
ls -la gpurun_out
