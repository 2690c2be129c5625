#!/bin/bash
# Runs on the GPU box (under gpurun): bench (both arms + C5 roofline run), launch list, full ncu captures of K1, K2 and the
# streamed GN evaluation.  Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into profiles/ back here.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
python bench.py --workload c5 > gpurun_out/bench_c5_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
for W in c1 c3; do
  python bench.py --workload $W --steps 20 --warmup 3 --no-hbm-run > gpurun_out/bench_${W}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
  python bench.py --workload $W --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${W}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
done
tail -c 900 gpurun_out/bench_c5_${TAG}.json
for W in stereo stereo_track; do
  python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_stereo_track_${TAG}.csv \
    python bench.py --workload stereo_track --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
# launch list of the bench command (short run): per-launch device time, compare SHARES
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm-run > gpurun_out/ncu_launches_${TAG}.log 2>&1
# full captures: --kernels-only launches K1,K2 over the WHOLE batch three times; take the second occurrence
ncu --set full --clock-control none --import-source on -k regex:hamming_knn2 -s 1 -c 1 -o gpurun_out/prof_k1_${TAG} -f \
    python bench.py --kernels-only > gpurun_out/ncu_k1_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:track_solve -s 1 -c 1 -o gpurun_out/prof_k2_${TAG} -f \
    python bench.py --kernels-only > gpurun_out/ncu_k2_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_eval_stream -s 2 -c 1 -o gpurun_out/prof_gn_${TAG} -f \
    python bench.py --workload c5 > gpurun_out/ncu_gn_${TAG}.log 2>&1
ls -la gpurun_out | tail -20
