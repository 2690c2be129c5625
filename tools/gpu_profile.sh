#!/bin/bash
# Runs on the GPU box (under gpurun): the bench lines of every workload (both arms where a CPU arm exists), launch lists of the
# C2 and C5 bench commands, one `ncu --set full` capture of each kernel of the hot path.  Outputs land in gpurun_out/;
# tools/summarise_profiles.py turns them into the tracked summaries under profiles/ back here.
#   usage: tools/gpu_profile.sh <tag> [quick]      quick: launch lists + full captures only (no bench lines)
set -x
mkdir -p gpurun_out
TAG=${1:-r2b}
QUICK=${2:-}
if [ -z "$QUICK" ]; then
  python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  tail -c 400 gpurun_out/bench_${TAG}.json
  python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
  python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_c5_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
  PLSTVO_STREAM_SOLVE=0 python bench.py --steps 30 --warmup 5 --no-cpu --no-hbm-run > gpurun_out/bench_k2_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
  for W in c1 c3; do
    python bench.py --workload $W --steps 20 --warmup 3 --no-hbm-run > gpurun_out/bench_${W}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
    python bench.py --workload $W --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${W}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
  done
  for W in stereo stereo_track; do
    python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
  done
fi
# launch lists (short runs): per-launch device time under ncu is cold-cache and serialised -> compare SHARES
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm-run > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c5_${TAG}.csv \
    python bench.py --workload c5 --steps 3 --warmup 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_stereo_track_${TAG}.csv \
    python bench.py --workload stereo_track --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
# full captures, one launch each (-s skips the warm-up launches of that kernel); C2 command unless said otherwise
C2="python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm-run"
cap() {   # name regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_${name}_${TAG} -f "$@" \
      > gpurun_out/ncu_${name}_${TAG}.log 2>&1
}
cap tc_hamming tc_hamming 4 $C2
cap tc_expand tc_expand 4 $C2
cap tc_resolve tc_resolve 4 $C2
cap stream_prepare stream_prepare 4 $C2
cap gn_loop gn_loop_stream 9 $C2          # an odd index: a stage-2 launch (the longer one)
cap stream_outlier stream_outlier 4 $C2
cap stream_finalize stream_finalize 4 $C2
PLSTVO_STREAM_SOLVE=0 cap track_solve track_solve 4 $C2
cap gn_loop_c5 gn_loop_stream 5 python bench.py --workload c5 --steps 3 --warmup 2
cap stream_outlier_c5 stream_outlier 2 python bench.py --workload c5 --steps 3 --warmup 2
cap stream_prepare_c5 stream_prepare 2 python bench.py --workload c5 --steps 3 --warmup 2
cap tc_hamming_c5 tc_hamming 2 python bench.py --workload c5 --steps 3 --warmup 2
cap gn_eval gn_eval_stream 2 python bench.py --workload c5_sweep
# summarise here (ncu is on the box), keep two reports for reading at home, drop the rest (gpurun copies back <= 64 MiB)
PROF_OUT=gpurun_out/profiles python tools/summarise_profiles.py ${TAG} > gpurun_out/summarise_${TAG}.log 2>&1
mkdir -p gpurun_out/keep
mv gpurun_out/prof_tc_hamming_${TAG}.ncu-rep gpurun_out/prof_gn_loop_c5_${TAG}.ncu-rep gpurun_out/keep/ 2>/dev/null
rm -f gpurun_out/prof_*_${TAG}.ncu-rep
du -sh gpurun_out; ls gpurun_out/profiles | head -40
