#!/bin/bash
# Round profile set (run under gpurun on one B200):  tools/profile_round.sh <tag>
#   1. the driver's bench line (bench.py defaults)                       -> gpurun_out/bench_<tag>.json
#   2. every launch of a short bench run with its device time (ncu)      -> gpurun_out/launches_<tag>.csv
#   3. one full ncu capture of each kernel of a cfg-5 step (2 waves)     -> gpurun_out/prof_<tag>.ncu-rep
# Each ncu run follows a plain run of the same command that exited 0.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-chi-sweep"
$CMD > $OUT/plain_launches_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
CMD2="python tools/run_one.py cfg5 f32 37888"
$CMD2 > $OUT/plain_full_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 8 -c 8 -o $OUT/prof_$TAG $CMD2 > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -2 $OUT/ncu_full_$TAG.log
