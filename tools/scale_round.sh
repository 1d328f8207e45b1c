#!/bin/bash
# Multi-GPU set of a round (run under `gpurun --gpus 8` on one box):  tools/scale_round.sh <tag>
#   cfg 5 weak at N = 8; cfg 5 strong (global 131072) at N = 2, 4, 8; cfg 4 strong (global 65536) at N = 1, 8 (f32 and f64)
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-chi-sweep --no-parity"
python $B --gpus 1 > $OUT/scale_${TAG}_weak_n1.json 2> $OUT/scale_${TAG}_weak_n1.err
$TR --nproc-per-node 8 --master-port 29501 $B --gpus 8 > $OUT/scale_${TAG}_weak_n8.json 2> $OUT/scale_${TAG}_weak_n8.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port 2951$N $B --gpus $N --scaling strong > $OUT/scale_${TAG}_strong_n$N.json 2> $OUT/scale_${TAG}_strong_n$N.err
done
for DT in f32 f64; do
  python tools/scale_cfg4.py --dtype $DT > $OUT/scale_${TAG}_cfg4_${DT}_n1.json 2> $OUT/scale_${TAG}_cfg4_${DT}_n1.err
  $TR --nproc-per-node 8 --master-port 29531 tools/scale_cfg4.py --dtype $DT > $OUT/scale_${TAG}_cfg4_${DT}_n8.json 2> $OUT/scale_${TAG}_cfg4_${DT}_n8.err
  $TR --nproc-per-node 8 --master-port 29532 tools/scale_cfg4.py --dtype $DT --no-overlap > $OUT/scale_${TAG}_cfg4_${DT}_n8_noov.json 2> $OUT/scale_${TAG}_cfg4_${DT}_n8_noov.err
done
grep -h '^{' $OUT/scale_${TAG}_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'workload' in d: print('cfg4', d['dtype'], d['n_gpus'], d['allreduce'], round(d['ms_per_step'], 3), 'ms', round(d['samples_per_s']), 'samples/s')
    else: print('cfg5', d['scaling'], d['n_gpus'], round(d['ms_per_step'], 2), 'ms', round(d['value']), 'samples/s', 'e2e', round(d['e2e']['value']))
"
