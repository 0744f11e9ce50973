#!/bin/bash
# Sweep of the rows per launch of the tensor-core kernels (train and ranking eval): smaller sub-batches keep the
# intermediates of a sub-batch (H1, H2, dZ) in the 126 MB L2 between the kernels that produce and consume them.
# usage (on the GPU box): bash tools/sub_batch_sweep.sh > gpurun_out/sub_batch_sweep.log
for pair in "655360 1048576" "327680 262144" "163840 131072" "81920 65536" "40960 32768"; do
  set -- $pair
  MR_TC_SUB_BATCH_ROWS=$1 MR_EVAL_SUB_BATCH_ROWS=$2 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('train_rows=$1 eval_rows=$2 ms_per_step=%.3f eval_ms=%.3f users/s=%.3g phases=%s eval_phases=%s' % (d['ms_per_step'], d['eval']['ms'], d['eval']['value'], {k: round(v,3) for k,v in d['phase_ms_per_step'].items()}, {k: round(v,3) for k,v in d['eval']['phase_ms'].items()}))"
done
