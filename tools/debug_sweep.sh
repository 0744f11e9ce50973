#!/bin/bash
# Diagnostics: phase times of the ML-20M step under the MR_TC_DEBUG switches of the tcgen05 kernels
# (1 skip weight copies, 2 skip A/Z global loads, 4 skip convert+st.shared, 8 skip epilogue global stores,
# 32 skip the epilogue).  Results are wrong by construction; only the phase times matter.
for d in ${@:-0 2 4 6 1 8 32 47}; do
  MR_TC_DEBUG=$d python bench.py --lean --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); p=l['phase_ms_per_step']
print('debug=%3s step %.3f ms | ' % ('$d', l['ms_per_step']) + ' '.join('%s %.3f' % (k.replace('tc_dense_','').replace('tc_',''), v) for k,v in p.items()))"
done
