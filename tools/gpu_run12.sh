#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_pointwise_gpu.py tests/test_model_gpu.py -x -q -m gpu > $O/pytest12.log 2>&1; tail -2 $O/pytest12.log
UB_WIDEN_BF16=0 python tools/bench_bf16_input.py > $O/bf16_in_direct3.txt 2>&1; cat $O/bf16_in_direct3.txt
( time python bench.py ) > $O/bench12.json 2> $O/bench12.err
tail -4 $O/bench12.err
python -c "
import json
d=json.loads(open('$O/bench12.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['clocks'])
print('cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
print(d['secondary'])
"
