#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_final4.log 2>&1; echo "pytest rc=$?" >> $O/pytest_final4.log
tail -4 $O/pytest_final4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/ab_step.sh tools/ab/libubssfp_prev.so 2 2>&1 | tee $O/r02k_step_ab.txt
( time timeout 900 python bench.py ) > $O/bench_final4.json 2> $O/bench_final4.err; tail -4 $O/bench_final4.err
python -c "
import json
d=json.loads(open('$O/bench_final4.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['clocks'], 'launches', d['gpu_launches'])
r=d['roofline']; print(r['kernel'], r['achieved'], r['frac'])
for k,v in list(r['classes'].items())[:8]: print('  ',k, v['ms_per_step'], v['achieved'], v['unit'], v['frac'])
print(d['secondary'])
"
