#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_final.log 2>&1; echo "pytest rc=$?" >> $O/pytest_final.log
tail -4 $O/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -6 | cut -c1-400
( time timeout 900 python bench.py ) > $O/bench_final.json 2> $O/bench_final.err; tail -4 $O/bench_final.err
python -c "
import json
d=json.loads(open('$O/bench_final.json').read().strip().splitlines()[-1])
print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['clocks'], 'launches', d['gpu_launches'])
r=d['roofline']; print(r['kernel'], r['achieved'], r['frac'], 'step frac sustained', r['step_frac_of_sustained_peak'])
for k,v in r['classes'].items(): print('  ',k, v['ms_per_step'], v['achieved'], v['unit'], v['frac'])
print('cpu', d['cpu_baseline']['value'], list(d['cpu_baseline']['legs'].keys()))
"
