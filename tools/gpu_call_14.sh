#!/bin/bash
# balanced unit distribution in K1: whole GPU suite, default / C1 / logging bench lines against the round-1 tree, timeline
set -u
O=gpurun_out
mkdir -p $O
timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 | tee $O/pytest_gpu_14.txt
line() { python -c "
import json
d=json.loads(open('$O/$1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-16s ms/step %.4f kernel_ms %.4f frac %.4f pipelined %s p50 %.4f' % ('$1', d['ms_per_step'], r['kernel_ms'], r['frac'], (d.get('pipelined') or {}).get('ms_per_step'), d['e2e']['p50_step_latency_ms']))" | tee -a $O/call14.txt; }
for w in mppi_ode_1m mppi_ode_c1 mppi_ode_1m_log; do
  (cd _ab_r01 && python bench.py --workload $w > ../$O/c14_r01_$w.json 2>/dev/null); line c14_r01_$w
  python bench.py --workload $w > $O/c14_new_$w.json 2> $O/c14_new_$w.err; line c14_new_$w
done
python bench.py --rollouts 125000 > $O/c14_new_125k.json 2> $O/c14_new_125k.err; line c14_new_125k
(cd _ab_r01 && python bench.py --rollouts 125000 > ../$O/c14_r01_125k.json 2>/dev/null); line c14_r01_125k
python tools/k1_trace.py 1000000 1 2>&1 | tail -11 | tee $O/k1_trace_14.txt
echo done
