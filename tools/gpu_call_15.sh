#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 | tee $O/pytest_gpu_15.txt
line() { python -c "
import json
d=json.loads(open('$O/$1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-16s ms/step %.4f kernel_ms %.4f frac %.4f pipelined %s p50 %.4f' % ('$1', d['ms_per_step'], r['kernel_ms'], r['frac'], (d.get('pipelined') or {}).get('ms_per_step'), d['e2e']['p50_step_latency_ms']))" | tee -a $O/call15.txt; }
for w in mppi_ode_1m mppi_ode_1m_log mppi_ode_c1; do
  python bench.py --workload $w > $O/c15_new_$w.json 2> $O/c15_new_$w.err; line c15_new_$w
done
CTK_NO_PDL=1 python bench.py --workload mppi_ode_c1 > $O/c15_nopdl_c1.json 2> $O/c15_nopdl_c1.err; line c15_nopdl_c1
echo done
