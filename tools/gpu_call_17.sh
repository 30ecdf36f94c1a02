#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -x 2>&1 | tail -6 | tee $O/pytest_gpu_17.txt
line() { python -c "
import json
d=json.loads(open('$O/$1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-16s ms/step %.4f kernel_ms %.4f frac %.4f pipelined %s p50 %.4f' % ('$1', d['ms_per_step'], r['kernel_ms'], r['frac'], (d.get('pipelined') or {}).get('ms_per_step'), d['e2e']['p50_step_latency_ms']))" | tee -a $O/call17.txt; }
python bench.py --workload mppi_ode_c1 > $O/c17_c1.json 2> $O/c17_c1.err; line c17_c1
python bench.py > $O/c17_default.json 2> $O/c17_default.err; line c17_default
echo done
