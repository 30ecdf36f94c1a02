#!/bin/bash
# after the K1 register fix (launch bounds per instantiation, plain two-level cost sum) and the owner + helper exact tcgen05 engine:
# whole GPU suite, A/B against the round-1 tree on the same box, exact-engine bench + ncu summary
set -u
O=gpurun_out
P=/tmp/ctk_prof
mkdir -p $O $P
timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee $O/pytest_gpu_8.txt
python bench.py --workload mppi_mlp_c4 --mlp-engine tcgen05 --steps 10 --warmup 3 > $O/bench_mlp_exact_v3.json 2> $O/bench_mlp_exact_v3.err
for i in 1 2; do
  (cd _ab_r01 && python bench.py > ../$O/ab8_r01_default_$i.json 2> ../$O/ab8_r01_default_$i.err)
  python bench.py > $O/ab8_r02_default_$i.json 2> $O/ab8_r02_default_$i.err
done
(cd _ab_r01 && python bench.py --workload mppi_ode_1m_log > ../$O/ab8_r01_log.json 2> ../$O/ab8_r01_log.err)
python bench.py --workload mppi_ode_1m_log > $O/ab8_r02_log.json 2> $O/ab8_r02_log.err
(cd _ab_r01 && python bench.py --workload mppi_ode_c1 > ../$O/ab8_r01_c1.json 2> ../$O/ab8_r01_c1.err)
python bench.py --workload mppi_ode_c1 > $O/ab8_r02_c1.json 2> $O/ab8_r02_c1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mppi_rollout_kernel -s 4 -c 1 -o $P/prof_mlp_tc_r02_exact_v3 -f python bench.py --workload mppi_mlp_c4 --mlp-engine tcgen05 --steps 3 --warmup 3 > $O/ncu_mlp_exact_v3.log 2>&1
python tools/ncu_summary.py $P/prof_mlp_tc_r02_exact_v3.ncu-rep > $O/prof_mlp_tc_r02_exact_v3_summary.txt 2>&1
for f in $O/bench_mlp_exact_v3.json $O/ab8_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('  ms/step', d['ms_per_step'], 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'), 'pipelined', (d.get('pipelined') or {}).get('ms_per_step'))
"; done
echo done
