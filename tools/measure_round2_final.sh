#!/bin/bash
# Final single-GPU pass of round 2 (what DESIGN.md quotes): smoke, parity suite, bench lines, latency table, ncu launch list and
# ncu --set full summaries of K1 and the exact tcgen05 engine.  The .ncu-rep files stay on the box.
set -u
O=gpurun_out
P=/tmp/ctk_prof
mkdir -p $O $P
rm -f $O/parity_errors.txt
python __graft_entry__.py --smoke 2>&1 | tail -7 | tee $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 | tee $O/pytest_gpu.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --rollouts 125000 > $O/bench_125k_share075.json 2> /dev/null
CTK_K1_FINISHER_SHARE=1.0 python bench.py --rollouts 125000 > $O/bench_125k_share100.json 2> /dev/null
for w in mppi_ode_c1 cem_ode_c2 rpgd_ode_c3 mppi_ode_1m_log cem_ode_large; do
  python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err
done
for e in tcgen05 tcgen05_bf16 tcgen05_fast; do
  python bench.py --workload mppi_mlp_c4 --mlp-engine $e --steps 10 --warmup 3 > $O/bench_mppi_mlp_c4_$e.json 2> $O/bench_mppi_mlp_c4_$e.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/latency_table.py 200 > $O/latency_table.txt 2>&1
python tools/k1_trace.py 2000 1 2>&1 | tail -11 > $O/k1_trace_c1.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_bench_steps3.csv python bench.py --steps 3 --warmup 3 > $O/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mppi_ode_kernel -s 4 -c 1 -o $P/prof_k1_r02 -f python bench.py --steps 3 --warmup 3 > $O/ncu_k1.log 2>&1
python tools/ncu_summary.py $P/prof_k1_r02.ncu-rep > $O/prof_k1_r02_summary.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mppi_rollout_kernel -s 4 -c 1 -o $P/prof_mlp_exact -f python bench.py --workload mppi_mlp_c4 --mlp-engine tcgen05 --steps 3 --warmup 3 > $O/ncu_mlp_exact.log 2>&1
python tools/ncu_summary.py $P/prof_mlp_exact.ncu-rep > $O/prof_mlp_tc_r02_exact_summary.txt 2>&1
for f in $O/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('%-44s ms/step %.4f kernel_ms %s frac %s pipelined %s p50 %s' % ('$f', d['ms_per_step'], r.get('kernel_ms'), r.get('frac'), (d.get('pipelined') or {}).get('ms_per_step'), d['e2e'].get('p50_step_latency_ms')))"; done
du -sh $O
echo done
