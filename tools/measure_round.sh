#!/bin/bash
# One GPU-box pass that refreshes every number DESIGN.md quotes: parity suite, bench lines of all workloads, latency table,
# ncu launch list of the default bench and ncu --set full captures of the kernels changed this round.
#   gpurun --timeout 1500 -- 'bash tools/measure_round.sh'
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_errors.txt
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_gpu.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err
for w in mppi_ode_c1 cem_ode_c2 rpgd_ode_c3 mppi_ode_1m_log; do
  python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err
done
python bench.py --workload mppi_mlp_c4 --steps 10 --warmup 3 > $O/bench_mppi_mlp_c4.json 2> $O/bench_mppi_mlp_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/latency_table.py 200 > $O/latency_table.txt 2>&1
python tools/step_breakdown.py 300 > $O/step_breakdown.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r01c.csv python bench.py --steps 3 --warmup 3 > $O/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cem_c2_r01c.csv python bench.py --workload cem_ode_c2 --steps 3 --warmup 3 > $O/ncu_launch_cem.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_rpgd_c3_r01c.csv python bench.py --workload rpgd_ode_c3 --steps 3 --warmup 3 > $O/ncu_launch_rpgd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cem_tick_kernel -s 4 -c 1 -o $O/prof_cem_tick_r01 -f python bench.py --workload cem_ode_c2 --steps 3 --warmup 3 > $O/ncu_cem.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rpgd_grad -s 4 -c 1 -o $O/prof_rpgd_coef_r01 -f python bench.py --workload rpgd_ode_c3 --steps 3 --warmup 3 > $O/ncu_rpgd.log 2>&1
echo done
