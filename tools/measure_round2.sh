#!/bin/bash
# One GPU-box pass (round 2) that refreshes every single-GPU number DESIGN.md quotes: smoke, parity suite, bench lines of all
# workloads, latency table, ncu launch lists and ncu --set full captures of the kernels changed this round.
#   gpurun --timeout 1700 -- 'bash tools/measure_round2.sh'
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_errors.txt
P=/tmp/ctk_prof   # .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): only their text summaries travel back
mkdir -p $P
python __graft_entry__.py --smoke 2>&1 | tail -6 | tee $O/smoke.txt
timeout 1100 python -m pytest tests -m gpu -q --timeout 600 ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -25 | tee $O/pytest_gpu.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err
for w in mppi_ode_c1 cem_ode_c2 rpgd_ode_c3 mppi_ode_1m_log cem_ode_large; do
  python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err
done
for e in tcgen05 tcgen05_bf16 tcgen05_fast; do
  python bench.py --workload mppi_mlp_c4 --mlp-engine $e --steps 10 --warmup 3 > $O/bench_mppi_mlp_c4_$e.json 2> $O/bench_mppi_mlp_c4_$e.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/latency_table.py 200 > $O/latency_table.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_bench_steps3.csv python bench.py --steps 3 --warmup 3 > $O/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_mlp_c4.csv python bench.py --workload mppi_mlp_c4 --mlp-engine tcgen05_fast --steps 3 --warmup 3 > $O/ncu_launch_mlp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_cem_large.csv python bench.py --workload cem_ode_large --steps 3 --warmup 3 > $O/ncu_launch_ceml.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mppi_ode_kernel -s 4 -c 1 -o $P/prof_k1_r02 -f python bench.py --steps 3 --warmup 3 > $O/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mppi_rollout_kernel -s 4 -c 1 -o $P/prof_mlp_tc_r02_fast -f python bench.py --workload mppi_mlp_c4 --mlp-engine tcgen05_fast --steps 3 --warmup 3 > $O/ncu_mlp_fast.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mppi_rollout_kernel -s 4 -c 1 -o $P/prof_mlp_tc_r02_exact -f python bench.py --workload mppi_mlp_c4 --mlp-engine tcgen05 --steps 3 --warmup 3 > $O/ncu_mlp_exact.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cem_ode_kernel -s 4 -c 1 -o $P/prof_cem_large_r02 -f python bench.py --workload cem_ode_large --steps 3 --warmup 3 > $O/ncu_ceml.log 2>&1
for r in prof_k1_r02 prof_mlp_tc_r02_fast prof_mlp_tc_r02_exact prof_cem_large_r02; do
  python tools/ncu_summary.py $P/$r.ncu-rep > $O/${r}_summary.txt 2>&1
done
du -sh $O
echo done
