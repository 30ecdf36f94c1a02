#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_errors.txt
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv | tee $O/gpu.txt
python -c "import os; print('cpus', os.cpu_count())" | tee -a $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 | tee $O/pytest_gpu.txt
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --workload cem_ode_c2 > $O/bench_cem_c2_coop.json 2> $O/bench_cem_c2_coop.err
CTK_CEM_NO_COOP=1 python bench.py --workload cem_ode_c2 > $O/bench_cem_c2_nocoop.json 2> $O/bench_cem_c2_nocoop.err
python tools/k1_trace.py 1000000 > $O/k1_trace_1m.txt 2>&1
python tools/k1_trace.py 125000 > $O/k1_trace_125k.txt 2>&1
python tools/k1_trace.py 125000 0 > $O/k1_trace_125k_noflush.txt 2>&1
echo done
