#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_errors.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu_full.txt 2>&1
tail -30 $O/pytest_gpu_full.txt
python tools/k1_chain_trace.py 125000 > $O/k1_chain_125k.txt 2>&1
python tools/k1_chain_trace.py 1000000 > $O/k1_chain_1m.txt 2>&1
python tools/k1_trace.py 125000 > $O/k1_trace_125k.txt 2>&1; CTK_NO_HANDOVER=1 python tools/k1_chain_trace.py 125000 > $O/k1_chain_125k_nohandover.txt 2>&1
python tools/k1_trace.py 1000000 > $O/k1_trace_1m.txt 2>&1
python bench.py > $O/bench_default.json 2> $O/bench_default.err
for e in ; do
  python bench.py --workload mppi_mlp_c4 --mlp-engine $e --steps 10 --warmup 3 > $O/bench_mlp_$e.json 2> $O/bench_mlp_$e.err
done
echo done
