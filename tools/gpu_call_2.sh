#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_errors.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -80 | tee $O/pytest_gpu.txt
python tools/k1_chain_trace.py 125000 > $O/k1_chain_125k.txt 2>&1
CTK_K1_ILP=2 python tools/k1_chain_trace.py 125000 > $O/k1_chain_125k_ilp2.txt 2>&1
CTK_K1_FINISHER_SHARE=1.0 python tools/k1_chain_trace.py 125000 > $O/k1_chain_125k_share1.txt 2>&1
python tools/k1_chain_trace.py 1000000 > $O/k1_chain_1m.txt 2>&1
python tools/k1_trace.py 125000 > $O/k1_trace_125k.txt 2>&1
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --rollouts 125000 > $O/bench_125k.json 2> $O/bench_125k.err
echo done
