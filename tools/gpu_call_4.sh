#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_errors.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu_full.txt 2>&1
tail -8 $O/pytest_gpu_full.txt
python tools/k1_chain_trace.py 125000 > $O/k1_chain_125k.txt 2>&1
python tools/k1_chain_trace.py 1000000 > $O/k1_chain_1m.txt 2>&1
python bench.py --steps 20 --warmup 5 > $O/scale_1.json 2> $O/scale_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/scale_2.json 2> $O/scale_2.err
tail -3 $O/scale_2.err
echo done
