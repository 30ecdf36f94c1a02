#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "shard or exchange or fused or cem" 2>&1 | tail -15 | tee $O/pytest_gpu_shard.txt
python bench.py --workload cem_ode_large --steps 10 --warmup 3 > $O/cem_large_1.json 2> $O/cem_large_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --workload cem_ode_large --gpus 2 --steps 10 --warmup 3 > $O/cem_large_2.json 2> $O/cem_large_2.err
tail -5 $O/cem_large_1.err $O/cem_large_2.err
echo done
