#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29601 tools/k1_chain_trace_dist.py 1000000 > $O/chain8_hops1.txt 2>&1
CTK_EXCHANGE_HOPS=2 $TR --master-port 29602 tools/k1_chain_trace_dist.py 1000000 > $O/chain8_hops2.txt 2>&1
CTK_EXCHANGE_HOPS=2 $TR --master-port 29603 bench.py --gpus 8 --steps 20 --warmup 5 > $O/scale_8_hops2.json 2> $O/scale_8_hops2.err
grep -v "^\*\|OMP\|NCCL" $O/chain8_hops1.txt $O/chain8_hops2.txt
echo done
