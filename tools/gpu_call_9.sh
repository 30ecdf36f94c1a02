#!/bin/bash
# (2 GPUs) pinning tests after restoring the compensated sums; K1 phase timeline round-1 tree vs current; sharded chain timeline at
# 125k rollouts per GPU (the per-GPU load of C5 on 8 GPUs) with one and two exchange hops; 2-GPU bench lines
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "production or top_m or fused or shard" 2>&1 | tail -8 | tee $O/pytest_gpu_9.txt
for n in 1000000 2000; do
  echo "== r01 tree N=$n" | tee -a $O/k1_trace_ab.txt; (cd _ab_r01 && python tools/k1_trace.py $n 1) 2>&1 | tail -11 | tee -a $O/k1_trace_ab.txt
  echo "== current N=$n" | tee -a $O/k1_trace_ab.txt; python tools/k1_trace.py $n 1 2>&1 | tail -11 | tee -a $O/k1_trace_ab.txt
done
python bench.py > $O/bench9_default.json 2> $O/bench9_default.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
for hops in 2 1; do
  echo "== chain timeline, 2 GPUs x 125k, hops=$hops" | tee -a $O/chain_trace_2gpu.txt
  CTK_EXCHANGE_HOPS=$hops $TR 29531 tools/k1_chain_trace_dist.py 250000 40 2>&1 | grep -v -i "warn\|^$" | tail -12 | tee -a $O/chain_trace_2gpu.txt
done
$TR 29533 bench.py --gpus 2 --rollouts 250000 > $O/bench9_2gpu_250k.json 2> $O/bench9_2gpu_250k.err
$TR 29534 bench.py --gpus 2 > $O/bench9_2gpu_1m.json 2> $O/bench9_2gpu_1m.err
$TR 29535 bench.py --gpus 2 --workload cem_ode_large > $O/bench9_2gpu_cem_large.json 2> $O/bench9_2gpu_cem_large.err
for f in $O/bench9_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('  ms/step', d['ms_per_step'], 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'), 'pipelined', (d.get('pipelined') or {}).get('ms_per_step'), 'e2e p50', d['e2e'].get('p50_step_latency_ms'))
"; done
echo done
