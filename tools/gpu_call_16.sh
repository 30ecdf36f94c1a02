#!/bin/bash
# 8 GPUs: C5 strong-scaling line, sharded-CEM line, chain timelines with two and one exchange hops
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$TR 29541 bench.py --gpus 8 > $O/scale_8_r02.json 2> $O/scale_8_r02.err
for hops in 2 1; do
  echo "== chain timeline, 8 GPUs x 125k, hops=$hops" | tee -a $O/chain_trace_8gpu.txt
  CTK_EXCHANGE_HOPS=$hops $TR 29542 tools/k1_chain_trace_dist.py 1000000 40 2>&1 | grep -v -i "warn\|^$\|\*\*\*\|OMP_NUM" | tail -9 | tee -a $O/chain_trace_8gpu.txt
done
CTK_EXCHANGE_HOPS=1 $TR 29543 bench.py --gpus 8 > $O/scale_8_r02_hops1.json 2> $O/scale_8_r02_hops1.err
$TR 29544 bench.py --gpus 8 --workload cem_ode_large > $O/scale_8_cem_large_r02.json 2> $O/scale_8_cem_large_r02.err
$TR 29545 bench.py --gpus 8 --workload mppi_mlp_c4 --mlp-engine tcgen05 --steps 10 --warmup 3 > $O/scale_8_mlp_c4_r02.json 2> $O/scale_8_mlp_c4_r02.err
for f in scale_8_r02 scale_8_r02_hops1 scale_8_cem_large_r02 scale_8_mlp_c4_r02; do python -c "
import json
d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1])
print('%-24s value %.4g ms/step %.4f pipelined %s e2e %.4g p50 %.4f' % ('$f', d['value'], d['ms_per_step'], (d.get('pipelined') or {}).get('ms_per_step'), d['e2e']['value'], d['e2e']['p50_step_latency_ms']))"; done
echo done
