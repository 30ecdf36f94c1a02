#!/bin/bash
# K1: ncu --set full of the round-1 tree's kernel and the current one on the same box (instruction counts per code region), plus the
# current build without programmatic dependent launch
set -u
O=gpurun_out
P=/tmp/ctk_prof
mkdir -p $O $P
CTK_NO_PDL=1 python bench.py > $O/var_nopdl.json 2> $O/var_nopdl.err
CTK_NO_HANDOVER=1 python bench.py > $O/var_nohandover.json 2> $O/var_nohandover.err
for f in nopdl nohandover; do python -c "
import json
d=json.loads(open('$O/var_$f.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s ms/step %.4f kernel_ms %.4f frac %.4f' % ('$f', d['ms_per_step'], r['kernel_ms'], r['frac']))" | tee -a $O/variants.txt; done
(cd _ab_r01 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:mppi_ode_kernel -s 4 -c 1 -o $P/prof_k1_r01tree -f python bench.py --steps 3 --warmup 3 > ../$O/ncu_k1_r01tree.log 2>&1)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mppi_ode_kernel -s 4 -c 1 -o $P/prof_k1_cur -f python bench.py --steps 3 --warmup 3 > $O/ncu_k1_cur.log 2>&1
python tools/ncu_summary.py $P/prof_k1_r01tree.ncu-rep > $O/prof_k1_r01tree_summary.txt 2>&1
python tools/ncu_summary.py $P/prof_k1_cur.ncu-rep > $O/prof_k1_r02b_summary.txt 2>&1
ncu -i $P/prof_k1_r01tree.ncu-rep --page raw --csv > $O/prof_k1_r01tree_raw.csv 2>/dev/null
ncu -i $P/prof_k1_cur.ncu-rep --page raw --csv > $O/prof_k1_cur_raw.csv 2>/dev/null
echo done
