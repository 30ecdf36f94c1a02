#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
V=$PWD/control_toolkit_b200/variants
run() {  # label, lib, env
  env $3 CTK_LIB=$2 python bench.py --steps 20 --warmup 5 > $O/exp_$1.json 2> $O/exp_$1.err
  python -c "
import json
d=json.loads(open('$O/exp_$1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s ms/step %.4f kernel_ms %.4f frac %.4f pipelined %s' % ('$1', d['ms_per_step'], r['kernel_ms'], r['frac'], (d.get('pipelined') or {}).get('ms_per_step')))" | tee -a $O/exp.txt
}
(cd _ab_3a9ba66 && python bench.py > ../$O/exp_3a9.json 2>/dev/null); python -c "
import json
d=json.loads(open('$O/exp_3a9.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s ms/step %.4f kernel_ms %.4f frac %.4f' % ('3a9ba66', d['ms_per_step'], r['kernel_ms'], r['frac']))" | tee -a $O/exp.txt
run regular "" "A=1"
run exp1 $V/libctk_exp1.so "CTK_K1_FINISHER_SHARE=1.0"
run exp2 $V/libctk_exp2.so "A=1"
run exp12 $V/libctk_exp12.so "A=1"
run exp15 $V/libctk_exp15.so "CTK_K1_FINISHER_SHARE=1.0"
run exp15k $V/libctk_exp15k.so "CTK_K1_FINISHER_SHARE=1.0"
echo done
