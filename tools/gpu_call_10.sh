#!/bin/bash
# K1 variants (tools/build_variants.sh) against the regular build and the round-1 tree on one box
set -u
O=gpurun_out
mkdir -p $O
V=$PWD/control_toolkit_b200/variants
run() {  # label, lib ('' = regular)
  CTK_LIB=$2 python bench.py --steps 20 --warmup 5 > $O/var_$1.json 2> $O/var_$1.err
  python -c "
import json
d=json.loads(open('$O/var_$1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s ms/step %.4f kernel_ms %.4f frac %.4f pipelined %.4f' % ('$1', d['ms_per_step'], r['kernel_ms'], r['frac'], d['pipelined']['ms_per_step']))" | tee -a $O/variants.txt
}
(cd _ab_r01 && python bench.py > ../$O/var_r01.json 2>/dev/null); python -c "
import json
d=json.loads(open('$O/var_r01.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s ms/step %.4f kernel_ms %.4f frac %.4f' % ('r01', d['ms_per_step'], r['kernel_ms'], r['frac']))" | tee -a $O/variants.txt
run regular ""
for v in kahan1024 kahan960 plain1024 plain960 dsum896 dsum1024; do run $v $V/libctk_$v.so; done
run regular2 ""
CTK_LIB=$V/libctk_dsum896.so timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "production and mppi" 2>&1 | tail -4 | tee $O/pytest_dsum.txt
echo done
