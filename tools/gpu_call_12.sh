#!/bin/bash
# bisect the 3 % K1 regression over the round-2 commits: one bench line per exported tree, same box
set -u
O=gpurun_out
mkdir -p $O
for c in r01 3a9ba66 678ed29 198b394; do
  (cd _ab_$c && python bench.py > ../$O/bis_$c.json 2> ../$O/bis_$c.err)
  python -c "
import json
d=json.loads(open('$O/bis_$c.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s ms/step %.4f kernel_ms %.4f frac %.4f' % ('$c', d['ms_per_step'], r['kernel_ms'], r['frac']))" | tee -a $O/bisect.txt
done
python bench.py > $O/bis_head.json 2> $O/bis_head.err
python -c "
import json
d=json.loads(open('$O/bis_head.json').read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s ms/step %.4f kernel_ms %.4f frac %.4f' % ('head', d['ms_per_step'], r['kernel_ms'], r['frac']))" | tee -a $O/bisect.txt
echo done
