#!/bin/bash
# A/B runs on ONE box need an older tree next to the current one: export a commit into _ab_<tag>/ (git-ignored, travels to the GPU box
# with the snapshot) and build its library.  The round-2 A/B scripts (tools/gpu_call_{7..14}.sh, profiles/k1_ab_r02.txt) ran against
#     bash tools/export_tree.sh 9401463 r01        # the round-1 tree -> _ab_r01/
set -eu
cd "$(dirname "$0")/.."
commit=$1; tag=$2
rm -rf "_ab_$tag"; mkdir -p "_ab_$tag"
git archive "$commit" | tar -x -C "_ab_$tag"
cp -f MEASURED_PEAKS.json "_ab_$tag/" 2>/dev/null || true
(cd "_ab_$tag" && python -m control_toolkit_b200.build > /dev/null && echo "built _ab_$tag/control_toolkit_b200/libctk_b200.so")
