#!/usr/bin/env bash
# Build the experiment branches' kernels next to the main extension, so that one gpurun call can A/B them:
#   bash tools/build_variants.sh            # -> pipegoose_b200/_C_pdl.so, pipegoose_b200/_C_coresident.so
#   PIPEGOOSE_B200_EXT=pdl PIPEGOOSE_B200_PDL=1 python bench.py --gpus 1
#   PIPEGOOSE_B200_EXT=coresident PIPEGOOSE_B200_DP_OVERLAP_CTAS=-148 python bench.py --gpus 2 --tp 1
# The branches only change csrc/; the Python of the working tree is used with every variant.  The .so files are
# git-ignored but travel with the gpurun snapshot.  Runs on the CPU-only box (nvcc cross-compiles sm_100a).
set -euo pipefail
cd "$(dirname "$0")/.."
grep -qx ".wt/" .git/info/exclude 2>/dev/null || echo ".wt/" >> .git/info/exclude
for name in "${@:-pdl coresident}"; do
  for v in $name; do
    rm -rf ".wt/$v"; git worktree prune
    git worktree add -q ".wt/$v" "exp/$v"
    # experiment branches may be behind main: take main's host glue (bindings / build script), keep the branch's kernels
    (cd ".wt/$v" && python -c "
import sys; sys.path.insert(0, '.')
from pipegoose_b200.csrc.build import build
print('built', build(verbose=False))")
    cp ".wt/$v/pipegoose_b200/_C.so" "pipegoose_b200/_C_$v.so"
    git worktree remove --force ".wt/$v"
    echo "variant $v -> pipegoose_b200/_C_$v.so"
  done
done
git worktree prune
